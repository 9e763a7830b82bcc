"""Host-side cost of one step (enqueue only) next to the device time of the same step: is the loop host-bound?

Every step is timed on the host from call to return with the GPU IDLE at the start (a synchronize before each call), so no launch
ever blocks on a full queue or on the pinned-buffer ring: the number is the pure host cost of GridMapBuilder.step(lazy=True) +
forward('navigation') (input staging, vpid tables, one graph launch).  The device time is the graph-replay loop of bench.py."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Step
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
step = Step(dev, seed=0)
for _ in range(5):
    step.run_resident()
torch.cuda.synchronize()


def host_cost(fn, n=100):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    return statistics.median(ts) * 1e3, min(ts) * 1e3


for name, fn in (("resident (inputs in HBM)", step.run_resident), ("serial e2e (host inputs, incl. the blocking .cpu() of the logits)", step.run_e2e)):
    med, best = host_cost(fn)
    print("host time per step, %s: median %.3f ms, min %.3f ms" % (name, med, best))
n = 200
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
for _ in range(n):
    step.run_resident()
e.record(); torch.cuda.synchronize()
print("device-bound loop: %.3f ms/step" % (s.elapsed_time(e) / n))
import cProfile, pstats
pr = cProfile.Profile()
for _ in range(50):
    torch.cuda.synchronize()
    pr.enable(); step.run_resident(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

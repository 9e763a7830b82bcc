"""Host-side time of one resident step (enqueue only) vs device time: is the loop host-bound?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Step
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
step = Step(dev, seed=0)
for _ in range(5):
    step.run_resident()
torch.cuda.synchronize()
import cProfile, pstats
n = 200
torch.cuda._sleep(2_000_000_000 // 4)          # park the GPU ~0.25 s so that the host runs ahead
t0 = time.perf_counter()
for _ in range(n):
    step.run_resident()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue per step: %.3f ms; total incl. drain %.3f ms/step" % ((t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    step.run_resident()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

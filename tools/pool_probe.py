"""Per-CTA timeline of one gridmm_pool launch at the bench shape (B = 32, T = 8): setup cycles, rows, episode switches, per-role
cycle counters, globaltimer entry / exit; and the kernel time by several timing methods.  Output: gpurun_out/pool_probe.csv + stdout.

    python tools/pool_probe.py [tag]
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Step, B
from gridmm_b200 import ops, _lib
from gridmm_b200.env import GridBatch

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
lib = _lib.load()
lib.gridmm_debug_set_pool_counters.argtypes = [ctypes.c_void_p]
step = Step(dev, seed=0); step.model.use_cuda_graph = False
step.run_resident(); torch.cuda.synchronize()
m = step.model; grid = GridBatch(step.builder)
pooled = m.buf("pooled16", (B * 196, 768), torch.float16, zero=True)
text_ws = ops.pool_text_ws(dev, B, 768)
if os.environ.get("POOL_COST") or os.environ.get("POOL_SNAP"):
    lib.gridmm_debug_set_pool_plan.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.gridmm_debug_set_pool_plan(int(os.environ.get("POOL_COST", -1)), int(os.environ.get("POOL_SNAP", -1)))
    print("plan: episode cost %s, snap %s" % (os.environ.get("POOL_COST"), os.environ.get("POOL_SNAP")))
if os.environ.get("POOL_EXP"):
    lib.gridmm_debug_set_pool_exp.argtypes = [ctypes.c_int]
    lib.gridmm_debug_set_pool_exp(int(os.environ["POOL_EXP"]))
    print("experiment", os.environ["POOL_EXP"], "(results are garbage)")
plan_ws = ops.pool_plan(grid.cell_start, 196, B, 768)
fn = lambda: ops.pool(grid.slab, 768, grid.slots, grid.t_cap, grid.slot_rows, grid.view_rows, grid.tok_off, grid.perm, grid.cap,
                      grid.cell_start, grid.cell_rank, 196, None, 80, B, pooled, text_ws=text_ws, text_ws_ready=True,
                      pool_ws_buf=plan_ws, plan_ready=True)
ref = pooled.clone()
fn(); torch.cuda.synchronize()
print("equal to the step's result:", bool(torch.equal(pooled, ref)), " max |diff| %.3e" % (pooled.float() - ref.float()).abs().max().item())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nv = int(grid.cell_start[:, -1].sum().item())


def cold_l2():
    """L2 as the step leaves it for the pooling kernel: the features evicted, the text operand (written by the text_proj GEMM just
    before) and the cell tables resident"""
    flush.zero_()
    text_ws.float().sum(); grid.cell_start.sum(); grid.cell_rank.sum(); plan_ws.sum()


def ev_time(n, cold):
    best = 1e30
    for _ in range(5):
        if cold:
            cold_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / n * 1e3)
    return best


def graph_time(n, cold):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        if cold:
            cold_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / n * 1e3)
    return best


plan_fn, fn_pool = (lambda: ops.pool_plan(grid.cell_start, 196, B, 768)), fn
fn = plan_fn
print("plan kernel: %.1f us (events, 8 launches) | %.1f us (graph, 8 launches)" % (ev_time(8, False), graph_time(8, False)))
fn = fn_pool
rec = plan_ws.cpu().numpy().view("int32")[:148 * 8].reshape(148, 8)
print("plan: CTAs with a head piece %d, tail piece %d, inside one cell %d; longest chain %d" %
      (int((rec[:, 7] & 1).sum()), int((rec[:, 7] & 2 > 0).sum()), int((rec[:, 7] & 4 > 0).sum()), int(max(rec[:, 4].max(), rec[:, 6].max()))))
for n in (1, 8):
    print("events, %d launches: warm %.1f us, flushed %.1f us | graph: warm %.1f us, flushed %.1f us" %
          (n, ev_time(n, False), ev_time(n, True), graph_time(n, False), graph_time(n, True)), flush=True)
dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
for rep in range(2):
    cold_l2(); torch.cuda.synchronize()
    lib.gridmm_debug_set_pool_counters(dbg.data_ptr())
    fn(); torch.cuda.synchronize()
    lib.gridmm_debug_set_pool_counters(None)
d = dbg.cpu()
vb = torch.zeros(B + 1, dtype=torch.int64); vb[1:] = grid.cell_start[:, -1].cpu().cumsum(0)
setup = d[:, 13] & 0xfffff; g1 = (d[:, 13] >> 20) & 0xfffff; g0 = (d[:, 13] >> 40) & 0xfffff
t0 = d[:, 14] - d[:, 14].min(); t1 = d[:, 15] - d[:, 14].min()
names = ["prod_tot", "prod_wait_empty", "prod_text", "mma_tot", "mma_wait_afull", "mma_wait_dempty", "red_tot", "red_wait_dfull",
         "red_text", "red_softmax", "pool_tot", "pool_wait", "pool_loop"]
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/pool_%s.csv" % tag, "w") as f:
    f.write("cta,g0,g1,rows,episodes,setup_cyc,entry_ns,exit_ns," + ",".join(names) + "\n")
    for c in range(148):
        eps = int(((vb[:-1] < g1[c]) & (vb[1:] > g0[c])).sum())
        f.write("%d,%d,%d,%d,%d,%d,%d,%d," % (c, g0[c], g1[c], g1[c] - g0[c], eps, setup[c], t0[c], t1[c]) +
                ",".join(str(int(d[c, i])) for i in range(13)) + "\n")
print("valid rows %d; in-kernel span (first entry -> last exit) %.1f us; entry spread %.1f us; exit: min %.1f mean %.1f max %.1f us"
      % (nv, t1.max().item() / 1e3, t0.max().item() / 1e3, t1.min().item() / 1e3, t1.float().mean().item() / 1e3, t1.max().item() / 1e3))
print("setup cycles: mean %.0f max %.0f" % (setup.float().mean().item(), setup.max().item()))
df = d.float()
print(" | ".join("%s %.0f/%.0f" % (n_, df[:, i].mean().item(), df[:, i].max().item()) for i, n_ in enumerate(names)))
rows = (g1 - g0).float()
print("rows per CTA: mean %.0f min %.0f max %.0f" % (rows.mean().item(), rows.min().item(), rows.max().item()))

# stage hand-over trace of CTAs 0..3 (cycles relative to the CTA's first stamp)
lib.gridmm_debug_set_pool_trace.argtypes = [ctypes.c_void_p]
tr = torch.zeros(4, 64, 8, dtype=torch.int64, device=dev)
cold_l2(); torch.cuda.synchronize()
lib.gridmm_debug_set_pool_trace(tr.data_ptr()); fn(); torch.cuda.synchronize(); lib.gridmm_debug_set_pool_trace(None)
tr = tr.cpu()
import numpy as np
np.save("gpurun_out/pool_trace_%s.npy" % tag, tr.numpy())
for c in range(2):
    x = tr[c]; n = int((x[:, 7] > 0).sum()); t0 = int(x[0, 0])
    print("CTA %d: %d tiles; columns: slot free seen | copies issued | tile landed seen | MMAs issued | relevance ready seen | numerators | weights | pooled" % (c, n))
    for i in list(range(min(n, 12))) + list(range(max(12, n - 3), n)):
        print("  tile %2d: " % i + " ".join("%7d" % (int(v) - t0) for v in x[i]))

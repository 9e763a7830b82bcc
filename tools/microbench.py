"""Per-kernel timings + in-kernel cycle counters (role totals / mbarrier waits) for the GEMM and pooling kernels."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Step, B, T
from gridmm_b200 import ops, _lib

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.load()
lib.gridmm_debug_set_gemm_counters.argtypes = [ctypes.c_void_p]
lib.gridmm_debug_set_pool_counters.argtypes = [ctypes.c_void_p]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(2_000_000)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3   # us


def gemm_case(M, N, K, act=0, res=False, f32=False):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) * 0.02).half()
    bias = torch.zeros(N, device=dev)
    o16 = None if f32 else torch.empty(M, N, device=dev, dtype=torch.float16)
    o32 = torch.randn(M, N, device=dev) if (f32 or res) else None
    fn = lambda: ops.linear(a, w, bias=bias, residual=o32 if res else None, out_f32=o32, out_f16=o16, act=act)
    us = timeit(fn)
    dbg = torch.zeros(148, 8, dtype=torch.int64, device=dev)
    lib.gridmm_debug_set_gemm_counters(dbg.data_ptr())
    fn(); torch.cuda.synchronize()
    lib.gridmm_debug_set_gemm_counters(None)
    d = dbg.float()
    d = d[d[:, 2] > 0].mean(0)
    tf = 2.0 * M * N * K / us / 1e6
    print("GEMM M=%5d N=%5d K=%5d act=%d res=%d: %7.1f us %7.1f TF | prod tot %6.0f wait_empty %6.0f | mma tot %6.0f wait_full %6.0f wait_acc %6.0f | epi tot %6.0f wait_tmem %6.0f"
          % (M, N, K, act, res, us, tf, d[0], d[1], d[2], d[3], d[4], d[5], d[6]), flush=True)


lib.gridmm_debug_set_gemm_pairs.argtypes = [ctypes.c_int]
if "gemm" in sys.argv or len(sys.argv) == 1:
    for pairs in (0, 1):
        lib.gridmm_debug_set_gemm_pairs(pairs)
        print("--- CTA pairs (cta_group::2):", "on" if pairs else "off", flush=True)
        gemm_case(6912, 2304, 768)
        gemm_case(6912, 3072, 768, act=1)
        gemm_case(6912, 768, 3072, res=True)
        gemm_case(6912, 768, 768, res=True)
        gemm_case(9472, 6144, 768)
        gemm_case(8192, 8192, 8192)
    lib.gridmm_debug_set_gemm_pairs(1)
if "gemm_all" in sys.argv:
    gemm_case(6912, 2304, 768)
    gemm_case(6912, 3072, 768, act=1)
    gemm_case(6912, 768, 3072, res=True)
    gemm_case(6912, 768, 768, res=True)
    gemm_case(9472, 6144, 768)
    gemm_case(1824, 768, 768)
    gemm_case(1824, 2304, 768)
    gemm_case(1824, 3072, 768, act=1)
    gemm_case(1824, 768, 3072, res=True)
    gemm_case(8192, 8192, 8192)

if "pool" in sys.argv or len(sys.argv) == 1:
    step = Step(dev, seed=0)
    step.model.use_cuda_graph = False
    step.run_resident(); torch.cuda.synchronize()
    m = step.model; g = step.builder
    from gridmm_b200.env import GridBatch
    grid = GridBatch(g)
    tp16 = m.buf("tp16", (B * 80, 768), torch.float16)
    pooled = m.buf("pooled16", (B * 196, 768), torch.float16, zero=True)
    for ctas in (148, 296):
        fn = lambda: ops.pool(grid.slab, 768, grid.slots, grid.t_cap, grid.slot_rows, grid.view_rows, grid.tok_off, grid.perm, grid.cap,
                              grid.cell_start, grid.cell_rank, 196, tp16, 80, B, pooled, num_ctas=min(ctas, 148))
        us = timeit(fn, 10)
        nv = int(grid.cell_start[:, -1].sum().item())
        print("pool: %.1f us, %.0f GB/s (valid rows %d)" % (us, nv * 768 * 2 / us / 1e3, nv), flush=True)
        break
    lib.gridmm_debug_set_pool_mode.argtypes = [ctypes.c_int]
    for mode in (0, 1, 2, 3):
        lib.gridmm_debug_set_pool_mode(mode)
        us = timeit(fn, 10)
        dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
        lib.gridmm_debug_set_pool_counters(dbg.data_ptr())
        fn(); torch.cuda.synchronize()
        lib.gridmm_debug_set_pool_counters(None)
        d = dbg.float().mean(0)
        print("mode %d: %.1f us | gather0 tot %.0f wait_empty %.0f land %.0f | gather1 tot %.0f wait_empty %.0f land %.0f | "
              "mma tot %.0f wait_afull %.0f wait_dempty %.0f red %.0f | epi tot %.0f wait_dfull %.0f text %.0f | pool tot %.0f wait_pfull %.0f loop %.0f"
              % ((mode, us) + tuple(d[:16].tolist())), flush=True)
    lib.gridmm_debug_set_pool_mode(0)

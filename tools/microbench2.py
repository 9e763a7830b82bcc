"""Device-time microbenchmarks, timed by replaying a CUDA graph of N back-to-back launches (no host launch latency in the
numbers): every distinct GEMM / attention / LayerNorm shape of the B=32, T=8 navigation step, and the pooling kernel with
per-CTA cycle statistics.

    python tools/microbench2.py [gemm] [attn] [ln] [pool]
"""
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
want = set(sys.argv[1:]) or {"gemm", "gemmln", "attn", "ln", "pool"}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_time(fn, n=20, reps=5, cold=False):
    """us per launch: n launches captured in one graph, replayed `reps` times (best), optionally with an L2 flush before each
    replay (then n should be 1)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / n * 1e3)
    return best


def gemm_case(M, N, K, act=0, res=False, f32=False, tag=""):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) * 0.02).half()
    bias = torch.zeros(N, device=dev)
    o16 = None if (f32 or res) else torch.empty(M, N, device=dev, dtype=torch.float16)
    o32 = torch.randn(M, N, device=dev) if (f32 or res) else None
    fn = lambda: ops.linear(a, w, bias=bias, residual=o32 if res else None, out_f32=o32, out_f16=o16, act=act)
    us = graph_time(fn)
    us1 = graph_time(fn, n=1, reps=5)
    tf = 2.0 * M * N * K / us / 1e6
    print("GEMM %-22s M=%5d N=%5d K=%5d act=%d res=%d f32=%d: %7.1f us (x20) %7.1f us (single) %7.1f TF"
          % (tag, M, N, K, act, res, f32, us, us1, tf), flush=True)
    return us


if "gemm" in want:
    S, Q, KC, L = 216, 57, 296, 80
    tot = 0.0
    tot += gemm_case(B * L, 768, 768, tag="text_proj")
    tot += gemm_case(B * 196, 768, 768, f32=True, tag="grid_proj")
    tot += gemm_case(B * S, 2304, 768, tag="map qkv") * 2
    tot += gemm_case(B * S, 768, 768, res=True, tag="map out-proj") * 3
    tot += gemm_case(B * S, 768, 768, tag="map xattn q")
    tot += gemm_case(B * S, 3072, 768, act=1, tag="map ffn1") * 2
    tot += gemm_case(B * S, 768, 3072, res=True, tag="map ffn2") * 2
    tot += gemm_case(B * L, 1536, 768, tag="txt kv")
    tot += gemm_case(B * KC, 6144, 768, tag="fusion kv x4")
    tot += gemm_case(B * Q, 768, 768, tag="x q") * 4
    tot += gemm_case(B * Q, 768, 768, res=True, tag="x out-proj") * 8
    tot += gemm_case(B * Q, 2304, 768, tag="x qkv") * 4
    tot += gemm_case(B * Q, 3072, 768, act=1, tag="x ffn1") * 4
    tot += gemm_case(B * Q, 768, 3072, res=True, tag="x ffn2") * 4
    tot += gemm_case(B * 20, 768, 2304, act=2, f32=True, tag="cls head G") * 2
    tot += gemm_case(B * 37, 768, 2304, act=2, f32=True, tag="cls head V")
    tot += gemm_case(B, 768, 4608, act=2, f32=True, tag="cls fuse")
    print("GEMM sum over the step's 42 launches: %.1f us" % tot, flush=True)
    # CTA pairs (cta_group::2, 256x256 tiles) against single-CTA tiles on the shapes where the choice is close
    lib.gridmm_debug_set_gemm_pairs.argtypes = [ctypes.c_int]
    for on in (0, 1):
        lib.gridmm_debug_set_gemm_pairs(on)
        print("--- CTA pairs %s" % ("on" if on else "off"), flush=True)
        gemm_case(B * Q, 3072, 768, act=1, tag="x ffn1"); gemm_case(B * Q, 2304, 768, tag="x qkv")
        gemm_case(B * S, 768, 768, tag="map xattn q"); gemm_case(B * L, 1536, 768, tag="txt kv")

    # 256 x 384 pair tiles (opt-in) on the two shapes they were written for
    lib.gridmm_debug_set_gemm_384.argtypes = [ctypes.c_int]
    for on in (1, 0):
        lib.gridmm_debug_set_gemm_384(on)
        print("--- 256x384 pair tiles %s" % ("on" if on else "off"), flush=True)
        gemm_case(B * Q, 3072, 768, act=1, tag="x ffn1"); gemm_case(B * Q, 2304, 768, tag="x qkv")
    # the packed map sequence at this workload's row count (~4500 of 6912 padded rows)
    print("--- map-sized GEMMs over 4480 rows (packed map sequence)", flush=True)
    gemm_case(4480, 2304, 768, tag="map qkv (packed)"); gemm_case(4480, 3072, 768, act=1, tag="map ffn1 (packed)")
    gemm_case(4480, 768, 768, tag="map xattn q (packed)")

if "gemmln" in want:
    def ln_case(M, K, tag):
        a = torch.randn(M, K, device=dev).half(); w = (torch.randn(768, K, device=dev) * 0.02).half()
        bias = torch.zeros(768, device=dev); g_ = torch.ones(768, device=dev); b_ = torch.zeros(768, device=dev)
        x32 = torch.randn(M, 768, device=dev); x16 = torch.empty(M, 768, device=dev, dtype=torch.float16)
        us = graph_time(lambda: ops.linear_ln(a, w, bias, x32, g_, b_, 1e-12, out_f32=x32, out_f16=x16))
        print("GEMM+LN %-14s M=%5d K=%5d: %6.1f us  %6.1f TF" % (tag, M, K, us, 2.0 * M * 768 * K / us / 1e6), flush=True)
        return us
    tot = ln_case(B * 216, 768, "map out-proj") * 3 + ln_case(B * 216, 3072, "map ffn2") * 2 + ln_case(B * 57, 768, "x out-proj") * 8 + ln_case(B * 57, 3072, "x ffn2") * 4
    print("GEMM+LN sum over the step's 17 launches: %.1f us" % tot, flush=True)
    lib.gridmm_debug_set_ln_cluster.argtypes = [ctypes.c_int]
    for cl in (2, 4, 6):
        lib.gridmm_debug_set_ln_cluster(cl)
        print("--- forced cluster %d%s" % (cl, " (two CTA pairs)" if cl == 4 else ""), flush=True)
        ln_case(B * 216, 768, "map out-proj"); ln_case(B * 216, 3072, "map ffn2"); ln_case(B * 57, 768, "x out-proj"); ln_case(B * 57, 3072, "x ffn2")
    lib.gridmm_debug_set_ln_cluster(0)

if "attn" in want:
    def attn_case(Sq, Sk, tag):
        q = torch.randn(B * Sq, 768, device=dev).half(); k = torch.randn(B * Sk, 768, device=dev).half()
        v = torch.randn(B * Sk, 768, device=dev).half(); o = torch.empty(B * Sq, 768, device=dev, dtype=torch.float16)
        m = torch.ones(B, Sk, dtype=torch.uint8, device=dev)
        fn = lambda: ops.attention(q, k, v, o, m, -10000.0, B, 12, Sq, Sk)
        us = graph_time(fn)
        fl = 4.0 * B * 12 * Sq * Sk * 64
        print("ATTN %-14s Sq=%3d Sk=%3d: %6.1f us  %6.1f TF" % (tag, Sq, Sk, us, fl / us / 1e6), flush=True)
        return us
    lib.gridmm_debug_set_attn_legacy.argtypes = [ctypes.c_int]
    for legacy in (1, 2, 3, 0):
        lib.gridmm_debug_set_attn_legacy(legacy)
        print("--- attention kernel:", {1: "mma.sync", 2: "tcgen05", 3: "dispatch by shape, tcgen05 head-pair kernel for <= 64 queries",
                                        0: "dispatch by shape (default)"}[legacy], flush=True)
        attn_case(57, 208, "x cross packed")           # the step's packed context averages 193 keys: head-pair kernel when dispatched by shape
        tot = attn_case(216, 216, "map self") * 2 + attn_case(216, 80, "map x txt") + attn_case(57, 296, "x cross") * 4 + attn_case(57, 57, "x self") * 4
        print("ATTN sum over the step's 11 launches: %.1f us" % tot, flush=True)

if "ln" in want:
    for rows in (B * 216, B * 57):
        x = torch.randn(rows, 768, device=dev); g_ = torch.ones(768, device=dev); b_ = torch.zeros(768, device=dev)
        o32 = torch.empty_like(x); o16 = torch.empty(rows, 768, device=dev, dtype=torch.float16)
        us = graph_time(lambda: ops.layernorm(x, g_, b_, 1e-12, out_f32=o32, out_f16=o16))
        print("LN rows=%5d: %6.1f us  (%.0f GB/s of 10 B/elt)" % (rows, us, rows * 768 * 10 / us / 1e3), flush=True)

if "pool" in want:
    if os.environ.get("GRIDMM_POOL_SPLIT") in ("0", "1"):
        lib.gridmm_debug_set_pool_split.argtypes = [ctypes.c_int]
        lib.gridmm_debug_set_pool_split(int(os.environ["GRIDMM_POOL_SPLIT"]))
        print("pool: split weights =", os.environ["GRIDMM_POOL_SPLIT"])
    step = Step(dev, seed=0)
    step.model.use_cuda_graph = False
    step.run_resident(); torch.cuda.synchronize()
    m = step.model; g = step.builder
    from gridmm_b200.env import GridBatch
    grid = GridBatch(g)
    ref_pooled = m.buf("pooled16", (B * 196, 768), torch.float16, zero=True).clone()
    pooled = m.buf("pooled16", (B * 196, 768), torch.float16, zero=True)
    text_ws = ops.pool_text_ws(dev, B, 768)          # filled by the step above (text_proj epilogue)
    plan_ws = ops.pool_plan(grid.cell_start, 196, B, 768)      # the step runs it behind the grid update, off the critical path
    fn = lambda: ops.pool(grid.slab, 768, grid.slots, grid.t_cap, grid.slot_rows, grid.view_rows, grid.tok_off, grid.perm, grid.cap,
                          grid.cell_start, grid.cell_rank, 196, None, 80, B, pooled, text_ws=text_ws, text_ws_ready=True,
                          pool_ws_buf=plan_ws, plan_ready=True)
    fn(); torch.cuda.synchronize()
    print("pool: re-run equals the step's result:", bool(torch.equal(pooled, ref_pooled)))
    nv = int(grid.cell_start[:, -1].sum().item())
    cs = grid.cell_start.cpu()
    sizes = (cs[:, 1:] - cs[:, :-1]).flatten()
    print("pool: valid rows %d; cell sizes: max %d, mean(nonempty) %.1f, cells>256 rows: %d" %
          (nv, int(sizes.max()), float(sizes[sizes > 0].float().mean()), int((sizes > 256).sum())), flush=True)
    us_w = graph_time(fn, n=1, reps=5, cold=False)
    us_c = graph_time(fn, n=1, reps=5, cold=True)
    dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    lib.gridmm_debug_set_pool_counters(dbg.data_ptr())
    fn(); torch.cuda.synchronize()
    lib.gridmm_debug_set_pool_counters(None)
    d = dbg.float()
    names = ["prod tot", "prod wait_empty", "prod text", "mma tot", "mma wait_afull", "mma wait_dempty", "red tot", "red wait_dfull",
             "red text", "red max+softmax", "pool tot", "pool wait", "pool loop"]
    print("pool: %.1f us L2-warm, %.1f us after L2 flush (%.0f GB/s feature bytes)" % (us_w, us_c, nv * 1536 / us_c / 1e3), flush=True)
    print("   " + " | ".join("%s mean %.0f max %.0f" % (n_, d[:, i].mean().item(), d[:, i].max().item()) for i, n_ in enumerate(names)), flush=True)

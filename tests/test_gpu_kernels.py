"""GPU: each sm_100a kernel, called through the C ABI (gridmm_b200/_lib.py -> include/gridmm_b200.h), against a plain
torch fp32 reference of the same op (floating point) or the oracle (integer grid work, bit-exact)."""
import math

import numpy as np
import pytest
import torch

from gridmm_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 768, 768), (6912, 2304, 768), (1000, 768, 3072), (32, 768, 1536),
                                   (6272, 768, 768), (9472, 6144, 768), (6913, 768, 3072), (12801, 256, 192),
                                   (1824, 3072, 768), (912, 3072, 768), (1824, 2304, 768), (912, 2304, 768), (1025, 768, 3072),
                                   (257, 1152, 128)])
def test_linear_tcgen05_plain(M, N, K):
    """The last rows of shapes include the 57-query GEMMs of the fusion encoder (whole batch and half batches)."""
    from gridmm_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).half().to(_dev())
    w = (torch.randn(N, K, generator=g) * 0.05).half().to(_dev())
    bias = torch.randn(N, generator=g).to(_dev())
    o32 = torch.empty(M, N, device=_dev())
    o16 = torch.empty(M, N, device=_dev(), dtype=torch.float16)
    ops.linear(a, w, bias=bias, out_f32=o32, out_f16=o16)
    ref = a.float() @ w.float().t() + bias
    torch.cuda.synchronize()
    err = (o32 - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err      # fp32 accumulate, order differs from cuBLAS only
    assert (o16.float() - ref).abs().max().item() < 4e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_epilogues(act):
    from gridmm_b200 import ops
    M, N, K = 517, 256, 192
    g = torch.Generator().manual_seed(act)
    a = torch.randn(M, K, generator=g).half().to(_dev())
    w = (torch.randn(N, K, generator=g) * 0.1).half().to(_dev())
    bias = torch.randn(N, generator=g).to(_dev())
    res = torch.randn(M, N, generator=g).to(_dev())
    out = res.clone()
    ops.linear(a, w, bias=bias, residual=out, out_f32=out, act=act)       # in-place residual, as the model uses it
    y = a.float() @ w.float().t() + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = torch.relu(y)
    ref = y + res
    torch.cuda.synchronize()
    assert (out - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("B,L", [(32, 80), (3, 37), (1, 128), (3, 136), (2, 200), (4, 250), (1, 256)])
def test_linear_lanes_layout(B, L):
    """text_proj epilogue that writes gridmm_pool's lane-major operand: unit u of (b, t) at ((b*96 + u)*128 + t) * 8 halves;
    positions 128.. of a long text (--max_instr_len 200 / 250) in a second [B, 96, 128] block."""
    from gridmm_b200 import ops
    g = torch.Generator().manual_seed(B * 131 + L)
    a = torch.randn(B * L, 768, generator=g).half().to(_dev())
    w = (torch.randn(768, 768, generator=g) * 0.05).half().to(_dev())
    bias = torch.randn(768, generator=g).to(_dev())
    nblk = (L + 127) // 128
    ws = torch.zeros(nblk * B * 128 * 768, dtype=torch.float16, device=_dev())
    ops.linear_lanes(a, w, bias, ws, L)
    ref16 = torch.empty(B * L, 768, device=_dev(), dtype=torch.float16)
    ops.linear(a, w, bias=bias, out_f16=ref16)
    torch.cuda.synchronize()
    # [blk, b, u, t, 8] -> [b, blk * 128 + t, u * 8]
    full = ws.view(nblk, B, 96, 128, 8).permute(1, 0, 3, 2, 4).reshape(B, nblk * 128, 768)
    got = full[:, :L].reshape(B * L, 768)
    assert torch.equal(got, ref16)                                   # same kernel, same rounding: bitwise
    assert full[:, L:].abs().max().item() == 0 if L < nblk * 128 else True
    ref = a.float() @ w.float().t() + bias
    assert (got.float() - ref).abs().max().item() < 4e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,K,cl,raw", [(6912, 768, 2, False), (6912, 3072, 2, False), (1824, 768, 6, False), (1824, 3072, 6, True),
                                        (300, 768, 0, True), (2560, 768, 0, False), (6913, 768, 0, False), (57, 3072, 2, True),
                                        (6912, 768, 4, False), (6912, 3072, 4, True), (300, 768, 4, False), (6913, 768, 4, True)])
def test_linear_ln_fused(M, K, cl, raw):
    """GEMM + bias + residual + LayerNorm in one cluster kernel vs torch fp32 (in-place residual stream, as the model uses it)."""
    import ctypes
    from gridmm_b200 import ops, _lib
    lib = _lib.load()
    lib.gridmm_debug_set_ln_cluster.argtypes = [ctypes.c_int]
    g = torch.Generator().manual_seed(M + K + cl)
    a = torch.randn(M, K, generator=g).half().to(_dev())
    w = (torch.randn(768, K, generator=g) * 0.03).half().to(_dev())
    bias = torch.randn(768, generator=g).to(_dev())
    res = (torch.randn(M, 768, generator=g) * 2 + 0.5).to(_dev())
    gamma = (1 + 0.1 * torch.randn(768, generator=g)).to(_dev()); beta = (0.1 * torch.randn(768, generator=g)).to(_dev())
    x32 = res.clone()
    x16 = torch.empty(M, 768, device=_dev(), dtype=torch.float16)
    lib.gridmm_debug_set_ln_cluster(cl)
    try:
        ops.linear_ln(a, w, bias, x32, gamma, beta, 1e-12, out_f32=x32, out_f16=x16, f32_raw=raw)
    finally:
        lib.gridmm_debug_set_ln_cluster(0)
    v = a.float() @ w.float().t() + bias + res
    y = torch.nn.functional.layer_norm(v, (768,), gamma, beta, 1e-12)
    torch.cuda.synchronize()
    assert (x32 - (v if raw else y)).abs().max().item() < (4e-3 if raw else 2e-3)
    assert (x16.float() - y).abs().max().item() < 6e-3


def test_linear_strided_views():
    """A operand / outputs addressed through row pitches (fused QKV buffers)."""
    from gridmm_b200 import ops
    g = torch.Generator().manual_seed(5)
    big = torch.randn(260, 3 * 768, generator=g).half().to(_dev())
    w = (torch.randn(768, 768, generator=g) * 0.05).half().to(_dev())
    out = torch.zeros(260, 1536, device=_dev(), dtype=torch.float16)
    ops.linear(big[:, 768:1536], w, out_f16=out[:, 768:])
    ref = big[:, 768:1536].float() @ w.float().t()
    torch.cuda.synchronize()
    assert (out[:, 768:].float() - ref).abs().max().item() < 2e-2
    assert out[:, :768].abs().max().item() == 0


@pytest.mark.parametrize("legacy", [2, 1, 0, 3])
@pytest.mark.parametrize("B,Sq,Sk,neg", [(2, 216, 216, float("-inf")), (3, 57, 296, -10000.0), (2, 64, 80, -10000.0), (1, 5, 7, -10000.0),
                                         (2, 130, 320, -10000.0), (1, 40, 400, float("-inf")), (4, 57, 57, -10000.0),
                                         (3, 57, 221, -10000.0), (2, 64, 256, float("-inf")), (2, 33, 250, -10000.0)])
def test_attention(B, Sq, Sk, neg, legacy):
    """legacy = 2: the tcgen05 kernel (Sk <= 320; Sk = 400 exercises its fall-back), 1: the mma.sync kernel, 0: dispatch by shape,
    3: dispatch by shape with the tcgen05 head-pair kernel for Sq <= 64 and Sk <= 256."""
    import ctypes
    from gridmm_b200 import ops, _lib
    lib = _lib.load()
    lib.gridmm_debug_set_attn_legacy.argtypes = [ctypes.c_int]
    lib.gridmm_debug_set_attn_legacy(legacy)
    g = torch.Generator().manual_seed(Sq * 7 + Sk)
    q = torch.randn(B * Sq, 768, generator=g).half().to(_dev())
    kv = torch.randn(B * Sk, 1536, generator=g).half().to(_dev())
    lens = torch.randint(1, Sk + 1, (B,), generator=g)
    lens[0] = Sk
    kmask = (torch.arange(Sk)[None, :] < lens[:, None]).to(torch.uint8).to(_dev())
    out = torch.empty(B * Sq, 768, device=_dev(), dtype=torch.float16)
    try:
        ops.attention(q, kv[:, :768], kv[:, 768:], out, kmask, neg, B, 12, Sq, Sk)
    finally:
        lib.gridmm_debug_set_attn_legacy(0)
    qh = q.float().view(B, Sq, 12, 64).permute(0, 2, 1, 3)
    kh = kv[:, :768].float().view(B, Sk, 12, 64).permute(0, 2, 1, 3)
    vh = kv[:, 768:].float().view(B, Sk, 12, 64).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / 8.0
    add = torch.zeros(B, Sk, device=_dev()).masked_fill(kmask == 0, neg)[:, None, None, :]
    ref = (torch.softmax(s + add, -1) @ vh).permute(0, 2, 1, 3).reshape(B * Sq, 768)
    torch.cuda.synchronize()
    assert (out.float() - ref).abs().max().item() < 6e-3      # P and the output are rounded to fp16


def test_layernorm_and_rows():
    from gridmm_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(1001, 768, generator=g) * 3 + 1).to(_dev())
    gamma = torch.randn(768, generator=g).to(_dev())
    beta = torch.randn(768, generator=g).to(_dev())
    for eps in (1e-12, 1e-5):
        o32 = torch.empty_like(x)
        o16 = torch.empty(1001, 768, device=_dev(), dtype=torch.float16)
        ops.layernorm(x, gamma, beta, eps, out_f32=o32, out_f16=o16)
        ref = torch.nn.functional.layer_norm(x, (768,), gamma, beta, eps)
        torch.cuda.synchronize()
        assert (o32 - ref).abs().max().item() < 2e-5
        assert (o16.float() - ref).abs().max().item() < 1e-2
    # copy_rows: x viewed as [7, 143] rows per episode -> rows 3..12 of each to offset 20 of a 40-row sequence
    o32 = torch.zeros(7 * 40, 768, device=_dev())
    ops.copy_rows(x, 143, 3, 10, 7, 40, 20, out_f32=o32)
    torch.cuda.synchronize()
    assert torch.equal(o32.view(7, 40, 768)[:, 20:30], x.view(7, 143, 768)[:, 3:13])
    assert o32.view(7, 40, 768)[:, :20].abs().max().item() == 0


def test_pos_embed():
    from gridmm_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, G, S = 3, 11, 30
    feat = torch.randn(B * G, 7, generator=g).to(_dev())
    w = torch.randn(768, 7, generator=g).to(_dev()); b = torch.randn(768, generator=g).to(_dev())
    gamma = torch.randn(768, generator=g).to(_dev()); beta = torch.randn(768, generator=g).to(_dev())
    base = torch.randn(B * G, 768, generator=g).to(_dev())
    table = torch.randn(100, 768, generator=g).to(_dev())
    idx = torch.randint(0, 100, (B * G,), generator=g).to(_dev())
    out = torch.zeros(B * S, 768, device=_dev())
    ops.pos_embed(feat, w.t().contiguous(), b, gamma, beta, 1e-12, out, None, G, S, 19, base=base, table=table, idx=idx)
    ref = base + table[idx] + torch.nn.functional.layer_norm(feat @ w.t() + b, (768,), gamma, beta, 1e-12)
    torch.cuda.synchronize()
    assert (out.view(B, S, 768)[:, 19:] - ref.view(B, G, 768)).abs().max().item() < 2e-5


def _run_builder(ep, grid_w=14, geometry="r2r"):
    from gridmm_b200.env import GridMapBuilder
    B, T = ep["pos"].shape[:2]
    gb = GridMapBuilder(B, feat_dim=ep["clip"].shape[-1], grid_w=grid_w, geometry=geometry, max_steps=4)   # forces a regrow
    per_step = []
    for t in range(T):
        grid = gb.step(ep["depth_sub"][:, t], ep["clip"][:, t], ep["pos"][:, t], ep["heading"][:, t])
        torch.cuda.synchronize()
        per_step.append(grid.grid_map_numpy())
    return gb, grid, per_step


@pytest.mark.parametrize("case", H.GRID_CASES + [dict(seed=13, batch=8, steps=15)], ids=lambda c: "s%d" % c["seed"])
def test_grid_update_bit_exact(case):
    """cell ids bit-exact vs the oracle at every step (and, through tests/golden, vs the reference itself)."""
    import os
    ep = synth.make_episodes(case["batch"], case["steps"], seed=case["seed"], dim=768)
    cells, fts, halfs, pos = H.oracle_grid(ep)
    gb, grid, per_step = _run_builder(ep)
    gold_path = os.path.join(H.GOLD, "grid_r2r_s%d.npz" % case["seed"])
    gold = np.load(gold_path) if os.path.exists(gold_path) else None
    B, T = case["batch"], case["steps"]
    for t in range(T):
        for b in range(B):
            got = per_step[t][b]
            assert got.dtype == np.float64 and got.shape == (588 * (t + 1),)
            assert np.array_equal(got.astype(np.int32), cells[b][t]), "b=%d t=%d" % (b, t)
            if gold is not None:
                assert np.array_equal(got.astype(np.int16), gold["cell_b%d_t%d" % (b, t)])
    # window + polar features
    np.testing.assert_array_equal(grid.half_len.cpu().numpy(), np.array(halfs, dtype=np.float32))
    np.testing.assert_allclose(grid.pos_fts.cpu().numpy(), np.stack(pos), atol=2e-6, rtol=0)
    # sorted layout: perm is a stable sort of the valid points by cell
    perm = grid.perm.cpu().numpy(); cs = grid.cell_start.cpu().numpy(); cr = grid.cell_rank.cpu().numpy()
    ne = grid.n_nonempty.cpu().numpy()
    for b in range(B):
        c = cells[b][T - 1]
        valid = np.nonzero(c >= 0)[0]
        order = valid[np.argsort(c[valid], kind="stable")]
        assert np.array_equal(perm[b, :len(order)], order)
        counts = np.bincount(c[valid], minlength=196)
        assert np.array_equal(cs[b], np.concatenate([[0], np.cumsum(counts)]))
        rank = np.where(counts > 0, np.cumsum(counts > 0) - 1, -1)
        assert np.array_equal(cr[b], rank) and ne[b] == (counts > 0).sum()
    # features in the reference's row order
    got_fts = grid.grid_fts_torch()
    for b in range(B):
        assert torch.equal(got_fts[b].cpu(), torch.from_numpy(np.ascontiguousarray(fts[b])))


def test_grid_update_pretraining_paths_bit_exact():
    """SURVEY 8a row 19, grid half: whole ground-truth paths with the pretraining dataset's headings, against the golden cell ids
    its own getGlobalMap produced (tests/golden/grid_pretrain_s51.npz, pretrain_src/data/dataset.py:351-473) -- bit-exact at
    every step, plus gridmap_pos_fts and the target_patch_id label computed from the device's half_len."""
    import os
    from gridmm_b200.env import GridMapBuilder, target_patch_id
    case = H.PRETRAIN_GRID_CASE
    gold = np.load(os.path.join(H.GOLD, "grid_pretrain_s%d.npz" % case["seed"]))
    ep = synth.make_episodes(case["batch"], case["steps"], seed=case["seed"], dim=768)
    ep["heading"] = synth.pretrain_headings(ep)
    B, T = case["batch"], case["steps"]
    gb = GridMapBuilder(B, max_steps=4)
    for t in range(T):
        grid = gb.step(ep["depth_sub"][:, t], ep["clip"][:, t], ep["pos"][:, t], ep["heading"][:, t])
        torch.cuda.synchronize()
        cells = grid.grid_map_numpy()
        half = grid.half_len.cpu().numpy()
        pos_fts = grid.pos_fts.cpu().numpy()
        for b in range(B):
            assert np.array_equal(cells[b].astype(np.int16), gold["cell_b%d_t%d" % (b, t)]), "b=%d t=%d" % (b, t)
            np.testing.assert_allclose(pos_fts[b], gold["pos_fts_b%d" % b][t], atol=2e-6, rtol=0)
            nxt = ep["pos"][b, t + 1] if t + 1 < T else None
            assert target_patch_id(ep["pos"][b, t], nxt, float(ep["heading"][b, t]), half[b]) == int(gold["target_b%d" % b][t])
    # the one-call form the pretraining loader would use gives the same final state
    last = [c.copy() for c in cells]
    grid2 = gb.run_trajectory(ep["depth_sub"], ep["clip"], ep["pos"], ep["heading"])
    torch.cuda.synchronize()
    for b, c in enumerate(grid2.grid_map_numpy()):
        assert np.array_equal(c, last[b])
    fts = grid2.grid_fts_torch()
    for b in range(B):
        assert torch.equal(fts[b].cpu(), torch.from_numpy(np.ascontiguousarray(ep["clip"][b][:, :, 1:].reshape(-1, 768))))


def test_grid_update_8x8_config1():
    """BASELINE config 1: 8x8 grid, 512-d features, one viewpoint (the reference hard-codes 14/768; oracle is parametric)."""
    ep = synth.make_episodes(1, 1, seed=1, dim=512)
    cells, fts, halfs, pos = H.oracle_grid(ep, grid_w=8)
    gb, grid, per_step = _run_builder(ep, grid_w=8)
    assert np.array_equal(per_step[0][0].astype(np.int32), cells[0][0])
    assert per_step[0][0].max() <= 63
    np.testing.assert_allclose(grid.pos_fts.cpu().numpy()[0], pos[0], atol=2e-6, rtol=0)


@pytest.mark.parametrize("geometry", ["r2r_ce", "rxr_ce"])
def test_grid_update_ce_geometry(geometry):
    """Continuous-env conventions (Policy_ViewSelection_GridMap.py:632-641, 689-825): R2R-CE (90-degree camera, MAX_DIST 25) and
    RxR-CE (79 degrees, MAX_DIST 40)."""
    from oracle import grid_oracle as go
    geom = go.CEGeometry if geometry == "r2r_ce" else go.RxRCEGeometry
    ep = synth.make_episodes(4, 5, seed=4, dim=768)
    ep["depth_sub"] = (ep["depth_sub"].astype(np.float32) / 4000.0).astype(np.float32)       # CE depth is metres
    cells, fts, halfs, pos = H.oracle_grid(ep, geom=geom)
    gb, grid, per_step = _run_builder(ep, geometry=geometry)
    for t in range(5):
        for b in range(4):
            assert np.array_equal(per_step[t][b].astype(np.int32), cells[b][t]), "b=%d t=%d" % (b, t)
    np.testing.assert_allclose(grid.pos_fts.cpu().numpy(), np.stack(pos), atol=2e-6, rtol=0)


def _oracle_pool(fts, cell, tp16, n_cells=196):
    """float64 statement of vilmodel.py:797-807 in feature space, with the SAME fp16-rounded text_fts the kernel sees."""
    x = torch.from_numpy(np.ascontiguousarray(fts)).double()
    w = (x @ tp16.double().t()).max(-1)[0]
    out = torch.zeros(n_cells, x.shape[1], dtype=torch.float64)
    cell = torch.from_numpy(cell.astype(np.int64))
    for c in range(n_cells):
        sel = cell == c
        if sel.any():
            out[c] = (torch.softmax(w[sel], 0)[:, None] * x[sel]).sum(0)
    return out, w


@pytest.mark.parametrize("B,T,L,D,gw", [(3, 2, 80, 768, 14), (8, 8, 80, 768, 14), (2, 15, 40, 768, 14), (1, 1, 16, 512, 8), (37, 3, 24, 768, 14),
                                        (3, 2, 128, 768, 14), (3, 2, 129, 768, 14), (3, 3, 136, 768, 14), (5, 4, 200, 768, 14),
                                        (2, 8, 250, 768, 14), (2, 2, 256, 768, 14), (2, 2, 200, 512, 8)])
@pytest.mark.parametrize("ctas", [0, 13], ids=["cta_per_sm", "13_ctas"])
def test_pool_vs_oracle(B, T, L, D, gw, ctas):
    """L > 128 (the reference's --max_instr_len 200 / 250, vilmodel.py:798 takes the max over ALL positions): two passes of the
    kernel, the first over positions 128.. only produces row maxima.  ctas: the work plan cuts the sorted rows into one range per
    CTA; with one CTA per SM on these small batches nearly every cell is pooled in pieces by several CTAs and merged."""
    from gridmm_b200 import ops
    ep = synth.make_episodes(B, T, seed=B * 100 + T, dim=D)
    cells, fts, halfs, pos = H.oracle_grid(ep, grid_w=gw)
    gb, grid, _ = _run_builder(ep, grid_w=gw)
    nc = gw * gw
    g = torch.Generator().manual_seed(L)
    tp = (torch.randn(B, L, D, generator=g) * 0.55).half()
    pooled = torch.zeros(B * nc, D, device=_dev(), dtype=torch.float16)
    w_out = torch.zeros(B, grid.cap, device=_dev())
    ops.pool(grid.slab, D, grid.slots, grid.t_cap, grid.slot_rows, grid.view_rows, grid.tok_off, grid.perm, grid.cap,
             grid.cell_start, grid.cell_rank, nc, tp.to(_dev()).view(B * L, D), L, B, pooled, w_out=w_out, num_ctas=ctas)
    torch.cuda.synchronize()
    pooled = pooled.view(B, nc, D).float().cpu(); w_out = w_out.cpu()
    perm = grid.perm.cpu().numpy(); cr = grid.cell_rank.cpu().numpy(); cs = grid.cell_start.cpu().numpy()
    for b in range(B):
        ref, w = _oracle_pool(fts[b], cells[b][T - 1], tp[b], nc)
        nv = cs[b, nc]
        werr = (w_out[b, :nv].double() - w[perm[b, :nv]]).abs().max().item()
        assert werr < 2e-3, "relevance max differs: %.3e" % werr          # fp32 tensor-core accumulate vs float64
        for c in range(nc):
            if cr[b, c] >= 0:
                err = (pooled[b, cr[b, c]].double() - ref[c]).abs().max().item()
                assert err < 6e-3, "b=%d cell=%d err=%.3e" % (b, c, err)   # fp16 output (ulp 4e-3 at |x|~4) + exp rounding


@pytest.mark.parametrize("ctas", [0, 3, 40, 300], ids=["cta_per_sm", "3_ctas", "40_ctas", "300_ctas"])
@pytest.mark.parametrize("sizes", ["tiny", "mixed", "one_big"])
def test_pool_handmade_cells(sizes, ctas):
    """gridmm_pool on a hand-made layout (no grid builder): tiny cells (> 8 cells per 32-row tile -> several passes of the
    8-slot HMMA pooling), a cell spanning many tiles (and, cut by the work plan, many CTAs: the pieces are merged by the last one
    to arrive -- 300 CTAs are more than the SMs hold at once, so no piece may wait for another), an episode without any valid
    point, ragged text length."""
    from gridmm_b200 import ops
    rng = np.random.default_rng({"tiny": 1, "mixed": 2, "one_big": 3}[sizes])
    B, nc, D, L, t_cap = 3, 196, 768, 37, 2
    cap = t_cap * 588
    slab = torch.randn(B * t_cap * 588, D, generator=torch.Generator().manual_seed(7)).half()
    slots = (torch.arange(B, dtype=torch.int32)[:, None] * t_cap + torch.arange(t_cap, dtype=torch.int32)[None, :]).contiguous()
    perm = torch.zeros(B, cap, dtype=torch.int32)
    cell_start = torch.zeros(B, nc + 1, dtype=torch.int32)
    cell_rank = torch.full((B, nc), -1, dtype=torch.int32)
    members = []
    for b in range(B):
        if b == 1:                                   # no valid point at all
            members.append({})
            continue
        pts = rng.permutation(cap)
        if sizes == "tiny":
            counts = rng.integers(0, 3, nc)          # 0..2 rows per cell
        elif sizes == "mixed":
            counts = rng.integers(0, 12, nc)
        else:
            counts = np.zeros(nc, dtype=np.int64); counts[77] = 700; counts[3] = 1; counts[190] = 2
        mem, pos, rank = {}, 0, 0
        for c in range(nc):
            n = int(counts[c])
            cell_start[b, c] = pos
            if n:
                mem[c] = np.sort(pts[pos:pos + n])
                perm[b, pos:pos + n] = torch.from_numpy(mem[c].astype(np.int32))
                cell_rank[b, c] = rank
                rank += 1
            pos += n
        cell_start[b, nc] = pos
        members.append(mem)
    tp = (torch.randn(B, L, D, generator=torch.Generator().manual_seed(11)) * 0.3).half()
    dev = _dev()
    pooled = torch.zeros(B * nc, D, device=dev, dtype=torch.float16)
    w_out = torch.zeros(B, cap, device=dev)
    ops.pool(slab.to(dev), D, slots.to(dev), t_cap, 588, 49, 0, perm.to(dev), cap, cell_start.to(dev), cell_rank.to(dev), nc,
             tp.to(dev).view(B * L, D), L, B, pooled, w_out=w_out, num_ctas=ctas)
    torch.cuda.synchronize()
    # a second launch on the same workspace (the arrival counters must be back at zero) gives the same bits
    again = torch.zeros_like(pooled)
    ops.pool(slab.to(dev), D, slots.to(dev), t_cap, 588, 49, 0, perm.to(dev), cap, cell_start.to(dev), cell_rank.to(dev), nc,
             tp.to(dev).view(B * L, D), L, B, again, num_ctas=ctas)
    torch.cuda.synchronize()
    assert torch.equal(again, pooled)
    pooled = pooled.view(B, nc, D).float().cpu()
    for b in range(B):
        x_b = slab[b * cap:(b + 1) * cap].double()
        for c, idx in members[b].items():
            x = x_b[torch.from_numpy(idx)]
            w = (x @ tp[b].double().t()).max(-1)[0]
            ref = (torch.softmax(w, 0)[:, None] * x).sum(0)
            err = (pooled[b, int(cell_rank[b, c])].double() - ref).abs().max().item()
            assert err < 6e-3, "b=%d cell=%d n=%d err=%.3e" % (b, c, len(idx), err)
    assert pooled[1].abs().max().item() == 0


def test_pool_batch_without_any_valid_point():
    """Every episode empty (no valid depth at all): the plan hands out empty ranges, nothing is written, nothing hangs."""
    from gridmm_b200 import ops
    B, nc, D, L, t_cap = 2, 196, 768, 20, 1
    cap = t_cap * 588
    dev = _dev()
    slab = torch.randn(B * cap, D, device=dev).half()
    slots = torch.arange(B * t_cap, dtype=torch.int32, device=dev).view(B, t_cap)
    perm = torch.zeros(B, cap, dtype=torch.int32, device=dev)
    cell_start = torch.zeros(B, nc + 1, dtype=torch.int32, device=dev)
    cell_rank = torch.full((B, nc), -1, dtype=torch.int32, device=dev)
    tp = torch.randn(B * L, D, device=dev).half()
    pooled = torch.full((B * nc, D), 7.0, device=dev, dtype=torch.float16)
    ops.pool(slab, D, slots, t_cap, 588, 49, 0, perm, cap, cell_start, cell_rank, nc, tp, L, B, pooled)
    torch.cuda.synchronize()
    assert bool((pooled == 7.0).all())


@pytest.mark.parametrize("ctas", [0, 7, 300])
def test_pool_plan_invariants(ctas):
    """gridmm_pool_plan: the ranges tile the sorted valid rows in order, a cut is either on a cell boundary or further than the
    snap distance from both neighbours, and the chain records of a cut cell name exactly the CTAs that hold a piece of it."""
    from gridmm_b200 import ops, _lib
    rng = np.random.default_rng(5)
    B, nc = 5, 196
    counts = rng.integers(0, 60, (B, nc)); counts[2] = 0; counts[3, 50] = 2500; counts[4, :] = 0; counts[4, 9] = 4000
    cs = np.zeros((B, nc + 1), dtype=np.int32); cs[:, 1:] = np.cumsum(counts, 1)
    dev = _dev()
    ws = ops.pool_plan(torch.from_numpy(cs).to(dev), nc, B, 768, num_ctas=ctas)
    torch.cuda.synchronize()
    G = ctas or torch.cuda.get_device_properties(dev).multi_processor_count
    w = ws.cpu().numpy().view(np.int32)
    rec = w[:G * 8].reshape(G, 8); vb = w[G * 8:G * 8 + B + 1]; cnt = w[G * 8 + B + 1:G * 8 + B + 1 + G]
    assert np.array_equal(vb, np.concatenate([[0], np.cumsum(cs[:, -1])])) and not cnt.any()
    g0, g1 = rec[:, 0], rec[:, 1]
    assert g0[0] == 0 and g1[-1] == vb[-1] and np.array_equal(g0[1:], g1[:-1]) and (g1 >= g0).all()
    bounds = np.unique(np.concatenate([vb[b] + cs[b] for b in range(B)]))
    def cell_of(row):      # [lo, hi) of the cell that holds global row `row`
        i = np.searchsorted(bounds, row, side="right")
        return int(bounds[i - 1]), int(bounds[i])
    for c in range(G):
        first, last, n, tlast, tn, fl = rec[c, 2], rec[c, 3], rec[c, 4], rec[c, 5], rec[c, 6], rec[c, 7]
        if g1[c] == g0[c]:
            assert fl == 0
            continue
        head = g0[c] not in bounds
        tail = g1[c] not in bounds
        if head:
            lo, hi = cell_of(g0[c])
            assert min(g0[c] - lo, hi - g0[c]) > 16      # closer than the snap distance: the cut would sit on the boundary
            pieces = [k for k in range(G) if g1[k] > g0[k] and g0[k] < hi and g1[k] > lo]
            assert (fl & 1) and first == pieces[0] and last == pieces[-1] and n == len(pieces)
        else:
            assert not (fl & 1)
        if tail:
            lo, hi = cell_of(g1[c] - 1)
            if head and cell_of(g0[c]) == (lo, hi):
                assert fl & 4 and not (fl & 2)
            else:
                pieces = [k for k in range(G) if g1[k] > g0[k] and g0[k] < hi and g1[k] > lo]
                assert (fl & 2) and pieces[0] == c and tlast == pieces[-1] and tn == len(pieces)
        else:
            assert not (fl & 6)





@pytest.mark.parametrize("S,pair", [(216, 0), (150, 0), (150, 3), (216, 3)])
def test_kv_index_and_varlen_attention(S, pair):
    """Packed fusion context: gridmm_kv_index positions, and attention over the packed keys == masked attention over the padded ones
    (pair = 3 and S = 150: <= 256 keys per episode, the tcgen05 head-pair kernel; otherwise the mma.sync kernel)."""
    import ctypes
    from gridmm_b200 import ops, _lib
    lib = _lib.load()
    lib.gridmm_debug_set_attn_legacy.argtypes = [ctypes.c_int]
    B, L, Sq = 5, 80, 57
    KC = S + L
    g = torch.Generator().manual_seed(9)
    map_mask = (torch.rand(B, S, generator=g) < 0.6).to(torch.uint8); map_mask[:, -3:] = 1
    txt_mask = (torch.arange(L)[None, :] < torch.randint(20, L + 1, (B, 1), generator=g)).to(torch.uint8)
    dev = _dev()
    kv_pos = torch.zeros(B * KC, dtype=torch.int32, device=dev); kv_off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    kv_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    ops.kv_index(map_mask.to(dev), txt_mask.to(dev), kv_pos, kv_off, kv_cnt, B, S, L)
    full = torch.cat([map_mask, txt_mask], 1).flatten().bool()
    want_pos = torch.where(full, torch.cumsum(full.int(), 0) - 1, torch.full_like(full.int(), -1))
    assert torch.equal(kv_pos.cpu(), want_pos.int())
    cnt = torch.cat([map_mask, txt_mask], 1).sum(1).int()
    assert torch.equal(kv_cnt.cpu(), cnt) and int(kv_off[-1]) == int(cnt.sum())
    assert torch.equal(kv_off[:-1].cpu(), (torch.cumsum(cnt, 0) - cnt).int())
    # attention: padded + mask vs packed
    q = torch.randn(B * Sq, 768, generator=g).half().to(dev)
    kv = torch.randn(B * KC, 1536, generator=g).half().to(dev)
    kmask = torch.cat([map_mask, txt_mask], 1).to(dev)
    ref = torch.empty(B * Sq, 768, device=dev, dtype=torch.float16)
    ops.attention(q, kv[:, :768], kv[:, 768:], ref, kmask, -10000.0, B, 12, Sq, KC)
    packed = torch.zeros_like(kv)
    packed[: int(cnt.sum())] = kv[full.to(dev)]
    out = torch.empty_like(ref)
    lib.gridmm_debug_set_attn_legacy(pair)
    try:
        ops.attention_varlen(q, packed[:, :768], packed[:, 768:], out, kv_off, kv_cnt, KC, B, 12, Sq)
    finally:
        lib.gridmm_debug_set_attn_legacy(0)
    torch.cuda.synchronize()
    assert (out.float() - ref.float()).abs().max().item() < 2e-3
    # GEMM over the first kv_off[B] rows only
    w = (torch.randn(256, 1536, generator=g) * 0.05).half().to(dev)
    bias = torch.randn(256, generator=g).to(dev)
    o = torch.full((B * KC, 256), 7.0, device=dev, dtype=torch.float16)
    ops.linear_rows(packed, w, bias, o, kv_off[B:])
    torch.cuda.synchronize()
    n = int(cnt.sum())
    want = packed[:n].float() @ w.float().t() + bias
    assert (o[:n].float() - want).abs().max().item() < 4e-3 * max(1.0, want.abs().max().item())
    tail = ((n + 255) // 256) * 256           # rows of tiles that hold no valid row stay untouched
    assert (o[tail:] == 7.0).all()


@pytest.mark.parametrize("B,S,L,keys", [(5, 216, 80, "self"), (5, 216, 80, "text"), (32, 216, 80, "self"), (3, 300, 250, "self"),
                                        (3, 300, 250, "text"), (2, 40, 24, "self")])
def test_attention_ragged(B, S, L, keys):
    """gridmm_attention_ragged_f16: packed query rows (and, for self-attention, packed keys with a validity mask and a per-key
    score bias = log multiplicity of a de-duplicated key) against plain fp32 attention per episode."""
    from gridmm_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + S + L)
    cnt = torch.randint(max(1, S // 3), S + 1, (B,), generator=g).int()
    cnt[0] = S
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(cnt, 0)
    total = int(off[-1])
    rows = B * S
    q = torch.zeros(rows, 768, dtype=torch.float16); q[:total] = torch.randn(total, 768, generator=g).half()
    out = torch.full((rows, 768), 3.0, dtype=torch.float16, device=dev)
    for neg in (-10000.0, float("-inf")):
        if keys == "self":
            kv = torch.zeros(rows, 1536, dtype=torch.float16); kv[:total] = torch.randn(total, 1536, generator=g).half()
            kvalid = (torch.rand(rows, generator=g) < 0.8).to(torch.uint8)
            kvalid[off[:-1].long()] = 1                                          # at least one valid key per episode
            kbias = torch.where(torch.rand(rows, generator=g) < 0.1, torch.rand(rows, generator=g) * 3.0, torch.zeros(rows))
            ops.attention_ragged(q.to(dev), kv[:, :768].to(dev), kv[:, 768:].to(dev), out, off.to(dev), cnt.to(dev), S, kvalid.to(dev), neg,
                                 B, 12, S, k_off=off.to(dev), k_cnt=cnt.to(dev), kbias=kbias.to(dev))
        else:
            kv = torch.randn(B * L, 1536, generator=g).half()
            lens = torch.randint(1, L + 1, (B,), generator=g); lens[0] = L
            tmask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.uint8)
            ops.attention_ragged(q.to(dev), kv[:, :768].to(dev), kv[:, 768:].to(dev), out, off.to(dev), cnt.to(dev), S, tmask.to(dev), neg,
                                 B, 12, L, k_rows=L)
        torch.cuda.synchronize()
        got = out.float().cpu()
        for b in range(B):
            r0, n = int(off[b]), int(cnt[b])
            qh = q[r0:r0 + n].float().view(n, 12, 64).permute(1, 0, 2)
            if keys == "self":
                kk, vv = kv[r0:r0 + n, :768], kv[r0:r0 + n, 768:]
                add = torch.where(kvalid[r0:r0 + n] > 0, kbias[r0:r0 + n], torch.full((n,), neg))
            else:
                kk, vv = kv[b * L:(b + 1) * L, :768], kv[b * L:(b + 1) * L, 768:]
                add = torch.zeros(L).masked_fill(tmask[b] == 0, neg)
            kh = kk.float().view(-1, 12, 64).permute(1, 0, 2)
            vh = vv.float().view(-1, 12, 64).permute(1, 0, 2)
            ref = (torch.softmax(qh @ kh.transpose(-1, -2) / 8.0 + add[None, None, :], -1) @ vh).permute(1, 0, 2).reshape(n, 768)
            err = (got[r0:r0 + n] - ref).abs().max().item()
            assert err < 4e-3, "b=%d neg=%s err=%.3e" % (b, neg, err)
        assert (got[total:] == 3.0).all()                # rows past the packed batch are never written


def test_linear_device_row_count():
    """m_dev: the GEMM kernels process the first *m_dev rows only (packed operand whose size only the GPU knows); every epilogue
    variant of gridmm_linear_f16 and both cluster shapes of gridmm_linear_ln_f16."""
    from gridmm_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(77)
    M, K = 6912, 768
    a = torch.randn(M, K, generator=g).half().to(dev)
    res = torch.randn(M, 768, generator=g).to(dev)
    gamma, beta = (torch.rand(768, generator=g) + 0.5).to(dev), torch.randn(768, generator=g).to(dev)
    for n_rows in (4500, 129, 1, 6912, 1800):
        md = torch.tensor([n_rows], dtype=torch.int32, device=dev)
        for N, act in ((2304, ops.ACT_NONE), (3072, ops.ACT_GELU), (768, ops.ACT_NONE)):
            w = (torch.randn(N, K, generator=g) * 0.03).half().to(dev)
            bias = torch.randn(N, generator=g).to(dev)
            o = torch.full((M, N), 5.0, device=dev, dtype=torch.float16)
            full = torch.empty(M, N, device=dev, dtype=torch.float16)
            ops.linear(a, w, bias, out_f16=o, act=act, m_dev=md)
            ops.linear(a, w, bias, out_f16=full, act=act)
            torch.cuda.synchronize()
            assert torch.equal(o[:n_rows], full[:n_rows])                      # row results do not depend on the row count
            assert (o[n_rows:] == 5.0).all()
        w = (torch.randn(768, K, generator=g) * 0.03).half().to(dev)
        bias = torch.randn(768, generator=g).to(dev)
        o32 = torch.full((M, 768), 5.0, device=dev); o16 = torch.full((M, 768), 5.0, device=dev, dtype=torch.float16)
        f32 = torch.empty(M, 768, device=dev); f16 = torch.empty(M, 768, device=dev, dtype=torch.float16)
        ops.linear_ln(a, w, bias, res, gamma, beta, 1e-12, out_f32=o32, out_f16=o16, m_dev=md)
        ops.linear_ln(a, w, bias, res, gamma, beta, 1e-12, out_f32=f32, out_f16=f16)
        torch.cuda.synchronize()
        assert torch.equal(o32[:n_rows], f32[:n_rows]) and torch.equal(o16[:n_rows], f16[:n_rows])
        assert (o32[n_rows:] == 5.0).all() and (o16[n_rows:] == 5.0).all()


def test_packed_map_inputs_equal_padded():
    """gridmm_map_index + gridmm_map_inputs_packed against gridmm_map_inputs (the padded layout with the reference's compaction quirk):
    same rows, the quirk's zero-vector slots represented once with key bias log z, all gmap rows kept; and the packed fusion context
    (gridmm_kv_index_packed / gridmm_fusion_inputs_packed) holds the rows of the padded one up to that de-duplication."""
    from gridmm_b200 import ops
    dev = _dev()
    B, T, G, L, V = 9, 4, 11, 30, 37
    NC, S = 196, 196 + 11
    ep = synth.make_episodes(B, T, seed=321, dim=768)
    ep["depth_sub"][3] = 0                                    # an episode without any valid point
    gb, grid, _ = _run_builder(ep)
    g = torch.Generator().manual_seed(5)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dev)      # noqa: E731
    proj = rnd(B * NC, 768)
    gw, gb_, gg, gbt = rnd(5, 768), rnd(768), rnd(768), rnd(768)
    mw, mb, mg, mbt = rnd(7, 768), rnd(768), rnd(768), rnd(768)
    ng, nb = rnd(768), rnd(768)
    gmap_pos, gmap_img, table = rnd(B * G, 7), rnd(B * G, 768), rnd(100, 768)
    step_ids = torch.randint(0, 100, (B * G,), generator=g).to(dev)
    gmask = (torch.rand(B, G, generator=g) < 0.7).to(torch.uint8).to(dev); gmask[:, 0] = 1
    map32 = torch.zeros(B * S, 768, device=dev); map16 = torch.zeros(B * S, 768, device=dev, dtype=torch.float16)
    map_mask = torch.zeros(B, S, dtype=torch.uint8, device=dev)
    ops.map_inputs(proj, grid.pos_fts, grid.cell_rank, grid.n_nonempty, gw, gb_, gg, gbt, gmap_pos, mw, mb, mg, mbt, gmap_img, table,
                   step_ids, gmask, ng, nb, 1e-5, map32, map16, map_mask, B, NC, S)
    i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)      # noqa: E731
    m_off, m_info, m_goff, cor = i32(B + 1), i32(4, B), i32(B), i32(B, NC)
    m_logz = torch.zeros(B, device=dev)
    p32 = torch.zeros(B * S, 768, device=dev); p16 = torch.zeros(B * S, 768, device=dev, dtype=torch.float16)
    kvalid = torch.zeros(B * S, dtype=torch.uint8, device=dev); kbias = torch.zeros(B * S, device=dev)
    ops.map_index(grid.cell_rank, grid.n_nonempty, B, NC, G, m_off, m_info, m_logz, cor, m_goff)
    ops.map_inputs_packed(proj, grid.pos_fts, cor, m_off, m_info, m_logz, gw, gb_, gg, gbt, gmap_pos, mw, mb, mg, mbt, gmap_img, table,
                          step_ids, gmask, ng, nb, 1e-5, p32, p16, kvalid, kbias, B, NC, G)
    torch.cuda.synchronize()
    off, info = m_off.cpu().numpy(), m_info.cpu().numpy()
    mm = map_mask.cpu().numpy().astype(bool)
    k_all = grid.n_nonempty.cpu().numpy()
    pad32, pad16 = map32.view(B, S, 768), map16.view(B, S, 768)
    for b in range(B):
        k, v, n, z = info[:, b]
        assert k == k_all[b] and n == v + G and off[b + 1] - off[b] == n and int(m_goff[b]) == off[b] + v
        assert z == mm[b, k:NC].sum() and v == k + (1 if z > 0 else 0)
        r0 = off[b]
        assert torch.equal(p32[r0:r0 + k], pad32[b, :k]) and torch.equal(p16[r0:r0 + k], pad16[b, :k])
        assert torch.equal(p32[r0 + v:r0 + n], pad32[b, NC:]) and torch.equal(p16[r0 + v:r0 + n], pad16[b, NC:])
        assert (kvalid[r0:r0 + v] == 1).all() and torch.equal(kvalid[r0 + v:r0 + n], gmask[b])
        if z > 0:
            slot = k + int(np.argmax(mm[b, k:NC]))           # any flagged zero-vector slot of the padded layout
            assert p32[r0 + k].abs().max().item() == 0 and torch.equal(p16[r0 + k], pad16[b, slot])
            assert abs(float(kbias[r0 + k]) - np.log(z)) < 1e-6
        assert float(kbias[r0:r0 + k].abs().max()) == 0 if k else True
    # packed fusion context
    KC = S + L
    txt32 = rnd(B * L, 768)
    tmask = (torch.arange(L)[None, :] < torch.randint(5, L + 1, (B, 1), generator=g)).to(torch.uint8).to(dev)
    vmask = torch.ones(B, V, dtype=torch.uint8, device=dev)
    kv_src, kv_off, kv_cnt = i32(B * KC), i32(B + 1), i32(B)
    kv_bias = torch.zeros(B * KC, device=dev)
    ops.kv_index_packed(m_off, kvalid, kbias, tmask, B, L, kv_src, kv_bias, kv_off, kv_cnt)
    x32 = torch.zeros(B * (G + V), 768, device=dev); x16 = torch.zeros(B * (G + V), 768, device=dev, dtype=torch.float16)
    kv16 = torch.zeros(B * KC, 768, device=dev, dtype=torch.float16)
    qm = torch.zeros(B, G + V, dtype=torch.uint8, device=dev)
    vp = (rnd(B * V, 14), rnd(14, 768), rnd(768), rnd(768), rnd(768), rnd(B * V, 768))
    ops.fusion_inputs_packed(p32, txt32, kv_src, kv_off, m_goff, gmask, vmask, x32, x16, kv16, qm, vp, B, L, G, V, B * KC)
    # padded reference of the same inputs
    kv_pos, kv_off2, kv_cnt2 = i32(B * KC), i32(B + 1), i32(B)
    ops.kv_index(map_mask, tmask, kv_pos, kv_off2, kv_cnt2, B, S, L)
    x32b = torch.zeros_like(x32); x16b = torch.zeros_like(x16); kv16b = torch.zeros_like(kv16)
    kvm = torch.zeros(B, KC, dtype=torch.uint8, device=dev); qmb = torch.zeros_like(qm)
    ops.fusion_inputs(map32, txt32, map_mask, tmask, gmask, vmask, x32b, x16b, kv16b, kvm, qmb, B, S, L, G, V, kv_pos=kv_pos, vp=vp)
    torch.cuda.synchronize()
    assert torch.equal(x32, x32b) and torch.equal(x16, x16b) and torch.equal(qm, qmb)
    o1, c1, o2, c2 = kv_off.cpu().numpy(), kv_cnt.cpu().numpy(), kv_off2.cpu().numpy(), kv_cnt2.cpu().numpy()
    for b in range(B):
        k, v, n, z = info[:, b]
        assert c1[b] == c2[b] - max(z - 1, 0)
        a = kv16[o1[b]:o1[b] + c1[b]]; ref = kv16b[o2[b]:o2[b] + c2[b]]
        assert torch.equal(a[:k], ref[:k])
        assert torch.equal(a[v:], ref[k + z:])               # valid gmap rows, then the valid text rows
        bias = kv_bias[o1[b]:o1[b] + c1[b]]
        if z > 0:
            assert a[k].abs().max().item() == 0 and abs(float(bias[k]) - np.log(z)) < 1e-6
        assert float(bias.abs().sum()) == (np.log(z) if z > 1 else 0.0) or abs(float(bias.abs().sum()) - np.log(max(z, 1))) < 1e-6


@pytest.mark.parametrize("M,N,act", [(1824, 3072, 1), (912, 2304, 0), (300, 1152, 2), (1824, 768, 0)])
def test_linear_384_wide_pair_tiles_equal_the_other_schedules(M, N, act):
    """256 x 384 pair tiles against the 128/256-wide schedules of the same kernel (debug hook): same K order per output element,
    so the fp32 accumulators agree to rounding and the fp16 outputs almost everywhere bitwise."""
    import ctypes
    from gridmm_b200 import ops, _lib
    lib = _lib.load()
    lib.gridmm_debug_set_gemm_384.argtypes = [ctypes.c_int]
    dev = _dev()
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, 768, generator=g).half().to(dev)
    w = (torch.randn(N, 768, generator=g) * 0.05).half().to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    outs = []
    for on in (1, 0):
        lib.gridmm_debug_set_gemm_384(on)
        try:
            o32 = torch.empty(M, N, device=dev); o16 = torch.empty(M, N, device=dev, dtype=torch.float16)
            ops.linear(a, w, bias, residual=res, out_f32=o32, out_f16=o16, act=act)
            torch.cuda.synchronize()
        finally:
            lib.gridmm_debug_set_gemm_384(0)
        outs.append((o32, o16))
    y = a.float() @ w.float().t() + bias
    y = torch.nn.functional.gelu(y) if act == 1 else (torch.relu(y) if act == 2 else y)
    ref = y + res
    for o32, o16 in outs:
        assert (o32 - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    assert (outs[0][0] - outs[1][0]).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


def test_adamw_and_grad_clip_match_the_reference_optimizer():
    """gridmm_grad_sumsq + gridmm_adamw_step against the reference's update rule restated in float64: clip_grad_norm_
    (pretrain_src/train_r2r.py:281-285) followed by AdamW with decoupled weight decay and bias correction
    (pretrain_src/optim/adamw.py:57-104), several steps, both parameter groups, gradients pre-summed over `world` ranks."""
    from gridmm_b200 import ops
    dev = _dev()
    g_ = torch.Generator().manual_seed(3)
    n = 1_000_003
    p0 = torch.randn(n, generator=g_)
    lr, b1, b2, eps, world, max_norm = 5e-5, 0.9, 0.98, 1e-6, 8, 5.0
    for wd in (0.01, 0.0):
        p = torch.zeros(n + 1, device=dev)[:n]           # odd length: the scalar tail of the vectorised kernels
        p.copy_(p0)
        m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
        rp, rm, rv = p0.double(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
        sumsq = torch.zeros(1, device=dev)
        for step in range(1, 4):
            g = torch.randn(n, generator=g_) * (40.0 if step == 2 else 0.002) * world      # step 2 is clipped
            gd = g.to(dev)
            sumsq.zero_()
            ops.grad_sumsq(gd[: n // 2], sumsq); ops.grad_sumsq(gd[n // 2:][:0], sumsq)      # empty range: no-op
            ops.grad_sumsq(gd[n // 2:].clone(), sumsq)                                       # accumulates over ranges
            ops.adamw_step(p, gd, m, v, lr, b1, b2, eps, wd, step, grad_scale=1.0 / world, sumsq=sumsq, max_norm=max_norm)
            # reference
            ga = g.double() / world
            norm = ga.norm()
            coef = max_norm / (norm + 1e-6)
            if coef < 1:
                ga = ga * coef
            rm = b1 * rm + (1 - b1) * ga
            rv = b2 * rv + (1 - b2) * ga * ga
            step_size = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
            rp = rp - step_size * rm / (rv.sqrt() + eps)
            if wd > 0:
                rp = rp - lr * wd * rp
            torch.cuda.synchronize()
            assert abs(float(sumsq.sqrt()) / world - float(norm)) < 1e-3 * float(norm)
        assert (p.cpu().double() - rp).abs().max().item() < 2e-6
        assert (m.cpu().double() - rm).abs().max().item() < 1e-6 * max(1.0, float(rm.abs().max()))
        assert (v.cpu().double() - rv).abs().max().item() < 1e-5 * max(1.0, float(rv.abs().max()))

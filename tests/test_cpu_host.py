"""CPU: host-side logic, the C-ABI surface, and the 'no CPU fallback' contract."""
import ctypes
import os

import numpy as np
import pytest
import torch

from gridmm_b200 import synth
from tests import helpers as H


def test_library_builds_loads_and_exports_header_symbols():
    from gridmm_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _lib.header_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "include/gridmm_b200.h declares %s but the library does not export it" % n
    assert set(_lib._SIGS) <= set(names)
    assert _lib.load().gridmm_abi_version() == 4


def test_param_spec_matches_reference_layout():
    from gridmm_b200.model import GlocalTextPathNavCMT, NavConfig, param_spec
    cfg = NavConfig(num_l_layers=1, num_pano_layers=1, obj_feat_size=768)
    m = GlocalTextPathNavCMT(cfg)
    sd = m.state_dict()
    assert list(sd.keys()) == list(param_spec(cfg).keys())
    # the keys SURVEY 8(b) lists as the drop-in contract
    for k in ("text_proj.weight", "grid_proj.bias", "grid_pos_embeddings.0.weight", "grid_pos_embeddings.1.bias",
              "grid_encoder.layers.0.self_attn.in_proj_weight", "grid_encoder.norm.weight",
              "grid_txt_encoder.x_layers.0.visual_attention.att.query.weight",
              "grid_txt_encoder.x_layers.0.visn_self_att.self.key.bias", "local_encoder.encoder.x_layers.3.visn_output.dense.weight",
              "local_encoder.vp_pos_embeddings.0.weight", "global_encoder.gmap_step_embeddings.weight",
              "global_encoder.sprel_linear.weight", "global_sap_head.net.3.weight", "sap_fuse_linear.net.0.weight",
              "og_head.net.2.bias"):
        assert k in sd, k
    assert tuple(sd["grid_encoder.layers.0.self_attn.in_proj_weight"].shape) == (2304, 768)
    assert tuple(sd["sap_fuse_linear.net.0.weight"].shape) == (768, 1536)
    total = sum(v.numel() for v in GlocalTextPathNavCMT(NavConfig(obj_feat_size=768)).state_dict().values())
    assert total == 161596423          # parameter count of the reference model (probed in the authoring container)


def test_fuse_index_equals_reference_loops():
    """build_fuse_index + the kernel's arithmetic (restated in numpy) == the reference's vpid loops (oracle.fuse_logits)."""
    from gridmm_b200.model import build_fuse_index
    from oracle import model_oracle as mo
    for seed in range(5):
        nav = synth.to_torch(synth.make_nav_inputs(16, seed=seed, gmap_len=14, n_views=36))
        B, G = nav["gmap_masks"].shape
        V = nav["vp_masks"].shape[1]
        g = torch.Generator().manual_seed(seed)
        gl = torch.randn(B, G, generator=g).masked_fill(nav["gmap_visited_masks"] | ~nav["gmap_masks"], float("-inf"))
        ll = torch.randn(B, V, generator=g).masked_fill(~nav["vp_nav_masks"], float("-inf"))
        ref = mo.fuse_logits(gl, ll, nav["gmap_vpids"], nav["gmap_visited_masks"], nav["vp_cand_vpids"])
        src, bw = build_fuse_index(nav["gmap_vpids"], nav["gmap_visited_masks"], nav["vp_cand_vpids"], G, V)
        got = gl.clone()
        got[:, 0] += ll[:, 0]
        for i in range(B):
            s = torch.zeros(())
            for v in range(1, V):
                if bw[i, v]:
                    s = s + ll[i, v]
            for j in range(1, G):
                if src[i, j] >= 0:
                    got[i, j] += ll[i, src[i, j]]
                elif src[i, j] == -2:
                    got[i, j] += s
        assert torch.equal(got, ref)
        # the mask-free tables the product passes to gridmm_nav_logits2 (no D2H read of the visited mask) + the kernel's rule
        # (heads.cu: a candidate is "already visited" when its node's flag is set; node_src only applies to unvisited nodes)
        src2, bw2 = _fuse_index_from_maps(nav["gmap_vpids"], nav["gmap_visited_masks"].numpy(), nav["vp_cand_vpids"], G, V)
        assert np.array_equal(src2, src) and np.array_equal(bw2, bw)


def _fuse_index_from_maps(gmap_vpids, vis, vp_cand_vpids, G, V):
    """numpy restatement of what nav_logits2_kernel derives from build_fuse_maps' tables and the device-side visited flags."""
    from gridmm_b200.model import build_fuse_maps
    node_src, cand_node = build_fuse_maps(gmap_vpids, vp_cand_vpids, G, V)
    B = len(gmap_vpids)
    fuse_src = np.full((B, G), -1, np.int32)
    bw = np.zeros((B, V), np.uint8)
    for i in range(B):
        for v in range(1, V):
            bw[i, v] = 1 if (cand_node[i, v] >= 0 and vis[i, cand_node[i, v]]) else 0
        for j in range(1, G):
            fuse_src[i, j] = -1 if vis[i, j] else node_src[i, j]
    return fuse_src, bw


def test_host_pose_rounding_contract():
    """trig evaluated in double then rounded to fp32 (python float x np.float32 array, env.py:119-120,347-348)."""
    import math
    from gridmm_b200.env import GEOMETRIES
    g = GEOMETRIES["r2r"]
    off = np.array([-6 / 7, -4 / 7, -2 / 7, 0., 2 / 7, 4 / 7, 6 / 7] * 7, np.float32) * math.tan(math.pi / 6)
    assert off.dtype == np.float32 and np.array_equal(off[:7], g.off7)


def test_no_cpu_path():
    """The product path must fail loudly without a GPU instead of computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gridmm_b200 import ops, _lib
    from gridmm_b200.env import GridMapBuilder
    with pytest.raises(RuntimeError):
        GridMapBuilder(2)
    with pytest.raises(_lib.GridmmError):
        ops.linear(torch.zeros(128, 64, dtype=torch.float16), torch.zeros(128, 64, dtype=torch.float16),
                   out_f16=torch.zeros(128, 128, dtype=torch.float16))


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "gridmm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_head_pack_and_table_restate_cls_prediction():
    """Host side of the grouped ClsPrediction launch: the stacked split weights / gamma*w2 / (c1, c0) constants and the per-tile
    table reproduce ClsPrediction (vilmodel.py:663-674) when the kernel's arithmetic is restated in torch on the CPU:
    logit = rstd * (S3 - mean * c1) + c0 over r = ReLU(x W^T + b)."""
    from gridmm_b200.model import GlocalTextPathNavCMT, NavConfig
    from oracle import model_oracle as mo
    cfg = NavConfig(num_l_layers=1, num_pano_layers=1, obj_feat_size=768)
    m = GlocalTextPathNavCMT(cfg).eval()
    with torch.no_grad():
        for p in m.parameters():          # non-trivial LayerNorm weights / biases
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    w, bias, gw2, consts, fuse_gw2, n_groups, n_names = m.head_pack(True)
    assert n_names == 4 and n_groups == 6 and tuple(w.shape) == (6 * 768, 3 * 768) and tuple(gw2.shape) == (6 * 768,)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x = torch.randn(9, 768)
    hi = x.half(); lo = (x - hi.float()).half()
    a = torch.cat([hi, lo, hi], 1).float()                      # what gridmm_head_rows writes
    for gi, name in enumerate(["global_sap_head", "local_sap_head", "grid_sap_head", "og_head"]):
        r = torch.relu(a @ w[gi * 768:(gi + 1) * 768].float().t() + bias[gi * 768:(gi + 1) * 768])
        s1, s2, s3 = r.sum(1), (r * r).sum(1), (r * gw2[gi * 768:(gi + 1) * 768]).sum(1)
        mean = s1 / 768
        logit = torch.rsqrt(s2 / 768 - mean * mean + 1e-12) * (s3 - mean * consts[gi, 0]) + consts[gi, 1]
        ref = mo.cls_head(sd, name, x).squeeze(-1)
        assert (logit - ref).abs().max().item() < 2e-4, name
    # fuse head: the two raw-product groups are the K halves of sap_fuse_linear.net.0
    g0, v0 = torch.randn(5, 768), torch.randn(5, 768)
    split = lambda t: torch.cat([t.half(), (t - t.half().float()).half(), t.half()], 1).float()      # noqa: E731
    h = split(g0) @ w[4 * 768:5 * 768].float().t() + split(v0) @ w[5 * 768:6 * 768].float().t()
    r = torch.relu(h + sd["sap_fuse_linear.net.0.bias"])
    mean = r.mean(1)
    logit = torch.rsqrt((r * r).mean(1) - mean * mean + 1e-12) * ((r * fuse_gw2).sum(1) - mean * consts[4, 0]) + consts[4, 1]
    ref = mo.cls_head(sd, "sap_fuse_linear", torch.cat([g0, v0], 1)).squeeze(-1)
    assert (logit - ref).abs().max().item() < 2e-4
    # tile table: A rows / output rows are 128-aligned, groups do not overlap, the object head re-reads the local rows
    grp, tiles, a0, a_rows, o_obj, out_rows = m.head_table(32, 20, 57, True, torch.device("cpu"))
    t = grp.numpy()
    assert tiles == t.shape[0] and (t[:, 0] % 128 == 0).all() and (t[:, 2] % 128 == 0).all() and t[:, 0].max() < a_rows
    assert len(set(t[:, 2].tolist())) == tiles and t[:, 2].max() + 128 <= out_rows
    obj = t[t[:, 1] == 3 * 768]
    loc = t[t[:, 1] == 1 * 768]
    assert (obj[:, 0] == loc[:, 0]).all() and (obj[:, 2] >= o_obj).all() and (t[t[:, 3] == 1][:, 1] >= 4 * 768).all()


def test_pretraining_param_spec_equals_reference_state_dict():
    """NavConfig(pretrain_trunk=True, use_lang2visn_attn=True): names, shapes and order of the parameter tree equal the reference
    pretraining trunk's state_dict (pretrain_src/model/vilmodel.py:640-666), recorded by oracle/make_golden.py next to the golden
    outputs -- so `load_state_dict(strict=True)` of a pretraining checkpoint's `bert.*` tensors works."""
    import json
    from gridmm_b200.model import GlocalTextPathNavCMT, NavConfig, param_spec
    kw = H.PRETRAIN_MODEL_CASE["model"]
    ref = json.load(open(os.path.join(H.GOLD, "pretrain_small_spec.json")))
    cfg = NavConfig(pretrain_trunk=True, use_lang2visn_attn=True, **kw)
    spec = param_spec(cfg)
    assert list(spec.keys()) == list(ref.keys())
    assert {k: list(v[0]) for k, v in spec.items()} == ref
    m = GlocalTextPathNavCMT(cfg)
    assert list(m.state_dict().keys()) == list(ref.keys())
    for absent in ("global_sap_head.net.0.weight", "global_encoder.sprel_linear.weight", "sap_fuse_linear.net.0.weight"):
        assert absent not in spec
    # the navigation model's tree is untouched by the new flags
    assert len(param_spec(NavConfig())) == 373


def test_gmap_aggregation_glue_equals_oracle():
    """GlocalTextPathNavCMT._aggregate_gmap (torch gathers, device-agnostic host glue of forward_pretrain) against the oracle's
    restatement of GlobalMapEncoder._aggregate_gmap_features (pretrain_src/model/vilmodel.py:578-612)."""
    from gridmm_b200.model import GlocalTextPathNavCMT
    from oracle import pretrain_oracle as po
    case = H.PRETRAIN_MODEL_CASE
    pb = synth.make_pretrain_batch(case["batch"], seed=case["seed"], txt_len=case["txt_len"], max_steps=case["max_steps"])
    n_tot = sum(pb["traj_step_lens"])
    g = torch.Generator().manual_seed(5)
    pano = torch.randn(n_tot, 36, 768, generator=g)
    lens = torch.from_numpy(pb["traj_vp_view_lens"]).clone()
    lens[1] = 30                                        # a shorter panorama exercises the validity mask
    m = GlocalTextPathNavCMT(H.make_config(pretrain_trunk=True, use_lang2visn_attn=True, **case["model"]))
    got = m._aggregate_gmap(pano, lens, pb["traj_step_lens"], pb["traj_vpids"], pb["traj_cand_vpids"], pb["gmap_vpids"])
    ref = po.aggregate_gmap_features(list(torch.split(pano, pb["traj_step_lens"], 0)), list(torch.split(lens, pb["traj_step_lens"], 0)),
                                     pb["traj_vpids"], pb["traj_cand_vpids"], pb["gmap_vpids"])
    assert got.shape == ref.shape and got[:, 0].abs().max() == 0
    assert (got - ref).abs().max().item() < 1e-6


def _stub_ops(monkeypatch):
    """Replace every kernel wrapper of gridmm_b200.ops by a recorder (host-logic tests; nothing is computed)."""
    import types
    from gridmm_b200 import ops
    calls = []
    for name in dir(ops):
        fn = getattr(ops, name)
        if isinstance(fn, types.FunctionType) and not name.startswith("_") and name != "pool_text_ws":
            monkeypatch.setattr(ops, name, (lambda n: (lambda *a, **k: calls.append(n)))(name))
    monkeypatch.setattr(ops, "pool_text_ws", lambda dev, B, D, L=128: torch.zeros((L + 127) // 128 * B * 128 * D, dtype=torch.float16))
    return calls


def test_pretrain_sap_heads_glue_masks_and_candidates(monkeypatch):
    """forward_pretrain(task="sap", heads=True): the host side of forward_sap (pretrain_src/model/pretrain_cmt.py:244-269) --
    navigable mask from the last panorama's nav types, candidates of the logit fusion from traj_cand_vpids[i][-1] -- must hand
    the navigation kernels the same masks / index arrays as the oracle's restatement uses; and the wrapper's state_dict loads
    after the reference's own key remap (vlnbert_init.py:19-27) with only the MLM head left over."""
    import json
    from gridmm_b200.model import GlocalTextPathNavCMT, build_fuse_index, remap_pretrained_keys
    calls = _stub_ops(monkeypatch)
    case = H.PRETRAIN_MODEL_CASE
    cfg = H.make_config(use_lang2visn_attn=True, graph_sprels=False, **case["model"])
    model = GlocalTextPathNavCMT(cfg).eval()
    shapes = json.load(open(os.path.join(H.GOLD, "pretrain_heads_small_spec.json")))
    sd = remap_pretrained_keys({k: torch.zeros(v) for k, v in shapes.items()})
    res = model.load_state_dict(sd, strict=False)
    assert not res.missing_keys and all(k.startswith("mlm_head.") for k in res.unexpected_keys) and len(res.unexpected_keys) == 6
    batch = H.pretrain_batch(case)
    pb = synth.make_pretrain_batch(case["batch"], seed=case["seed"], txt_len=case["txt_len"], max_steps=case["max_steps"])
    lab = synth.make_pretrain_labels(pb, seed=case["seed"])
    batch["gmap_visited_masks"] = torch.from_numpy(lab["gmap_visited_masks"])
    out = model.forward_pretrain(batch, task="sap", heads=True)
    B, G, V = case["batch"], int(batch["gmap_lens"].max()), 37
    assert out["fused_logits"].shape == (B, G) and out["local_logits"].shape == (B, V) and out["grid_logits"].shape == (B, G)
    assert calls.count("cls_heads") == 1 and calls.count("nav_logits2") == 1
    staged = {k[0]: v for k, v in model._ws.items() if k[0].startswith("in_")}
    nav_types_last = torch.stack([t[-1] for t in torch.split(batch["traj_nav_types"], batch["traj_step_lens"], 0)], 0)
    want_nav = torch.cat([torch.ones(B, 1, dtype=torch.bool), nav_types_last[:, :V - 1] == 1], 1)
    assert torch.equal(staged["in_vp_nav"].bool(), want_nav)
    assert torch.equal(staged["in_gmap_visited"].bool(), batch["gmap_visited_masks"])
    cands = [[None] + list(c[-1]) for c in batch["traj_cand_vpids"]]
    src, bw = build_fuse_index(batch["gmap_vpids"], batch["gmap_visited_masks"], cands, G, V)
    vis = batch["gmap_visited_masks"].numpy()
    node_src, cand_node = staged["in_fuse_src"].numpy(), staged["in_cand_node"].numpy()
    assert np.array_equal(np.where(vis, -1, node_src), src)
    assert np.array_equal((cand_node >= 0) & np.take_along_axis(vis, np.maximum(cand_node, 0), 1), bw.astype(bool))
    with pytest.raises(ValueError):
        GlocalTextPathNavCMT(H.make_config(pretrain_trunk=True, use_lang2visn_attn=True, **case["model"])).forward_pretrain(
            batch, task="sap", heads=True)


def test_host_glue_launch_sequence_with_stubbed_kernels(monkeypatch):
    """Host logic without a GPU: every `ops.*` kernel wrapper is replaced by a recorder, the model lives on the CPU, and the
    Python glue of forward('navigation'), its intermediates path and forward_pretrain (both tasks) must run to the end with
    consistent shapes.  Pins the launch budget of the navigation step: 59 kernel launches at one launch per wrapper call."""
    from gridmm_b200.model import GlocalTextPathNavCMT
    calls = _stub_ops(monkeypatch)
    ep_kw, nav_kw, model_kw = H.NAV_CASES["r2r_small"]
    model = GlocalTextPathNavCMT(H.make_config(**model_kw)).eval()
    ep = synth.make_episodes(dim=768, **ep_kw)
    cells, fts, _, pos = H.oracle_grid(ep)
    nav = H.nav_batch(ep_kw, nav_kw, cells, fts, pos)
    out = model("navigation", nav)
    B, G, V = ep_kw["batch"], nav_kw["gmap_len"], 1 + nav_kw["n_views"]
    assert out["fused_logits"].shape == (B, G) and out["local_logits"].shape == (B, V) and out["obj_logits"] is None
    assert out["gmap_embeds"].shape == (B, G, 768) and out["vp_embeds"].shape == (B, V, 768)
    # cell_sort (list path only) + the 59 launches after the grid build = the 60-launch step of bench.py, where grid_update replaces
    # cell_sort (+ one gridmm_copy_segments staging launch there); gridmm_map_index is the launch the packed map sequence adds,
    # gridmm_pool_plan the one the pooling kernel's work plan adds (behind the grid update, beside the text branch)
    assert calls.count("cell_sort") == 1 and len(calls) == 60, (len(calls), calls)
    assert calls.count("pool_plan") == 1
    assert calls.count("map_index") == 1 and calls.count("map_inputs_packed") == 1 and calls.count("attention_ragged") == 3
    model.ragged_map = False                    # the padded layout (used for map sequences longer than 320 rows) stays available
    calls.clear()
    model("navigation", nav)
    assert len(calls) == 59 and calls.count("map_inputs") == 1 and "map_index" not in calls
    model.ragged_map = True
    assert calls.count("linear_ln") == 17 and calls.count("pool") == 1 and calls.count("cls_heads") == 1
    calls.clear()
    inter = model("navigation", nav, return_intermediates=True)
    assert inter["map_embeds"].shape[0] == B and "pooled" in inter
    # pretraining trunk, both exits
    case = H.PRETRAIN_MODEL_CASE
    pm = GlocalTextPathNavCMT(H.make_config(pretrain_trunk=True, use_lang2visn_attn=True, **case["model"])).eval()
    batch = H.pretrain_batch(case)
    calls.clear()
    gmap_e, vp_e, grid_g = pm.forward_pretrain(batch, task="sap")
    Gp = int(batch["gmap_lens"].max())
    assert gmap_e.shape == (case["batch"], Gp, 768) and vp_e.shape == (case["batch"], 37, 768) and grid_g.shape == gmap_e.shape
    assert "cls_heads" not in calls and "nav_logits2" not in calls          # the trunk stops before any action head
    calls.clear()
    txt = pm.forward_pretrain(batch, task="mlm")
    assert txt.shape == (case["batch"], case["txt_len"], 768)
    assert "linear_rows" not in calls                                       # the MLM exit leaves before the fusion encoder's K/V projection
    # REVERIE / SOON batches: object tokens behind the views of every panorama (host glue: packed panorama rows, lens = views + objects)
    ocase = H.PRETRAIN_OBJ_CASE
    om = GlocalTextPathNavCMT(H.make_config(pretrain_trunk=True, use_lang2visn_attn=True, **ocase["model"])).eval()
    obatch = H.pretrain_batch(ocase)
    calls.clear()
    gmap_e, vp_e, grid_g = om.forward_pretrain(obatch, task="sap")
    assert vp_e.shape == (ocase["batch"], 1 + 36 + ocase["n_objs"], 768) and gmap_e.shape[1] == int(obatch["gmap_lens"].max())


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm): one JSON line with the contract's keys,
    measured on the oracle port of the reference algorithm; and the CUDA arm refuses to run without a GPU instead of falling
    back to anything."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "nav-steps/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0 and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("configs[1]")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert abs(line["cpu_baseline"]["value"] - line["value"]) < 1e-9
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode != 0 and "no CPU path" in (r.stdout + r.stderr)


def test_bench_flop_model_matches_the_survey():
    """bench.py's roofline numerator: 2mnk over the GEMM launches of one B=32 step.  SURVEY 8(d) puts the whole step at
    ~14.4 GFLOP per episode-step including the attention cores (4 S^2 D etc., ~0.6 G) -- the GEMM share must sit just below."""
    import bench
    per_sample = bench.gemm_flops_per_step() / bench.B
    assert 13.0e9 < per_sample < 14.4e9
    packed = bench.gemm_flops_per_step(kv_rows=bench.B * 217)        # ~217 valid context rows per episode instead of 296
    assert packed < bench.gemm_flops_per_step() and packed > 0.9 * bench.gemm_flops_per_step()


def test_subsample_depth_reads_the_reference_pixels():
    """SURVEY 8a row 1 (host indexing, no GPU): views 12..23 and pixels 9 + 18 i of a [36,128,128] map (r2r/env.py:278-285);
    19 + 36 i of the CE policy's [12,256,256] stack.  Everything else is noise here, so a wrong stride or view range fails."""
    from gridmm_b200.env import GridMapBuilder
    rng = np.random.default_rng(3)
    full = rng.integers(1, 60000, size=(2, 36, 128, 128)).astype(np.uint16)
    c = np.array([9 + 18 * i for i in range(7)])
    want = np.stack([[full[b, 12 + v][c[:, None], c[None, :]].reshape(49) for v in range(12)] for b in range(2)])
    got = GridMapBuilder.subsample_depth(full)
    assert got.shape == (2, 12, 49) and np.array_equal(got, want)
    assert np.array_equal(GridMapBuilder.subsample_depth(full[0]), want[0])              # one episode, no batch axis
    ce = rng.random((2, 12, 256, 256)).astype(np.float32)
    c = np.array([19 + 36 * i for i in range(7)])
    want = np.stack([[ce[b, v][c[:, None], c[None, :]].reshape(49) for v in range(12)] for b in range(2)])
    assert np.array_equal(GridMapBuilder.subsample_depth(ce, ce=True), want)
    assert np.array_equal(GridMapBuilder.subsample_depth(synth.expand_depth(want[0].astype(np.uint16))), want[0].astype(np.uint16))


def test_wrappers_and_ctypes_table_agree_with_the_header():
    """Three statements of every entry point must agree without a GPU: the prototype in include/gridmm_b200.h, the ctypes
    signature table (_lib._SIGS) and the number of arguments the tensor-level wrapper in ops.py passes."""
    import ast
    import re
    from gridmm_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "gridmm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = dict(re.findall(r"\bint\s+(gridmm_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S))
    for name, sig in _lib._SIGS.items():
        assert name in protos, name
        n_hdr = len([a for a in protos[name].split(",") if a.strip() and a.strip() != "void"])
        assert n_hdr == len(sig), "%s: header has %d parameters, ctypes table %d" % (name, n_hdr, len(sig))
    tree = ast.parse(open(os.path.join(root, "gridmm_b200", "ops.py")).read())
    seen = set()
    for node in ast.walk(tree):
        if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "call" and node.args
                and isinstance(node.args[0], ast.Constant)):
            name = node.args[0].value
            seen.add(name)
            assert len(node.args) - 1 == len(_lib._SIGS[name]), "%s: ops.py passes %d arguments, ctypes table has %d" % (
                name, len(node.args) - 1, len(_lib._SIGS[name]))
    assert seen == set(_lib._SIGS), set(_lib._SIGS) ^ seen


def test_feature_files_read_the_reference_formats(tmp_path):
    """gridmm_b200/io.py against what the reference's DepthFeaturesDB / SemanticFeaturesDB (map_nav_src/r2r/env.py:80-113) and the
    depth sub-sampling of getGlobalMap (:279-285) would produce from the same stored arrays (an in-memory store stands in for
    h5py.File: this image has no h5py)."""
    import json
    from gridmm_b200.io import FeatureFiles
    from gridmm_b200.env import GridMapBuilder
    rng = np.random.default_rng(3)
    keys = [("scanA", "vp1"), ("scanA", "vp2"), ("scanB", "vp9")]
    clip_store, depth_store, info = {}, {}, {}
    for i, (s, v) in enumerate(keys):
        k = "%s_%s" % (s, v)
        clip_store[k] = rng.standard_normal((12, 50 + i, 768))                     # written as "float" (float64), >= 50 tokens
        d = rng.integers(0, 40000, (36, 128, 128)).astype(np.float64)
        depth_store[k] = d[..., None] if i == 1 else (d.reshape(36, -1) if i == 2 else d)      # the three layouts seen in the wild
        info[k] = {"x": float(i), "y": 2.0 * i, "z": 0.5}
    vinfo = tmp_path / "viewpoint_info.json"
    vinfo.write_text(json.dumps(info))
    stores = {"clip.h5": clip_store, "depth.h5": depth_store}
    ff = FeatureFiles("clip.h5", "depth.h5", str(vinfo), open_fn=lambda path: stores[path])
    for i, (s, v) in enumerate(keys):
        k = "%s_%s" % (s, v)
        # SemanticFeaturesDB: f[key][...][:, :50].astype(float16)
        assert np.array_equal(ff.clip_tokens(s, v), clip_store[k][:, :50].astype(np.float16))
        # DepthFeaturesDB + getGlobalMap's sub-sampling: depth[:, idx][:, :, idx].reshape(36, -1), rows 12..23 are used
        full = np.asarray(depth_store[k]).reshape(36, 128, 128).astype(np.uint16)
        idx = np.array([9 + 18 * j for j in range(7)])
        ref = full[:, idx][:, :, idx].reshape(36, -1)[12:24]
        assert np.array_equal(GridMapBuilder.subsample_depth(ff.depth_map(s, v)), ref)
        assert ff.position(s, v) == (float(i), 2.0 * i)
    depth, clip, pos = ff.step_inputs(keys)
    assert depth.shape == (3, 12, 49) and depth.dtype == np.uint16 and clip.shape == (3, 12, 50, 768) and clip.dtype == np.float16
    assert pos.shape == (3, 2) and ff.step_inputs(keys, with_clip=False)[1] is None

    class FakeDB(dict):
        def put(self, k, fts):
            self[k] = fts
    db = FakeDB()
    assert ff.preload(db, keys) == 3 and ff.preload(db, keys) == 0 and db["scanB_vp9"].shape == (12, 50, 768)
    with pytest.raises(ValueError):
        FeatureFiles("clip.h5", "depth.h5", open_fn=lambda path: {"a_b": np.zeros((12, 50, 512))}).clip_tokens("a", "b")

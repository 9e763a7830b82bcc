"""Generate tests/golden/*.npz by running the UNMODIFIED reference (MrZihan/GridMM, /root/reference) in-process.

TEST INFRASTRUCTURE.  Run in the authoring container only (`python -m oracle.make_golden`); the fixtures are committed
because /root/reference does not exist on the GPU box.  Inputs and weights are NOT stored: they are regenerated from
seeds by gridmm_b200/synth.py (numpy PCG64 streams are platform independent), only the reference's outputs are.

  grid_r2r_s{seed}.npz   EnvBatch.getGlobalMap (map_nav_src/r2r/env.py:267-374) over T consecutive steps:
                         cell ids per step (int16, -1 masked) and gridmap_pos_fts of the last step
  grid_pretrain_s{seed}.npz  the pretraining dataset's getGlobalMap (pretrain_src/data/dataset.py:351-473) over whole ground-truth
                         paths (as get_traj_pano_fts :482-507 drives it): cell ids per step, gridmap_pos_fts and target_patch_id
  pretrain_small.npz     the pretraining trunk GlocalTextPathCMT.forward / forward_mlm (pretrain_src/model/vilmodel.py:668-855,
                         fp16 pooling) on one collated batch: gmap / vp embeddings, grid-encoded gmap rows, MLM text states
  nav_{name}.npz         GlocalTextPathNavCMT.forward('navigation', ...) (map_nav_src/models/vilmodel.py:782-918):
                         all five logit tensors + gmap/vp embeddings
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gridmm_b200 import synth  # noqa: E402
from oracle import _refshim    # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

GRID_CASES = [dict(seed=11, batch=3, steps=6), dict(seed=12, batch=2, steps=15)]
NAV_CASES = {
    # name: (episode kwargs, nav-input kwargs, model kwargs)
    "r2r_small": (dict(batch=2, steps=3, seed=21), dict(txt_len=32, gmap_len=12, n_views=36, n_objs=0), dict(obj_feat_size=0)),
    "reverie_small": (dict(batch=2, steps=2, seed=22), dict(txt_len=24, gmap_len=10, n_views=36, n_objs=8), dict(obj_feat_size=768)),
}
MODEL_KW = dict(num_l_layers=1, num_pano_layers=1, num_x_layers=4)


def reference_grid(ep):
    """Run the reference getGlobalMap for every episode/step of `ep`; returns per-step cell arrays etc."""
    B, T = ep["pos"].shape[:2]
    env, _ = _refshim.load_reference_env(B)
    cells = [[None] * T for _ in range(B)]
    fts, pos_fts = [None] * B, [None] * B
    for t in range(T):
        for b in range(B):
            f, gm, pf = _refshim.ref_step(env, b, "scan%d" % b, "vp%d" % t, float(ep["heading"][b, t]),
                                          synth.expand_depth(ep["depth_sub"][b, t]), ep["clip"][b, t], ep["pos"][b, t])
            cells[b][t] = gm.astype(np.int16)
            fts[b], pos_fts[b] = f, pf
    return cells, fts, pos_fts


def make_grid():
    for case in GRID_CASES:
        ep = synth.make_episodes(case["batch"], case["steps"], seed=case["seed"], dim=768)
        cells, _, pos_fts = reference_grid(ep)
        out = {"pos_fts_last": np.stack(pos_fts).astype(np.float32)}
        for b in range(case["batch"]):
            for t in range(case["steps"]):
                out["cell_b%d_t%d" % (b, t)] = cells[b][t]
        path = os.path.join(GOLD, "grid_r2r_s%d.npz" % case["seed"])
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


PRETRAIN_GRID_CASE = dict(seed=51, batch=3, steps=7)


def make_pretrain_grid():
    case = PRETRAIN_GRID_CASE
    ep = synth.make_episodes(case["batch"], case["steps"], seed=case["seed"], dim=768)
    heads = synth.pretrain_headings(ep)
    ds = _refshim.load_reference_pretrain_data()
    out = {}
    for b in range(case["batch"]):
        T = case["steps"]
        depth = [synth.expand_depth(ep["depth_sub"][b, t]) for t in range(T)]
        clip = [ep["clip"][b, t].astype(np.float32) for t in range(T)]           # SemanticFeaturesDB returns float32 (dataset.py:78)
        cells, pos_fts, targets, fts = _refshim.ref_pretrain_traj(ds, "scan%d" % b, [float(x) for x in heads[b]], depth, clip,
                                                                   ep["pos"][b])
        assert fts.shape == (588 * T, 768) and np.array_equal(fts.astype(np.float16)[:588], ep["clip"][b, 0][:, 1:].reshape(-1, 768))
        for t in range(T):
            out["cell_b%d_t%d" % (b, t)] = cells[t].astype(np.int16)
        out["pos_fts_b%d" % b] = np.stack(pos_fts).astype(np.float32)            # [T,196,5]
        out["target_b%d" % b] = np.array(targets, np.int32)
    path = os.path.join(GOLD, "grid_pretrain_s%d.npz" % case["seed"])
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), {k: out[k].tolist() for k in out if k.startswith("target")})


PRETRAIN_MODEL_CASE = dict(seed=61, batch=3, max_steps=4, txt_len=40, model=dict(num_l_layers=2, num_pano_layers=2, num_x_layers=4))


def make_pretrain_model():
    """pretrain_small.npz: the pretraining trunk's `forward` and `forward_mlm` (pretrain_src/model/vilmodel.py:668-855) on a
    collated batch whose grids come from the pretraining dataset's own getGlobalMap; pretrain_small_spec.json: parameter names
    and shapes of that model (the weights themselves are regenerated from the seed by synth.make_weights)."""
    import json
    case = PRETRAIN_MODEL_CASE
    B, seed = case["batch"], case["seed"]
    model = _refshim.load_reference_pretrain_model(**case["model"])
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    w = synth.make_weights(shapes, seed=seed)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    assert model.grid_proj.weight.dtype == torch.float16
    pb = synth.make_pretrain_batch(B, seed=seed, txt_len=case["txt_len"], max_steps=case["max_steps"])
    ep = synth.make_episodes(B, case["max_steps"], seed=seed, dim=768)
    heads = synth.pretrain_headings(ep)
    ds = _refshim.load_reference_pretrain_data()
    gfts, gmap, gpos = [], [], []
    for b in range(B):
        T = pb["traj_step_lens"][b]
        cells, pos_fts, _, fts = _refshim.ref_pretrain_traj(
            ds, "scan%d" % b, [float(x) for x in heads[b, :T]], [synth.expand_depth(ep["depth_sub"][b, t]) for t in range(T)],
            [ep["clip"][b, t].astype(np.float32) for t in range(T)], ep["pos"][b, :T])
        gfts.append(torch.from_numpy(fts).to(torch.float16))                # tasks.py:95
        gmap.append(torch.from_numpy(cells[-1]).to(torch.int64))             # tasks.py:96
        gpos.append(pos_fts[-1])
    t = lambda k: torch.from_numpy(pb[k])
    args = (t("txt_ids"), t("txt_lens"), t("traj_view_img_fts"), None, t("traj_loc_fts"), t("traj_nav_types"), pb["traj_step_lens"],
            t("traj_vp_view_lens"), None, pb["traj_vpids"], pb["traj_cand_vpids"], t("gmap_lens"), t("gmap_step_ids"),
            t("gmap_pos_fts"), t("gmap_pair_dists"), pb["gmap_vpids"], t("vp_pos_fts"), gfts, gmap)
    gp = torch.from_numpy(np.stack(gpos).astype(np.float32))
    with torch.no_grad():
        gmap_e, vp_e, grid_g = model.forward(*args, None, gp)
        txt = model.forward_mlm(*args, gp)
    save = {"gmap_embeds": gmap_e.numpy(), "vp_embeds": vp_e.numpy(), "grid_gmap_embeds": grid_g.numpy(), "mlm_txt_embeds": txt.numpy()}
    path = os.path.join(GOLD, "pretrain_small.npz")
    np.savez_compressed(path, **save)
    json.dump(shapes, open(os.path.join(GOLD, "pretrain_small_spec.json"), "w"), indent=0)
    print("wrote", path, os.path.getsize(path), {k: tuple(v.shape) for k, v in save.items()})


PRETRAIN_OBJ_CASE = dict(seed=62, batch=3, max_steps=3, txt_len=32, n_objs=6,
                         model=dict(num_l_layers=1, num_pano_layers=2, num_x_layers=2, obj_feat_size=768))


def make_pretrain_obj():
    """pretrain_obj_small.npz: the pretraining trunk's `forward` on a REVERIE-style batch WITH object tokens
    (`traj_obj_img_fts` / `traj_vp_obj_lens`, pretrain_src/model/vilmodel.py:496-512, 722-731)."""
    import json
    case = PRETRAIN_OBJ_CASE
    B, seed = case["batch"], case["seed"]
    model = _refshim.load_reference_pretrain_model(**case["model"])
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    w = synth.make_weights(shapes, seed=seed)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    pb = synth.make_pretrain_batch(B, seed=seed, txt_len=case["txt_len"], max_steps=case["max_steps"], n_objs=case["n_objs"])
    ep = synth.make_episodes(B, case["max_steps"], seed=seed, dim=768)
    heads = synth.pretrain_headings(ep)
    ds = _refshim.load_reference_pretrain_data()
    gfts, gmap, gpos = [], [], []
    for b in range(B):
        T = pb["traj_step_lens"][b]
        cells, pos_fts, _, fts = _refshim.ref_pretrain_traj(
            ds, "scan%d" % b, [float(x) for x in heads[b, :T]], [synth.expand_depth(ep["depth_sub"][b, t]) for t in range(T)],
            [ep["clip"][b, t].astype(np.float32) for t in range(T)], ep["pos"][b, :T])
        gfts.append(torch.from_numpy(fts).to(torch.float16))
        gmap.append(torch.from_numpy(cells[-1]).to(torch.int64))
        gpos.append(pos_fts[-1])
    t = lambda k: torch.from_numpy(pb[k])
    args = (t("txt_ids"), t("txt_lens"), t("traj_view_img_fts"), t("traj_obj_img_fts"), t("traj_loc_fts"), t("traj_nav_types"),
            pb["traj_step_lens"], t("traj_vp_view_lens"), t("traj_vp_obj_lens"), pb["traj_vpids"], pb["traj_cand_vpids"], t("gmap_lens"),
            t("gmap_step_ids"), t("gmap_pos_fts"), t("gmap_pair_dists"), pb["gmap_vpids"], t("vp_pos_fts"), gfts, gmap)
    gp = torch.from_numpy(np.stack(gpos).astype(np.float32))
    with torch.no_grad():
        gmap_e, vp_e, grid_g = model.forward(*args, None, gp)
    save = {"gmap_embeds": gmap_e.numpy(), "vp_embeds": vp_e.numpy(), "grid_gmap_embeds": grid_g.numpy()}
    path = os.path.join(GOLD, "pretrain_obj_small.npz")
    np.savez_compressed(path, **save)
    json.dump(shapes, open(os.path.join(GOLD, "pretrain_obj_small_spec.json"), "w"), indent=0)
    print("wrote", path, os.path.getsize(path), {k: tuple(v.shape) for k, v in save.items()})


def make_pretrain_heads():
    """pretrain_heads_small.npz: `GlocalTextPathCMTPreTraining.forward_sap` (logits and per-sample losses) and `.forward_mlm`
    (vocabulary scores at the masked positions) on the batch of pretrain_small.npz (pretrain_src/model/pretrain_cmt.py:128-292);
    pretrain_heads_small_spec.json: the wrapper's parameter names and shapes."""
    import json
    case = PRETRAIN_MODEL_CASE
    B, seed = case["batch"], case["seed"]
    model = _refshim.load_reference_pretrain_wrapper(**case["model"])
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    w = synth.make_weights(shapes, seed=seed)
    w["mlm_head.predictions.decoder.weight"] = w["bert.embeddings.word_embeddings.weight"]       # tie_weights, pretrain_cmt.py:68-71
    model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    pb = synth.make_pretrain_batch(B, seed=seed, txt_len=case["txt_len"], max_steps=case["max_steps"])
    lab = synth.make_pretrain_labels(pb, seed=seed)
    ep = synth.make_episodes(B, case["max_steps"], seed=seed, dim=768)
    heads = synth.pretrain_headings(ep)
    ds = _refshim.load_reference_pretrain_data()
    gfts, gmap, gpos = [], [], []
    for b in range(B):
        T = pb["traj_step_lens"][b]
        cells, pos_fts, _, fts = _refshim.ref_pretrain_traj(
            ds, "scan%d" % b, [float(x) for x in heads[b, :T]], [synth.expand_depth(ep["depth_sub"][b, t]) for t in range(T)],
            [ep["clip"][b, t].astype(np.float32) for t in range(T)], ep["pos"][b, :T])
        gfts.append(torch.from_numpy(fts).to(torch.float16))
        gmap.append(torch.from_numpy(cells[-1]).to(torch.int64))
        gpos.append(pos_fts[-1])
    t = lambda k: torch.from_numpy(pb[k])
    common = (t("txt_ids"), t("txt_lens"), t("traj_view_img_fts"), None, t("traj_loc_fts"), t("traj_nav_types"), pb["traj_step_lens"],
              t("traj_vp_view_lens"), None, pb["traj_vpids"], pb["traj_cand_vpids"], t("gmap_lens"), t("gmap_step_ids"),
              t("gmap_pos_fts"), t("gmap_pair_dists"), pb["gmap_vpids"], t("vp_pos_fts"))
    gp = torch.from_numpy(np.stack(gpos).astype(np.float32))
    lt = {k: torch.from_numpy(v) for k, v in lab.items()}
    with torch.no_grad():
        sap_args = common + (lt["gmap_visited_masks"], lt["global_act_labels"], lt["local_act_labels"], gfts, gmap, None, gp)
        gl, ll, fused, _, _ = model.forward_sap(*sap_args, False)
        losses = model.forward_sap(*sap_args, True)
        scores = model.forward_mlm(*common, lt["txt_labels"], gfts, gmap, None, gp, False)
    save = {"global_logits": gl.numpy(), "local_logits": ll.numpy(), "fused_logits": fused.numpy(), "sap_losses": losses.numpy(),
            "mlm_scores": scores.numpy().astype(np.float32)}
    path = os.path.join(GOLD, "pretrain_heads_small.npz")
    np.savez_compressed(path, **save)
    json.dump(shapes, open(os.path.join(GOLD, "pretrain_heads_small_spec.json"), "w"), indent=0)
    print("wrote", path, os.path.getsize(path), {k: tuple(v.shape) for k, v in save.items()}, losses.tolist())


def reference_nav(ep_kw, nav_kw, model_kw):
    from gridmm_b200.model import NavConfig, param_spec
    ep = synth.make_episodes(dim=768, **ep_kw)
    cells, fts, pos_fts = reference_grid(ep)
    B, T = ep["pos"].shape[:2]
    kw = dict(MODEL_KW); kw.update(model_kw)
    model = _refshim.load_reference_model(seed=0, **kw)
    spec = param_spec(NavConfig(**kw))
    ref_keys = list(model.state_dict().keys())
    assert ref_keys == list(spec.keys()), "param_spec order/keys differ from the reference state_dict"
    w = synth.make_weights({k: v[0] for k, v in spec.items()}, seed=ep_kw["seed"])
    model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    nav = synth.to_torch(synth.make_nav_inputs(ep_kw["batch"], seed=ep_kw["seed"], **nav_kw))
    nav["grid_fts"] = [torch.from_numpy(np.ascontiguousarray(f)) for f in fts]
    nav["grid_map"] = [torch.from_numpy(cells[b][T - 1].astype(np.float64)) for b in range(B)]
    nav["gridmap_pos_fts"] = torch.from_numpy(np.stack(pos_fts).astype(np.float32))
    with torch.no_grad():
        outs = model("navigation", nav)
    return outs


AUX_CASES = {
    # forward('language') and forward('panorama'): name -> (seed, model kwargs, input kwargs)
    "lang_small": (31, dict(num_l_layers=2), dict(batch=3, txt_len=24)),
    "pano_r2r": (32, dict(num_pano_layers=2), dict(batch=3, n_views=36, n_objs=0)),
    "pano_reverie": (33, dict(num_pano_layers=2, obj_feat_size=768), dict(batch=3, n_views=36, n_objs=8)),
}


def make_aux():
    from gridmm_b200.model import NavConfig, param_spec
    for name, (seed, model_kw, in_kw) in AUX_CASES.items():
        kw = dict(MODEL_KW); kw.update(model_kw)
        model = _refshim.load_reference_model(seed=0, **kw)
        spec = param_spec(NavConfig(**kw))
        assert list(model.state_dict().keys()) == list(spec.keys())
        w = synth.make_weights({k: v[0] for k, v in spec.items()}, seed=seed)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
        with torch.no_grad():
            if name.startswith("lang"):
                out = model("language", synth.to_torch(synth.make_lang_inputs(seed=seed, **in_kw)))
                save = {"txt_embeds": out.numpy()}
            else:
                embeds, masks = model("panorama", synth.to_torch(synth.make_pano_inputs(seed=seed, **in_kw)))
                save = {"pano_embeds": embeds.numpy(), "pano_masks": masks.numpy()}
        path = os.path.join(GOLD, "aux_%s.npz" % name)
        np.savez_compressed(path, **save)
        print("wrote", path, os.path.getsize(path), {k: tuple(v.shape) for k, v in save.items()})


CE_GRID_CASE = dict(seed=41, batch=3, steps=6)
CE_NAV_CASE = (dict(batch=3, steps=3, seed=42), dict(txt_len=24, gmap_len=10, n_views=12, n_objs=0))


def reference_ce_grid(ep, dataset="R2R", max_dist=25):
    B, T = ep["pos"].shape[:2]
    g = _refshim.load_reference_ce_grid(B, dataset=dataset, max_dist=max_dist)
    dep = (ep["depth_sub"].astype(np.float32) / 4000.0).astype(np.float32)          # CE depth is float32 metres
    cells = [[None] * T for _ in range(B)]
    fts, pos_fts = [None] * B, [None] * B
    for t in range(T):
        for b in range(B):
            f, gm, pf = _refshim.ref_ce_step(g, b, float(ep["heading"][b, t]), synth.expand_depth_ce(dep[b, t]),
                                             ep["clip"][b, t], ep["pos"][b, t])
            cells[b][t] = gm.astype(np.int16)
            fts[b], pos_fts[b] = f, pf
    return cells, fts, pos_fts


def make_ce():
    """Continuous-env variant: the reference's CE getGlobalMap (compiled from its source text) and the CE
    copy of GlocalTextPathNavCMT.forward('navigation', 14-tuple)."""
    from gridmm_b200.model import NavConfig, param_spec
    case = CE_GRID_CASE
    ep = synth.make_episodes(case["batch"], case["steps"], seed=case["seed"], dim=768)
    cells, _, pos_fts = reference_ce_grid(ep)
    out = {"pos_fts_last": np.stack(pos_fts).astype(np.float32)}
    for b in range(case["batch"]):
        for t in range(case["steps"]):
            out["cell_b%d_t%d" % (b, t)] = cells[b][t]
    path = os.path.join(GOLD, "grid_ce_s%d.npz" % case["seed"])
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))

    ep_kw, nav_kw = CE_NAV_CASE
    ep = synth.make_episodes(dim=768, **ep_kw)
    cells, fts, pos_fts = reference_ce_grid(ep)
    B, T = ep["pos"].shape[:2]
    kw = dict(MODEL_KW, graph_sprels=False)            # the CE copy of GlobalMapEncoder has no sprel_linear
    model = _refshim.load_reference_ce_model(**kw)
    spec = param_spec(NavConfig(**kw))
    w = synth.make_weights({k: v[0] for k, v in spec.items()}, seed=ep_kw["seed"])
    res = model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith(("clip.", "visual_encoder.")) for k in res.missing_keys), res.missing_keys
    nav = synth.to_torch(synth.make_nav_inputs(ep_kw["batch"], seed=ep_kw["seed"], **nav_kw))
    cand = [int(x) for x in nav["vp_nav_masks"].sum(1)]
    tup = (nav["txt_embeds"], nav["txt_masks"], nav["gmap_img_embeds"], nav["gmap_step_ids"], nav["gmap_pos_fts"],
           nav["gmap_masks"], nav["vp_img_embeds"], nav["vp_pos_fts"], nav["vp_masks"], nav["vp_nav_masks"],
           [torch.from_numpy(np.ascontiguousarray(f)) for f in fts],
           [torch.from_numpy(cells[b][T - 1].astype(np.float64)) for b in range(B)],
           torch.from_numpy(np.stack(pos_fts).astype(np.float32)), cand)
    with torch.no_grad():
        logits = model("navigation", tup)
    path = os.path.join(GOLD, "nav_ce_small.npz")
    np.savez_compressed(path, fused_logits=logits.numpy(), candidate_lengths=np.array(cand))
    print("wrote", path, os.path.getsize(path), tuple(logits.shape))


def make_nav():
    for name, (ep_kw, nav_kw, model_kw) in NAV_CASES.items():
        outs = reference_nav(ep_kw, nav_kw, model_kw)
        save = {k: v.numpy() for k, v in outs.items() if v is not None}
        path = os.path.join(GOLD, "nav_%s.npz" % name)
        np.savez_compressed(path, **save)
        print("wrote", path, os.path.getsize(path), {k: tuple(v.shape) for k, v in save.items()})


RXR_CE_GRID_CASE = dict(seed=43, batch=3, steps=4)


def make_rxr_ce():
    """RxR-CE conventions of the same source (DATASET = 'RxR': 79-degree camera, Policy_ViewSelection_GridMap.py:635-638; MAX_DIST
    40, :282-285): cell ids per step and gridmap_pos_fts of the last step."""
    case = RXR_CE_GRID_CASE
    ep = synth.make_episodes(case["batch"], case["steps"], seed=case["seed"], dim=768)
    cells, _, pos_fts = reference_ce_grid(ep, dataset="RxR", max_dist=40)
    out = {"pos_fts_last": np.stack(pos_fts).astype(np.float32)}
    for b in range(case["batch"]):
        for t in range(case["steps"]):
            out["cell_b%d_t%d" % (b, t)] = cells[b][t]
    path = os.path.join(GOLD, "grid_rxrce_s%d.npz" % case["seed"])
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    if not _refshim.available():
        raise SystemExit("reference not available at %s" % _refshim.REF_ROOT)
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    which = sys.argv[1:] or ["grid", "nav", "aux", "ce", "rxr_ce", "pretrain", "pretrain_model", "pretrain_heads", "pretrain_obj"]
    if "grid" in which:
        make_grid()
    if "nav" in which:
        make_nav()
    if "aux" in which:
        make_aux()
    if "ce" in which:
        make_ce()
    if "rxr_ce" in which:
        make_rxr_ce()
    if "pretrain" in which:
        make_pretrain_grid()
    if "pretrain_model" in which:
        make_pretrain_model()
    if "pretrain_heads" in which:
        make_pretrain_heads()
    if "pretrain_obj" in which:
        make_pretrain_obj()

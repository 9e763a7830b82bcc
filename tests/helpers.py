"""Shared builders for the parity tests: seeded inputs (gridmm_b200/synth.py), weights, oracle runs."""
import os

import numpy as np
import torch

from gridmm_b200 import synth
from gridmm_b200.model import NavConfig, param_spec

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_KW = dict(num_l_layers=1, num_pano_layers=1, num_x_layers=4)     # the nav path never touches lang/pano encoders

# must mirror oracle/make_golden.py
GRID_CASES = [dict(seed=11, batch=3, steps=6), dict(seed=12, batch=2, steps=15)]
NAV_CASES = {
    "r2r_small": (dict(batch=2, steps=3, seed=21), dict(txt_len=32, gmap_len=12, n_views=36, n_objs=0), dict(obj_feat_size=0)),
    "reverie_small": (dict(batch=2, steps=2, seed=22), dict(txt_len=24, gmap_len=10, n_views=36, n_objs=8), dict(obj_feat_size=768)),
}


AUX_CASES = {
    "lang_small": (31, dict(num_l_layers=2), dict(batch=3, txt_len=24)),
    "pano_r2r": (32, dict(num_pano_layers=2), dict(batch=3, n_views=36, n_objs=0)),
    "pano_reverie": (33, dict(num_pano_layers=2, obj_feat_size=768), dict(batch=3, n_views=36, n_objs=8)),
}


PRETRAIN_GRID_CASE = dict(seed=51, batch=3, steps=7)      # mirrors oracle/make_golden.py
PRETRAIN_MODEL_CASE = dict(seed=61, batch=3, max_steps=4, txt_len=40, model=dict(num_l_layers=2, num_pano_layers=2, num_x_layers=4))
PRETRAIN_OBJ_CASE = dict(seed=62, batch=3, max_steps=3, txt_len=32, n_objs=6,
                         model=dict(num_l_layers=1, num_pano_layers=2, num_x_layers=2, obj_feat_size=768))     # mirrors make_golden.py
CE_GRID_CASE = dict(seed=41, batch=3, steps=6)
RXR_CE_GRID_CASE = dict(seed=43, batch=3, steps=4)         # mirrors oracle/make_golden.py
CE_NAV_CASE = (dict(batch=3, steps=3, seed=42), dict(txt_len=24, gmap_len=10, n_views=12, n_objs=0))


def ce_episodes(ep_kw):
    """CE episodes: same generator, depth converted to float32 metres (the CE policy's depth sensor)."""
    ep = synth.make_episodes(ep_kw["batch"], ep_kw["steps"], seed=ep_kw["seed"], dim=768)
    ep["depth_sub"] = (ep["depth_sub"].astype(np.float32) / 4000.0).astype(np.float32)
    return ep


def ce_nav_tuple(ep_kw, nav_kw, cells, fts, pos, device="cpu"):
    """The 14-tuple of Policy_ViewSelection_GridMap.py:622-623."""
    nav = synth.to_torch(synth.make_nav_inputs(ep_kw["batch"], seed=ep_kw["seed"], **nav_kw), device)
    T = ep_kw["steps"]
    cand = [int(x) for x in nav["vp_nav_masks"].sum(1)]
    return (nav["txt_embeds"], nav["txt_masks"], nav["gmap_img_embeds"], nav["gmap_step_ids"], nav["gmap_pos_fts"],
            nav["gmap_masks"], nav["vp_img_embeds"], nav["vp_pos_fts"], nav["vp_masks"], nav["vp_nav_masks"],
            [torch.from_numpy(np.ascontiguousarray(f)).to(device) for f in fts],
            [torch.from_numpy(cells[b][T - 1].astype(np.float64)).to(device) for b in range(len(fts))],
            torch.from_numpy(np.stack(pos).astype(np.float32)).to(device), cand)


def make_config(**model_kw):
    kw = dict(MODEL_KW)
    kw.update(model_kw)
    return NavConfig(**kw)


def make_weights(cfg, seed):
    spec = param_spec(cfg)
    return synth.make_weights({k: v[0] for k, v in spec.items()}, seed=seed)


def oracle_grid(ep, grid_w=14, geom=None):
    """oracle.grid_oracle over all episodes/steps -> (cells[b][t] int32, fts[b] f16[N,D], half[b], pos_fts[b])."""
    from oracle import grid_oracle as go
    geom = geom or go.R2RGeometry
    B, T = ep["pos"].shape[:2]
    cells = [[None] * T for _ in range(B)]
    fts, halfs, pos = [None] * B, [None] * B, [None] * B
    for b in range(B):
        st = go.GridState()
        for t in range(T):
            f, c, h = go.grid_step(st, ep["depth_sub"][b, t], ep["clip"][b, t], ep["pos"][b, t], float(ep["heading"][b, t]),
                                   grid_w=grid_w, geom=geom)
            cells[b][t] = c
        fts[b], halfs[b] = f, h
        pos[b] = go.gridmap_pos_fts(h, grid_w, geom)
    return cells, fts, halfs, pos


def nav_batch(ep_kw, nav_kw, cells, fts, pos, device="cpu"):
    """Reference-format 'navigation' batch (map_nav_src/r2r/agent.py:163-205)."""
    nav = synth.to_torch(synth.make_nav_inputs(ep_kw["batch"], seed=ep_kw["seed"], **nav_kw), device)
    T = ep_kw["steps"]
    nav["grid_fts"] = [torch.from_numpy(np.ascontiguousarray(f)).to(device) for f in fts]
    nav["grid_map"] = [torch.from_numpy(cells[b][T - 1].astype(np.float64)).to(device) for b in range(len(fts))]
    nav["gridmap_pos_fts"] = torch.from_numpy(np.stack(pos).astype(np.float32)).to(device)
    return nav


def finite_close(a, b, atol):
    """same -inf pattern and |a-b| <= atol on finite entries; returns max abs error."""
    a = torch.as_tensor(a).float().cpu()
    b = torch.as_tensor(b).float().cpu()
    fa, fb = torch.isfinite(a), torch.isfinite(b)
    assert torch.equal(fa, fb), "finite/-inf pattern differs"
    assert torch.equal(a[~fa], b[~fb]), "non-finite entries differ"
    err = (a[fa] - b[fb]).abs().max().item() if fa.any() else 0.0
    assert err <= atol, "max abs error %.3e > %.1e" % (err, atol)
    return err


def pretrain_batch(case):
    """The collated pretraining batch of oracle/make_golden.py's PRETRAIN_MODEL_CASE, with the grid tensors from the oracle's grid
    build over each path (bit-identical to the pretraining dataset's, see test_grid_oracle_matches_pretraining_dataset)."""
    from oracle import grid_oracle as go
    B = case["batch"]
    pb = synth.make_pretrain_batch(B, seed=case["seed"], txt_len=case["txt_len"], max_steps=case["max_steps"], n_objs=case.get("n_objs", 0))
    ep = synth.make_episodes(B, case["max_steps"], seed=case["seed"], dim=768)
    heads = synth.pretrain_headings(ep)
    batch = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
    gfts, gmap, gpos = [], [], []
    for b in range(B):
        st = go.GridState()
        for t in range(pb["traj_step_lens"][b]):
            f, c, h = go.grid_step(st, ep["depth_sub"][b, t], ep["clip"][b, t], ep["pos"][b, t], float(heads[b, t]))
        gfts.append(torch.from_numpy(np.ascontiguousarray(f)))                       # fp16 [588 T_b, 768]
        gmap.append(torch.from_numpy(c.astype(np.int64)))
        gpos.append(go.gridmap_pos_fts(h))
    batch.update(grid_fts=gfts, grid_map=gmap, gridmap_pos_fts=torch.from_numpy(np.stack(gpos).astype(np.float32)))
    return batch

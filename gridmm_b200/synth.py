"""Seeded synthetic inputs of the shapes BASELINE.json names.

There is no Matterport data and no checkpoint in this environment, so every
test and the bench run on these.  The same generator feeds the reference (in the
authoring container, to produce tests/golden/*), the oracle and the CUDA path,
so all three see bit-identical inputs.

Shapes follow the reference's on-disk formats:
  depth     uint16[36,128,128], 0.25 mm units, 0 = invalid
            (map_nav_src/r2r/env.py:80-95, 278-285) -- only views 12..23 and the
            7x7 patch-centre pixels 9+18*i are read, so the generator draws the
            sub-sampled uint16[12,49] and `expand_depth` scatters it into a full map.
  clip      float16[12,50,768] CLIP ViT patch tokens, token 0 = CLS
            (map_nav_src/r2r/env.py:98-113, 296-304)
  pose      viewpoint x,y (python floats, viewpoint_info.json) and heading (rad).
"""
import math

import numpy as np

N_VIEWS = 12            # horizon views 12..23 of the 36 (env.py:289)
N_PATCH = 49            # 7x7 patch centres per view (env.py:279-281)
PTS_PER_VP = N_VIEWS * N_PATCH   # 588
PATCH_CENTRES = np.array([9 + 18 * i for i in range(7)])


def make_episodes(batch, steps, seed=0, dim=768, zero_frac=0.10):
    """Per-episode, per-step viewpoint observations.

    Returns a dict of numpy arrays:
      depth_sub u16[B,T,12,49]   clip f16[B,T,12,50,dim]
      pos f64[B,T,2]             heading f64[B,T]
    """
    rng = np.random.default_rng(seed)
    depth = rng.integers(2000, 20001, size=(batch, steps, N_VIEWS, N_PATCH)).astype(np.uint16)
    depth[rng.random(depth.shape) < zero_frac] = 0
    clip = rng.standard_normal((batch, steps, N_VIEWS, N_PATCH + 1, dim), dtype=np.float32).astype(np.float16)
    step_xy = rng.uniform(1.0, 3.0, size=(batch, steps, 2)) * rng.choice([-1.0, 1.0], size=(batch, steps, 2))
    start = rng.uniform(-10.0, 10.0, size=(batch, 1, 2))
    pos = start + np.cumsum(step_xy, axis=1)
    heading = rng.uniform(0.0, 2.0 * math.pi, size=(batch, steps))
    return {"depth_sub": depth, "clip": clip, "pos": pos, "heading": heading}


def pretrain_headings(ep):
    """Headings as the pretraining dataset sets them along a ground-truth path: (viewidx % 12) * 30 degrees of the candidate
    view that led to each viewpoint (pretrain_src/data/dataset.py:497-499); the first viewpoint keeps the start heading."""
    B, T = ep["pos"].shape[:2]
    rng = np.random.default_rng(1000 + int(B) * 31 + int(T))
    h = ep["heading"].astype(np.float64).copy()
    h[:, 1:] = rng.integers(0, 12, size=(B, T - 1)) * math.radians(30)
    return h


def expand_depth(depth_sub):
    """uint16[12,49] -> the reference's uint16[36,128,128] map (zeros elsewhere)."""
    full = np.zeros((36, 128, 128), dtype=np.uint16)
    sub = depth_sub.reshape(N_VIEWS, 7, 7)
    full[12:24][:, PATCH_CENTRES[:, None], PATCH_CENTRES[None, :]] = sub
    return full


PATCH_CENTRES_CE = np.array([19 + 36 * i for i in range(7)])


def expand_depth_ce(depth_sub_f32):
    """float32[12,49] metres -> the CE policy's float32[12,256,256] depth stack (zeros elsewhere)."""
    full = np.zeros((12, 256, 256), dtype=np.float32)
    full[:, PATCH_CENTRES_CE[:, None], PATCH_CENTRES_CE[None, :]] = depth_sub_f32.reshape(N_VIEWS, 7, 7)
    return full


def make_nav_inputs(batch, seed=0, txt_len=80, gmap_len=20, n_views=36, n_objs=0,
                    dim=768, min_txt=20, ragged=True):
    """Everything `forward('navigation')` needs except the grid tensors
    (map_nav_src/r2r/agent.py:96-205, map_nav_src/models/vilmodel.py:782-786).

    Returns a dict of numpy arrays / python lists (reference batch keys).
    """
    rng = np.random.default_rng(seed + 7919)
    B, L, G = batch, txt_len, gmap_len
    V = 1 + n_views + n_objs
    f32 = np.float32
    txt_embeds = rng.standard_normal((B, L, dim), dtype=f32)
    txt_lens = rng.integers(min_txt, L + 1, size=B) if ragged else np.full(B, L)
    txt_lens[0] = L
    txt_masks = np.arange(L)[None, :] < txt_lens[:, None]

    gmap_lens = rng.integers(max(4, G // 2), G + 1, size=B) if ragged else np.full(B, G)
    gmap_lens[0] = G
    gmap_masks = np.arange(G)[None, :] < gmap_lens[:, None]
    gmap_img_embeds = rng.standard_normal((B, G, dim), dtype=f32)
    gmap_img_embeds[:, 0] = 0.0                     # [stop] token (agent.py:129-131)
    gmap_img_embeds *= gmap_masks[:, :, None]       # pad_tensors_wgrad zero padding
    gmap_step_ids = rng.integers(0, 16, size=(B, G)).astype(np.int64) * gmap_masks
    gmap_step_ids[:, 0] = 0
    gmap_pos_fts = rng.standard_normal((B, G, 7), dtype=f32) * gmap_masks[:, :, None]
    gmap_visited = np.zeros((B, G), dtype=bool)
    gmap_vpids, vp_cand_vpids = [], []
    n_cands = rng.integers(2, 6, size=B)
    vp_lens = np.full(B, V)
    if ragged and n_objs > 0:
        vp_lens = 1 + n_views + rng.integers(0, n_objs + 1, size=B)
        vp_lens[0] = V
    vp_masks = np.arange(V)[None, :] < vp_lens[:, None]
    vp_nav_masks = np.zeros((B, V), dtype=bool)
    vp_nav_masks[:, 0] = True
    vp_obj_masks = np.zeros((B, V), dtype=bool) if n_objs > 0 else None
    for b in range(B):
        g = int(gmap_lens[b])
        n_vis = int(rng.integers(1, max(2, g // 2)))
        gmap_visited[b, 1:1 + n_vis] = True       # enc_full_graph order: [stop]+visited+unvisited
        ids = [None] + ["vp%d_%d" % (b, j) for j in range(1, g)] + [None] * (G - g)
        gmap_vpids.append(ids[:g])
        nc = int(n_cands[b])
        vp_nav_masks[b, 1:1 + nc] = True
        # candidates: a mix of unvisited gmap nodes, visited nodes and (rarely) nodes absent from gmap
        pool_unvis = ids[1 + n_vis:g]
        pool_vis = ids[1:1 + n_vis]
        cands = []
        for j in range(nc):
            r = rng.random()
            if r < 0.6 and len(pool_unvis) > 0:
                cands.append(pool_unvis[int(rng.integers(len(pool_unvis)))])
            elif len(pool_vis) > 0:
                cands.append(pool_vis[int(rng.integers(len(pool_vis)))])
            else:
                cands.append("ghost%d_%d" % (b, j))
        vp_cand_vpids.append([None] + cands)
        if vp_obj_masks is not None:
            vp_obj_masks[b, 1 + n_views:int(vp_lens[b])] = True
    vp_img_embeds = rng.standard_normal((B, V, dim), dtype=f32)
    vp_img_embeds[:, 0] = 0.0                       # [stop] (agent.py:176-178)
    vp_img_embeds *= vp_masks[:, :, None]
    vp_pos_fts = rng.standard_normal((B, V, 14), dtype=f32) * vp_masks[:, :, None]
    return {
        "txt_embeds": txt_embeds, "txt_masks": txt_masks,
        "gmap_img_embeds": gmap_img_embeds, "gmap_step_ids": gmap_step_ids,
        "gmap_pos_fts": gmap_pos_fts, "gmap_masks": gmap_masks,
        "gmap_pair_dists": np.zeros((B, G, G), dtype=f32),
        "gmap_visited_masks": gmap_visited, "gmap_vpids": gmap_vpids,
        "vp_img_embeds": vp_img_embeds, "vp_pos_fts": vp_pos_fts, "vp_masks": vp_masks,
        "vp_nav_masks": vp_nav_masks, "vp_obj_masks": vp_obj_masks,
        "vp_cand_vpids": vp_cand_vpids,
    }


def make_lang_inputs(batch, seed=0, txt_len=80, vocab=30522, min_txt=8):
    """`forward('language')` inputs (map_nav_src/r2r/agent.py:62-77): padded token ids + masks."""
    rng = np.random.default_rng(seed + 31337)
    lens = rng.integers(min_txt, txt_len + 1, size=batch)
    lens[0] = txt_len
    ids = rng.integers(1, vocab, size=(batch, txt_len)).astype(np.int64)
    masks = np.arange(txt_len)[None, :] < lens[:, None]
    ids = ids * masks                                   # pad id 0
    return {"txt_ids": ids, "txt_masks": masks}


def make_pano_inputs(batch, seed=0, n_views=36, n_objs=0, dim=768, loc_dim=7):
    """`forward('panorama')` inputs (map_nav_src/r2r/agent.py:79-129; reverie/agent_obj.py adds object tokens)."""
    rng = np.random.default_rng(seed + 4242)
    f32 = np.float32
    view_lens = np.full(batch, n_views, dtype=np.int64)
    out = {"view_img_fts": rng.standard_normal((batch, n_views, dim), dtype=f32), "view_lens": view_lens}
    if n_objs > 0:
        obj_lens = rng.integers(0, n_objs + 1, size=batch).astype(np.int64)
        obj_lens[0] = n_objs
        obj = rng.standard_normal((batch, n_objs, dim), dtype=f32)
        obj *= (np.arange(n_objs)[None, :] < obj_lens[:, None])[:, :, None]
        out.update(obj_img_fts=obj, obj_lens=obj_lens)
        n = int((view_lens + obj_lens).max())
    else:
        out.update(obj_img_fts=None, obj_lens=None)
        n = n_views
    out["loc_fts"] = rng.standard_normal((batch, n, loc_dim), dtype=f32)
    nav_types = rng.integers(0, 2, size=(batch, n)).astype(np.int64)
    if n_objs > 0:
        for b in range(batch):
            nav_types[b, n_views:n_views + int(out["obj_lens"][b])] = 2
            nav_types[b, n_views + int(out["obj_lens"][b]):] = 0
    out["nav_types"] = nav_types
    return out


def make_pretrain_batch(batch, seed=0, txt_len=40, max_steps=4, n_views=36, n_cands=3, dim=768, vocab=30522, min_txt=8, n_objs=0):
    """One collated pretraining batch (n_objs > 0: with REVERIE / SOON object tokens: `traj_obj_img_fts`, `traj_vp_obj_lens`, nav
    type 2, pretrain_src/model/vilmodel.py:487-512), in the layout pretrain_src/data/tasks.py's collate functions hand to
    `GlocalTextPathCMT.forward` / `forward_mlm` (pretrain_src/model/vilmodel.py:668-674, 767-771), minus the grid tensors
    (those come from the grid build over the same paths).  Episode b walks steps[b] viewpoints "v{b}_{t}"; each panorama has
    n_views views of which the first n_cands are navigable candidates: the next viewpoint of the path, the previous one, and
    fresh unvisited nodes "u{b}_{t}_{j}".  gmap = [stop] + visited + unvisited (tasks.py order)."""
    rng = np.random.default_rng(seed + 15485863)
    f32 = np.float32
    B, L = batch, txt_len
    steps = rng.integers(2, max_steps + 1, size=B)
    steps[0] = max_steps
    txt_lens = rng.integers(min_txt, L + 1, size=B)
    txt_lens[0] = L
    txt_ids = rng.integers(1000, vocab, size=(B, L)).astype(np.int64) * (np.arange(L)[None, :] < txt_lens[:, None])
    n_tot = int(steps.sum())
    traj_view_img_fts = rng.standard_normal((n_tot, n_views, dim), dtype=f32)
    traj_loc_fts = rng.standard_normal((n_tot, n_views, 7), dtype=f32)
    traj_nav_types = np.zeros((n_tot, n_views), dtype=np.int64)
    traj_nav_types[:, :n_cands] = 1
    traj_vpids, traj_cand_vpids, gmap_vpids, gmap_step = [], [], [], []
    for b in range(B):
        T = int(steps[b])
        vps = ["v%d_%d" % (b, t) for t in range(T)]
        cands_b, unvisited = [], []
        for t in range(T):
            c = []
            if t + 1 < T:
                c.append(vps[t + 1])
            if t > 0:
                c.append(vps[t - 1])
            j = 0
            while len(c) < n_cands:
                u = "u%d_%d_%d" % (b, t, j)
                c.append(u); unvisited.append(u); j += 1
            cands_b.append(c)
        traj_vpids.append(vps)
        traj_cand_vpids.append(cands_b)
        gmap_vpids.append([None] + vps + unvisited)
        gmap_step.append([0] + list(range(1, T + 1)) + [0] * len(unvisited))
    gmap_lens = np.array([len(g) for g in gmap_vpids], dtype=np.int64)
    G = int(gmap_lens.max())
    gmap_masks = np.arange(G)[None, :] < gmap_lens[:, None]
    gmap_step_ids = np.zeros((B, G), dtype=np.int64)
    for b in range(B):
        gmap_step_ids[b, :gmap_lens[b]] = gmap_step[b]
    gmap_pos_fts = rng.standard_normal((B, G, 7), dtype=f32) * gmap_masks[:, :, None]
    vp_pos_fts = rng.standard_normal((B, 1 + n_views, 14), dtype=f32)
    out = {
        "txt_ids": txt_ids, "txt_lens": txt_lens.astype(np.int64),
        "traj_view_img_fts": traj_view_img_fts, "traj_loc_fts": traj_loc_fts, "traj_nav_types": traj_nav_types,
        "traj_step_lens": [int(x) for x in steps], "traj_vp_view_lens": np.full(n_tot, n_views, dtype=np.int64),
        "traj_vpids": traj_vpids, "traj_cand_vpids": traj_cand_vpids,
        "gmap_lens": gmap_lens, "gmap_step_ids": gmap_step_ids, "gmap_pos_fts": gmap_pos_fts,
        "gmap_pair_dists": np.zeros((B, G, G), dtype=f32), "gmap_vpids": gmap_vpids, "vp_pos_fts": vp_pos_fts,
    }
    if n_objs > 0:
        # drawn AFTER everything above, so that the object-free batch of the same seed is unchanged
        obj_lens = rng.integers(0, n_objs + 1, size=n_tot).astype(np.int64)
        obj_lens[0] = n_objs                                  # the padded panorama length is n_views + n_objs
        obj_valid = np.arange(n_objs)[None, :] < obj_lens[:, None]
        out["traj_obj_img_fts"] = rng.standard_normal((n_tot, n_objs, dim), dtype=f32) * obj_valid[:, :, None]
        out["traj_vp_obj_lens"] = obj_lens
        out["traj_loc_fts"] = np.concatenate([traj_loc_fts, rng.standard_normal((n_tot, n_objs, 7), dtype=f32) * obj_valid[:, :, None]], 1)
        out["traj_nav_types"] = np.concatenate([traj_nav_types, 2 * obj_valid.astype(np.int64)], 1)
        out["vp_pos_fts"] = np.concatenate([vp_pos_fts, rng.standard_normal((B, n_objs, 14), dtype=f32)], 1)
    return out


def make_pretrain_labels(pb, seed=0, n_masked=2):
    """Task inputs that go with make_pretrain_batch: SAP (pretrain_src/data/tasks.py sap collate: gmap_visited_masks,
    global / local action labels, here the first unvisited node and the first candidate view) and MLM (txt_labels: the original
    id at n_masked masked positions per instruction, -1 elsewhere)."""
    rng = np.random.default_rng(seed + 32452843)
    B, G = pb["gmap_step_ids"].shape
    visited = np.zeros((B, G), dtype=bool)
    glob = np.zeros(B, dtype=np.int64)
    for b in range(B):
        T = pb["traj_step_lens"][b]
        visited[b, 1:1 + T] = True
        glob[b] = 1 + T
    txt_labels = np.full(pb["txt_ids"].shape, -1, dtype=np.int64)
    for b in range(B):
        pos = rng.choice(int(pb["txt_lens"][b]), size=n_masked, replace=False)
        txt_labels[b, pos] = pb["txt_ids"][b, pos]
    return {"gmap_visited_masks": visited, "global_act_labels": glob, "local_act_labels": np.ones(B, dtype=np.int64),
            "txt_labels": txt_labels}


def to_torch(nav, device="cpu"):
    """numpy nav-input dict -> torch tensors with the reference's dtypes."""
    import torch
    out = {}
    for k, v in nav.items():
        if isinstance(v, np.ndarray):
            out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(device)
        else:
            out[k] = v
    return out


def make_weights(shapes, seed=0):
    """Deterministic N(0,0.02) weights (LayerNorm gammas 1 + N(0,0.05)) for a
    {name: shape} spec, drawn in sorted-name order from numpy's PCG64 so that the
    authoring container (reference + golden vectors) and the GPU box build
    bit-identical state_dicts without shipping a 640 MB checkpoint."""
    rng = np.random.default_rng(seed + 104729)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        w = rng.standard_normal(shape, dtype=np.float32)
        if len(shape) == 1 and name.endswith(".weight"):
            out[name] = (1.0 + 0.05 * w).astype(np.float32)
        else:
            out[name] = (0.02 * w).astype(np.float32)
    return out

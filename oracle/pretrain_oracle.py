"""CPU restatement (torch) of the PRETRAINING trunk of GridMM -- TEST INFRASTRUCTURE.

Only tests/ and oracle/make_golden.py import this module; the product path (gridmm_b200/) never does.

Parity status: PINNED IN THE AUTHORING CONTAINER against the reference's own `GlocalTextPathCMT.forward` and `.forward_mlm`
(pretrain_src/model/vilmodel.py:668-764, 767-855) imported in-process (oracle/_refshim.py, torch 2.11 CPU, fp16 pooling executed by
torch's CPU half kernels); the reference's outputs on seeded inputs are committed as tests/golden/pretrain_small.npz and
tests/test_oracle_golden.py re-checks this file against them.  No CUDA path consumes this oracle yet (out of scope).

What differs from the navigation forward (oracle/model_oracle.py):
  * the model embeds the whole trajectory itself: every panorama of the path goes through ImageEmbeddings + the pano encoder
    (vilmodel.py:487-530), the gmap node features are averages over those tokens (GlobalMapEncoder._aggregate_gmap_features,
    :578-612) and the vp tokens are the last panorama (+ [stop]) (LocalVPEncoder.vp_input_embedding, :541-555);
  * relevance, grid_proj, the per-cell softmax and the weighted sum run in fp16 (:685-699; grid_proj is an fp16 Linear, :664);
  * `forward_mlm` swaps the roles in the fusion encoder: text tokens are the queries, [gmap'; vp] the context, through the
    lang_* halves of GraphLXRTXLayer (forward_lang2visn, :404-415).
The task heads on top of the trunk (`GlocalTextPathCMTPreTraining.forward_sap` / `.forward_mlm`,
pretrain_src/model/pretrain_cmt.py:128-153, 214-292) are restated by `sap` and `mlm_scores`; they take the wrapper's state_dict
(trunk under `bert.`) and are pinned by tests/golden/pretrain_heads_small.npz.
Functional over a plain state_dict with the reference's key names.
"""
import torch

from oracle import model_oracle as mo


def trajectory_embeddings(sd, batch, n_pano_layers=2):
    """ImageEmbeddings.forward (vilmodel.py:487-530; object tokens :496-512 when the batch carries `traj_obj_img_fts`) over all
    panoramas of all paths: [sum_T, V, 768] -> per episode a [T_b, V, 768] block and its token counts (views + objects)."""
    obj = batch.get("traj_obj_img_fts")
    pano = {"view_img_fts": batch["traj_view_img_fts"], "view_lens": batch["traj_vp_view_lens"], "obj_img_fts": obj,
            "obj_lens": batch.get("traj_vp_obj_lens") if obj is not None else None,
            "loc_fts": batch["traj_loc_fts"], "nav_types": batch["traj_nav_types"]}
    x, _ = mo.panorama(sd, pano, n_layers=n_pano_layers)            # same arithmetic as forward('panorama')
    steps = list(batch["traj_step_lens"])
    lens = batch["traj_vp_view_lens"] + (batch["traj_vp_obj_lens"] if obj is not None else 0)
    return list(torch.split(x, steps, 0)), list(torch.split(lens, steps, 0))


def aggregate_gmap_features(split_embeds, split_lens, traj_vpids, traj_cand_vpids, gmap_vpids):
    """GlobalMapEncoder._aggregate_gmap_features (vilmodel.py:578-612): a visited node is the mean of its own panorama's tokens,
    an unvisited node the mean of the candidate-view tokens that pointed at it (collected only while it was unvisited)."""
    out = []
    for i, emb in enumerate(split_embeds):
        lens = split_lens[i]
        valid = torch.arange(int(lens.max()))[None, :] < lens[:, None]
        e = emb[:, :int(lens.max())] * valid[:, :, None]
        visited, unvisited = {}, {}
        for t in range(e.shape[0]):
            visited[traj_vpids[i][t]] = e[t].sum(0) / lens[t]
            for j, vp in enumerate(traj_cand_vpids[i][t]):
                if vp not in visited:
                    unvisited.setdefault(vp, []).append(e[t][j])
        rows = [visited[vp] if vp in visited else torch.stack(unvisited[vp], 0).mean(0) for vp in gmap_vpids[i][1:]]
        out.append(torch.stack(rows, 0))
    G = max(r.shape[0] for r in out)
    padded = torch.stack([torch.nn.functional.pad(r, (0, 0, 0, G - r.shape[0])) for r in out], 0)
    return torch.cat([torch.zeros(len(out), 1, padded.shape[2]), padded], 1)          # [stop] first


def _inputs(sd, batch, n_l_layers, n_pano_layers):
    """Everything `forward` / `forward_mlm` compute before the grid pooling and the encoders, as a navigation-style batch."""
    txt_masks = torch.arange(batch["txt_ids"].shape[1])[None, :] < batch["txt_lens"][:, None]
    txt = mo.language(sd, {"txt_ids": batch["txt_ids"], "txt_masks": txt_masks}, n_layers=n_l_layers)
    split_embeds, split_lens = trajectory_embeddings(sd, batch, n_pano_layers)
    gmap_img = aggregate_gmap_features(split_embeds, split_lens, batch["traj_vpids"], batch["traj_cand_vpids"], batch["gmap_vpids"])
    gmap_masks = torch.arange(int(batch["gmap_lens"].max()))[None, :] < batch["gmap_lens"][:, None]
    # LocalVPEncoder.vp_input_embedding (:541-555): [stop] + the last panorama's tokens, lens + 1
    last = [e[-1] for e in split_embeds]
    vp_lens = torch.stack([l[-1] + 1 for l in split_lens], 0)
    V = int(vp_lens.max())
    vp_img = torch.cat([torch.zeros(len(last), 1, last[0].shape[1]), torch.stack(last, 0)], 1)[:, :V]
    vp_masks = torch.arange(V)[None, :] < vp_lens[:, None]
    return {"txt_embeds": txt, "txt_masks": txt_masks, "gmap_img_embeds": gmap_img, "gmap_step_ids": batch["gmap_step_ids"],
            "gmap_pos_fts": batch["gmap_pos_fts"], "gmap_masks": gmap_masks, "vp_img_embeds": vp_img,
            "vp_pos_fts": batch["vp_pos_fts"], "vp_masks": vp_masks, "grid_fts": batch["grid_fts"], "grid_map": batch["grid_map"],
            "gridmap_pos_fts": batch["gridmap_pos_fts"]}


def forward(sd, batch, n_l_layers=9, n_pano_layers=2, n_x_layers=4):
    """GlocalTextPathCMT.forward (pretrain_src/model/vilmodel.py:668-764) -> (gmap_embeds, vp_embeds, grid-encoded gmap rows)."""
    nav = _inputs(sd, batch, n_l_layers, n_pano_layers)
    gmap_e, vp_e, gmap2, _ = mo._nav_trunk(sd, nav, n_x_layers, 196, fts_dtype=torch.float16)
    return gmap_e, vp_e, gmap2


def forward_mlm(sd, batch, n_l_layers=9, n_pano_layers=2, n_x_layers=4):
    """GlocalTextPathCMT.forward_mlm (:767-855): text tokens attend to [gmap'; vp] through forward_lang2visn (:404-415)."""
    nav = _inputs(sd, batch, n_l_layers, n_pano_layers)
    _, vp_in, gmap2, _ = mo._nav_trunk(sd, nav, n_x_layers, 196, fts_dtype=torch.float16, stop_after_map=True)
    ctx = torch.cat([gmap2, vp_in], 1)
    ctx_add = mo.neg_mask(torch.cat([nav["gmap_masks"], nav["vp_masks"]], 1))
    txt, txt_add = nav["txt_embeds"], mo.neg_mask(nav["txt_masks"])
    for i in range(n_x_layers):
        p = "local_encoder.encoder.x_layers.%d" % i
        q = p + ".visual_attention"
        a = mo._attend(mo._lin(sd, q + ".att.query", txt), mo._lin(sd, q + ".att.key", ctx), mo._lin(sd, q + ".att.value", ctx), ctx_add)
        txt = mo._ln(sd, q + ".output.LayerNorm", mo._lin(sd, q + ".output.dense", a) + txt, 1e-12)
        txt = mo.bert_self_block(sd, p + ".lang_self_att", txt, txt_add)
        txt = mo.bert_ffn(sd, p + ".lang_inter", p + ".lang_output", txt)
    return txt


def _trunk_sd(sd):
    return {k[5:]: v for k, v in sd.items() if k.startswith("bert.")}


def sap(sd, batch, labels, n_l_layers=9, n_pano_layers=2, n_x_layers=4):
    """forward_sap (pretrain_cmt.py:214-292) -> (global_logits, local_logits, fused_logits, per-sample loss).
    The heads, masks and the logit fusion are the navigation forward's (model_oracle.navigation); what differs is where the masks
    come from: navigable = nav type 1 of the LAST panorama (:244-250), candidates = traj_cand_vpids[i][-1] (:259)."""
    import torch.nn.functional as F
    gmap_e, vp_e, gmap2 = forward(_trunk_sd(sd), batch, n_l_layers, n_pano_layers, n_x_layers)
    ninf = float("-inf")
    fw = torch.sigmoid(mo.cls_head(sd, "sap_fuse_linear", torch.cat([gmap_e[:, 0], vp_e[:, 0]], 1))) \
        if "sap_fuse_linear.net.0.weight" in sd else 0.5
    gmap_masks = torch.arange(gmap_e.shape[1])[None, :] < batch["gmap_lens"][:, None]
    visited = labels["gmap_visited_masks"]
    gl = (mo.cls_head(sd, "global_sap_head", gmap_e).squeeze(2) * fw).masked_fill(visited, ninf).masked_fill(~gmap_masks, ninf)
    gr = mo.cls_head(sd, "grid_sap_head", gmap2).squeeze(2).masked_fill(visited, ninf).masked_fill(~gmap_masks, ninf)
    ll = mo.cls_head(sd, "local_sap_head", vp_e).squeeze(2) * (1 - fw)
    last_types = [t[-1] for t in torch.split(batch["traj_nav_types"], list(batch["traj_step_lens"]), 0)]
    not_nav = torch.stack(last_types, 0)[:, :ll.shape[1] - 1] != 1
    ll = ll.masked_fill(torch.cat([torch.zeros(len(last_types), 1, dtype=torch.bool), not_nav], 1), ninf)
    cands = [[None] + list(c[-1]) for c in batch["traj_cand_vpids"]]
    fused = mo.fuse_logits(gl, ll, batch["gmap_vpids"], visited, cands)
    ga, la = labels["global_act_labels"], labels["local_act_labels"]
    losses = [F.cross_entropy(x, y, reduction="none") for x, y in ((gl, ga), (ll, la), (fused, ga), (gr, ga))]
    n_go = int((ga != 0).sum())
    stop_rate = (int((ga == 0).sum()) / n_go) if n_go else 1.0                         # :279-287
    for x, y in zip(losses, (ga, la, ga, ga)):
        x[y == 0] = x[y == 0] / stop_rate if (y == 0).any() else x[y == 0]
    return gl, ll, fused, sum(losses)


def mlm_scores(sd, batch, txt_labels, n_l_layers=9, n_pano_layers=2, n_x_layers=4):
    """forward_mlm of the wrapper (pretrain_cmt.py:128-153): vocabulary scores at the masked positions only
    (BertOnlyMLMHead, pretrain_src/model/vilmodel.py:262-303; decoder weight tied to the word embeddings)."""
    txt = forward_mlm(_trunk_sd(sd), batch, n_l_layers, n_pano_layers, n_x_layers)
    h = txt[txt_labels != -1]
    p = "mlm_head.predictions"
    h = mo._lin(sd, p + ".transform.dense", h)
    h = h * 0.5 * (1.0 + torch.erf(h / 2.0 ** 0.5))
    h = mo._ln(sd, p + ".transform.LayerNorm", h, 1e-12)
    return torch.nn.functional.linear(h, sd[p + ".decoder.weight"]) + sd[p + ".bias"]

"""CPU restatement (torch, fp32) of GridMM's per-step navigation forward -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (gridmm_b200/) never imports it.

Parity status: PINNED IN THE AUTHORING CONTAINER against the reference's own
`GlocalTextPathNavCMT.forward('navigation', ...)` imported in-process
(oracle/_refshim.py); outputs of the reference on seeded inputs are committed under
tests/golden/ (oracle/make_golden.py generates them) and
tests/test_oracle_golden.py re-checks this file against them on any machine.

Functional over a plain `state_dict` (name -> tensor) with the reference's key names.
Follows:
  forward_navigation_per_step   map_nav_src/models/vilmodel.py:782-918
  TransformerEncoderLayer.forward_pre / TransformerEncoder   map_nav_src/models/transformer.py:60-87,170-182
  GraphLXRTXLayer.forward       map_nav_src/models/vilmodel.py:399-414
  BertOutAttention / BertSelfAttention / BertSelfOutput / BertIntermediate / BertOutput  vilmodel.py:95-209,317-379
  extend_neg_masks              map_nav_src/models/ops.py:25-34
  ClsPrediction                 vilmodel.py:663-674
  forward_panorama_per_step     vilmodel.py:736-780
  forward_text                  vilmodel.py:730-734
"""
import math

import torch
import torch.nn.functional as F

NH = 12


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _ln(sd, p, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _heads(x):
    B, S, D = x.shape
    return x.view(B, S, NH, D // NH).permute(0, 2, 1, 3)


def _attend(q, k, v, add_mask):
    """softmax(q k^T / sqrt(dh) + mask) v  (vilmodel.py:128-153, 346-367)."""
    s = torch.matmul(_heads(q), _heads(k).transpose(-1, -2)) / math.sqrt(q.shape[-1] // NH)
    if add_mask is not None:
        s = s + add_mask
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, _heads(v))
    B, H, S, Dh = o.shape
    return o.permute(0, 2, 1, 3).reshape(B, S, H * Dh)


def neg_mask(masks):
    """ops.py:25-34: (1 - mask) * -10000, shape [B,1,1,L]."""
    return (1.0 - masks[:, None, None, :].float()) * -10000.0


def prenorm_encoder(sd, p, x, key_valid, n_layers):
    """create_transformer_encoder(..., norm=True) (ops.py:11-23): pre-norm layers of
    nn.MultiheadAttention + FFN(F.gelu), LayerNorm eps 1e-5 inside the layer
    (transformer.py:144-145) and a final BertLayerNorm eps 1e-12.
    key_valid: bool[B,S]; padded keys get -inf (key_padding_mask)."""
    add = torch.zeros(key_valid.shape, dtype=x.dtype).masked_fill(~key_valid, float("-inf"))[:, None, None, :]
    for i in range(n_layers):
        q = "%s.layers.%d" % (p, i)
        h = _ln(sd, q + ".norm1", x, 1e-5)
        qkv = F.linear(h, sd[q + ".self_attn.in_proj_weight"], sd[q + ".self_attn.in_proj_bias"])
        qq, kk, vv = qkv.chunk(3, dim=-1)
        a = _attend(qq, kk, vv, add)
        x = x + _lin(sd, q + ".self_attn.out_proj", a)
        h = _ln(sd, q + ".norm2", x, 1e-5)
        x = x + _lin(sd, q + ".linear2", F.gelu(_lin(sd, q + ".linear1", h)))
    return _ln(sd, p + ".norm", x, 1e-12)


def bert_self_block(sd, p, x, add_mask):
    """BertAttention (vilmodel.py:172-182): self-attention + dense + residual LN(1e-12)."""
    a = _attend(_lin(sd, p + ".self.query", x), _lin(sd, p + ".self.key", x), _lin(sd, p + ".self.value", x), add_mask)
    return _ln(sd, p + ".output.LayerNorm", _lin(sd, p + ".output.dense", a) + x, 1e-12)


def bert_ffn(sd, p_inter, p_out, x):
    """BertIntermediate + BertOutput (vilmodel.py:184-209), erf-GELU."""
    h = _lin(sd, p_inter + ".dense", x)
    h = h * 0.5 * (1.0 + torch.erf(h / math.sqrt(2.0)))
    return _ln(sd, p_out + ".LayerNorm", _lin(sd, p_out + ".dense", h) + x, 1e-12)


def lxrt_layer(sd, p, ctx, ctx_add, x, x_add):
    """GraphLXRTXLayer.forward (vilmodel.py:399-414), graph_sprels=None on this path."""
    q = p + ".visual_attention"
    a = _attend(_lin(sd, q + ".att.query", x), _lin(sd, q + ".att.key", ctx), _lin(sd, q + ".att.value", ctx), ctx_add)
    x = _ln(sd, q + ".output.LayerNorm", _lin(sd, q + ".output.dense", a) + x, 1e-12)
    x = bert_self_block(sd, p + ".visn_self_att", x, x_add)
    return bert_ffn(sd, p + ".visn_inter", p + ".visn_output", x)


def crossmodal_encoder(sd, p, n_layers, ctx, ctx_masks, x, x_masks):
    """CrossmodalEncoder.forward (vilmodel.py:459-468)."""
    ca, xa = neg_mask(ctx_masks), neg_mask(x_masks)
    for i in range(n_layers):
        x = lxrt_layer(sd, "%s.x_layers.%d" % (p, i), ctx, ca, x, xa)
    return x


def cls_head(sd, p, x):
    """ClsPrediction (vilmodel.py:663-674): Linear, ReLU, LN(1e-12), Linear->1."""
    h = torch.relu(_lin(sd, p + ".net.0", x))
    return _lin(sd, p + ".net.3", _ln(sd, p + ".net.2", h, 1e-12))


def grid_pool(sd, txt_embeds, grid_fts, grid_map, n_cells=196, fts_dtype=torch.float32):
    """vilmodel.py:793-807.  grid_fts: list of f16[N,D]; grid_map: list of [N] cell ids (-1 = masked).
    Returns (grid_map_input f32[B,n_cells,D], nonempty i64[B,n_cells]).
    The relevance max runs over ALL L text positions, padding included (:798)."""
    B = len(grid_fts)
    D = sd["grid_proj.weight"].shape[0]
    out = torch.zeros(B, n_cells, D)
    nonempty = torch.zeros(B, n_cells, dtype=torch.long)
    text_fts = _lin(sd, "text_proj", txt_embeds).permute(0, 2, 1)
    for b in range(B):
        x = grid_fts[b].to(fts_dtype)
        w, _ = (x @ text_fts[b].to(fts_dtype)).max(dim=-1)
        p = F.linear(x, sd["grid_proj.weight"].to(fts_dtype), sd["grid_proj.bias"].to(fts_dtype))
        gm = grid_map[b]
        for i in range(n_cells):
            sel = gm == i
            if int(sel.sum()) == 0:
                continue
            nonempty[b, i] = 1
            out[b, i] = (p[sel] * torch.softmax(w[sel], dim=-1).unsqueeze(-1)).sum(-2).float()
    return out, nonempty


def compact_cells(grid_map_input, nonempty):
    """vilmodel.py:813-823 including the aliasing quirk: `grid_mask`
    is a view of `grid_masks[b]`, so after `grid_masks[b,:k]=1` the second `.sum()` is
    re-evaluated on the modified row: k' = k + count(S in [k,196)) and the row ends up as
    [0,k) U (S n [k,k')) where S is the original non-empty set."""
    B, n_cells, D = grid_map_input.shape
    masks = nonempty.clone()
    max_cell = int(nonempty.sum(1).max()) if B > 0 else 0
    embeds = torch.zeros(B, max_cell, D)
    for b in range(B):
        row = masks[b]
        k = int(row.sum())
        embeds[b, :k] = grid_map_input[b][row == 1]
        row[:k] = 1
        k2 = int(row.sum())
        row[k2:] = 0
    return embeds, masks[:, :max_cell].bool(), max_cell


def fuse_logits(global_logits, local_logits, gmap_vpids, gmap_visited_masks, vp_cand_vpids):
    """vilmodel.py:881-899."""
    fused = global_logits.clone()
    fused[:, 0] += local_logits[:, 0]
    for i in range(fused.shape[0]):
        visited = set(vp for vp, m in zip(gmap_vpids[i], gmap_visited_masks[i]) if m)
        tmp, bw = {}, 0
        for j, cand in enumerate(vp_cand_vpids[i]):
            if j > 0:
                if cand in visited:
                    bw = bw + local_logits[i, j]
                else:
                    tmp[cand] = local_logits[i, j]
        for j, vp in enumerate(gmap_vpids[i]):
            if j > 0 and vp not in visited:
                fused[i, j] += tmp[vp] if vp in tmp else bw
    return fused


def _nav_trunk(sd, batch, n_x_layers, n_cells, fts_dtype=torch.float32, stop_after_map=False):
    """Everything of forward_navigation_per_step up to the fused [gmap; vp] embeddings (vilmodel.py:788-856; identical
    in VLN_CE/vlnce_baselines/models/gridmap/vilmodel.py:714-789)."""
    txt, txt_masks = batch["txt_embeds"], batch["txt_masks"]
    gmap_masks = batch["gmap_masks"]
    gmi, nonempty = grid_pool(sd, txt, batch["grid_fts"], batch["grid_map"], n_cells, fts_dtype=fts_dtype)
    pos = _ln(sd, "grid_pos_embeddings.1", _lin(sd, "grid_pos_embeddings.0", batch["gridmap_pos_fts"]), 1e-12)
    gmi = gmi + pos                                                                     # :816
    cells, cell_masks, C = compact_cells(gmi, nonempty)                                 # :813-823
    gmap = batch["gmap_img_embeds"] + sd["global_encoder.gmap_step_embeddings.weight"][batch["gmap_step_ids"]] \
        + _ln(sd, "global_encoder.gmap_pos_embeddings.1",
              _lin(sd, "global_encoder.gmap_pos_embeddings.0", batch["gmap_pos_fts"]), 1e-12)  # :828-830
    vp = batch["vp_img_embeds"] + _ln(sd, "local_encoder.vp_pos_embeddings.1",
                                      _lin(sd, "local_encoder.vp_pos_embeddings.0", batch["vp_pos_fts"]), 1e-12)  # :833
    m = torch.cat([cells, gmap], 1)
    mm = torch.cat([cell_masks, gmap_masks], 1)
    m = prenorm_encoder(sd, "grid_encoder", m, mm, 1)                                   # :840
    m = crossmodal_encoder(sd, "grid_txt_encoder", 1, txt, txt_masks, m, mm)            # :841
    gmap2 = m[:, C:]
    if stop_after_map:          # the pretraining MLM path leaves the trunk here (pretrain_src/model/vilmodel.py:836-838)
        return None, vp, gmap2, {"map_embeds": m}
    kv = torch.cat([m, txt], 1)
    kvm = torch.cat([mm, txt_masks], 1)
    q = torch.cat([gmap2, vp], 1)
    qm = torch.cat([gmap_masks, batch["vp_masks"]], 1)
    q = crossmodal_encoder(sd, "local_encoder.encoder", n_x_layers, kv, kvm, q, qm)     # :853
    G = gmap_masks.shape[1]
    inter = {"grid_map_input": gmi, "nonempty": nonempty, "grid_map_embeds": cells, "grid_masks": cell_masks, "map_embeds": m}
    return q[:, :G], q[:, G:], gmap2, inter


def navigation(sd, batch, n_x_layers=4, n_cells=196, return_intermediates=False):
    """forward_navigation_per_step (vilmodel.py:782-918), eval mode (dropout = identity)."""
    gmap_masks = batch["gmap_masks"]
    gm_e, vp_e, gmap2, inter = _nav_trunk(sd, batch, n_x_layers, n_cells)
    if "sap_fuse_linear.net.0.weight" in sd:
        fw = torch.sigmoid(cls_head(sd, "sap_fuse_linear", torch.cat([gm_e[:, 0], vp_e[:, 0]], 1)))
    else:
        fw = 0.5
    ninf = float("-inf")
    gl = cls_head(sd, "global_sap_head", gm_e).squeeze(2) * fw
    gl = gl.masked_fill(batch["gmap_visited_masks"], ninf).masked_fill(~gmap_masks, ninf)
    gr = cls_head(sd, "grid_sap_head", gmap2).squeeze(2)
    gr = gr.masked_fill(batch["gmap_visited_masks"], ninf).masked_fill(~gmap_masks, ninf)
    ll = cls_head(sd, "local_sap_head", vp_e).squeeze(2) * (1 - fw)
    ll = ll.masked_fill(~batch["vp_nav_masks"], ninf)
    fused = fuse_logits(gl, ll, batch["gmap_vpids"], batch["gmap_visited_masks"], batch["vp_cand_vpids"])
    if batch.get("vp_obj_masks") is not None:
        ol = cls_head(sd, "og_head", vp_e).squeeze(2).masked_fill(~batch["vp_obj_masks"], ninf)   # :903-905
    else:
        ol = None
    outs = {"gmap_embeds": gm_e, "vp_embeds": vp_e, "global_logits": gl, "local_logits": ll,
            "fused_logits": fused, "obj_logits": ol, "grid_logits": gr}
    if return_intermediates:
        outs.update(inter)
    return outs


def navigation_ce(sd, batch, n_x_layers=4, n_cells=196):
    """Continuous-env head: VLN_CE/vlnce_baselines/models/gridmap/vilmodel.py:710-800.  Same trunk; the action logits are
    `global * w + local * (1 - w)` on the first max(candidate_lengths) slots, masked by vp_nav_masks (:791-800)."""
    gm_e, vp_e, _, _ = _nav_trunk(sd, batch, n_x_layers, n_cells)
    fw = torch.sigmoid(cls_head(sd, "sap_fuse_linear", torch.cat([gm_e[:, 0], vp_e[:, 0]], 1)))
    maxc = int(max(batch["candidate_lengths"]))
    ninf = float("-inf")
    nav = ~batch["vp_nav_masks"][:, :maxc]
    gl = (cls_head(sd, "global_sap_head", gm_e).squeeze(2) * fw)[:, :maxc].masked_fill(nav, ninf)
    ll = (cls_head(sd, "local_sap_head", vp_e).squeeze(2) * (1 - fw))[:, :maxc].masked_fill(nav, ninf)
    return gl + ll


def panorama(sd, batch, n_layers=2):
    """forward_panorama_per_step (vilmodel.py:736-780); obj_linear is None when
    obj_feat_size == image_feat_size (vilmodel.py:479-483)."""
    view = _ln(sd, "img_embeddings.img_layer_norm", _lin(sd, "img_embeddings.img_linear", batch["view_img_fts"]), 1e-12)
    view_lens = batch["view_lens"]
    if batch.get("obj_img_fts") is not None:
        obj = _ln(sd, "img_embeddings.img_layer_norm", _lin(sd, "img_embeddings.img_linear", batch["obj_img_fts"]), 1e-12)
        obj_lens = batch["obj_lens"]
        rows = [torch.cat([view[b, :int(view_lens[b])], obj[b, :int(obj_lens[b])]], 0) for b in range(view.shape[0])]
        n = max(r.shape[0] for r in rows)
        img = torch.stack([F.pad(r, (0, 0, 0, n - r.shape[0])) for r in rows], 0)
        lens = view_lens + obj_lens
    else:
        img, lens = view, view_lens
    x = img + _ln(sd, "img_embeddings.loc_layer_norm", _lin(sd, "img_embeddings.loc_linear", batch["loc_fts"]), 1e-12) \
        + sd["img_embeddings.nav_type_embedding.weight"][batch["nav_types"]] \
        + sd["embeddings.token_type_embeddings.weight"][1]
    x = _ln(sd, "img_embeddings.layer_norm", x, 1e-12)
    masks = torch.arange(x.shape[1])[None, :] < lens[:, None]
    x = prenorm_encoder(sd, "img_embeddings.pano_encoder", x, masks, n_layers)
    return x, masks


def language(sd, batch, n_layers=9):
    """forward_text (vilmodel.py:730-734): BertEmbeddings + n_layers BertLayer."""
    ids, masks = batch["txt_ids"], batch["txt_masks"]
    L = ids.shape[1]
    x = sd["embeddings.word_embeddings.weight"][ids] + sd["embeddings.position_embeddings.weight"][:L][None] \
        + sd["embeddings.token_type_embeddings.weight"][0]
    x = _ln(sd, "embeddings.LayerNorm", x, 1e-12)
    add = neg_mask(masks)
    for i in range(n_layers):
        p = "lang_encoder.layer.%d" % i
        x = bert_self_block(sd, p + ".attention", x, add)
        x = bert_ffn(sd, p + ".intermediate", p + ".output", x)
    return x

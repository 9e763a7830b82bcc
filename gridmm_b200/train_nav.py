"""Trainable counterpart of the navigation model: `forward(mode, batch)` with autograd, for the reference's fine-tuning loop
(map_nav_src/r2r/agent.py: train() -> rollout(train_ml=...) -> loss.backward(), r2r/agent_base.py:203-208).

The inference model (gridmm_b200/model.py) runs the whole step in hand-written kernels and records no autograd graph; it refuses
train().  This module computes the SAME functions (map_nav_src/models/vilmodel.py:730-918; models/model.py:21-40) the way
gridmm_b200/train_model.py does for pretraining: every nn.Linear whose sizes are multiples of 128 runs forward / dgrad / wgrad on
the tcgen05 GEMM (LinearFn), the attention cores on torch's fused attention, LayerNorm / pooling / heads are torch ops, so
loss.backward() works and the dropout sites of the reference are active in train() mode.  Same `state_dict` keys as the reference
(and as the inference model: weights can be handed back and forth with load_state_dict).
"""
import collections
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .model import HID, NavConfig, _Holder, build_fuse_index, param_spec
from .train_model import PretrainModel, _WeightCache, _bool_masks


class TrainableNavCMT(PretrainModel):
    """GlocalTextPathNavCMT (vilmodel.py:676-939) under autograd.  Building blocks (linear, LayerNorm, attention, pre-norm and
    LXRT layers, ClsPrediction, cell compaction with the mask-aliasing quirk) are inherited from PretrainModel."""

    def __init__(self, config=None, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, **kw):
        nn.Module.__init__(self)
        self.p_hid, self.p_att = float(hidden_dropout_prob), float(attention_probs_dropout_prob)
        self.config = config if config is not None else NavConfig(**kw)
        if getattr(self.config, "pretrain_trunk", False):
            raise ValueError("TrainableNavCMT is the fine-tuning model (action heads); the pretraining model is PretrainModel")
        self._spec = param_spec(self.config)
        for name, (shape, kind) in self._spec.items():
            t = torch.empty(shape).normal_(0.0, 0.02) if kind == "w" else (torch.ones(shape) if kind == "g" else torch.zeros(shape))
            mod, parts = self, name.split(".")
            for p in parts[:-1]:
                if not hasattr(mod, p):
                    mod.add_module(p, _Holder())
                mod = getattr(mod, p)
            mod.register_parameter(parts[-1], nn.Parameter(t))
        self._cache = _WeightCache()
        self.weights_updated = self._cache.invalidate        # call after every optimizer step (fp16 weight copies are cached)
        self.use_native_linear = True
        self.use_fused_attention = os.environ.get("GRIDMM_TRAIN_SDPA", "1") != "0"
        self.split_precision = False

    # plain nn.Module (de)serialisation: no tied decoder here
    def load_state_dict(self, state_dict, strict=True):
        self._cache.invalidate()
        return nn.Module.load_state_dict(self, state_dict, strict=strict)

    def state_dict(self, *a, **k):
        return nn.Module.state_dict(self, *a, **k)

    # ------------------------------------------------------------------ modes
    CE_KEYS = ("txt_embeds", "txt_masks", "gmap_img_embeds", "gmap_step_ids", "gmap_pos_fts", "gmap_masks", "vp_img_embeds",
               "vp_pos_fts", "vp_masks", "vp_nav_masks", "grid_fts", "grid_map", "gridmap_pos_fts", "candidate_lengths")

    def forward(self, mode, batch):
        if mode == "navigation" and isinstance(batch, (tuple, list)):
            # continuous-env calling convention: the 14-tuple of Policy_ViewSelection_GridMap.py:622-623
            return self.forward_navigation_ce(dict(zip(self.CE_KEYS, batch)))
        batch = collections.defaultdict(lambda: None, batch)
        if mode == "language":
            return self.forward_text(batch["txt_ids"], batch["txt_masks"])
        if mode == "panorama":
            return self.forward_panorama(batch["view_img_fts"], batch["obj_img_fts"], batch["loc_fts"], batch["nav_types"],
                                         batch["view_lens"], batch["obj_lens"])
        if mode == "navigation":
            return self.forward_navigation(batch)
        raise NotImplementedError("wrong mode: %s" % mode)

    def forward_text(self, txt_ids, txt_masks):
        """vilmodel.py:730-734: BertEmbeddings + lang_encoder."""
        B, L = txt_ids.shape
        e = "embeddings"
        x = F.embedding(txt_ids, self.P(e + ".word_embeddings.weight")) + self.P(e + ".position_embeddings.weight")[:L][None] + \
            self.P(e + ".token_type_embeddings.weight")[0]
        x = self.drop(self.ln(e + ".LayerNorm", x, self.config.layer_norm_eps), self.p_hid)
        add = self.neg_mask(txt_masks.bool())
        for i in range(self.config.num_l_layers):
            p = "lang_encoder.layer.%d" % i
            x = self.bert_self(p + ".attention", x, add)
            x = self.bert_ffn(p + ".intermediate", p + ".output", x)
        return x

    def forward_panorama(self, view_img_fts, obj_img_fts, loc_fts, nav_types, view_lens, obj_lens):
        """vilmodel.py:736-780 (object tokens behind the views; obj_linear only exists when obj_feat_size != image_feat_size)."""
        ie = "img_embeddings"
        view = self.ln(ie + ".img_layer_norm", self.lin(ie + ".img_linear", view_img_fts), 1e-12)
        if obj_img_fts is not None:
            own = (ie + ".obj_linear.weight") in self._spec
            o = self.ln(ie + (".obj_layer_norm" if own else ".img_layer_norm"),
                        self.lin(ie + (".obj_linear" if own else ".img_linear"), obj_img_fts), 1e-12)
            rows = [torch.cat([view[i, :int(view_lens[i])], o[i, :int(obj_lens[i])]], 0) for i in range(view.shape[0])]
            n = max(r.shape[0] for r in rows)
            img = torch.stack([F.pad(r, (0, 0, 0, n - r.shape[0])) for r in rows], 0)
            lens = view_lens + obj_lens
        else:
            img, lens = view, view_lens
        loc = self.ln(ie + ".loc_layer_norm", F.linear(loc_fts, self.P(ie + ".loc_linear.weight"), self.P(ie + ".loc_linear.bias")), 1e-12)
        x = img + loc + F.embedding(nav_types, self.P(ie + ".nav_type_embedding.weight")) + \
            self.P("embeddings.token_type_embeddings.weight")[1]
        x = self.drop(self.ln(ie + ".layer_norm", x, 1e-12), self.p_hid)
        masks = _bool_masks(lens.to(x.device), x.shape[1])
        if self.config.num_pano_layers > 0:
            x = self.prenorm(ie + ".pano_encoder", self.config.num_pano_layers, x, masks)
        return x, masks

    def nav_grid_pool(self, txt, grid_fts, grid_map):
        """vilmodel.py:793-807 in fp32: per episode w = max_l <x, text_proj(txt)_l> over ALL positions (padding included), per
        cell a softmax over its points; grid_proj after the convex combination (it commutes with it)."""
        dev, B, nc = txt.device, txt.shape[0], self.config.grid_w * self.config.grid_w
        tp = self.lin("text_proj", txt, split=True)                                 # [B, L, 768]
        lens = [int(g.shape[0]) for g in grid_map]
        x = torch.cat([torch.as_tensor(f).to(dev) for f in grid_fts], 0).float()      # [n, 768]
        cell = torch.cat([torch.as_tensor(c).to(dev) for c in grid_map], 0).long()
        lens_t = torch.tensor(lens, device=dev)
        ep = torch.repeat_interleave(torch.arange(B, device=dev), lens_t)
        slot = torch.arange(x.shape[0], device=dev) - (torch.cumsum(lens_t, 0) - lens_t)[ep]
        xpad = torch.zeros(B, max(lens), HID, device=dev).index_put((ep, slot), x)
        w = torch.bmm(xpad, tp.transpose(1, 2)).max(-1)[0][ep, slot]
        nb = B * nc
        ids = torch.where(cell >= 0, cell + ep * nc, torch.full_like(cell, nb))
        m = torch.full((nb + 1,), float("-inf"), device=dev).scatter_reduce(0, ids, w.detach(), "amax", include_self=True)
        e = torch.exp(w - m[ids])
        z = torch.zeros(nb + 1, device=dev).index_add(0, ids, e)
        pooled = torch.zeros(nb + 1, HID, device=dev).index_add(0, ids, (e / z[ids])[:, None] * x)[:nb]
        nonempty = torch.zeros(nb + 1, dtype=torch.bool, device=dev)
        nonempty[ids] = True
        nonempty = nonempty[:nb]
        proj = self.lin("grid_proj", pooled) * nonempty[:, None]
        return proj.view(B, nc, HID), nonempty.view(B, nc)

    def compact(self, cells, nonempty):
        """vilmodel.py:813-823 with the mask-aliasing quirk, for any grid width (PretrainModel.compact assumes 196 cells)."""
        B, nc = nonempty.shape
        k = nonempty.sum(1)
        C = int(k.max()) if B else 0
        order = torch.sort((~nonempty).to(torch.int8), dim=1, stable=True)[1][:, :C]
        pos = torch.arange(nc, device=cells.device)[None, :]
        embeds = cells.gather(1, order[:, :, None].expand(B, C, HID)) * (pos[:, :C] < k[:, None])[:, :, None]
        tail = nonempty & (pos >= k[:, None])
        k2 = k + tail.sum(1)
        masks = ((pos < k[:, None]) | tail) & (pos < k2[:, None])
        return embeds, masks[:, :C], C

    def forward_navigation_ce(self, batch):
        """VLN_CE/vlnce_baselines/models/gridmap/vilmodel.py:710-800: same trunk, the action logits are global * w + local * (1 - w)
        on the first max(candidate_lengths) slots, masked by vp_nav_masks (:791-800)."""
        batch = collections.defaultdict(lambda: None, batch)
        gmap_e, vp_e, _, _ = self._nav_trunk(batch)
        fw = torch.sigmoid(self.cls_head("sap_fuse_linear", torch.cat([gmap_e[:, 0], vp_e[:, 0]], 1)))
        maxc = int(max(batch["candidate_lengths"]))
        nav = ~batch["vp_nav_masks"].bool()[:, :maxc]
        ninf = float("-inf")
        gl = (self.cls_head("global_sap_head", gmap_e).squeeze(2) * fw)[:, :maxc].masked_fill(nav, ninf)
        ll = (self.cls_head("local_sap_head", vp_e).squeeze(2) * (1 - fw))[:, :maxc].masked_fill(nav, ninf)
        return gl + ll

    def forward_navigation(self, batch):
        """vilmodel.py:782-918 (forward_navigation_per_step)."""
        cfg = self.config
        gmap_e, vp_e, gmap2, gmap_masks = self._nav_trunk(batch)
        dev = gmap_e.device
        G, V = gmap_e.shape[1], vp_e.shape[1]
        return self._nav_heads(batch, gmap_e, vp_e, gmap2, gmap_masks, G, V, dev)

    def _nav_trunk(self, batch):
        """Everything of forward_navigation_per_step up to the fused [gmap; vp] embeddings (vilmodel.py:788-856)."""
        cfg = self.config
        txt, txt_masks = batch["txt_embeds"], batch["txt_masks"].bool()
        dev = txt.device
        gmap_masks, vp_masks = batch["gmap_masks"].bool(), batch["vp_masks"].bool()
        grid_fts, grid_map, grid_pos = batch["grid_fts"], batch["grid_map"], batch["gridmap_pos_fts"]
        if batch["grid"] is not None:
            # a device-built grid (GridMapBuilder.step): the reference's per-episode views of it, all on the device
            g = batch["grid"]
            grid_fts = g.grid_fts_torch()
            n = g.n_pts.cpu().numpy()
            grid_map = [g.cell[i, :int(n[i])] for i in range(g.batch)]
            grid_pos = g.pos_fts
        cells, nonempty = self.nav_grid_pool(txt, grid_fts, grid_map)
        pos = self.ln("grid_pos_embeddings.1", F.linear(grid_pos.to(dev), self.P("grid_pos_embeddings.0.weight"),
                                                        self.P("grid_pos_embeddings.0.bias")), 1e-12)
        cell_embeds, cell_masks, C = self.compact(cells + pos, nonempty)
        ge, le = "global_encoder", "local_encoder"
        gmap = batch["gmap_img_embeds"] + F.embedding(batch["gmap_step_ids"], self.P(ge + ".gmap_step_embeddings.weight")) + \
            self.ln(ge + ".gmap_pos_embeddings.1", F.linear(batch["gmap_pos_fts"], self.P(ge + ".gmap_pos_embeddings.0.weight"),
                                                            self.P(ge + ".gmap_pos_embeddings.0.bias")), 1e-12)
        vp = batch["vp_img_embeds"] + self.ln(le + ".vp_pos_embeddings.1", F.linear(batch["vp_pos_fts"], self.P(le + ".vp_pos_embeddings.0.weight"),
                                                                                    self.P(le + ".vp_pos_embeddings.0.bias")), 1e-12)
        x = torch.cat([cell_embeds, gmap], 1)
        x_masks = torch.cat([cell_masks, gmap_masks], 1)
        x = self.prenorm("grid_encoder", 1, x, x_masks)
        x = self.lxrt("grid_txt_encoder.x_layers.0", txt, self.neg_mask(txt_masks), x, self.neg_mask(x_masks))
        gmap2 = x[:, C:]
        ctx = torch.cat([x, txt], 1)
        ctx_add = self.neg_mask(torch.cat([x_masks, txt_masks], 1))
        q = torch.cat([gmap2, vp], 1)
        q_add = self.neg_mask(torch.cat([gmap_masks, vp_masks], 1))
        for i in range(cfg.num_x_layers):
            q = self.lxrt(le + ".encoder.x_layers.%d" % i, ctx, ctx_add, q, q_add)
        G = gmap.shape[1]
        return q[:, :G], q[:, G:], gmap2, gmap_masks

    def _nav_heads(self, batch, gmap_e, vp_e, gmap2, gmap_masks, G, V, dev):
        """Heads and logit fusion (vilmodel.py:859-907)."""
        cfg = self.config
        ninf = float("-inf")
        fw = torch.sigmoid(self.cls_head("sap_fuse_linear", torch.cat([gmap_e[:, 0], vp_e[:, 0]], 1))) if cfg.glocal_fuse else 0.5
        visited = batch["gmap_visited_masks"].bool()
        gl = (self.cls_head("global_sap_head", gmap_e).squeeze(2) * fw).masked_fill(visited, ninf).masked_fill(~gmap_masks, ninf)
        gr = self.cls_head("grid_sap_head", gmap2).squeeze(2).masked_fill(visited, ninf).masked_fill(~gmap_masks, ninf)
        ll = (self.cls_head("local_sap_head", vp_e).squeeze(2) * (1 - fw)).masked_fill(~batch["vp_nav_masks"].bool(), ninf)
        src, bw = build_fuse_index(batch["gmap_vpids"], visited, batch["vp_cand_vpids"], G, V)
        src_t = torch.from_numpy(src.astype(np.int64)).to(dev)
        bw_t = torch.from_numpy(bw).to(dev).bool()
        zero = torch.zeros_like(ll)
        back = torch.where(bw_t, ll, zero).sum(1, keepdim=True)                    # local logits of the already-visited candidates
        add = torch.where(src_t >= 0, torch.where(torch.isfinite(ll), ll, zero).gather(1, src_t.clamp(min=0)), torch.zeros_like(gl)) + \
            torch.where(src_t == -2, back.expand_as(gl), torch.zeros_like(gl))
        first = torch.zeros_like(gl)
        first[:, 0] = 1.0
        fused = gl + add + first * ll[:, :1]
        ol = None
        if batch["vp_obj_masks"] is not None:
            ol = self.cls_head("og_head", vp_e).squeeze(2).masked_fill(~batch["vp_obj_masks"].bool(), ninf)
        return {"gmap_embeds": gmap_e, "vp_embeds": vp_e, "global_logits": gl, "local_logits": ll, "fused_logits": fused,
                "obj_logits": ol, "grid_logits": gr}


class VLNBertTrainable(nn.Module):
    """map_nav_src/models/model.py:12-40 for training: the object an agent holds as `self.vln_bert` when it fine-tunes.
    `drop_env` (feature dropout on the panorama inputs, args.feat_dropout) is applied in train() mode like the reference does."""

    def __init__(self, args=None, config=None):
        super().__init__()
        from .model import nav_config_from_args
        if config is None:
            config = nav_config_from_args(args)
        self.args = args
        kw = {}
        if args is not None:
            for a in ("hidden_dropout_prob", "attention_probs_dropout_prob"):
                if hasattr(args, a):
                    kw[a] = getattr(args, a)
        self.vln_bert = TrainableNavCMT(config, **kw)
        self.drop_env = nn.Dropout(p=float(getattr(args, "feat_dropout", 0.0)) if args is not None else 0.0)

    def forward(self, mode, batch):
        batch = collections.defaultdict(lambda: None, batch)
        if mode == "panorama":
            batch["view_img_fts"] = self.drop_env(batch["view_img_fts"])
            if batch["obj_img_fts"] is not None:
                batch["obj_img_fts"] = self.drop_env(batch["obj_img_fts"])
        if mode in ("language", "panorama", "navigation"):
            return self.vln_bert(mode, batch)
        raise NotImplementedError("wrong mode: %s" % mode)

"""Trainable form of the reference's pretraining model (BASELINE config 5): `GlocalTextPathCMTPreTraining`
(pretrain_src/model/pretrain_cmt.py:37-292) over the trunk `GlocalTextPathCMT` (pretrain_src/model/vilmodel.py:640-855) for the
two proxy tasks of the headline config, MLM and SAP, with the reference's parameter names (`bert.*`, `mlm_head.*`,
`{global,local,grid}_sap_head.*`, `sap_fuse_linear.*`) so that its checkpoints load and save unchanged.

What runs where (and what does not run on this package's kernels yet):
  * every nn.Linear with >= 64 input features -- > 95 % of the step's FLOPs: forward, data gradient and weight gradient are the
    tcgen05 GEMM kernel of the navigation step (`LinearFn`: fp16 operands, fp32 accumulate; dx = dy . W and dW = dy^T . x get their
    transposed fp16 operands from gridmm_cast_transpose_f16, the bias gradient from gridmm_colsum_f32);
  * the gradient all-reduce, the gradient-norm clip and AdamW: gridmm_b200.train (flat buckets over NCCL, fused kernels);
  * LayerNorm, softmax attention cores, GELU, the per-cell softmax pooling and the losses are PyTorch ops under autograd in this
    file: their backward kernels are not written.  The inference forward of the same model is the all-native
    `GlocalTextPathNavCMT.forward_pretrain`.
"""
import collections
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .model import HID, HEADS, NavConfig, _Holder, _cls, param_spec


# ----------------------------------------------------------------------------------------------- Linear on the tcgen05 GEMM
class _WeightCache:
    """fp16 copy and fp16 transpose of every weight LinearFn touches, refreshed when the parameter's version counter or storage changes
    (torch optimizers) or after invalidate() (this package's fused update)."""

    def __init__(self):
        self.c = {}
        self.generation = 0

    def invalidate(self):
        """The parameters changed behind autograd's back (gridmm_adamw_step writes the flat buffer through a raw pointer):
        GradientStep calls this after every update (`after_step`)."""
        self.generation += 1

    def get(self, w):
        key = id(w)
        e = self.c.get(key)
        stamp = (w._version, w.data_ptr(), self.generation)
        if e is None or e[0] is not w or e[1] != stamp:
            N, K = w.shape
            fresh = e is None or e[0] is not w or e[2].device != w.device
            w16 = torch.empty(N, K, dtype=torch.float16, device=w.device) if fresh else e[2]
            w16t = torch.empty(K, N, dtype=torch.float16, device=w.device) if fresh else e[3]
            ops.cast_transpose(w.detach(), dst=w16, dst_t=w16t)
            e = (w, stamp, w16, w16t)             # holding w keeps id(w) from being reused by another tensor
            self.c[key] = e
        return e[2], e[3]


class LinearFn(torch.autograd.Function):
    """y = x W^T + b with all three GEMMs (forward, dx = dy W, dW = dy^T x) on gridmm_linear_f16."""

    @staticmethod
    def forward(ctx, x, weight, bias, cache):
        N, K = weight.shape
        x2 = x.reshape(-1, K)
        if x2.dtype != torch.float32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        M = x2.shape[0]
        m_pad = (M + 63) // 64 * 64
        x16 = torch.empty(M, K, dtype=torch.float16, device=x.device)
        x16t = torch.empty(K, m_pad, dtype=torch.float16, device=x.device)      # operand of the weight gradient
        ops.cast_transpose(x2, dst=x16, dst_t=x16t)
        w16, _ = cache.get(weight)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        ops.linear(x16, w16, bias.detach() if bias is not None else None, out_f32=y)
        ctx.save_for_backward(x16t, weight)
        ctx.cache, ctx.has_bias, ctx.in_shape, ctx.M = cache, bias is not None, x.shape, M
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x16t, weight = ctx.saved_tensors
        N, K = weight.shape
        M = ctx.M
        dy2 = dy.reshape(M, N)
        if dy2.dtype != torch.float32 or not dy2.is_contiguous():
            dy2 = dy2.float().contiguous()
        m_pad = x16t.shape[1]
        dy16 = torch.empty(M, N, dtype=torch.float16, device=dy.device)
        dy16t = torch.empty(N, m_pad, dtype=torch.float16, device=dy.device)
        ops.cast_transpose(dy2, dst=dy16, dst_t=dy16t)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            _, w16t = ctx.cache.get(weight)                         # [K, N]: dx[M, K] = dy16[M, N] . (W^T)[K, N]^T
            dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
            ops.linear(dy16, w16t, None, out_f32=dx)
            dx = dx.view(ctx.in_shape)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
            ops.linear(dy16t, x16t, None, out_f32=dw)               # dW[N, K] = dy^T[N, M] . (x^T)[K, M]^T
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, dtype=torch.float32, device=dy.device)
            ops.colsum(dy2, db)
        return dx, dw, db, None


def _bool_masks(lens, n):
    return torch.arange(n, device=lens.device)[None, :] < lens[:, None]


class PretrainModel(nn.Module):
    """`GlocalTextPathCMTPreTraining(config)` with pretrain_tasks = ['mlm', 'sap'] (pretrain_cmt.py:37-66)."""

    def __init__(self, config=None, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, **kw):
        super().__init__()
        # dropout sites of the reference in training mode (pretrain_src/config/r2r_model_config.json: both 0.1)
        self.p_hid, self.p_att = float(hidden_dropout_prob), float(attention_probs_dropout_prob)
        kw.setdefault("pretrain_trunk", True)
        kw.setdefault("use_lang2visn_attn", True)
        self.config = config if config is not None else NavConfig(**kw)
        spec = collections.OrderedDict()
        for k, v in param_spec(self.config).items():
            spec["bert." + k] = v
        spec["mlm_head.predictions.bias"] = ((self.config.vocab_size,), "b")
        spec["mlm_head.predictions.transform.dense.weight"] = ((HID, HID), "w")
        spec["mlm_head.predictions.transform.dense.bias"] = ((HID,), "b")
        spec["mlm_head.predictions.transform.LayerNorm.weight"] = ((HID,), "g")
        spec["mlm_head.predictions.transform.LayerNorm.bias"] = ((HID,), "b")
        for h in ("global_sap_head", "local_sap_head", "grid_sap_head"):
            _cls(h, HID, spec)
        if self.config.glocal_fuse:
            _cls("sap_fuse_linear", 2 * HID, spec)
        self._spec = spec
        for name, (shape, kind) in spec.items():
            t = torch.empty(shape).normal_(0.0, 0.02) if kind == "w" else (torch.ones(shape) if kind == "g" else torch.zeros(shape))
            mod, parts = self, name.split(".")
            for p in parts[:-1]:
                if not hasattr(mod, p):
                    mod.add_module(p, _Holder())
                mod = getattr(mod, p)
            mod.register_parameter(parts[-1], nn.Parameter(t))
        self._cache = _WeightCache()
        self.weights_updated = self._cache.invalidate        # GradientStep(..., after_step=[model.weights_updated])
        self.use_native_linear = True

    # ---- reference checkpoints carry the tied decoder weight as its own key (pretrain_cmt.py:68-71)
    def load_state_dict(self, state_dict, strict=True):
        sd = {k: v for k, v in state_dict.items() if k != "mlm_head.predictions.decoder.weight"}
        return super().load_state_dict(sd, strict=strict)

    def state_dict(self, *a, **k):
        sd = super().state_dict(*a, **k)
        sd["mlm_head.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
        return sd

    # ------------------------------------------------------------------ building blocks
    def P(self, name):
        mod = self
        for p in name.split("."):
            mod = getattr(mod, p)
        return mod

    def lin(self, pre, x, weight=None, bias=None):
        w = self.P(pre + ".weight") if weight is None else weight
        b = (self.P(pre + ".bias") if weight is None else bias)
        N, K = w.shape
        if self.use_native_linear and x.is_cuda and K % 128 == 0 and N % 128 == 0:
            return LinearFn.apply(x, w, b, self._cache)
        return F.linear(x, w, b)

    def drop(self, x, p):
        return F.dropout(x, p, True) if (self.training and p > 0.0) else x

    def ln(self, pre, x, eps):
        return F.layer_norm(x, (HID,), self.P(pre + ".weight"), self.P(pre + ".bias"), eps)

    @staticmethod
    def _heads(x):
        B, S, _ = x.shape
        return x.view(B, S, HEADS, HID // HEADS).permute(0, 2, 1, 3)

    def attend(self, q, k, v, add_mask, p_drop=None):
        """softmax(q k^T / sqrt(64) + mask) v, dropout on the probabilities in training mode (vilmodel.py:95-153, 317-368)."""
        s = torch.matmul(self._heads(q), self._heads(k).transpose(-1, -2)) / math.sqrt(HID // HEADS)
        if add_mask is not None:
            s = s + add_mask
        o = torch.matmul(self.drop(torch.softmax(s, -1), self.p_att if p_drop is None else p_drop), self._heads(v))
        B, H, S, Dh = o.shape
        return o.permute(0, 2, 1, 3).reshape(B, S, H * Dh)

    @staticmethod
    def neg_mask(masks):
        return (1.0 - masks[:, None, None, :].float()) * -10000.0              # extend_neg_masks, ops.py:25-34

    def gelu(self, h):
        return h * 0.5 * (1.0 + torch.erf(h / math.sqrt(2.0)))

    def bert_self(self, pre, x, add):
        a = self.attend(self.lin(pre + ".self.query", x), self.lin(pre + ".self.key", x), self.lin(pre + ".self.value", x), add)
        return self.ln(pre + ".output.LayerNorm", self.drop(self.lin(pre + ".output.dense", a), self.p_hid) + x, self.config.layer_norm_eps)

    def bert_ffn(self, pi, po, x):
        h = self.gelu(self.lin(pi + ".dense", x))
        return self.ln(po + ".LayerNorm", self.drop(self.lin(po + ".dense", h), self.p_hid) + x, self.config.layer_norm_eps)

    def cross(self, pre, x, ctx, ctx_add):
        a = self.attend(self.lin(pre + ".att.query", x), self.lin(pre + ".att.key", ctx), self.lin(pre + ".att.value", ctx), ctx_add)
        return self.ln(pre + ".output.LayerNorm", self.drop(self.lin(pre + ".output.dense", a), self.p_hid) + x, self.config.layer_norm_eps)

    def lxrt(self, pre, ctx, ctx_add, x, x_add):
        """GraphLXRTXLayer.forward (vilmodel.py:387-402)."""
        x = self.cross(pre + ".visual_attention", x, ctx, ctx_add)
        x = self.bert_self(pre + ".visn_self_att", x, x_add)
        return self.bert_ffn(pre + ".visn_inter", pre + ".visn_output", x)

    def prenorm(self, pre, n_layers, x, key_valid):
        """create_transformer_encoder(config, n, norm=True): pre-norm layers + final LayerNorm (transformer.py:60-87, 170-182)."""
        add = torch.zeros(key_valid.shape, dtype=x.dtype, device=x.device).masked_fill(~key_valid, float("-inf"))[:, None, None, :]
        for i in range(n_layers):
            q = "%s.layers.%d" % (pre, i)
            h = self.ln(q + ".norm1", x, 1e-5)
            qkv = self.lin(None, h, self.P(q + ".self_attn.in_proj_weight"), self.P(q + ".self_attn.in_proj_bias"))
            qq, kk, vv = qkv.chunk(3, -1)
            x = x + self.drop(self.lin(q + ".self_attn.out_proj", self.attend(qq, kk, vv, add, self.p_hid)), self.p_hid)
            h = self.ln(q + ".norm2", x, 1e-5)
            x = x + self.drop(self.lin(q + ".linear2", self.drop(F.gelu(self.lin(q + ".linear1", h)), self.p_hid)), self.p_hid)
        return self.ln(pre + ".norm", x, 1e-12)

    def cls_head(self, pre, x):
        """ClsPrediction (vilmodel.py:628-638)."""
        h = torch.relu(self.lin(pre + ".net.0", x))
        return F.linear(self.ln(pre + ".net.2", h, 1e-12), self.P(pre + ".net.3.weight"), self.P(pre + ".net.3.bias"))

    # ------------------------------------------------------------------ trunk pieces
    def text(self, txt_ids, txt_masks):
        """BertEmbeddings + lang_encoder (vilmodel.py:62-93, 416-440)."""
        B, L = txt_ids.shape
        e = "bert.embeddings"
        x = self.P(e + ".word_embeddings.weight")[txt_ids] + self.P(e + ".position_embeddings.weight")[:L][None] + \
            self.P(e + ".token_type_embeddings.weight")[0]
        x = self.drop(self.ln(e + ".LayerNorm", x, self.config.layer_norm_eps), self.p_hid)
        add = self.neg_mask(txt_masks)
        for i in range(self.config.num_l_layers):
            p = "bert.lang_encoder.layer.%d" % i
            x = self.bert_self(p + ".attention", x, add)
            x = self.bert_ffn(p + ".intermediate", p + ".output", x)
        return x

    def panoramas(self, batch, dev):
        """ImageEmbeddings.forward over every panorama of every path (vilmodel.py:487-530); object tokens behind the views."""
        ie = "bert.img_embeddings"
        view = self.ln(ie + ".img_layer_norm", self.lin(ie + ".img_linear", batch["traj_view_img_fts"].to(dev)), 1e-12)
        vlen = batch["traj_vp_view_lens"].to(dev)
        obj = batch.get("traj_obj_img_fts")
        if obj is not None:
            own = (ie + ".obj_linear.weight") in self._spec or ("bert." + ie[5:] + ".obj_linear.weight") in self._spec
            o = self.ln(ie + (".obj_layer_norm" if own else ".img_layer_norm"),
                        self.lin(ie + (".obj_linear" if own else ".img_linear"), obj.to(dev)), 1e-12)
            olen = batch["traj_vp_obj_lens"].to(dev)
            rows = [torch.cat([view[i, :int(vlen[i])], o[i, :int(olen[i])]], 0) for i in range(view.shape[0])]
            n = max(r.shape[0] for r in rows)
            img = torch.stack([F.pad(r, (0, 0, 0, n - r.shape[0])) for r in rows], 0)
            lens = vlen + olen
        else:
            img, lens = view, vlen
        loc = self.ln(ie + ".loc_layer_norm", F.linear(batch["traj_loc_fts"].to(dev), self.P(ie + ".loc_linear.weight"),
                                                       self.P(ie + ".loc_linear.bias")), 1e-12)
        x = img + loc + self.P(ie + ".nav_type_embedding.weight")[batch["traj_nav_types"].to(dev)] + \
            self.P("bert.embeddings.token_type_embeddings.weight")[1]
        x = self.drop(self.ln(ie + ".layer_norm", x, 1e-12), self.p_hid)
        masks = _bool_masks(lens, x.shape[1])
        if self.config.num_pano_layers > 0:
            x = self.prenorm(ie + ".pano_encoder", self.config.num_pano_layers, x, masks)
        return x, lens

    @staticmethod
    def aggregate_gmap(pano, lens, step_lens, traj_vpids, traj_cand_vpids, gmap_vpids):
        """GlobalMapEncoder._aggregate_gmap_features (vilmodel.py:578-612)."""
        rows, row0 = [], 0
        for i, T in enumerate(step_lens):
            e, n = pano[row0:row0 + T], lens[row0:row0 + T]
            row0 += T
            e = e * _bool_masks(n, e.shape[1])[:, :, None]
            own, seen = {}, {}
            for t in range(T):
                own[traj_vpids[i][t]] = e[t].sum(0) / n[t]
                for j, vp in enumerate(traj_cand_vpids[i][t]):
                    if vp not in own:
                        seen.setdefault(vp, []).append(e[t, j])
            rows.append(torch.stack([own[vp] if vp in own else torch.stack(seen[vp], 0).mean(0) for vp in gmap_vpids[i][1:]], 0))
        G = 1 + max(r.shape[0] for r in rows)
        return torch.stack([F.pad(r, (0, 0, 1, G - 1 - r.shape[0])) for r in rows], 0)

    def grid_pool(self, txt, batch, dev):
        """vilmodel.py:685-700 (fp32 here; the reference pools in fp16): per episode w = max_l <x, text_proj(txt)_l>, per cell a
        softmax over its points; grid_proj applied after the convex combination (it commutes with it)."""
        B = txt.shape[0]
        tp = self.lin("bert.text_proj", txt)                                       # [B, L, 768]
        xs, ws, ids = [], [], []
        for b in range(B):
            cell = torch.as_tensor(batch["grid_map"][b]).to(dev).long()
            keep = cell >= 0
            x = torch.as_tensor(batch["grid_fts"][b]).to(dev)[keep].float()
            xs.append(x)
            ws.append((x @ tp[b].t()).max(-1)[0])
            ids.append(cell[keep] + b * 196)
        x, w, ids = torch.cat(xs, 0), torch.cat(ws, 0), torch.cat(ids, 0)
        m = torch.full((B * 196,), float("-inf"), device=dev).scatter_reduce(0, ids, w.detach(), "amax", include_self=True)
        e = torch.exp(w - m[ids])
        z = torch.zeros(B * 196, device=dev).index_add(0, ids, e)
        pooled = torch.zeros(B * 196, HID, device=dev).index_add(0, ids, (e / z[ids])[:, None] * x)
        nonempty = torch.zeros(B * 196, dtype=torch.bool, device=dev)
        nonempty[ids] = True
        proj = self.lin("bert.grid_proj", pooled) * nonempty[:, None]
        return proj.view(B, 196, HID), nonempty.view(B, 196)

    def compact(self, cells, nonempty):
        """vilmodel.py:701-711 with the mask-aliasing quirk (valid = [0,k) u (S n [k,k')), truncated to C = max k)."""
        B = cells.shape[0]
        k = nonempty.sum(1)
        C = int(k.max()) if B else 0
        embeds = torch.zeros(B, C, HID, device=cells.device)
        masks = torch.zeros(B, C, dtype=torch.bool, device=cells.device)
        for b in range(B):
            kb = int(k[b])
            embeds[b, :kb] = cells[b][nonempty[b]]
            k2 = kb + int(nonempty[b, kb:].sum())
            row = nonempty[b].clone()
            row[:kb] = True
            row[k2:] = False
            masks[b] = row[:C]
        return embeds, masks, C

    def trunk(self, batch, dev, stop_before_fusion=False):
        """GlocalTextPathCMT.forward up to the fused [gmap'; vp] embeddings (vilmodel.py:668-764)."""
        cfg = self.config
        txt_ids = batch["txt_ids"].to(dev)
        txt_masks = _bool_masks(batch["txt_lens"].to(dev), txt_ids.shape[1])
        txt = self.text(txt_ids, txt_masks)
        pano, lens = self.panoramas(batch, dev)
        step_lens = [int(x) for x in batch["traj_step_lens"]]
        gmap_img = self.aggregate_gmap(pano, lens, step_lens, batch["traj_vpids"], batch["traj_cand_vpids"], batch["gmap_vpids"])
        G = gmap_img.shape[1]
        gmap_masks = _bool_masks(batch["gmap_lens"].to(dev), G)
        ge = "bert.global_encoder"
        gmap = gmap_img + self.P(ge + ".gmap_step_embeddings.weight")[batch["gmap_step_ids"].to(dev)] + \
            self.ln(ge + ".gmap_pos_embeddings.1", F.linear(batch["gmap_pos_fts"].to(dev), self.P(ge + ".gmap_pos_embeddings.0.weight"),
                                                            self.P(ge + ".gmap_pos_embeddings.0.bias")), 1e-12)
        last = torch.tensor([sum(step_lens[:i + 1]) - 1 for i in range(len(step_lens))], device=dev)
        vp_lens = lens[last] + 1
        V = int(vp_lens.max())
        vp_img = torch.cat([torch.zeros(len(step_lens), 1, HID, device=dev), pano[last]], 1)[:, :V]
        vp_masks = _bool_masks(vp_lens, V)
        le = "bert.local_encoder"
        vp = vp_img + self.ln(le + ".vp_pos_embeddings.1", F.linear(batch["vp_pos_fts"].to(dev)[:, :V], self.P(le + ".vp_pos_embeddings.0.weight"),
                                                                    self.P(le + ".vp_pos_embeddings.0.bias")), 1e-12)
        # grid map: pooled cells + position embedding, compacted; map sequence = [cells ; gmap]
        cells, nonempty = self.grid_pool(txt, batch, dev)
        pos = self.ln("bert.grid_pos_embeddings.1", F.linear(batch["gridmap_pos_fts"].to(dev), self.P("bert.grid_pos_embeddings.0.weight"),
                                                             self.P("bert.grid_pos_embeddings.0.bias")), 1e-12)
        cell_embeds, cell_masks, C = self.compact(cells + pos, nonempty)
        x = torch.cat([cell_embeds, gmap], 1)
        x_masks = torch.cat([cell_masks, gmap_masks], 1)
        x = self.prenorm("bert.grid_encoder", 1, x, x_masks)
        x = self.lxrt("bert.grid_txt_encoder.x_layers.0", txt, self.neg_mask(txt_masks), x, self.neg_mask(x_masks))
        gmap2 = x[:, C:]
        if stop_before_fusion:
            return txt, txt_masks, gmap2, gmap_masks, vp, vp_masks
        ctx = torch.cat([x, txt], 1)
        ctx_add = self.neg_mask(torch.cat([x_masks, txt_masks], 1))
        q = torch.cat([gmap2, vp], 1)
        q_add = self.neg_mask(torch.cat([gmap_masks, vp_masks], 1))
        for i in range(cfg.num_x_layers):
            q = self.lxrt(le + ".encoder.x_layers.%d" % i, ctx, ctx_add, q, q_add)
        return q[:, :G], q[:, G:], gmap2, gmap_masks, vp_masks

    # ------------------------------------------------------------------ tasks
    def forward(self, batch, task, compute_loss=True):
        dev = next(self.parameters()).device
        if task.startswith("mlm"):
            return self.forward_mlm(batch, dev, compute_loss)
        if task.startswith("sap"):
            return self.forward_sap(batch, dev, compute_loss)
        raise ValueError("invalid task: %s (this module builds the mlm and sap proxy tasks)" % task)

    def forward_mlm(self, batch, dev, compute_loss=True):
        """pretrain_cmt.py:128-153 over GlocalTextPathCMT.forward_mlm (vilmodel.py:767-855: text queries over [gmap'; vp])."""
        cfg = self.config
        txt, txt_masks, gmap2, gmap_masks, vp, vp_masks = self.trunk(batch, dev, stop_before_fusion=True)
        ctx = torch.cat([gmap2, vp], 1)
        ctx_add = self.neg_mask(torch.cat([gmap_masks, vp_masks], 1))
        t_add = self.neg_mask(txt_masks)
        for i in range(cfg.num_x_layers):
            p = "bert.local_encoder.encoder.x_layers.%d" % i
            txt = self.cross(p + ".visual_attention", txt, ctx, ctx_add)
            txt = self.bert_self(p + ".lang_self_att", txt, t_add)
            txt = self.bert_ffn(p + ".lang_inter", p + ".lang_output", txt)
        labels = batch["txt_labels"].to(dev)
        h = txt[labels != -1]
        mp = "mlm_head.predictions"
        h = self.ln(mp + ".transform.LayerNorm", self.gelu(self.lin(mp + ".transform.dense", h)), self.config.layer_norm_eps)
        scores = self.lin(None, h, self.P("bert.embeddings.word_embeddings.weight"), None) + self.P(mp + ".bias")
        if not compute_loss:
            return scores
        return F.cross_entropy(scores, labels[labels != -1], reduction="none")

    def forward_sap(self, batch, dev, compute_loss=True):
        """pretrain_cmt.py:214-292."""
        gmap_e, vp_e, gmap2, gmap_masks, vp_masks = self.trunk(batch, dev)
        B, G = gmap_e.shape[:2]
        ninf = float("-inf")
        fw = torch.sigmoid(self.cls_head("sap_fuse_linear", torch.cat([gmap_e[:, 0], vp_e[:, 0]], 1))) if self.config.glocal_fuse else 0.5
        visited = batch["gmap_visited_masks"].to(dev)
        gl = (self.cls_head("global_sap_head", gmap_e).squeeze(2) * fw).masked_fill(visited, ninf).masked_fill(~gmap_masks, ninf)
        gr = self.cls_head("grid_sap_head", gmap2).squeeze(2).masked_fill(visited, ninf).masked_fill(~gmap_masks, ninf)
        ll = self.cls_head("local_sap_head", vp_e).squeeze(2) * (1 - fw)
        steps = [int(x) for x in batch["traj_step_lens"]]
        last = torch.tensor([sum(steps[:i + 1]) - 1 for i in range(B)])
        not_nav = batch["traj_nav_types"][last][:, :ll.shape[1] - 1].to(dev) != 1
        ll = ll.masked_fill(torch.cat([torch.zeros(B, 1, dtype=torch.bool, device=dev), not_nav], 1), ninf)
        # logit fusion (pretrain_cmt.py:256-273): candidates of the last panorama
        fused = gl.clone()
        fused[:, 0] = fused[:, 0] + ll[:, 0]
        vis_host = visited.cpu()
        for i in range(B):
            vp_i = batch["gmap_vpids"][i]
            done = set(vp for vp, m in zip(vp_i, vis_host[i]) if m)
            tmp, bw = {}, 0
            for j, cand in enumerate([None] + list(batch["traj_cand_vpids"][i][-1])):
                if j > 0:
                    if cand in done:
                        bw = bw + ll[i, j]
                    else:
                        tmp[cand] = ll[i, j]
            for j, vp in enumerate(vp_i):
                if j > 0 and vp not in done:
                    fused[i, j] = fused[i, j] + (tmp[vp] if vp in tmp else bw)
        if not compute_loss:
            return gl, ll, fused
        ga, la = batch["global_act_labels"].to(dev), batch["local_act_labels"].to(dev)
        # stop actions are re-weighted by the batch's stop / go ratio (pretrain_cmt.py:279-287)
        n_go, n_stop = int((ga != 0).sum()), int((ga == 0).sum())
        inv = 1.0 / (n_stop / n_go) if (n_go and n_stop) else 1.0
        wg = torch.where(ga == 0, torch.full_like(ga, inv, dtype=torch.float32), torch.ones_like(ga, dtype=torch.float32))
        wl = torch.where(la == 0, torch.full_like(la, inv, dtype=torch.float32), torch.ones_like(la, dtype=torch.float32))
        return F.cross_entropy(gl, ga, reduction="none") * wg + F.cross_entropy(ll, la, reduction="none") * wl + \
            F.cross_entropy(fused, ga, reduction="none") * wg + F.cross_entropy(gr, ga, reduction="none") * wg

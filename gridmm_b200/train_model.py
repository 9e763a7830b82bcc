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
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .model import HID, HEADS, NavConfig, _Holder, _cls, param_spec


# ----------------------------------------------------------------------------------------------- Linear on the tcgen05 GEMM
class _WeightCache:
    """fp16 copy and fp16 transpose of every weight LinearFn touches, refreshed when the parameter's version counter or storage changes
    (torch optimizers) or after invalidate() (this package's fused update)."""

    def __init__(self):
        self.c = {}
        self.lo = {}
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
            self.lo.pop(key, None)
        return e[2], e[3]

    def get_lo(self, w):
        """fp16 of the rounding residual w - fp16(w) (split-precision forward)."""
        w16, _ = self.get(w)
        key = id(w)
        lo = self.lo.get(key)
        if lo is None:
            lo = (w.detach() - w16.float()).half()
            self.lo[key] = lo
        return lo


class _Scratch:
    """One growing fp16 scratch buffer per (device, role) for the operands that live only inside a call (x16, dy16, dy16t).  Every
    kernel of the training path is enqueued on the device's current stream, so the next call's writes are ordered behind this
    call's reads."""
    bufs = {}

    @classmethod
    def get(cls, device, role, numel):
        key = (device, role)
        b = cls.bufs.get(key)
        if b is None or b.numel() < numel:
            b = torch.empty(max(numel, 1 << 20), dtype=torch.float16, device=device)
            cls.bufs[key] = b
        return b


_FN = {}


def _native(name, *args):
    """Direct call of a C entry (the argument checks of gridmm_b200.ops are skipped: LinearFn builds every operand itself)."""
    f = _FN.get(name)
    if f is None:
        f = _FN[name] = getattr(_lib.load(), name)
    rc = f(*args)
    if rc != 0:
        raise _lib.GridmmError("%s failed: %d" % (name, rc))


class LinearFn(torch.autograd.Function):
    """y = x W^T + b with all three GEMMs (forward, dx = dy W, dW = dy^T x) on gridmm_linear_f16 (one C call per direction:
    gridmm_linear_train_fwd / gridmm_linear_train_bwd, the casts, transposes and the bias gradient fused around the GEMMs)."""

    @staticmethod
    def forward(ctx, x, weight, bias, cache, split=False):
        N, K = weight.shape
        x2 = x.reshape(-1, K)
        if x2.dtype != torch.float32 or x2.stride(-1) != 1:
            x2 = x2.float().contiguous()
        M = x2.shape[0]
        m_pad = (M + 63) // 64 * 64
        x16t = torch.empty(K, m_pad, dtype=torch.float16, device=x.device)      # operand of the weight gradient
        w16, _ = cache.get(weight)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        b = bias.detach() if bias is not None else None
        if split:
            # split-precision forward: x = xh + xl, W = Wh + Wl in fp16 pairs, y = xh Wh^T + xl Wh^T + xh Wl^T (fp32 accumulate, the
            # dropped xl Wl^T term is 2^-22 relative): for the few layers whose rounding error is amplified downstream
            xh = x2.half()
            xl = (x2 - xh.float()).half()
            ops.cast_transpose(xh, dst_t=x16t)
            ops.linear(xh, w16, b, out_f32=y)
            ops.linear(xl, w16, None, residual=y, out_f32=y)
            ops.linear(xh, cache.get_lo(weight), None, residual=y, out_f32=y)
        else:
            _native("gridmm_linear_train_fwd", x2.data_ptr(), x2.stride(0), M, K, w16.data_ptr(), N, b.data_ptr() if b is not None else None,
                    y.data_ptr(), _Scratch.get(x.device, "x16", M * K).data_ptr(), x16t.data_ptr(), m_pad,
                    torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(x16t, weight)
        ctx.cache, ctx.has_bias, ctx.in_shape, ctx.M = cache, bias is not None, x.shape, M
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x16t, weight = ctx.saved_tensors
        N, K = weight.shape
        M = ctx.M
        dy2 = dy.reshape(M, N)
        if dy2.dtype != torch.float32 or dy2.stride(-1) != 1:
            dy2 = dy2.float().contiguous()
        m_pad = x16t.shape[1]
        dev = dy.device
        dx = torch.empty(M, K, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        dw = torch.empty(N, K, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        db = torch.empty(N, dtype=torch.float32, device=dev) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        w16t = ctx.cache.get(weight)[1] if dx is not None else None           # [K, N]: dx[M, K] = dy16[M, N] . (W^T)[K, N]^T
        _native("gridmm_linear_train_bwd", dy2.data_ptr(), dy2.stride(0), M, N, K, w16t.data_ptr() if w16t is not None else None,
                x16t.data_ptr(), m_pad, _Scratch.get(dev, "dy16", M * N).data_ptr(), _Scratch.get(dev, "dy16t", N * m_pad).data_ptr(),
                dx.data_ptr() if dx is not None else None, dw.data_ptr() if dw is not None else None,
                db.data_ptr() if db is not None else None, torch.cuda.current_stream().cuda_stream)
        return (dx.view(ctx.in_shape) if dx is not None else None), dw, db, None, None


def collate_plan(batch):
    """Index tensors (host) that turn the per-episode Python bookkeeping of the reference into a few gathers / index_adds:

      * gmap aggregation (GlobalMapEncoder._aggregate_gmap_features, vilmodel.py:578-612): node g of episode b is either the mean of
        the valid tokens of the (last) panorama taken at that viewpoint, or the mean of every candidate-view token that points at it;
        emitted as COO triplets (destination node, flat source token, weight);
      * logit fusion (pretrain_cmt.py:256-273): per unvisited node the local logit that points at it (the LAST such candidate, the
        reference fills a dict) or else the summed logits of the candidates that lead back to visited nodes;
      * the row of the last panorama of every path.

    Depends only on the batch's ids / lengths: it belongs to collation (the reference's DataLoader workers) and is cached in the batch
    under '_plan'."""
    steps = [int(x) for x in batch["traj_step_lens"]]
    B = len(steps)
    vlen = torch.as_tensor(batch["traj_vp_view_lens"]).cpu()
    olen = batch.get("traj_vp_obj_lens")
    lens = (vlen + torch.as_tensor(olen).cpu()) if olen is not None and batch.get("traj_obj_img_fts") is not None else vlen
    lens = [int(x) for x in lens]
    n_tok = int(max(lens)) if lens else 0
    G = 1 + max(len(g) - 1 for g in batch["gmap_vpids"])
    dst, src, wgt = [], [], []
    row0 = 0
    for b, T in enumerate(steps):
        own_t = {}
        for t in range(T):
            own_t[batch["traj_vpids"][b][t]] = t
        seen = {}
        for t in range(T):
            for j, vp in enumerate(batch["traj_cand_vpids"][b][t]):
                if vp not in own_t:
                    seen.setdefault(vp, []).append((row0 + t) * n_tok + j)
        for g, vp in enumerate(batch["gmap_vpids"][b]):
            if g == 0:
                continue
            if vp in own_t:
                t = own_t[vp]
                n = lens[row0 + t]
                base = (row0 + t) * n_tok
                dst.extend([b * G + g] * n); src.extend(range(base, base + n)); wgt.extend([1.0 / n] * n)
            else:
                toks = seen[vp]
                dst.extend([b * G + g] * len(toks)); src.extend(toks); wgt.extend([1.0 / len(toks)] * len(toks))
        row0 += T
    last = [sum(steps[:i + 1]) - 1 for i in range(B)]
    plan = {"G": G, "n_tok": n_tok, "agg_dst": torch.tensor(dst, dtype=torch.long), "agg_src": torch.tensor(src, dtype=torch.long),
            "agg_w": torch.tensor(wgt, dtype=torch.float32), "last": torch.tensor(last, dtype=torch.long), "steps": steps}
    n_per = [int(f.shape[0]) for f in batch["grid_fts"]] if "grid_fts" in batch else []
    if n_per:
        plan["pt_ep"] = torch.repeat_interleave(torch.arange(B), torch.tensor(n_per))
        plan["pt_slot"] = torch.cat([torch.arange(n) for n in n_per])
        plan["pt_max"] = max(n_per)
    labels = batch.get("txt_labels")
    if labels is not None:
        lab = torch.as_tensor(labels).cpu().reshape(-1)
        plan["mlm_pos"] = torch.nonzero(lab != -1).reshape(-1)
        plan["mlm_tgt"] = lab[plan["mlm_pos"]]
    vis = batch.get("gmap_visited_masks")
    if vis is not None:
        vis = torch.as_tensor(vis).cpu()
        V1 = 1 + n_tok                                   # local logits: [stop] + the tokens of the last panorama
        src_idx = torch.zeros(B, G, dtype=torch.long)
        use_src = torch.zeros(B, G, dtype=torch.bool)
        use_bw = torch.zeros(B, G, dtype=torch.bool)
        bw_mask = torch.zeros(B, V1, dtype=torch.bool)
        for b in range(B):
            vp_b = batch["gmap_vpids"][b]
            done = set(vp for vp, m in zip(vp_b, vis[b].tolist()) if m)
            tmp = {}
            for j, cand in enumerate(batch["traj_cand_vpids"][b][-1], start=1):
                if cand in done:
                    bw_mask[b, j] = True
                else:
                    tmp[cand] = j
            for g, vp in enumerate(vp_b):
                if g > 0 and vp not in done:
                    if vp in tmp:
                        src_idx[b, g], use_src[b, g] = tmp[vp], True
                    else:
                        use_bw[b, g] = True
        plan.update(fuse_src=src_idx, fuse_use_src=use_src, fuse_use_bw=use_bw, fuse_bw_mask=bw_mask)
    return plan


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
        self.use_fused_attention = os.environ.get("GRIDMM_TRAIN_SDPA", "1") != "0"      # torch SDPA for the attention cores (A/B switch)
        # optional split-precision forward (3 GEMMs, ~fp32 accuracy) of text_proj, whose output feeds the per-cell softmax.  Off by
        # default: on the parity batch it does not change the gradient error, which comes from ReLU units of the ClsPrediction heads
        # changing side under the fp16 rounding of ANY upstream operand (tools/diag_train_grad.py)
        self.split_precision = False

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

    def lin(self, pre, x, weight=None, bias=None, split=False):
        w = self.P(pre + ".weight") if weight is None else weight
        b = (self.P(pre + ".bias") if weight is None else bias)
        N, K = w.shape
        if self.use_native_linear and x.is_cuda and K % 128 == 0 and N % 128 == 0:
            return LinearFn.apply(x, w, b, self._cache, split and self.split_precision)
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
        p = self.p_att if p_drop is None else p_drop
        if self.use_fused_attention and q.is_cuda:
            # torch's fused fp32 attention (library kernel: one launch forward, one backward, instead of ~6 + ~10 eager ones in a
            # launch-bound step); same mathematics, dropout on the probabilities inside the kernel
            m = None if add_mask is None else add_mask.to(q.dtype).expand(q.shape[0], HEADS, q.shape[1], k.shape[1])
            o = F.scaled_dot_product_attention(self._heads(q), self._heads(k), self._heads(v), attn_mask=m,
                                               dropout_p=p if (self.training and p > 0.0) else 0.0)
        else:
            s = torch.matmul(self._heads(q), self._heads(k).transpose(-1, -2)) / math.sqrt(HID // HEADS)
            if add_mask is not None:
                s = s + add_mask
            o = torch.matmul(self.drop(torch.softmax(s, -1), p), self._heads(v))
        B, H, S, Dh = o.shape
        return o.permute(0, 2, 1, 3).reshape(B, S, H * Dh)

    @staticmethod
    def neg_mask(masks):
        return (1.0 - masks[:, None, None, :].float()) * -10000.0              # extend_neg_masks, ops.py:25-34

    def gelu(self, h):
        return F.gelu(h)                                                        # erf form, as the reference's `gelu` (vilmodel.py:32-40)

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
        x = F.embedding(txt_ids, self.P(e + ".word_embeddings.weight")) + self.P(e + ".position_embeddings.weight")[:L][None] + \
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
        x = img + loc + F.embedding(batch["traj_nav_types"].to(dev), self.P(ie + ".nav_type_embedding.weight")) + \
            self.P("bert.embeddings.token_type_embeddings.weight")[1]
        x = self.drop(self.ln(ie + ".layer_norm", x, 1e-12), self.p_hid)
        masks = _bool_masks(lens, x.shape[1])
        if self.config.num_pano_layers > 0:
            x = self.prenorm(ie + ".pano_encoder", self.config.num_pano_layers, x, masks)
        return x, lens

    @staticmethod
    def aggregate_gmap(pano, plan):
        """GlobalMapEncoder._aggregate_gmap_features (vilmodel.py:578-612) as one gather + one index_add (collate_plan)."""
        dev = pano.device
        B = len(plan["steps"])
        vals = pano.reshape(-1, HID).index_select(0, plan["agg_src"]) * plan["agg_w"][:, None]
        return torch.zeros(B * plan["G"], HID, device=dev).index_add(0, plan["agg_dst"], vals).view(B, plan["G"], HID)

    def grid_pool(self, txt, batch, plan, dev):
        """vilmodel.py:685-700 (fp32 here; the reference pools in fp16): per episode w = max_l <x, text_proj(txt)_l>, per cell a
        softmax over its points; grid_proj applied after the convex combination (it commutes with it).  One padded bmm for the
        relevance, segment softmax by scatter / index_add over (episode, cell) bins; masked points go to a spare bin."""
        B = txt.shape[0]
        tp = self.lin("bert.text_proj", txt, split=True)                           # [B, L, 768]
        x = torch.cat([torch.as_tensor(f).to(dev, non_blocking=True) for f in batch["grid_fts"]], 0).float()      # [n, 768]
        cell = torch.cat([torch.as_tensor(c).to(dev, non_blocking=True) for c in batch["grid_map"]], 0).long()
        ep, slot = plan["pt_ep"], plan["pt_slot"]
        xpad = torch.zeros(B, plan["pt_max"], HID, device=dev).index_put((ep, slot), x)
        w = torch.bmm(xpad, tp.transpose(1, 2)).max(-1)[0][ep, slot]               # [n]
        nb = B * 196
        ids = torch.where(cell >= 0, cell + ep * 196, torch.full_like(cell, nb))
        m = torch.full((nb + 1,), float("-inf"), device=dev).scatter_reduce(0, ids, w.detach(), "amax", include_self=True)
        e = torch.exp(w - m[ids])
        z = torch.zeros(nb + 1, device=dev).index_add(0, ids, e)
        pooled = torch.zeros(nb + 1, HID, device=dev).index_add(0, ids, (e / z[ids])[:, None] * x)[:nb]
        nonempty = torch.zeros(nb + 1, dtype=torch.bool, device=dev)
        nonempty[ids] = True
        nonempty = nonempty[:nb]
        proj = self.lin("bert.grid_proj", pooled) * nonempty[:, None]
        return proj.view(B, 196, HID), nonempty.view(B, 196)

    def compact(self, cells, nonempty):
        """vilmodel.py:701-711 with the mask-aliasing quirk (valid = [0,k) u (S n [k,k')), truncated to C = max k)."""
        B = cells.shape[0]
        k = nonempty.sum(1)
        C = int(k.max()) if B else 0                                             # the one host read (it sizes the sequence)
        order = torch.sort((~nonempty).to(torch.int8), dim=1, stable=True)[1][:, :C]        # non-empty cells first, in cell order
        pos = torch.arange(196, device=cells.device)[None, :]
        embeds = cells.gather(1, order[:, :, None].expand(B, C, HID)) * (pos[:, :C] < k[:, None])[:, :, None]
        tail = nonempty & (pos >= k[:, None])
        k2 = k + tail.sum(1)
        masks = ((pos < k[:, None]) | tail) & (pos < k2[:, None])
        return embeds, masks[:, :C], C

    def trunk(self, batch, dev, stop_before_fusion=False):
        """GlocalTextPathCMT.forward up to the fused [gmap'; vp] embeddings (vilmodel.py:668-764)."""
        cfg = self.config
        txt_ids = batch["txt_ids"].to(dev)
        txt_masks = _bool_masks(batch["txt_lens"].to(dev), txt_ids.shape[1])
        txt = self.text(txt_ids, txt_masks)
        pano, lens = self.panoramas(batch, dev)
        plan = batch.get("_plan")
        if plan is None or plan["agg_w"].device != dev:
            plan = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in (plan or collate_plan(batch)).items()}
            batch["_plan"] = plan
        assert plan["n_tok"] == pano.shape[1]
        gmap_img = self.aggregate_gmap(pano, plan)
        G = gmap_img.shape[1]
        gmap_masks = _bool_masks(batch["gmap_lens"].to(dev), G)
        ge = "bert.global_encoder"
        gmap = gmap_img + F.embedding(batch["gmap_step_ids"].to(dev), self.P(ge + ".gmap_step_embeddings.weight")) + \
            self.ln(ge + ".gmap_pos_embeddings.1", F.linear(batch["gmap_pos_fts"].to(dev), self.P(ge + ".gmap_pos_embeddings.0.weight"),
                                                            self.P(ge + ".gmap_pos_embeddings.0.bias")), 1e-12)
        last = plan["last"]
        vp_lens = lens[last] + 1
        V = int(vp_lens.max())
        vp_img = torch.cat([torch.zeros(len(plan["steps"]), 1, HID, device=dev), pano.index_select(0, last)], 1)[:, :V]
        vp_masks = _bool_masks(vp_lens, V)
        le = "bert.local_encoder"
        vp = vp_img + self.ln(le + ".vp_pos_embeddings.1", F.linear(batch["vp_pos_fts"].to(dev)[:, :V], self.P(le + ".vp_pos_embeddings.0.weight"),
                                                                    self.P(le + ".vp_pos_embeddings.0.bias")), 1e-12)
        # grid map: pooled cells + position embedding, compacted; map sequence = [cells ; gmap]
        cells, nonempty = self.grid_pool(txt, batch, plan, dev)
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
        plan = batch["_plan"]
        h = txt.reshape(-1, HID).index_select(0, plan["mlm_pos"])
        mp = "mlm_head.predictions"
        h = self.ln(mp + ".transform.LayerNorm", self.gelu(self.lin(mp + ".transform.dense", h)), self.config.layer_norm_eps)
        scores = self.lin(None, h, self.P("bert.embeddings.word_embeddings.weight"), None) + self.P(mp + ".bias")
        if not compute_loss:
            return scores
        return F.cross_entropy(scores, plan["mlm_tgt"], reduction="none")

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
        plan = batch["_plan"]
        not_nav = batch["traj_nav_types"].to(dev)[plan["last"]][:, :ll.shape[1] - 1] != 1
        ll = ll.masked_fill(torch.cat([torch.zeros(B, 1, dtype=torch.bool, device=dev), not_nav], 1), ninf)
        # logit fusion (pretrain_cmt.py:256-273) through collate_plan's index tensors
        V1 = ll.shape[1]
        llz = torch.where(torch.isfinite(ll), ll, torch.zeros_like(ll))              # masked views are never sources
        bw = (llz * plan["fuse_bw_mask"][:, :V1]).sum(1, keepdim=True)
        add = torch.where(plan["fuse_use_src"], llz.gather(1, plan["fuse_src"].clamp(max=V1 - 1)), torch.zeros_like(gl)) + \
            torch.where(plan["fuse_use_bw"], bw.expand_as(gl), torch.zeros_like(gl))
        first = torch.zeros_like(gl)
        first[:, 0] = 1.0
        fused = gl + add + first * ll[:, :1]
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

"""Tensor-level wrappers over the C ABI (include/gridmm_b200.h).  torch is used only for device memory and
streams; every function here launches hand-written sm_100a kernels and raises if that is impossible."""
import math

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
HID = 768


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.GridmmError("%s must be a CUDA tensor (gridmm_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise _lib.GridmmError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.stride(-1) != 1:
        raise _lib.GridmmError("%s must be contiguous in its last dimension" % name)


def linear(a16, w16, bias=None, residual=None, out_f32=None, out_f16=None, act=ACT_NONE, m_dev=None):
    """act(a16 @ w16.T + bias) + residual.  a16 [M,K] fp16 (row pitch = stride(0)), w16 [N,K] fp16.
    m_dev: optional device int32 tensor, the number of rows to process (packed / ragged operand)."""
    _chk(m_dev, torch.int32, "m_dev")
    _chk(a16, torch.float16, "a"); _chk(w16, torch.float16, "w")
    _chk(bias, torch.float32, "bias"); _chk(residual, torch.float32, "residual")
    _chk(out_f32, torch.float32, "out_f32"); _chk(out_f16, torch.float16, "out_f16")
    M, K = a16.shape
    N = w16.shape[0]
    assert w16.shape[1] == K
    _lib.call("gridmm_linear_f16", a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0), M, N, K,
              _lib.ptr(bias), _lib.ptr(residual), residual.stride(0) if residual is not None else 0,
              _lib.ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
              _lib.ptr(out_f16), out_f16.stride(0) if out_f16 is not None else 0, act, _lib.ptr(m_dev), _lib.stream_ptr())


def linear_ln(a16, w16, bias, residual, gamma, beta, eps, out_f32=None, out_f16=None, f32_raw=False, m_dev=None):
    """LayerNorm(a16 @ w16.T + bias + residual) in one kernel; out_f32 gets the un-normalised sum when f32_raw (pre-norm blocks)."""
    _chk(m_dev, torch.int32, "m_dev")
    _chk(a16, torch.float16, "a"); _chk(w16, torch.float16, "w"); _chk(bias, torch.float32, "bias")
    _chk(residual, torch.float32, "residual"); _chk(gamma, torch.float32, "gamma"); _chk(beta, torch.float32, "beta")
    _chk(out_f32, torch.float32, "out_f32"); _chk(out_f16, torch.float16, "out_f16")
    M, K = a16.shape
    N = w16.shape[0]
    assert w16.shape[1] == K
    _lib.call("gridmm_linear_ln_f16", a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0), M, N, K, _lib.ptr(bias),
              _lib.ptr(residual), residual.stride(0) if residual is not None else 0, gamma.data_ptr(), beta.data_ptr(), float(eps),
              _lib.ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
              _lib.ptr(out_f16), out_f16.stride(0) if out_f16 is not None else 0, int(bool(f32_raw)), _lib.ptr(m_dev),
              _lib.stream_ptr())


def attention(q, k, v, out, kmask, mask_neg, batch, heads, sq, sk, q_rows=None, k_rows=None):
    """q [batch*q_rows, >=heads*64] fp16 views (column offset folded into the view), likewise k, v; out fp16."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _chk(t, torch.float16, n)
    _chk(kmask, torch.uint8, "kmask")
    _lib.call("gridmm_attention_f16", q.data_ptr(), q.stride(0), q_rows or sq, k.data_ptr(), k.stride(0), v.data_ptr(),
              v.stride(0), k_rows or sk, out.data_ptr(), out.stride(0), kmask.data_ptr(), float(mask_neg), batch, heads, sq,
              sk, 1.0 / math.sqrt(64.0), _lib.stream_ptr())


def layernorm(x, gamma, beta, eps, out_f32=None, out_f16=None):
    _chk(x, torch.float32, "x"); _chk(out_f32, torch.float32, "out_f32"); _chk(out_f16, torch.float16, "out_f16")
    rows = x.shape[0]
    _lib.call("gridmm_layernorm", x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), float(eps),
              _lib.ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
              _lib.ptr(out_f16), out_f16.stride(0) if out_f16 is not None else 0, rows, x.shape[1], _lib.stream_ptr())


def copy_rows(x, in_rows_per_b, in_off, rows_per_b, batch, out_rows_per_b, out_off, out_f32=None, out_f16=None):
    """out[b, out_off + r] = x[b, in_off + r] for r < rows_per_b; x fp32 [batch*in_rows_per_b, 768]."""
    _chk(x, torch.float32, "x"); _chk(out_f32, torch.float32, "out_f32"); _chk(out_f16, torch.float16, "out_f16")
    _lib.call("gridmm_copy_rows", x.data_ptr(), x.stride(0), in_rows_per_b, in_off,
              _lib.ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
              _lib.ptr(out_f16), out_f16.stride(0) if out_f16 is not None else 0,
              out_rows_per_b, out_off, rows_per_b, batch, x.shape[1], _lib.stream_ptr())


def map_inputs(proj, pos_fts, cell_rank, n_nonempty, w, bias, gamma, beta, gmap_pos, gw, gbias, ggamma, gbeta, gmap_img, step_table,
               step_ids, gmap_mask, norm_gamma, norm_beta, norm_eps, map_f32, map_f16, map_mask, batch, n_cells, seq):
    _chk(proj, torch.float32, "proj"); _chk(pos_fts, torch.float32, "pos_fts"); _chk(cell_rank, torch.int32, "cell_rank")
    _chk(n_nonempty, torch.int32, "n_nonempty"); _chk(gmap_pos, torch.float32, "gmap_pos"); _chk(gmap_img, torch.float32, "gmap_img")
    _chk(step_ids, torch.int64, "step_ids"); _chk(gmap_mask, torch.uint8, "gmap_mask"); _chk(map_f32, torch.float32, "map_f32")
    _chk(map_f16, torch.float16, "map_f16"); _chk(map_mask, torch.uint8, "map_mask")
    for t_ in (gmap_pos, gmap_img, gw, map_f32, map_f16):
        assert t_.is_contiguous()
    _lib.call("gridmm_map_inputs", proj.data_ptr(), pos_fts.data_ptr(), cell_rank.data_ptr(), n_nonempty.data_ptr(), w.data_ptr(),
              bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(), gmap_pos.data_ptr(), gmap_pos.shape[1], gw.data_ptr(),
              gbias.data_ptr(), ggamma.data_ptr(), gbeta.data_ptr(), gmap_img.data_ptr(), step_table.data_ptr(), step_ids.data_ptr(),
              gmap_mask.data_ptr(), norm_gamma.data_ptr(), norm_beta.data_ptr(), float(norm_eps), map_f32.data_ptr(),
              map_f16.data_ptr(), map_mask.data_ptr(), batch, n_cells, seq, HID, _lib.stream_ptr())


def kv_index(map_mask, txt_mask, kv_pos, kv_off, kv_cnt, batch, S, L):
    _chk(map_mask, torch.uint8, "map_mask"); _chk(txt_mask, torch.uint8, "txt_mask")
    for t_ in (kv_pos, kv_off, kv_cnt):
        _chk(t_, torch.int32, "kv index")
    _lib.call("gridmm_kv_index", map_mask.data_ptr(), txt_mask.data_ptr(), batch, S, L, kv_pos.data_ptr(), kv_off.data_ptr(),
              kv_cnt.data_ptr(), _lib.stream_ptr())


def linear_rows(a16, w16, bias, out_f16, m_dev):
    """linear() over the first m_dev[0] rows (device int32 tensor)."""
    _chk(a16, torch.float16, "a"); _chk(w16, torch.float16, "w"); _chk(bias, torch.float32, "bias"); _chk(out_f16, torch.float16, "out_f16")
    _chk(m_dev, torch.int32, "m_dev")
    M, K = a16.shape
    _lib.call("gridmm_linear_f16_rows", a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0), M, w16.shape[0], K, _lib.ptr(bias),
              out_f16.data_ptr(), out_f16.stride(0), m_dev.data_ptr(), _lib.stream_ptr())


def attention_varlen(q, k, v, out, k_off, k_cnt, max_sk, batch, heads, sq, q_rows=None, k_bias=None):
    for t_, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _chk(t_, torch.float16, n)
    _chk(k_off, torch.int32, "k_off"); _chk(k_cnt, torch.int32, "k_cnt"); _chk(k_bias, torch.float32, "k_bias")
    _lib.call("gridmm_attention_varlen_f16", q.data_ptr(), q.stride(0), q_rows or sq, k.data_ptr(), k.stride(0), v.data_ptr(),
              v.stride(0), k_off.data_ptr(), k_cnt.data_ptr(), max_sk, min(k.shape[0], v.shape[0]), _lib.ptr(k_bias), out.data_ptr(),
              out.stride(0), batch, heads,
              sq, 1.0 / math.sqrt(64.0), _lib.stream_ptr())


def attention_ragged(q, k, v, out, q_off, q_cnt, max_sq, kmask, mask_neg, batch, heads, max_sk, k_off=None, k_cnt=None, k_rows=0,
                     kbias=None):
    """tcgen05 attention over packed query rows (rows q_off[b] .. + q_cnt[b] of q / out); keys packed too (k_off / k_cnt, kmask and
    kbias indexed by packed row) or regular (k_rows rows per episode, kmask [batch, max_sk])."""
    for t_, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _chk(t_, torch.float16, n)
    for t_, n in ((q_off, "q_off"), (q_cnt, "q_cnt"), (k_off, "k_off"), (k_cnt, "k_cnt")):
        _chk(t_, torch.int32, n)
    _chk(kmask, torch.uint8, "kmask"); _chk(kbias, torch.float32, "kbias")
    _lib.call("gridmm_attention_ragged_f16", q.data_ptr(), q.stride(0), q_off.data_ptr(), q_cnt.data_ptr(), max_sq, q.shape[0],
              k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), _lib.ptr(k_off), _lib.ptr(k_cnt), k_rows, max_sk, k.shape[0],
              kmask.data_ptr(), _lib.ptr(kbias), float(mask_neg), out.data_ptr(), out.stride(0), batch, heads, 1.0 / math.sqrt(64.0),
              _lib.stream_ptr())


def map_index(cell_rank, n_nonempty, batch, n_cells, G, m_off, m_info, m_logz, cell_of_rank, m_goff):
    for t_, n in ((cell_rank, "cell_rank"), (n_nonempty, "n_nonempty"), (m_off, "m_off"), (m_info, "m_info"),
                  (cell_of_rank, "cell_of_rank"), (m_goff, "m_goff")):
        _chk(t_, torch.int32, n)
    _chk(m_logz, torch.float32, "m_logz")
    _lib.call("gridmm_map_index", cell_rank.data_ptr(), n_nonempty.data_ptr(), batch, n_cells, G, m_off.data_ptr(), m_info.data_ptr(),
              m_logz.data_ptr(), cell_of_rank.data_ptr(), m_goff.data_ptr(), _lib.stream_ptr())


def map_inputs_packed(proj, pos_fts, cell_of_rank, m_off, m_info, m_logz, w, bias, gamma, beta, gmap_pos, gw, gbias, ggamma, gbeta,
                      gmap_img, step_table, step_ids, gmap_mask, norm_gamma, norm_beta, norm_eps, map_f32, map_f16, kvalid, kbias,
                      batch, n_cells, G):
    _chk(proj, torch.float32, "proj"); _chk(pos_fts, torch.float32, "pos_fts"); _chk(gmap_pos, torch.float32, "gmap_pos")
    _chk(gmap_img, torch.float32, "gmap_img"); _chk(step_ids, torch.int64, "step_ids"); _chk(gmap_mask, torch.uint8, "gmap_mask")
    _chk(map_f32, torch.float32, "map_f32"); _chk(map_f16, torch.float16, "map_f16"); _chk(kvalid, torch.uint8, "kvalid")
    _chk(kbias, torch.float32, "kbias")
    for t_ in (cell_of_rank, m_off, m_info):
        _chk(t_, torch.int32, "map index")
    for t_ in (gmap_pos, gmap_img, gw, map_f32, map_f16):
        assert t_.is_contiguous()
    _lib.call("gridmm_map_inputs_packed", proj.data_ptr(), pos_fts.data_ptr(), cell_of_rank.data_ptr(), m_off.data_ptr(),
              m_info.data_ptr(), m_logz.data_ptr(), w.data_ptr(), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
              gmap_pos.data_ptr(), gmap_pos.shape[1], gw.data_ptr(), gbias.data_ptr(), ggamma.data_ptr(), gbeta.data_ptr(),
              gmap_img.data_ptr(), step_table.data_ptr(), step_ids.data_ptr(), gmap_mask.data_ptr(), norm_gamma.data_ptr(),
              norm_beta.data_ptr(), float(norm_eps), map_f32.data_ptr(), map_f16.data_ptr(), kvalid.data_ptr(), kbias.data_ptr(),
              batch, n_cells, G, HID, _lib.stream_ptr())


def kv_index_packed(m_off, kvalid, kbias, txt_mask, batch, L, kv_src, kv_bias, kv_off, kv_cnt):
    for t_ in (m_off, kv_src, kv_off, kv_cnt):
        _chk(t_, torch.int32, "kv index")
    _chk(kvalid, torch.uint8, "kvalid"); _chk(txt_mask, torch.uint8, "txt_mask"); _chk(kbias, torch.float32, "kbias")
    _chk(kv_bias, torch.float32, "kv_bias")
    _lib.call("gridmm_kv_index_packed", m_off.data_ptr(), kvalid.data_ptr(), kbias.data_ptr(), txt_mask.data_ptr(), batch, L,
              kv_src.data_ptr(), kv_bias.data_ptr(), kv_off.data_ptr(), kv_cnt.data_ptr(), _lib.stream_ptr())


def fusion_inputs_packed(map32, txt32, kv_src, kv_off, m_goff, gmap_mask, vp_mask, x32, x16, kv16, q_mask, vp, batch, L, G, V,
                         kv_rows_max):
    """vp = (vp_pos [B*V, kin], w_t [kin, 768], bias, gamma, beta, vp_img [B*V, 768])."""
    for t_, n in ((gmap_mask, "gmap_mask"), (vp_mask, "vp_mask"), (q_mask, "q_mask")):
        _chk(t_, torch.uint8, n)
    _chk(map32, torch.float32, "map32"); _chk(txt32, torch.float32, "txt32"); _chk(x32, torch.float32, "x32")
    _chk(x16, torch.float16, "x16"); _chk(kv16, torch.float16, "kv16")
    for t_ in (kv_src, kv_off, m_goff):
        _chk(t_, torch.int32, "index")
    for t_ in (map32, txt32, x32, x16, kv16, vp[0], vp[5]):
        assert t_.is_contiguous()
    _lib.call("gridmm_fusion_inputs_packed", map32.data_ptr(), txt32.data_ptr(), kv_src.data_ptr(), kv_off.data_ptr(), m_goff.data_ptr(),
              gmap_mask.data_ptr(), vp_mask.data_ptr(), x32.data_ptr(), x16.data_ptr(), kv16.data_ptr(), q_mask.data_ptr(),
              vp[0].data_ptr(), vp[0].shape[1], vp[1].data_ptr(), vp[2].data_ptr(), vp[3].data_ptr(), vp[4].data_ptr(), vp[5].data_ptr(),
              batch, L, G, V, kv_rows_max, HID, _lib.stream_ptr())


def fusion_inputs(map32, txt32, map_mask, txt_mask, gmap_mask, vp_mask, x32, x16, kv16, kv_mask, q_mask, batch, S, L, G, V, kv_pos=None,
                  vp=None):
    """vp = (vp_pos [B*V, kin], w_t [kin, 768], bias, gamma, beta, vp_img [B*V, 768]) computes the vp tokens of x in the same launch."""
    for t_, n in ((map_mask, "map_mask"), (txt_mask, "txt_mask"), (gmap_mask, "gmap_mask"), (vp_mask, "vp_mask"), (kv_mask, "kv_mask"),
                  (q_mask, "q_mask")):
        _chk(t_, torch.uint8, n)
    _chk(map32, torch.float32, "map32"); _chk(txt32, torch.float32, "txt32"); _chk(x32, torch.float32, "x32")
    _chk(x16, torch.float16, "x16"); _chk(kv16, torch.float16, "kv16")
    for t_ in (map32, txt32, x32, x16, kv16, map_mask, txt_mask, gmap_mask, vp_mask, kv_mask, q_mask):
        assert t_.is_contiguous()
    _lib.call("gridmm_fusion_inputs", map32.data_ptr(), txt32.data_ptr(), map_mask.data_ptr(), txt_mask.data_ptr(), gmap_mask.data_ptr(),
              vp_mask.data_ptr(), x32.data_ptr(), x16.data_ptr(), kv16.data_ptr(), kv_mask.data_ptr(), q_mask.data_ptr(), _lib.ptr(kv_pos),
              _lib.ptr(vp[0]) if vp else None, vp[0].shape[1] if vp else 0, _lib.ptr(vp[1]) if vp else None,
              _lib.ptr(vp[2]) if vp else None, _lib.ptr(vp[3]) if vp else None, _lib.ptr(vp[4]) if vp else None,
              _lib.ptr(vp[5]) if vp else None, batch, S, L, G, V, HID, _lib.stream_ptr())


def split_rows(x, in_rows_per_b, in_off, rows_per_b, batch, out_f16, k_total):
    """out_f16[b*rows_per_b + r] = [hi | lo | hi](x[b, in_off + r]) at column blocks 0, k_total, 2*k_total."""
    _chk(x, torch.float32, "x"); _chk(out_f16, torch.float16, "out_f16")
    _lib.call("gridmm_split_rows", x.data_ptr(), x.stride(0), in_rows_per_b, in_off, out_f16.data_ptr(), out_f16.stride(0),
              k_total, rows_per_b, batch, x.shape[1], _lib.stream_ptr())


def pos_embed(feat, w, bias, gamma, beta, eps, out_f32, out_f16, in_rows_per_b, out_rows_per_b, out_row_off,
              base=None, table=None, idx=None):
    _chk(feat, torch.float32, "feat"); _chk(base, torch.float32, "base"); _chk(table, torch.float32, "table")
    _chk(idx, torch.int64, "idx"); _chk(out_f32, torch.float32, "out_f32"); _chk(out_f16, torch.float16, "out_f16")
    rows, kin = feat.shape
    if base is not None:
        assert base.is_contiguous()
    assert feat.is_contiguous() and w.is_contiguous()
    _lib.call("gridmm_pos_embed", feat.data_ptr(), kin, w.data_ptr(), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
              float(eps), _lib.ptr(base), _lib.ptr(table), _lib.ptr(idx), _lib.ptr(out_f32), _lib.ptr(out_f16),
              in_rows_per_b, out_rows_per_b, out_row_off, rows, HID, _lib.stream_ptr())


def text_embed(ids, word, pos, type0, gamma, beta, out_f32, out_f16, batch, L, eps=1e-12):
    _chk(ids, torch.int64, "ids"); _chk(out_f32, torch.float32, "out_f32"); _chk(out_f16, torch.float16, "out_f16")
    _lib.call("gridmm_text_embed", ids.data_ptr(), word.data_ptr(), pos.data_ptr(), type0.data_ptr(), gamma.data_ptr(),
              beta.data_ptr(), float(eps), _lib.ptr(out_f32), _lib.ptr(out_f16), batch, L, HID, _lib.stream_ptr())


def grid_assemble(proj, pos_fts, cell_rank, n_nonempty, w, bias, gamma, beta, map_f32, map_mask, batch, n_cells, seq):
    _chk(proj, torch.float32, "proj"); _chk(pos_fts, torch.float32, "pos_fts"); _chk(cell_rank, torch.int32, "cell_rank")
    _chk(n_nonempty, torch.int32, "n_nonempty"); _chk(map_f32, torch.float32, "map_f32"); _chk(map_mask, torch.uint8, "map_mask")
    _lib.call("gridmm_grid_assemble", proj.data_ptr(), pos_fts.data_ptr(), cell_rank.data_ptr(), n_nonempty.data_ptr(),
              w.data_ptr(), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(), map_f32.data_ptr(), map_mask.data_ptr(),
              batch, n_cells, seq, HID, _lib.stream_ptr())


def cls_tail(h, gamma, beta, w2, b2, logit):
    _chk(h, torch.float32, "h"); _chk(logit, torch.float32, "logit")
    assert h.is_contiguous()
    _lib.call("gridmm_cls_tail", h.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w2.data_ptr(), b2.data_ptr(),
              logit.data_ptr(), h.shape[0], h.shape[1], _lib.stream_ptr())


def nav_logits(raw_global, raw_grid, raw_local, raw_obj, raw_fuse, gmap_masks, gmap_visited, vp_nav_masks, vp_obj_masks,
               fuse_src, bw_mask, global_logits, grid_logits, local_logits, fused_logits, obj_logits, batch, G, V):
    for t, n in ((gmap_masks, "gmap_masks"), (gmap_visited, "gmap_visited"), (vp_nav_masks, "vp_nav_masks"),
                 (vp_obj_masks, "vp_obj_masks"), (bw_mask, "bw_mask")):
        _chk(t, torch.uint8, n)
    _chk(fuse_src, torch.int32, "fuse_src")
    _lib.call("gridmm_nav_logits", raw_global.data_ptr(), raw_grid.data_ptr(), raw_local.data_ptr(), _lib.ptr(raw_obj),
              _lib.ptr(raw_fuse), gmap_masks.data_ptr(), gmap_visited.data_ptr(), vp_nav_masks.data_ptr(),
              _lib.ptr(vp_obj_masks), fuse_src.data_ptr(), bw_mask.data_ptr(), global_logits.data_ptr(),
              grid_logits.data_ptr(), local_logits.data_ptr(), fused_logits.data_ptr(), _lib.ptr(obj_logits), batch, G, V,
              _lib.stream_ptr())


def head_rows(segs, batch, out_f16):
    """segs: list of (x fp32 [*, 768], in_rows_per_b, in_off, rows_per_b, out_row0[, row_off]) -> [hi | lo | hi] rows of out_f16
    [*, 2304]; row_off: optional int32 device tensor [batch], first input row of every episode (packed source)."""
    import ctypes
    n = len(segs)
    _chk(out_f16, torch.float16, "out_f16")
    for sgm in segs:
        _chk(sgm[0], torch.float32, "x")
    xs = (ctypes.c_void_p * n)(*[sgm[0].data_ptr() for sgm in segs])
    mk = lambda vals: (ctypes.c_int * n)(*[int(v) for v in vals])      # noqa: E731
    ldx, irb, ioff, rpb, or0 = (mk([sgm[0].stride(0) for sgm in segs]), mk([sgm[1] for sgm in segs]), mk([sgm[2] for sgm in segs]),
                                mk([sgm[3] for sgm in segs]), mk([sgm[4] for sgm in segs]))
    cast = lambda a: ctypes.cast(a, ctypes.c_void_p)                   # noqa: E731
    offs = [(sgm[5] if len(sgm) > 5 else None) for sgm in segs]
    for o_ in offs:
        _chk(o_, torch.int32, "row_off")
    roff = (ctypes.c_void_p * n)(*[(o_.data_ptr() if o_ is not None else None) for o_ in offs])
    _lib.call("gridmm_head_rows", n, cast(xs), cast(ldx), cast(irb), cast(ioff), cast(rpb), cast(or0),
              cast(roff) if any(o_ is not None for o_ in offs) else None, batch, out_f16.data_ptr(), out_f16.stride(0), HID,
              _lib.stream_ptr())


def cls_heads(a16, w16, groups, tiles_m, bias, gw2, grp, part, raw):
    _chk(a16, torch.float16, "a"); _chk(w16, torch.float16, "w"); _chk(bias, torch.float32, "bias"); _chk(gw2, torch.float32, "gw2")
    _chk(grp, torch.int32, "grp"); _chk(part, torch.float32, "part"); _chk(raw, torch.float32, "raw")
    _lib.call("gridmm_cls_heads_f16", a16.data_ptr(), a16.stride(0), a16.shape[0], w16.data_ptr(), w16.stride(0), groups, tiles_m,
              HID, a16.shape[1], bias.data_ptr(), gw2.data_ptr(), grp.data_ptr(), part.data_ptr(), raw.data_ptr(), _lib.stream_ptr())


def nav_logits2(part, fuse_raw, fuse_bias, fuse_gw2, row_fuse_g, row_fuse_v, consts, row_global, row_local, row_grid, row_obj,
                gmap_masks, gmap_visited, vp_nav_masks, vp_obj_masks, fuse_src, bw_mask, global_logits, grid_logits, local_logits,
                fused_logits, obj_logits, batch, G, V, cand_node=None):
    for t, n in ((gmap_masks, "gmap_masks"), (gmap_visited, "gmap_visited"), (vp_nav_masks, "vp_nav_masks"),
                 (vp_obj_masks, "vp_obj_masks"), (bw_mask, "bw_mask")):
        _chk(t, torch.uint8, n)
    _chk(fuse_src, torch.int32, "fuse_src"); _chk(part, torch.float32, "part"); _chk(consts, torch.float32, "consts")
    _chk(cand_node, torch.int32, "cand_node")
    _lib.call("gridmm_nav_logits2", part.data_ptr(), _lib.ptr(fuse_raw), _lib.ptr(fuse_bias), _lib.ptr(fuse_gw2), row_fuse_g, row_fuse_v,
              consts.data_ptr(), row_global, row_local, row_grid, row_obj,
              gmap_masks.data_ptr(), gmap_visited.data_ptr(), vp_nav_masks.data_ptr(), _lib.ptr(vp_obj_masks), fuse_src.data_ptr(),
              _lib.ptr(bw_mask), _lib.ptr(cand_node), global_logits.data_ptr(), grid_logits.data_ptr(), local_logits.data_ptr(),
              fused_logits.data_ptr(), _lib.ptr(obj_logits), batch, G, V, _lib.stream_ptr())


def ce_logits(raw_global, raw_local, raw_fuse, vp_nav_masks, fused, batch, G, V, maxc):
    _chk(vp_nav_masks, torch.uint8, "vp_nav_masks"); _chk(fused, torch.float32, "fused")
    _lib.call("gridmm_ce_logits", raw_global.data_ptr(), raw_local.data_ptr(), raw_fuse.data_ptr(), vp_nav_masks.data_ptr(),
              fused.data_ptr(), batch, G, V, maxc, _lib.stream_ptr())


def grid_update(batch, depth, depth_is_f32, depth_scale, pose, view_cs, active, off7_host, flip_y, negate_map_x, pos_mode,
                max_dist, grid_w, cap,
                wx, wy, valid, bounds, n_pts, cell, half_len, perm, cell_start, cell_rank, n_nonempty, pos_fts,
                new_slot=None, slots=None, t_cap=0):
    import ctypes
    off = (ctypes.c_float * 7)(*[float(x) for x in off7_host])
    _lib.call("gridmm_grid_update", batch, depth.data_ptr(), int(depth_is_f32), float(depth_scale), pose.data_ptr(),
              view_cs.data_ptr(), _lib.ptr(active), ctypes.cast(off, ctypes.c_void_p), int(flip_y), int(negate_map_x),
              int(pos_mode), float(max_dist), grid_w, cap, wx.data_ptr(), wy.data_ptr(), valid.data_ptr(), bounds.data_ptr(), n_pts.data_ptr(), cell.data_ptr(),
              half_len.data_ptr(), perm.data_ptr(), cell_start.data_ptr(), cell_rank.data_ptr(), n_nonempty.data_ptr(),
              pos_fts.data_ptr(), _lib.ptr(new_slot), _lib.ptr(slots), int(t_cap), _lib.stream_ptr())


_POOL_WS = {}


def pool_text_ws(device, batch, feat_dim, l_pad=128):
    """Persistent (CUDA-graph safe) lane-major text workspace of gridmm_pool: [ceil(l_pad/128)][batch, feat_dim/8, 128] 16-byte
    units (a text of 129..256 positions takes a second block, see include/gridmm_b200.h)."""
    blocks = (int(l_pad) + 127) // 128
    key = (device, batch, feat_dim, blocks)
    ws = _POOL_WS.get(key)
    if ws is None:
        ws = _POOL_WS[key] = torch.zeros(blocks * batch * 128 * feat_dim, dtype=torch.float16, device=device)
    return ws


def pool_w_scratch(device, batch, cap):
    """f32 [batch, cap] row maxima of the first pass over a text longer than 128 positions."""
    key = ("w", device, batch, cap)
    ws = _POOL_WS.get(key)
    if ws is None:
        ws = _POOL_WS[key] = torch.zeros(batch, cap, dtype=torch.float32, device=device)
    return ws


def pool_ws(device, batch, feat_dim, num_ctas=0):
    """Persistent workspace of gridmm_pool / gridmm_pool_plan: the work plan (one row range per CTA) and the partials of cells
    that several CTAs pool in pieces (include/gridmm_b200.h)."""
    key = ("ws", device, batch, feat_dim, num_ctas)
    ws = _POOL_WS.get(key)
    if ws is None:
        with torch.cuda.device(device):
            n = int(_lib.load().gridmm_pool_ws_bytes(batch, feat_dim, num_ctas))
        if n <= 0:
            raise _lib.GridmmError("gridmm_pool_ws_bytes(%d, %d, %d) failed" % (batch, feat_dim, num_ctas))
        ws = _POOL_WS[key] = torch.zeros((n + 15) // 16 * 16, dtype=torch.uint8, device=device)
    return ws


def pool_plan(cell_start, n_cells, batch, feat_dim, num_ctas=0, ws=None):
    """gridmm_pool_plan on the current stream; returns the workspace to hand to pool(..., pool_ws=ws, plan_ready=True)."""
    _chk(cell_start, torch.int32, "cell_start")
    if ws is None:
        ws = pool_ws(cell_start.device, batch, feat_dim, num_ctas)
    _lib.call("gridmm_pool_plan", cell_start.data_ptr(), n_cells, batch, feat_dim, num_ctas, ws.data_ptr(), _lib.stream_ptr())
    return ws


def linear_lanes(a16, w16, bias, out_lanes, rows_per_b):
    """text_proj written directly in gridmm_pool's lane-major layout (see include/gridmm_b200.h)."""
    _chk(a16, torch.float16, "a"); _chk(w16, torch.float16, "w"); _chk(bias, torch.float32, "bias")
    _chk(out_lanes, torch.float16, "out_lanes")
    M, K = a16.shape
    _lib.call("gridmm_linear_f16_lanes", a16.data_ptr(), a16.stride(0), w16.data_ptr(), w16.stride(0), M, w16.shape[0], K,
              _lib.ptr(bias), out_lanes.data_ptr(), rows_per_b, _lib.stream_ptr())


def pool(fts, feat_dim, slots, t_cap, slot_rows, view_rows, tok_off, perm, cap, cell_start, cell_rank, n_cells, text_fts, l_pad,
         batch, pooled, w_out=None, num_ctas=0, text_ws=None, text_ws_ready=False, pool_ws_buf=None, plan_ready=False):
    """text_fts [batch*l_pad, D] fp16, or None with text_ws_ready=True when `text_ws` was filled by linear_lanes.
    pool_ws_buf / plan_ready: the workspace pool_plan() already filled for this cell_start and num_ctas."""
    _chk(fts, torch.float16, "fts"); _chk(text_fts, torch.float16, "text_fts"); _chk(pooled, torch.float16, "pooled")
    _chk(slots, torch.int32, "slots"); _chk(perm, torch.int32, "perm")
    if text_ws is None:
        if text_ws_ready:
            raise _lib.GridmmError("text_ws_ready needs the workspace that linear_lanes filled")
        text_ws = pool_text_ws(fts.device, batch, feat_dim, l_pad)
    if text_ws.numel() < ((l_pad + 127) // 128) * batch * 128 * feat_dim:
        raise _lib.GridmmError("text_ws too small for %d text positions" % l_pad)
    w_scratch = pool_w_scratch(fts.device, batch, cap) if l_pad > 128 else None
    if pool_ws_buf is None:
        if plan_ready:
            raise _lib.GridmmError("plan_ready needs the workspace that pool_plan filled")
        pool_ws_buf = pool_ws(fts.device, batch, feat_dim, num_ctas)
    fts_rows = fts.numel() // feat_dim
    _lib.call("gridmm_pool", fts.data_ptr(), fts_rows, feat_dim, slots.data_ptr(), t_cap, slot_rows, view_rows, tok_off,
              perm.data_ptr(), cap, cell_start.data_ptr(), cell_rank.data_ptr(), n_cells, _lib.ptr(text_fts), l_pad, batch,
              text_ws.data_ptr(), int(bool(text_ws_ready)), pooled.data_ptr(), _lib.ptr(w_out), _lib.ptr(w_scratch),
              pool_ws_buf.data_ptr(), int(bool(plan_ready)), num_ctas, _lib.stream_ptr())


def gmap_update(pano_embeds, pano_masks, cur_slot, cand_slot, node_sum, node_cnt):
    """pano_embeds f32 [B, V, D], pano_masks u8/bool [B, V], cur_slot i32 [B], cand_slot i32 [B, V], node_sum f32 [B, N, D], node_cnt f32 [B, N]."""
    _chk(pano_embeds, torch.float32, "pano_embeds"); _chk(cur_slot, torch.int32, "cur_slot"); _chk(cand_slot, torch.int32, "cand_slot")
    _chk(node_sum, torch.float32, "node_sum"); _chk(node_cnt, torch.float32, "node_cnt")
    if pano_masks.dtype == torch.bool:
        pano_masks = pano_masks.view(torch.uint8)
    _chk(pano_masks, torch.uint8, "pano_masks")
    B, V, D = pano_embeds.shape
    _lib.call("gridmm_gmap_update", pano_embeds.data_ptr(), pano_masks.data_ptr(), V, D, cur_slot.data_ptr(), cand_slot.data_ptr(),
              node_sum.data_ptr(), node_cnt.data_ptr(), node_sum.shape[1], B, _lib.stream_ptr())


def gmap_gather(node_sum, node_cnt, slots, out):
    """slots i32 [B, G] -> out f32 [B, G, D] = sum / count (zero rows for slots < 0)."""
    _chk(node_sum, torch.float32, "node_sum"); _chk(node_cnt, torch.float32, "node_cnt"); _chk(slots, torch.int32, "slots")
    _chk(out, torch.float32, "out")
    B, N, D = node_sum.shape
    _lib.call("gridmm_gmap_gather", node_sum.data_ptr(), node_cnt.data_ptr(), N, D, slots.data_ptr(), slots.shape[1], B, out.data_ptr(),
              _lib.stream_ptr())


def cell_sort(batch, cell, n_pts, grid_w, cap, perm, cell_start, cell_rank, n_nonempty):
    _chk(cell, torch.int16, "cell"); _chk(n_pts, torch.int32, "n_pts")
    _lib.call("gridmm_cell_sort", batch, cell.data_ptr(), n_pts.data_ptr(), grid_w, cap, perm.data_ptr(), cell_start.data_ptr(),
              cell_rank.data_ptr(), n_nonempty.data_ptr(), _lib.stream_ptr())


def copy_segments(pairs):
    """pairs: list of (src, dst) tensors with equal byte counts (src: CUDA or pinned host, contiguous; dst: CUDA, contiguous).
    One kernel launch for all of them."""
    import ctypes
    pairs = [(s_, d_) for s_, d_ in pairs if d_.numel() > 0]
    for i0 in range(0, len(pairs), 24):
        chunk = pairs[i0:i0 + 24]
        n = len(chunk)
        for s_, d_ in chunk:
            if not d_.is_cuda or not (s_.is_cuda or s_.is_pinned()):
                raise _lib.GridmmError("copy_segments: destinations must be CUDA tensors, sources CUDA or pinned host tensors")
            if not (s_.is_contiguous() and d_.is_contiguous()) or s_.numel() * s_.element_size() != d_.numel() * d_.element_size():
                raise _lib.GridmmError("copy_segments: contiguous tensors of equal byte size expected")
        src = (ctypes.c_void_p * n)(*[s_.data_ptr() for s_, _ in chunk])
        dst = (ctypes.c_void_p * n)(*[d_.data_ptr() for _, d_ in chunk])
        nb = (ctypes.c_longlong * n)(*[d_.numel() * d_.element_size() for _, d_ in chunk])
        cast = lambda a: ctypes.cast(a, ctypes.c_void_p)                   # noqa: E731
        _lib.call("gridmm_copy_segments", n, cast(src), cast(dst), cast(nb), _lib.stream_ptr())


def grad_sumsq(g, out):
    """out[0] += sum(g ** 2) for a flat fp32 CUDA tensor (zero `out` first)."""
    _chk(g, torch.float32, "g"); _chk(out, torch.float32, "out")
    _lib.call("gridmm_grad_sumsq", g.data_ptr(), g.numel(), out.data_ptr(), _lib.stream_ptr())


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, sumsq=None, max_norm=0.0):
    """In-place AdamW step over flat fp32 CUDA tensors (pretrain_src/optim/adamw.py:57-104)."""
    for t_, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v"), (sumsq, "sumsq")):
        _chk(t_, torch.float32, n)
    assert p.numel() == g.numel() == m.numel() == v.numel() and p.is_contiguous() and g.is_contiguous()
    _lib.call("gridmm_adamw_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(lr), float(beta1), float(beta2),
              float(eps), float(weight_decay), int(step), float(grad_scale), _lib.ptr(sumsq), float(max_norm), _lib.stream_ptr())


def cast_transpose(src, dst=None, dst_t=None):
    """src [R, C] fp32 / fp16 -> dst [R, C] fp16 and / or dst_t [C, r_pad >= R] fp16 (columns past R zero-filled)."""
    if src.dtype not in (torch.float32, torch.float16):
        raise _lib.GridmmError("cast_transpose: fp32 or fp16 source expected")
    _chk(src, src.dtype, "src"); _chk(dst, torch.float16, "dst"); _chk(dst_t, torch.float16, "dst_t")
    R, C = src.shape
    _lib.call("gridmm_cast_transpose_f16", src.data_ptr(), int(src.dtype == torch.float16), src.stride(0), R, C, _lib.ptr(dst),
              dst.stride(0) if dst is not None else 0, _lib.ptr(dst_t), dst_t.stride(0) if dst_t is not None else 0,
              dst_t.shape[1] if dst_t is not None else R, _lib.stream_ptr())


def linear_train_fwd(x, w16, bias, y, x16, x16t):
    """Forward of a trainable nn.Linear: x fp32 [M, K] -> y fp32 [M, N] = x . w16^T + bias; fills x16 (scratch, >= M*K fp16 elements)
    and x16t [K, m_pad] (kept for linear_train_bwd)."""
    _chk(x, torch.float32, "x"); _chk(w16, torch.float16, "w16"); _chk(bias, torch.float32, "bias"); _chk(y, torch.float32, "y")
    _chk(x16, torch.float16, "x16"); _chk(x16t, torch.float16, "x16t")
    M, K = x.shape
    N = w16.shape[0]
    if x16.numel() < M * K or x16t.shape[0] != K or not (w16.is_contiguous() and y.is_contiguous() and x16t.is_contiguous()):
        raise _lib.GridmmError("linear_train_fwd: operand shapes")
    _lib.call("gridmm_linear_train_fwd", x.data_ptr(), x.stride(0), M, K, w16.data_ptr(), N, _lib.ptr(bias), y.data_ptr(), x16.data_ptr(),
              x16t.data_ptr(), x16t.shape[1], _lib.stream_ptr())


def linear_train_bwd(dy, w16t, x16t, dy16, dy16t, dx=None, dw=None, db=None):
    """Backward of the same layer: dy fp32 [M, N]; dx [M, K] = dy . W (needs w16t [K, N]), dw [N, K] = dy^T . x (needs x16t [K, m_pad]),
    db [N] = column sums of dy; dy16 / dy16t are scratch (>= M*N and N*m_pad fp16 elements)."""
    _chk(dy, torch.float32, "dy"); _chk(w16t, torch.float16, "w16t"); _chk(x16t, torch.float16, "x16t")
    _chk(dy16, torch.float16, "dy16"); _chk(dy16t, torch.float16, "dy16t")
    _chk(dx, torch.float32, "dx"); _chk(dw, torch.float32, "dw"); _chk(db, torch.float32, "db")
    M, N = dy.shape
    K, m_pad = x16t.shape
    if dy16.numel() < M * N or dy16t.numel() < N * m_pad or not x16t.is_contiguous() or (w16t is not None and tuple(w16t.shape) != (K, N)):
        raise _lib.GridmmError("linear_train_bwd: operand shapes")
    _lib.call("gridmm_linear_train_bwd", dy.data_ptr(), dy.stride(0), M, N, K, _lib.ptr(w16t), x16t.data_ptr(), m_pad, dy16.data_ptr(),
              dy16t.data_ptr(), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.stream_ptr())


def colsum(dy, out):
    """out[n] += sum_m dy[m, n] (fp32)."""
    _chk(dy, torch.float32, "dy"); _chk(out, torch.float32, "out")
    _lib.call("gridmm_colsum_f32", dy.data_ptr(), dy.stride(0), dy.shape[0], dy.shape[1], out.data_ptr(), _lib.stream_ptr())

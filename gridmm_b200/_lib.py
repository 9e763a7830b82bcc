"""ctypes binding of include/gridmm_b200.h -- the only door from Python into the CUDA kernels.

There is deliberately no fallback: if the shared library is missing it is built with nvcc, and if that fails
(or a kernel call returns non-zero) a GridmmError is raised.
"""
import ctypes
import os
import re

from . import build as _build

c_int, c_float, c_void_p, c_longlong = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong


class GridmmError(RuntimeError):
    pass


_ERR = {-1: "GRIDMM_ERR_SHAPE (unsupported size or alignment)", -2: "GRIDMM_ERR_DRIVER (tensor-map encode failed)",
        -3: "GRIDMM_ERR_ARG (null pointer / inconsistent arguments)"}

# name -> argument ctypes, in the order of include/gridmm_b200.h
_SIGS = {
    "gridmm_grid_update": [c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                           c_int, c_int,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "gridmm_cell_sort": [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "gridmm_pool": [c_void_p, c_longlong, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                    c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "gridmm_pool_plan": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "gridmm_gmap_update": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "gridmm_gmap_gather": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "gridmm_linear_ln_f16": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                             c_float, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "gridmm_head_rows": [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                         c_void_p],
    "gridmm_map_index": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "gridmm_map_inputs_packed": [c_void_p] * 11 + [c_int] + [c_void_p] * 10 + [c_float] + [c_void_p] * 4 + [c_int] * 4 + [c_void_p],
    "gridmm_attention_ragged_f16": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                    c_void_p, c_int, c_int, c_longlong, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int,
                                    c_float, c_void_p],
    "gridmm_kv_index_packed": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "gridmm_fusion_inputs_packed": [c_void_p] * 12 + [c_int] + [c_void_p] * 5 + [c_int] * 6 + [c_void_p],
    "gridmm_cls_heads_f16": [c_void_p, c_int, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p],
    "gridmm_nav_logits2": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 12 +
                          [c_int, c_int, c_int, c_void_p],
    "gridmm_copy_segments": [c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "gridmm_grad_sumsq": [c_void_p, c_longlong, c_void_p, c_void_p],
    "gridmm_cast_transpose_f16": [c_void_p, c_int, c_longlong, c_int, c_int, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_void_p],
    "gridmm_colsum_f32": [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p],
    "gridmm_linear_train_fwd": [c_void_p, c_longlong, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "gridmm_linear_train_bwd": [c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p],
    "gridmm_adamw_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_float, c_float, c_float, c_float, c_int, c_float,
                          c_void_p, c_float, c_void_p],
    "gridmm_linear_f16_lanes": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "gridmm_linear_f16": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                          c_void_p, c_int, c_int, c_void_p, c_void_p],
    "gridmm_attention_f16": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_float,
                             c_int, c_int, c_int, c_int, c_float, c_void_p],
    "gridmm_layernorm": [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p],
    "gridmm_copy_rows": [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                         c_void_p],
    "gridmm_map_inputs": [c_void_p] * 9 + [c_int] + [c_void_p] * 10 + [c_float] + [c_void_p] * 3 + [c_int] * 4 + [c_void_p],
    "gridmm_fusion_inputs": [c_void_p] * 13 + [c_int] + [c_void_p] * 5 + [c_int] * 6 + [c_void_p],
    "gridmm_kv_index": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "gridmm_linear_f16_rows": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "gridmm_attention_varlen_f16": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_int, c_float, c_void_p],
    "gridmm_split_rows": [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "gridmm_pos_embed": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "gridmm_text_embed": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                          c_void_p],
    "gridmm_grid_assemble": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_int, c_int, c_int, c_void_p],
    "gridmm_cls_tail": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "gridmm_ce_logits": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "gridmm_nav_logits": [c_void_p] * 16 + [c_int, c_int, c_int, c_void_p],
}

_lib = None


def header_symbols():
    """Every function name include/gridmm_b200.h declares (used by the CPU test-suite)."""
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gridmm_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"^\s*(?:int|long long|void)\s+(gridmm_\w+)\s*\(", txt, flags=re.M)))


def load():
    """Build (if needed) and dlopen the library; idempotent."""
    global _lib
    if _lib is not None:
        return _lib
    # build() is a no-op when the source digest matches the stamp next to the library (a stale binary after an edit of
    # csrc/*.cu would otherwise be loaded silently); concurrent ranks serialise on a file lock inside build()
    path = _build.build()
    lib = ctypes.CDLL(path)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    lib.gridmm_abi_version.restype = c_int
    lib.gridmm_pool_ws_bytes.argtypes = [c_int, c_int, c_int]
    lib.gridmm_pool_ws_bytes.restype = c_longlong
    lib.gridmm_launch_count.restype = c_longlong
    lib.gridmm_launch_count_reset.restype = None
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    """Raw cudaStream_t of torch's current stream on `device` (default: the current device)."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = _ERR.get(rc)
        if msg is None:
            msg = "cudaError %d" % rc
        raise GridmmError("%s failed: %s" % (name, msg))


def launch_count():
    return int(load().gridmm_launch_count())


def launch_count_reset():
    load().gridmm_launch_count_reset()

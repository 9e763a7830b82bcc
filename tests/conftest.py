import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionstart(session):
    """GRIDMM_POOL_SPLIT=0/1 runs the whole GPU suite with single / value + residual softmax weights in gridmm_pool (A/B of the
    accuracy cost of halving the pooling kernel's HMMA count)."""
    v = os.environ.get("GRIDMM_POOL_SPLIT")
    if v in ("0", "1"):
        try:
            import ctypes
            from gridmm_b200 import _lib
            lib = _lib.load()
            lib.gridmm_debug_set_pool_split.argtypes = [ctypes.c_int]
            lib.gridmm_debug_set_pool_split(int(v))
        except Exception:
            pass

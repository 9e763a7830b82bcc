"""Every native call of one training step (forward + backward, sap and mlm) checked on the spot against torch on the SAME fp16
operands: a GEMM / cast / column-sum that is wrong for some shape shows up here with its shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridmm_b200 import ops
from tests import helpers as H
from tests.test_gpu_train import _setup
worst = {}
_fwd, _bwd = ops.linear_train_fwd, ops.linear_train_bwd
def rec(kind, shape, err):
    k = (kind,) + tuple(shape)
    worst[k] = max(worst.get(k, 0.0), err)
def rel(a, b, scale=None):
    return (a - b).abs().max().item() / max((scale if scale is not None else b.abs().max().item()), 1e-30)
def fwd(x, w16, bias, y, x16, x16t):
    _fwd(x, w16, bias, y, x16, x16t)
    ref = x.half().float() @ w16.float().t()
    if bias is not None: ref = ref + bias
    rec("fwd", (x.shape[0], w16.shape[0], x.shape[1]), rel(y, ref))
    M = x.shape[0]
    rec("x16t", tuple(x.shape), (x16t[:, :M].float() - x.half().float().t()).abs().max().item() + x16t[:, M:].float().abs().sum().item())
def bwd(dy, w16t, x16t, dy16, dy16t, dx=None, dw=None, db=None):
    _bwd(dy, w16t, x16t, dy16, dy16t, dx=dx, dw=dw, db=db)
    M, N = dy.shape
    h = dy.half().float()
    if dx is not None: rec("dx", (M, N, x16t.shape[0]), rel(dx, h @ w16t.float().t()))
    if dw is not None: rec("dw", (M, N, x16t.shape[0]), rel(dw, h.t() @ x16t[:, :M].float().t()))
    if db is not None: rec("db", (M, N), rel(db, dy.sum(0), scale=dy.abs().sum(0).max().item()))
ops.linear_train_fwd, ops.linear_train_bwd = fwd, bwd
case = H.PRETRAIN_MODEL_CASE
for task in ("sap", "mlm"):
    model, w, batch = _setup(case)
    model = model.cuda().train()
    for rep in range(2):
        (model(batch, task).mean() * 1024.0).backward()
    torch.cuda.synchronize()
bad = sorted(worst.items(), key=lambda kv: -kv[1])
print("checked %d distinct (op, shape) keys; worst:" % len(worst))
for k, v in bad[:12]:
    print("   ", k, "%.3e" % v)

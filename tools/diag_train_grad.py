"""Where the SAP gradient difference comes from: the same model on the GPU with torch fp32 linears, under torch.autocast(fp16), and
with the fp16-operand tcgen05 linears (forward only / backward only / both), all against CPU autograd through the oracle."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridmm_b200 import train_model as tm
from tests import helpers as H
from tests.test_gpu_train import _setup, _oracle_grads
case = H.PRETRAIN_MODEL_CASE
Native = tm.LinearFn

class Mixed(torch.autograd.Function):
    """fp16-rounded operands in one direction only (emulated with torch matmuls on rounded copies)."""
    fwd16 = True
    bwd16 = False
    @staticmethod
    def forward(ctx, x, w, b, cache, split=False):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        r = (lambda t: t.half().float()) if (Mixed.fwd16 and not split) else (lambda t: t)
        y = r(x.reshape(-1, x.shape[-1])) @ r(w).t()
        if b is not None: y = y + b
        return y.view(*x.shape[:-1], w.shape[0])
    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        r = (lambda t: t.half().float()) if Mixed.bwd16 else (lambda t: t)
        dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
        dx = (r(dy2) @ r(w)).view(x.shape)
        dw = r(dy2).t() @ r(x2)
        return dx, dw, (dy2.sum(0) if ctx.has_bias else None), None, None

modes = [("torch fp32", None), ("native", Native), ("native", Native), ("native", Native),
         ("fp16 operands forward only", (True, False)), ("fp16 operands backward only", (False, True)), ("fp16 operands both (emulated)", (True, True))]
for task in ("sap", "mlm"):
    for name, mode in modes:
        try:
            model, w, batch = _setup(case)
            ref_loss, ref = _oracle_grads(w, batch, task, case)
            model = model.cuda().train()
            model.use_native_linear = mode not in (None, "autocast")
            if isinstance(mode, tuple):
                Mixed.fwd16, Mixed.bwd16 = mode
                tm.LinearFn = Mixed
            else:
                tm.LinearFn = Native
            with torch.autocast("cuda", dtype=torch.float16, enabled=(mode == "autocast")):
                loss = model(batch, task).mean()
            (loss.float() * 1024.0).backward()
            tot = sum(float(g.double().pow(2).sum()) for g in ref.values())
            err = 0.0
            for n, p in model.named_parameters():
                if n in ref and p.grad is not None:
                    err += float((p.grad.cpu() / 1024.0 - ref[n]).double().pow(2).sum())
            print("%s | %-32s loss %.6f (oracle %.6f)  whole-model gradient rel. error %.2e" % (task, name, float(loss), ref_loss, (err / tot) ** 0.5))
        except Exception:
            print("%s | %s FAILED: %s" % (task, name, traceback.format_exc().strip().splitlines()[-1]))

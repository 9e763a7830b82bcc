"""Where the SAP gradient difference comes from: the same model on the GPU with torch fp32 linears (use_native_linear = False)
and with the fp16-operand tcgen05 linears, both against CPU autograd through the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import helpers as H
from tests.test_gpu_train import _setup, _oracle_grads
case = H.PRETRAIN_MODEL_CASE
for task in ("sap", "mlm"):
    for native, scale in ((False, 1024.0), (True, 1024.0), (True, 65536.0), (True, 2.0 ** 20)):
        model, w, batch = _setup(case)
        ref_loss, ref = _oracle_grads(w, batch, task, case)
        model = model.cuda().train()
        model.use_native_linear = native
        loss = model(batch, task).mean()
        (loss * scale).backward()
        tot = sum(float(g.double().pow(2).sum()) for g in ref.values())
        err = 0.0
        for n, p in model.named_parameters():
            if n in ref and p.grad is not None:
                err += float((p.grad.cpu() / scale - ref[n]).double().pow(2).sum())
        print("%s native_linear=%s scale=%g: loss %.6f (oracle %.6f), whole-model gradient rel. error %.2e" % (task, native, scale, float(loss), ref_loss, (err / tot) ** 0.5))

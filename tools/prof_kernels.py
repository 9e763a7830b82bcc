"""Eager launches of the two headline kernels at bench shapes, for `ncu --set full` (one GPU, short)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Step, B, T
from gridmm_b200 import ops

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
step = Step(dev, seed=0)
step.model.use_cuda_graph = False
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for _ in range(2):
    step.run_resident()
torch.cuda.synchronize()
if which in ("gemm", "all"):
    # the FFN1-shaped GEMM of the map sequence: M = 32*216, N = 3072, K = 768 (+ GELU, fp16 out)
    M, N, K = B * 216, 3072, 768
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) * 0.02).half()
    bias = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.float16)
    for _ in range(3):
        ops.linear(a, w, bias=bias, out_f16=out, act=1)
    # QKV-shaped: N = 2304
    w2 = (torch.randn(2304, K, device=dev) * 0.02).half(); out2 = torch.empty(M, 2304, device=dev, dtype=torch.float16)
    for _ in range(3):
        ops.linear(a, w2, bias=bias[:2304], out_f16=out2)
torch.cuda.synchronize()
print("done")

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridmm_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
M, K = 6912, 768
a = torch.randn(M, K, device=dev).half()
for N, act, res in ((2304, 0, False), (768, 0, True)):
    w = (torch.randn(N, K, device=dev) * 0.02).half(); bias = torch.zeros(N, device=dev)
    o16 = None if res else torch.empty(M, N, device=dev, dtype=torch.float16)
    o32 = torch.randn(M, N, device=dev) if res else None
    for _ in range(3):
        ops.linear(a, w, bias=bias, residual=o32, out_f32=o32, out_f16=o16, act=act)
torch.cuda.synchronize(); print("done")

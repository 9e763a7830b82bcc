"""Fine-tuning step of the trainable navigation model (gridmm_b200/train_nav.py) at the bench shape (B = 32, T = 8): forward
('navigation') + imitation loss + backward + AdamW (torch), device-timed.  Not a BASELINE metric; reported in DESIGN.md next to
the inference step to show what the autograd path costs.

    python tools/bench_finetune.py [steps]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from bench import B, T
from gridmm_b200 import synth
from gridmm_b200.env import GridMapBuilder
from gridmm_b200.train_nav import TrainableNavCMT

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ep, nav_np = bench._inputs(0)
cfg, w = bench._weights()
model = TrainableNavCMT(cfg)
model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
model.to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
gb = GridMapBuilder(B, max_steps=T, device=dev)
for t in range(T):
    grid = gb.step(ep["depth_sub"][:, t], ep["clip"][:, t], ep["pos"][:, t], ep["heading"][:, t])
nav = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.to_torch(nav_np).items()}
nav["grid"] = grid
with torch.no_grad():
    fin = torch.isfinite(model("navigation", nav)["fused_logits"])
target = torch.tensor([int(torch.nonzero(fin[b])[-1]) for b in range(B)], device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    out = model("navigation", nav)
    loss = torch.nn.functional.cross_entropy(out["fused_logits"], target, reduction="sum") / B
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 40.0)
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(steps):
    loss = step()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / steps
print("fine-tuning step, B=%d T=%d: %.1f ms per step (%.0f nav-steps/s), loss %.4f" % (B, T, ms, B / ms * 1e3, float(loss)))

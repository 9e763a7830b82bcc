"""torch.profiler table of one pretraining step (sap and mlm): where the 100 ms go (host-bound or kernel-bound)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_pretrain as bp
from gridmm_b200.train import FlatParams, GradientStep
dev = torch.device("cuda:0")
model = bp.build_model(0).to(dev).train()
flat = FlatParams(model)
gs = GradientStep(flat, lr=5e-5, max_norm=5.0, after_step=[model.weights_updated])
batches = [bp.to_device(bp.make_batch(i), dev) for i in range(2)]
def step(i):
    gs.arm(); loss = model(batches[i % 2], ["mlm", "sap"][i % 2]).mean(); (loss * 1024).backward(); gs.step(loss_scale=1024.0)
for i in range(4): step(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))

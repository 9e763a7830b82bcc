import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Step
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
step = Step(dev, seed=0); step.model.use_cuda_graph = False
for _ in range(3):
    step.run_resident()
torch.cuda.synchronize(); print("done")

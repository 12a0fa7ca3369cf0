#!/usr/bin/env python
"""Kernel-time breakdown (torch.profiler, CUDA activities) of one UNetUndirected(3,8,3) diffusion training step."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile
from qiddm_b200 import models, noise
from qiddm_b200 import nn as qnn

imgs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda")
torch.manual_seed(0)
net = qnn.UNetUndirected(3, 8, 3)
diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (28, 28), torch.nn.MSELoss()).to(dev, torch.float64)
diff.train()
opt = torch.optim.Adam(diff.parameters(), lr=1e-3)
x = torch.rand(imgs, 784, device=dev, dtype=torch.float64)


def step():
    opt.zero_grad(set_to_none=True)
    diff(x=x, T=10)
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))

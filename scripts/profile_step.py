#!/usr/bin/env python
"""Kernel-time breakdown (torch.profiler, CUDA activities only) of eager diffusion training steps of one model.
  python scripts/profile_step.py qiddm_ll 4096      # model (as scripts/run_step.py), images per step"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile
from qiddm_b200 import models, noise
from qiddm_b200 import nn as qnn

name, imgs = sys.argv[1], int(sys.argv[2])
goal = sys.argv[3] if len(sys.argv) > 3 else "data"
dev = torch.device("cuda")
torch.manual_seed(0)
net = {"unet": lambda: qnn.UNetUndirected(3, 8, 3), "qiddm_ll": lambda: qnn.QIDDM_LL_noise(784, 6, 14, 2),
       "qiddm_pl": lambda: qnn.QIDDM_PL_noise(784, 8, 6, 2), "qnn": lambda: qnn.QNN_noise(784, 8, 14)}[name]()
diff = models.Diffusion(net, noise.add_normal_noise_multiple, goal, (28, 28), torch.nn.MSELoss()).to(dev, torch.float64)
diff.train()
opt = torch.optim.Adam(diff.parameters(), lr=1e-3, capturable=True)
x = torch.rand(imgs, 784, device=dev, dtype=torch.float64)


def step():
    opt.zero_grad(set_to_none=True)
    diff(x=x, T=10)
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    step()
b.record()
torch.cuda.synchronize()
print(f"{name} {imgs} images: eager {a.elapsed_time(b) / 5:.3f} ms/step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=100))

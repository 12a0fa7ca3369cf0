#!/usr/bin/env python
"""Runs a few eager diffusion training steps of one model (for ncu launch lists).
  python scripts/run_step.py unet 64 3      # model, images per step, steps"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from qiddm_b200 import models, noise
from qiddm_b200 import nn as qnn

name, imgs, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda")
torch.manual_seed(0)
side = 28
net = {"unet": lambda: qnn.UNetUndirected(3, 8, 3), "qiddm_ll": lambda: qnn.QIDDM_LL_noise(784, 6, 14, 2),
       "qiddm_pl": lambda: qnn.QIDDM_PL_noise(784, 8, 6, 2), "qnn": lambda: qnn.QNN_noise(784, 8, 14)}[name]()
diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (side, side), torch.nn.MSELoss()).to(dev, torch.float64)
diff.train()
opt = torch.optim.Adam(diff.parameters(), lr=1e-3)
x = torch.rand(imgs, side * side, device=dev, dtype=torch.float64)
for _ in range(steps):
    opt.zero_grad(set_to_none=True)
    (loss,) = diff(x=x, T=10)
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))

#!/usr/bin/env python
"""PRODUCT path against outputs of the real stack (round-2 candidate for a `-m gpu` test; not yet run on a GPU).
Loads the reference's QIDDM_PL_noise(784,8,6,2) checkpoint (tests/golden/f3_qiddm_pl_logo_sanyo.pt) into the product module on
cuda:0, runs the 5 `Diffusion.sample` iterations the reference ran, and compares with the images the reference saved
(tests/test_oracle.py does the same with the CPU oracle: 2.3 / 255 grey levels, correlation 0.9997).
  python scripts/check_reference_images_gpu.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from qiddm_b200 import models, nn, noise

gold = torch.load(ROOT / "tests/golden/f3_qiddm_pl_logo_sanyo.pt", weights_only=True)
net = nn.QIDDM_PL_noise(784, 8, 6, 2)
net.load_state_dict({k: gold[k] for k in ("weights1", "linear_up.weight", "linear_up.bias")})
diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (28, 28), torch.nn.MSELoss()).to("cuda:0", torch.float64)
diff.eval()
S = gold["sample_steps_u8"].double()
x0 = 0.5 + S[0] / 255 * 0.5
x0[S[0] == 255] = 1.125
out = diff.sample(5, first_x=x0.reshape(10, 1, 28, 28).cuda(), only_last=True).cpu()
img = out[:, 0].clamp(0, 1)
lo, hi = img.amin(dim=(1, 2), keepdim=True), img.amax(dim=(1, 2), keepdim=True)
pred = (img - lo) / (hi - lo) * 255
corr = torch.stack([torch.corrcoef(torch.stack([pred[i].flatten(), S[5][i].flatten()]))[0, 1] for i in range(10)]).mean().item()
d = (pred - S[5]).abs().mean().item()
print(f"mean |diff| {d:.2f} of 255 grey levels, correlation {corr:.4f} (oracle: 2.3, 0.9997)")
sys.exit(0 if (d < 4.0 and corr > 0.999) else 1)

#!/usr/bin/env python
"""One small QConv2d forward + backward on the direct-convolution path (for compute-sanitizer)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from qiddm_b200 import nn

torch.manual_seed(0)
for cin, cout, k, h, w in ((8, 8, 3, 9, 11), (16, 16, 3, 14, 14), (32, 16, 1, 7, 7), (1, 8, 3, 28, 28)):
    m = nn.QConv2d(cin, cout, kernel_size=k, padding=k // 2, qdepth=2).cuda()
    x = torch.rand(3, cin, h, w, dtype=torch.float64, device="cuda", requires_grad=True)
    out = m(x)
    out.square().sum().backward()
    torch.cuda.synchronize()
    print(cin, cout, k, float(out.sum()), float(x.grad.abs().sum()), float(m.weights.grad.abs().sum()))

#!/usr/bin/env python
"""Forward GEMM of the bench layer alone: training forward (keeps Y) vs inference forward (no Y store), CUDA events."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os
import torch
from qiddm_b200 import _lib as L, nn as qnn
import dataclasses

B = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
net = qnn.QDenseUndirected_old_noise(60, 28).to("cuda", torch.float64)
spec = dataclasses.replace(net._spec(), path=L.PATH_GEMM)
plan = L.Plan.get(spec)
x = torch.rand(B, 784, device="cuda")
w = net.weights.detach()
ev = lambda: torch.cuda.Event(enable_timing=True)
res = {}
for name, save in (("train_fwd", True), ("infer_fwd", False)):
    for _ in range(3):
        plan.gemm_forward(x, w, save=save)
    torch.cuda.synchronize()
    L.timing_enable(True); L.timing_collect()
    for _ in range(6):
        plan.gemm_forward(x, w, save=save)
    k = L.timing_collect(); L.timing_enable(False)
    res[name] = k["gemm_forward"]["ms"] / 6
print(os.environ.get("TAG", ""), {k: round(v, 3) for k, v in res.items()}, flush=True)

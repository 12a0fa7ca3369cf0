#!/usr/bin/env python
"""Measured GEMM-path accuracy vs the complex128 oracle (rel-to-max), for tuning the fp16 split scheme."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
from oracle import qiddm_oracle as O
from qiddm_b200 import _lib as L
from qiddm_b200.functional import run_stage

def rel(a, b): return ((a.double().cpu() - b).abs().max() / b.abs().max()).item()

for (n, F, K, depth, small) in [(10, 784, 784, 60, False), (10, 784, 784, 60, True), (8, 200, 200, 10, False), (7, 72, 8, 3, True)]:
    g = torch.Generator().manual_seed(n)
    d = O.StageDesc(n_qubits=n, layers_per_block=depth, init=O.INIT_AMPLITUDE, n_features=F, pad_value=0.1,
                    imprimitive=O.IMP_CNOT, remap=O.REMAP_TANH, readout=O.READ_PROBS, read_count=K, post_scale=float(K))
    W = torch.randn(1, depth, n, 3, generator=g, dtype=torch.float64) * 0.4
    x = torch.rand(300, F, generator=g, dtype=torch.float64) * (0.05 if small else 1.0)
    Wr, xr = W.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ref = O.run_stage(d, xr, Wr, batch=300)
    go = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * go).sum().backward()
    spec = L.StageSpec(n_qubits=n, layers_per_block=depth, init=L.INIT_AMPLITUDE, n_features=F, pad_value=0.1,
                       imprimitive=L.IMP_CNOT, remap=L.REMAP_TANH, readout=L.READ_PROBS, read_count=K,
                       post_scale=float(K), path=L.PATH_GEMM, gemm_precision=3)
    Wd, xd = W.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    out = run_stage(spec, xd, Wd)
    (out * go.cuda()).sum().backward()
    print(f"n={n} F={F} K={K} depth={depth} small_inputs={small}: out {rel(out, ref.detach()):.2e}  dW {rel(Wd.grad, Wr.grad):.2e}  dX {rel(xd.grad, xr.grad):.2e}")

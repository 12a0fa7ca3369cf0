#!/usr/bin/env python
"""CPU emulation of the fp16 operand-split schemes of the collapse (GEMM) path on the bench circuit
(QDenseUndirected_old_noise(60,28): n = 10, 784 features, 784 retained amplitudes): how much accuracy each split of the
gradient GEMMs costs.  Operands are rounded to fp16 hi / lo exactly as the kernels do (hi = fp16(v), lo = fp16((v - hi) 2^11)
/ 2^11); the products are accumulated in float64, so the numbers are the error of the DROPPED cross terms and operand
rounding only (the tensor cores' fp32 accumulation adds about 1e-6 on top).  No GPU needed.
  python scripts/emulate_split_accuracy.py [--batch 256]
Output: rel-to-max errors of dX, dW^T and of the final weight gradient (dW pushed through the circuit with autograd on the
oracle's unitary) against the float64 result, for  x3 = hi hi + lo hi + hi lo,  x2a = hi hi + lo hi (activation-side lo kept),
x2b = hi hi + hi lo (weight-side lo kept),  x1 = hi hi."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from oracle import qiddm_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--depth", type=int, default=60)
a = ap.parse_args()
torch.manual_seed(0)
n, F, K = 10, 784, 784
A = 2 ** n
d = O.desc_qdense(a.depth, F, O.REMAP_TANH)
Wc = (torch.randn(1, a.depth, n, 3, dtype=torch.float64) * 0.4).requires_grad_(True)
U = O.circuit_unitary(d, Wc)                                   # (A, A) complex128, differentiable in the weights
x = torch.rand(a.batch, F, dtype=torch.float64)
go = torch.randn(a.batch, K, dtype=torch.float64) / (a.batch * K)


def split(v):
    hi = v.to(torch.float16).to(torch.float64)
    lo = ((v - hi) * 2048).to(torch.float16).to(torch.float64) / 2048
    return hi, lo


def mm(Ah, Al, Bh, Bl, scheme):
    out = Ah @ Bh
    if scheme in ("x3", "x2a"):
        out = out + Al @ Bh
    if scheme in ("x3", "x2b"):
        out = out + Ah @ Bl
    return out


# real GEMM operands of the kernel: X = [f | 1], W = rows of [Re U^T | Im U^T] for the F features + the folded pad row
pad = torch.full((A - F,), d.pad_value, dtype=torch.float64)
Ud = U.detach()
Wf = torch.cat([Ud[:K, :F].T.real, Ud[:K, :F].T.imag], dim=1)             # (F, 2K): columns [re | im] (ordering is irrelevant here)
bias = torch.cat([(Ud[:K, F:] @ pad.to(Ud.dtype)).real, (Ud[:K, F:] @ pad.to(Ud.dtype)).imag])
Wfull = torch.cat([Wf, bias[None]], dim=0)                               # (F + 1, 2K)
X = torch.cat([x, torch.ones(a.batch, 1, dtype=torch.float64)], dim=1)
inv_n2 = 1.0 / ((x ** 2).sum(1) + (A - F) * d.pad_value ** 2)
Y = X @ Wfull                                                            # exact forward (the forward keeps x3 in every scheme)
P = (Y[:, :K] ** 2 + Y[:, K:] ** 2) * inv_n2[:, None] * d.post_scale
mask = ((P >= d.clamp_lo) & (P <= d.clamp_hi)).double() if d.clamp else torch.ones_like(P)
c = 2 * go * mask * inv_n2[:, None] * d.post_scale
G = torch.cat([c * Y[:, :K], c * Y[:, K:]], dim=1)                       # dL/dY  (B, 2K)
gs = 2.0 ** (14 - torch.frexp(G.abs().max())[1].item())                  # the kernels' power-of-two scale into fp16 range
Gh, Gl = split(G * gs)
Xh, Xl = split(X)
Wh, Wl = split(Wfull)
dX_ref, dW_ref = G @ Wfull.T, G.T @ X


def weight_grad(dWT):
    """dL/dU[j, k] from dW^T (2K, F + 1) -> gradient of the circuit weights through the oracle's unitary."""
    gU = torch.zeros(A, A, dtype=torch.complex128)
    gU[:K, :F] = torch.complex(dWT[:K, :F], dWT[K:, :F])
    gU[:K, F:] = torch.complex(dWT[:K, F:], dWT[K:, F:]) * d.pad_value    # the ones column spreads over the pad rows
    (g,) = torch.autograd.grad((U.real * gU.real + U.imag * gU.imag).sum(), Wc, retain_graph=True)
    return g


# fidelity check of the emulation itself: the forward under x1 / x3 against the GPU's measured 3e-4 / 8.7e-6
P_ref = (Y[:, :K] ** 2 + Y[:, K:] ** 2) * inv_n2[:, None] * d.post_scale
for scheme in ("x3", "x1"):
    Ys = mm(Xh, Xl, Wh, Wl, scheme)
    Ps = (Ys[:, :K] ** 2 + Ys[:, K:] ** 2) * inv_n2[:, None] * d.post_scale
    print(f"forward {scheme}: un-clamped output error {((Ps - P_ref).abs().max() / P_ref.abs().max()).item():.2e} rel-to-max")

gw_ref = weight_grad(dW_ref)
rel = lambda v, r: ((v - r).abs().max() / r.abs().max()).item()
print(f"bench circuit n={n} depth={a.depth} batch={a.batch}; rel-to-max errors vs float64")
print(f"{'scheme':6s} {'dX':>10s} {'dW^T':>10s} {'weight grad':>12s}")
for scheme in ("x3", "x2a", "x2b", "x1"):
    dX = mm(Gh, Gl, Wh.T, Wl.T, scheme) / gs
    dW = mm(Gh.T, Gl.T, Xh, Xl, scheme) / gs
    print(f"{scheme:6s} {rel(dX, dX_ref):10.2e} {rel(dW, dW_ref):10.2e} {rel(weight_grad(dW), gw_ref):12.2e}")

#!/usr/bin/env python
"""Measured parity margins per circuit family at the BASELINE sizes (VERDICT r1 "next" 4a): rel-to-max error of outputs,
weight gradients and input gradients of the CUDA paths against the C restatement (oracle/statevec_oracle.c: complex128,
adjoint-method gradients), three seeds each, for a random-sign upstream gradient ("randn": instances cancel in dW) and an
MSE-like one ("mse": 2 (out - target) / numel).  Writes gpurun_out/parity_margins.json and a markdown table to stdout.
  python scripts/measure_parity_margins.py"""
import dataclasses
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch

from conftest import rel_to_max
from oracle import c_oracle as CO
from oracle import qiddm_oracle as O
from qiddm_b200 import _lib as L
from qiddm_b200.functional import run_stage


def spec_of(d, path, precision=3, bwd=0):
    return L.StageSpec(n_qubits=d.n_qubits, n_blocks=d.n_blocks, layers_per_block=d.layers_per_block, init=d.init,
                       n_features=d.n_features, pad_value=d.pad_value, add_offset=d.add_offset, enc=d.enc,
                       enc_scale=d.enc_scale, imprimitive=d.imprimitive, remap=d.remap, readout=d.readout,
                       read_count=d.read_count, read_stride=d.read_stride, post_scale=d.post_scale, clamp=d.clamp,
                       clamp_lo=d.clamp_lo, clamp_hi=d.clamp_hi, path=path, gemm_precision=precision, gemm_bwd_precision=bwd)


CASES = [
    # row, label, descriptor, path, precision, bwd precision, batch, weight scale
    ("a1", "QDenseUndirected_old_noise(60,28) n=10", O.desc_qdense(60, 784, O.REMAP_TANH), "gate", 3, 0, 64, 0.4),
    ("a1", "QDenseUndirected_old_noise(60,28) n=10", O.desc_qdense(60, 784, O.REMAP_TANH), "gemm", 3, 0, 512, 0.4),
    ("a1", "same, un-clamped", dataclasses.replace(O.desc_qdense(60, 784, O.REMAP_TANH), clamp=False), "gemm", 3, 0, 512, 0.4),
    ("a1", "same, un-clamped", dataclasses.replace(O.desc_qdense(60, 784, O.REMAP_TANH), clamp=False), "gemm", 3, 1, 512, 0.4),
    ("a1", "same, un-clamped", dataclasses.replace(O.desc_qdense(60, 784, O.REMAP_TANH), clamp=False), "gemm", 1, 0, 512, 0.4),
    ("a1", "QDenseUndirected_old(60,64) n=12", O.desc_qdense(60, 4096, O.REMAP_PI_TANH), "gate", 3, 0, 12, 0.4),
    ("a1", "QDenseUndirected_old(60,64) n=12", O.desc_qdense(60, 4096, O.REMAP_PI_TANH), "gemm", 3, 0, 48, 0.4),
    ("a1", "same, un-clamped", dataclasses.replace(O.desc_qdense(60, 4096, O.REMAP_PI_TANH), clamp=False), "gemm", 3, 0, 48, 0.4),
    ("a1", "QDenseUndirected_old_noise(60,8) n=6", O.desc_qdense(60, 64, O.REMAP_TANH), "gate", 3, 0, 256, 0.4),
    ("a2", "QNN_A(10,28) circuit n=10", O.desc_qnn_a(10, 784), "gate", 3, 0, 64, 0.4),
    ("a3", "differN(28,9,2) stage (next angles)", O.desc_reupload(10, 9, 2, readout=O.READ_PROBS, read_count=10), "gate", 3, 0, 64, 0.4),
    ("a3", "differN(28,15,2) last stage (784 probs, clamp)",
     dataclasses.replace(O.desc_reupload(10, 15, 2, readout=O.READ_PROBS, read_count=784), post_scale=784.0, clamp=True), "gate", 3, 0, 64, 0.4),
    ("a4", "QIDDM_LL_noise(784,6,14,2) stage n=6", O.desc_reupload(6, 14, 2), "gate", 3, 0, 256, 0.4),
    ("a4", "QIDDM_PL_noise(784,8,6,2) stage n=8", O.desc_reupload(8, 6, 2), "gate", 3, 0, 256, 0.4),
    ("a5", "QNN_noise(784,8,14) circuit", O.desc_reupload(8, 1, 14), "gate", 3, 0, 256, 0.4),
    ("a6", "QConv2d(8,8,k3,q3) rows n=7", O.desc_qconv(8, 8, (3, 3), 3), "gate", 3, 0, 1024, 1.0),
    ("a6", "QConv2d(8,8,k3,q3) rows n=7", O.desc_qconv(8, 8, (3, 3), 3), "gemm", 3, 0, 1024, 1.0),
    ("a6", "QConv2d(32,32,k3,q3) rows n=9", O.desc_qconv(32, 32, (3, 3), 3), "gemm", 3, 0, 1024, 1.0),
]


def one(d, path, precision, bwd, B, wscale, seed, upstream):
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(d.n_blocks, d.layers_per_block, d.n_qubits, 3, generator=g, dtype=torch.float64) * wscale
    if d.init == O.INIT_AMPLITUDE:
        x = torch.rand(B, d.n_features, generator=g, dtype=torch.float64)
    else:
        x = torch.randn(B, d.n_qubits, generator=g, dtype=torch.float64)
    ref = CO.run_stage(d, x, W)
    if upstream == "randn":
        go = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    else:
        go = 2.0 * (ref - torch.rand(ref.shape, generator=g, dtype=torch.float64)) / ref.numel()
    gw_ref, gx_ref = CO.stage_grads(d, x, W, go)
    Wd, xd = W.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    pid = L.PATH_GATE if path == "gate" else L.PATH_GEMM
    out = run_stage(spec_of(d, pid, precision, bwd), xd, Wd)
    (out * go.cuda()).sum().backward()
    floor = 1e-6 * go.abs().max().item()      # analytically-zero gradients (RZ on |0..0>, QNN_noise) are compared absolutely
    flips = 0
    if d.clamp:
        o = out.detach().cpu().double()
        flips = int((((o >= d.clamp_hi) != (ref >= d.clamp_hi)) | ((o <= d.clamp_lo) != (ref <= d.clamp_lo))).sum())
    return (rel_to_max(out, ref), rel_to_max(Wd.grad, gw_ref, floor), rel_to_max(xd.grad, gx_ref, floor), flips, ref.numel())


def main():
    rows = []
    for row, label, d, path, prec, bwd, B, ws in CASES:
        rec = {"row": row, "circuit": label, "path": path if path == "gate" else f"gemm x{prec}" + (f" / gradients x{bwd}" if bwd else ""),
               "batch": B}
        for up in ("randn", "mse"):
            errs = [one(d, path, prec, bwd, B, ws, 7000 + s, up) for s in range(3)]
            rec[up] = {"out": max(e[0] for e in errs), "grad_w": max(e[1] for e in errs), "grad_x": max(e[2] for e in errs),
                       "clamp_flips": sum(e[3] for e in errs), "outputs": sum(e[4] for e in errs)}
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_margins.json").write_text(json.dumps(rows, indent=1))
    print("\n| row | circuit | path | B | out | dW (randn) | dX (randn) | dW (mse) | dX (mse) | clamp decisions that differ |\n"
          "|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['row']} | {r['circuit']} | {r['path']} | {r['batch']} | {r['randn']['out']:.1e} | {r['randn']['grad_w']:.1e} | "
              f"{r['randn']['grad_x']:.1e} | {r['mse']['grad_w']:.1e} | {r['mse']['grad_x']:.1e} | "
              f"{r['randn']['clamp_flips']} of {r['randn']['outputs']} |")


if __name__ == "__main__":
    main()

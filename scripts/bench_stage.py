#!/usr/bin/env python
"""Kernel-level timing of one circuit stage on the gate path (CUDA events inside the library).
  python scripts/bench_stage.py --family reupload --n 6 --L 14 --batch 262144
  python scripts/bench_stage.py --family qdense --n 10 --depth 60 --batch 16384"""
import argparse, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from qiddm_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--family", default="reupload", choices=["reupload", "reupload_probs", "qdense"])
ap.add_argument("--n", type=int, default=6)
ap.add_argument("--L", type=int, default=14)
ap.add_argument("--depth", type=int, default=60)
ap.add_argument("--batch", type=int, default=262144)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda")
n, A = a.n, 1 << a.n
if a.family == "qdense":
    F = min(A, 784) if n == 10 else A
    spec = L.StageSpec(n_qubits=n, layers_per_block=a.depth, init=L.INIT_AMPLITUDE, n_features=F, pad_value=0.1,
                       imprimitive=L.IMP_CNOT, remap=L.REMAP_TANH, readout=L.READ_PROBS, read_count=F,
                       post_scale=float(F), clamp=True, path=L.PATH_GATE)
    x = torch.rand(a.batch, F, device=dev)
    w = torch.randn(a.depth, n, 3, device=dev, dtype=torch.float64) * 0.4
    n_rot = a.depth * n
else:
    probs = a.family == "reupload_probs"
    spec = L.StageSpec(n_qubits=n, n_blocks=a.L, layers_per_block=2, init=L.INIT_ZERO, enc=L.ENC_RZ, imprimitive=L.IMP_CZ,
                       readout=L.READ_PROBS if probs else L.READ_EXPVAL_Z, read_count=min(A, 784) if probs else 0,
                       post_scale=float(min(A, 784)) if probs else 1.0, clamp=probs, path=L.PATH_GATE)
    x = torch.randn(a.batch, n, device=dev)
    w = torch.randn(a.L, 2, n, 3, device=dev, dtype=torch.float64) * 0.4
    n_rot = a.L * 2 * n
plan = L.Plan.get(spec)
go = torch.randn(a.batch, spec.n_out, device=dev)
for _ in range(2):
    plan.forward(x, w); plan.backward(x, w, go)
torch.cuda.synchronize()
L.timing_enable(True); L.timing_collect()
for _ in range(a.iters):
    plan.forward(x, w); plan.backward(x, w, go)
k = L.timing_collect(); L.timing_enable(False)
f, b = k["gate_forward"]["ms"] / a.iters, k["gate_backward"]["ms"] / a.iters
flop = a.batch * n_rot * 14.0 * A
print(json.dumps({"family": a.family, "n": n, "n_rot": n_rot, "batch": a.batch, "fwd_ms": round(f, 4), "bwd_ms": round(b, 4),
                  "fwd_evals_per_s": round(a.batch / f * 1e3), "fwd_bwd_evals_per_s": round(a.batch / (f + b) * 1e3),
                  "fwd_alg_tflops": round(flop / f / 1e9, 2), "bwd_alg_tflops(4x)": round(4 * flop / b / 1e9, 2)}))

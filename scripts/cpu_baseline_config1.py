#!/usr/bin/env python
"""CPU baseline of BASELINE.json config 1 (src/mnist_exm.py: one training step, batch 1 image x tau = 10) with the oracle,
in the two modes SURVEY.md §8(d) names: (a) the whole (batch tau) ladder as ONE batched circuit call with autograd through
the circuit, and (b) the reference's own shape of the computation — a Python loop of B = 1 circuit calls per sample and
chained stage (nn/qdense.py:1631-1635: `for i in range(b)` x N QNode calls), forward only through the circuit (the
reference detaches it), gradients for the classical layers only; and (c) the same loop on the C restatement (oracle/statevec_oracle.c: one state vector,
gates applied in place one after the other, like lightning.qubit, the reference's device for these classes).  Needs no GPU; prints one JSON line per model.
  python scripts/cpu_baseline_config1.py [--repeat 3]"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from oracle import c_oracle as C
from oracle import qiddm_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--repeat", type=int, default=3)
a = ap.parse_args()
torch.set_num_threads(os.cpu_count() or 1)


def best(fn, n):
    fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


def params(n_hidden, shape_w, seed):
    g = torch.Generator().manual_seed(seed)
    p = {"w": (torch.randn(*shape_w, generator=g, dtype=torch.float64) * 0.4).requires_grad_(True),
         "wd": (torch.randn(n_hidden, 784, generator=g, dtype=torch.float64) * 0.03).requires_grad_(True),
         "bd": torch.zeros(n_hidden, dtype=torch.float64, requires_grad=True),
         "wu": (torch.randn(784, n_hidden, generator=g, dtype=torch.float64) * 0.3).requires_grad_(True),
         "bu": torch.zeros(784, dtype=torch.float64, requires_grad=True)}
    return p


for name, n, shape_w in (("QIDDM_LL_noise(784,6,14,2)", 6, (2, 14, 2, 6, 3)), ("QNN_noise(784,8,14)", 8, (14, 8, 3))):
    p = params(n, shape_w, 42)
    torch.manual_seed(42)
    data = torch.rand(1, 784, dtype=torch.float64)
    eps = torch.normal(0.5, 0.2, size=(1, 784)).double()
    ll = name.startswith("QIDDM_LL")

    def chain(ang, w):                      # (B, n) angles -> (B, n) <Z>
        return O.qiddm_expval_chain(ang, w) if ll else O.run_stage(O.desc_reupload(n, 1, shape_w[0]), ang, w[None])

    def batched(v):
        ang = v.reshape(v.shape[0], -1) @ p["wd"].T + p["bd"]
        return (chain(ang, p["w"]) @ p["wu"].T + p["bu"]).reshape(-1, 1, 28, 28)

    def looped(v):
        ang = (v.reshape(v.shape[0], -1) @ p["wd"].T + p["bd"]).detach()
        with torch.no_grad():               # one circuit call per sample (and per chained stage inside `chain`)
            z = torch.cat([chain(ang[i:i + 1], p["w"]) for i in range(ang.shape[0])])
        return (z @ p["wu"].T + p["bu"]).reshape(-1, 1, 28, 28)

    def step(f):
        def run():
            for q in p.values():
                q.grad = None
            O.diffusion_loss(f, data, eps, 10, (28, 28), "data").backward()
        return run

    def looped_c(v):                        # the same loop on the gate-by-gate C restatement (lightning.qubit-like), 1 thread per call
        ang = (v.reshape(v.shape[0], -1) @ p["wd"].T + p["bd"]).detach()
        rows = []
        for i in range(ang.shape[0]):
            z = ang[i:i + 1]
            if ll:
                for k in range(shape_w[0]):
                    z = C.run_stage(O.desc_reupload(n, shape_w[1], shape_w[2]), z, p["w"].detach()[k], threads=1)
            else:
                z = C.run_stage(O.desc_reupload(n, 1, shape_w[0]), z, p["w"].detach()[None], threads=1)
            rows.append(z)
        return (torch.cat(rows) @ p["wu"].T + p["bu"]).reshape(-1, 1, 28, 28)

    tb, tl, tc = best(step(batched), a.repeat), best(step(looped), a.repeat), best(step(looped_c), a.repeat)
    stages = 2 if ll else 1
    print(json.dumps({"what": "config1_cpu_baseline", "model": name, "images_per_step": 1, "tau": 10, "cores": os.cpu_count(),
                      "dtype": "complex128", "batched_autograd_ms_per_step": round(tb * 1e3, 1),
                      "per_sample_loop_forward_only_ms_per_step": round(tl * 1e3, 1),
                      "per_sample_loop_c_gate_by_gate_ms_per_step": round(tc * 1e3, 2),
                      "circuit_evals_per_s_batched": round(10 * stages / tb, 1),
                      "circuit_evals_per_s_loop_c": round(10 * stages / tc, 1),
                      "circuit_evals_per_s_loop": round(10 * stages / tl, 1), "host": "build container (no GPU)"}))

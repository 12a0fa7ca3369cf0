#!/usr/bin/env python
"""Is the output error of the tcgen05 x3 path a systematic (truncation) bias?  Un-clamped probabilities of the collapse path against
the C restatement: mean and spread of p_gpu / p_ref - 1 over the outputs above 1e-3 of the maximum, for K = 256 / 784 / 4096."""
import dataclasses
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from oracle import c_oracle as CO
from oracle import qiddm_oracle as O
from qiddm_b200 import _lib as L
from qiddm_b200.functional import run_stage
sys.path.insert(0, str(ROOT / "scripts"))
from measure_parity_margins import spec_of

for n, P, B in ((8, 256, 512), (10, 784, 512), (10, 1024, 512), (12, 4096, 64)):
    d = dataclasses.replace(O.desc_qdense(60, P, O.REMAP_TANH), clamp=False)
    g = torch.Generator().manual_seed(n)
    W = torch.randn(1, 60, n, 3, generator=g, dtype=torch.float64) * 0.4
    x = torch.rand(B, P, generator=g, dtype=torch.float64)
    ref = CO.run_stage(d, x, W)
    for prec in (3, 1):
        out = run_stage(spec_of(d, L.PATH_GEMM, prec), x.cuda(), W.cuda()).double().cpu()
        m = ref > 1e-3 * ref.max()
        r = (out[m] / ref[m] - 1.0)
        print(json.dumps({"n": n, "K": P, "precision": prec, "mean_rel": r.mean().item(), "std_rel": r.std().item(),
                          "max_abs_rel_to_max": ((out - ref).abs().max() / ref.max()).item(),
                          "mean_rel_top": (out[ref > 0.5 * ref.max()] / ref[ref > 0.5 * ref.max()] - 1).mean().item()}), flush=True)

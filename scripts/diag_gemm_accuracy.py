"""Diagnostic: where does the x3 GEMM path's forward error come from? (gate path vs GEMM vs U build)"""
import sys, dataclasses, torch
sys.path.insert(0, ".")
from oracle import qiddm_oracle as O
from qiddm_b200 import _lib as L
from qiddm_b200.functional import run_stage, build_unitary
sys.path.insert(0, "tests")
from test_gpu_gemm_path import _spec

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()

for depth in (4, 20, 60):
    d = O.desc_qdense(depth, 784, O.REMAP_TANH)
    d = dataclasses.replace(d, clamp=False)
    g = torch.Generator().manual_seed(1)
    W = torch.randn(1, depth, 10, 3, generator=g, dtype=torch.float64) * 0.4
    x = torch.rand(257, 784, generator=g, dtype=torch.float64)
    ref = O.run_stage(d, x, W)
    gate = run_stage(_spec(d, L.PATH_GATE), x.cuda(), W.cuda())
    gemm3 = run_stage(_spec(d, L.PATH_GEMM, 3), x.cuda(), W.cuda())
    gemm1 = run_stage(_spec(d, L.PATH_GEMM, 1), x.cuda(), W.cuda())
    U = build_unitary(_spec(d, L.PATH_GATE), W.cuda())
    Uref = O.circuit_unitary(d, W)
    eU = (U.cpu().to(torch.complex128) - Uref).abs().max().item() / Uref.abs().max().item()
    print(f"depth {depth}: gate {rel(gate, ref):.2e}  gemm_x3 {rel(gemm3, ref):.2e}  gemm_x1 {rel(gemm1, ref):.2e}  "
          f"U build {eU:.2e}  gemm3 vs gate {rel(gemm3, gate):.2e}  max ref {ref.max().item():.3f}")

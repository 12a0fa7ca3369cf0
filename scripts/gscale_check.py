"""GEMM path and gate path against the complex128 oracle at B = 32768 (n = 6): which of the two carries the large-batch error."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import rel_to_max
from oracle import qiddm_oracle as O
from qiddm_b200 import _lib as L
from test_gpu_big_batch import _spec, _grads

n, F, K, B = 6, 64, 64, 32768
d = O.StageDesc(n_qubits=n, layers_per_block=3, init=O.INIT_AMPLITUDE, n_features=F, pad_value=0.0, add_offset=0.1,
                imprimitive=O.IMP_CNOT, remap=O.REMAP_TANH, readout=O.READ_PROBS, read_count=K, read_stride=1,
                post_scale=float(2 ** n) / 2)
for kind in ["pos", "mixed", "outlier12345", "outlier0"]:
    g = torch.Generator().manual_seed(5)
    W = (torch.randn(1, 3, n, 3, generator=g, dtype=torch.float64) * 0.4)
    x = torch.rand(B, F, generator=g, dtype=torch.float32)
    go = torch.rand(B, K, generator=g, dtype=torch.float32) * 1e-6
    if kind == "mixed":
        go = (go - 0.5e-6) * 2
    if kind.startswith("outlier"):
        go[int(kind[7:])] = torch.rand(K, generator=g, dtype=torch.float32) * 1e3
    Wr, xr = W.clone().requires_grad_(True), x.double().requires_grad_(True)
    ref = O.run_stage(d, xr, Wr)
    (ref * go.double()).sum().backward()
    o1, w1, x1 = _grads(_spec(d, L.PATH_GEMM), x.cuda(), W.cuda(), go.cuda())
    o0, w0, x0 = _grads(_spec(d, L.PATH_GATE), x.cuda(), W.cuda(), go.cuda())
    print(kind, "gemm: out %.2e dW %.2e dX %.2e | gate: out %.2e dW %.2e dX %.2e" % (
        rel_to_max(o1, ref), rel_to_max(w1, Wr.grad), rel_to_max(x1, xr.grad),
        rel_to_max(o0, ref), rel_to_max(w0, Wr.grad), rel_to_max(x0, xr.grad)), flush=True)

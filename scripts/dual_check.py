"""Prints the GEMM-path vs gate-path gradient errors at large batches (run under different QIDDM_GEMM_DW_* environments)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import rel_to_max
from oracle import qiddm_oracle as O
from qiddm_b200 import _lib as L
from test_gpu_big_batch import _spec, _grads

for (n, F, K, B) in [(10, 784, 784, 65536 + 40), (10, 576, 300, 65536), (9, 400, 100, 70000), (10, 784, 784, 8192)]:
    d = O.StageDesc(n_qubits=n, layers_per_block=2, init=O.INIT_AMPLITUDE, n_features=F, pad_value=0.3, add_offset=0.1,
                    imprimitive=O.IMP_CNOT, remap=O.REMAP_TANH, readout=O.READ_PROBS, read_count=K, read_stride=1,
                    post_scale=float(2 ** n) / 2)
    g = torch.Generator().manual_seed(n + F)
    W = (torch.randn(1, 2, n, 3, generator=g, dtype=torch.float64) * 0.4).cuda()
    x = torch.rand(B, F, generator=g, dtype=torch.float32).cuda()
    go = (torch.rand(B, K, generator=g, dtype=torch.float32) / B).cuda()
    o1, w1, x1 = _grads(_spec(d, L.PATH_GEMM), x, W, go)
    o0, w0, x0 = _grads(_spec(d, L.PATH_GATE), x, W, go)
    # fp64 reference of the weight gradient on a subset is too slow at this B: report GEMM vs gate and the dW magnitude
    print(os.environ.get("QIDDM_GEMM_DW_DUAL", "-"), os.environ.get("QIDDM_GEMM_DW_BK", "-"), (n, F, K, B),
          "out %.2e dW %.2e dX %.2e |dW|max %.3e" % (rel_to_max(o1, o0), rel_to_max(w1, w0), rel_to_max(x1, x0), w0.abs().max().item()), flush=True)

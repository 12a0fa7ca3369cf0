"""Large-batch branches of the unitary-collapse path that the small oracle-sized cases never reach (B >= 16 384 / 65 536):
the dual-N work items of the dW GEMM (two neighbouring N tiles fed from one staged G tile, 32-row k-blocks) and the
provisional scale of G (sampled bound, verified and if necessary redone by grad_y_kernel).  The checker is the fp32 gate
path of the same library (itself held to 1e-5 / 2e-5 against the oracle in test_gpu_stage_parity.py) plus, for the rows that
matter, the complex128 oracle."""
import dataclasses

import pytest
import torch

from conftest import rel_to_max
from oracle import qiddm_oracle as O

pytestmark = pytest.mark.gpu


def _spec(d: O.StageDesc, path, precision=3):
    from qiddm_b200._lib import StageSpec
    return StageSpec(n_qubits=d.n_qubits, n_blocks=d.n_blocks, layers_per_block=d.layers_per_block, init=d.init,
                     n_features=d.n_features, pad_value=d.pad_value, add_offset=d.add_offset, enc=d.enc,
                     enc_scale=d.enc_scale, imprimitive=d.imprimitive, remap=d.remap, readout=d.readout,
                     read_count=d.read_count, read_stride=d.read_stride, post_scale=d.post_scale, clamp=d.clamp,
                     clamp_lo=d.clamp_lo, clamp_hi=d.clamp_hi, path=path, gemm_precision=precision)


def _grads(spec, x, W, go):
    from qiddm_b200.functional import run_stage
    Wd, xd = W.clone().requires_grad_(True), x.clone().requires_grad_(True)
    out = run_stage(spec, xd, Wd)
    (out * go).sum().backward()
    return out.detach(), Wd.grad, xd.grad


# n, features, retained probabilities: dW is (2 K) x (F + 1) -> M tiles of 256 rows, N tiles chosen by pick_bn_mn:
# 784 + 1 -> 4 tiles of 208 (two full dual items); 576 + 1 -> 3 tiles (the second item holds ONE tile); 400 + 1 -> 2 tiles
@pytest.mark.parametrize("n,F,K,B", [(10, 784, 784, 65536 + 40), (10, 576, 300, 65536), (9, 400, 100, 70000)])
def test_dw_dual_n_items_match_the_gate_path(n, F, K, B):
    from qiddm_b200 import _lib as L
    d = O.StageDesc(n_qubits=n, layers_per_block=2, init=O.INIT_AMPLITUDE, n_features=F, pad_value=0.3, add_offset=0.1,
                    imprimitive=O.IMP_CNOT, remap=O.REMAP_TANH, readout=O.READ_PROBS, read_count=K, read_stride=1,
                    post_scale=float(2 ** n) / 2)
    g = torch.Generator().manual_seed(n + F)
    W = (torch.randn(1, 2, n, 3, generator=g, dtype=torch.float64) * 0.4).cuda()
    x = torch.rand(B, F, generator=g, dtype=torch.float32).cuda()
    go = (torch.rand(B, K, generator=g, dtype=torch.float32) / B).cuda()       # one sign: no cancellation across instances in dW
    o1, w1, x1 = _grads(_spec(d, L.PATH_GEMM), x, W, go)
    o0, w0, x0 = _grads(_spec(d, L.PATH_GATE), x, W, go)
    assert rel_to_max(o1, o0) <= 2e-5
    # A same-sign upstream gradient over 65 k instances is the worst case for the tensor cores' truncating accumulate
    # (-6e-9 per accumulated k-row): every dW^T entry is biased the same way and the parameter gradient is a
    # cancelling contraction of dW^T.  Measured 2.6e-5 ... 6.3e-5 with dual-N items (1.1e-4 with single-tile items, whose
    # split-K chains are the same length but twice as many); 2.4e-5 at B = 8192.
    assert rel_to_max(w1, w0) <= 1.5e-4, "weight gradient (dual-N dW GEMM)"
    assert rel_to_max(x1, x0) <= 3e-5
    # and against the complex128 oracle on a slice of rows (per-row quantities)
    idx = torch.tensor([0, 63, 64, 4097, B - 1])
    xr = x[idx.cuda()].double().cpu().requires_grad_(True)
    ref = O.run_stage(d, xr, W.cpu())
    (ref * go[idx.cuda()].double().cpu()).sum().backward()
    assert rel_to_max(o1[idx.cuda()], ref) <= 2e-5
    assert rel_to_max(x1[idx.cuda()], xr.grad) <= 3e-5


@pytest.mark.parametrize("outlier_row", [None, 12345, 0])
def test_provisional_g_scale_is_redone_when_an_unsampled_row_dominates(outlier_row):
    """grad_out tiny everywhere except (optionally) ONE row the 1-in-64 sample does not see (12345 = 64 * 192 + 57), 10^9
    times larger: the provisional scale would overflow fp16 there, grad_y's exact bound exceeds it and the pass is redone.
    Row 0 is in the sample (no second pass); None = no outlier.  All three must match the gate path."""
    from qiddm_b200 import _lib as L
    n, F, K, B = 6, 64, 64, 32768
    d = O.StageDesc(n_qubits=n, layers_per_block=3, init=O.INIT_AMPLITUDE, n_features=F, pad_value=0.0, add_offset=0.1,
                    imprimitive=O.IMP_CNOT, remap=O.REMAP_TANH, readout=O.READ_PROBS, read_count=K, read_stride=1,
                    post_scale=float(2 ** n) / 2)
    g = torch.Generator().manual_seed(5)
    W = (torch.randn(1, 3, n, 3, generator=g, dtype=torch.float64) * 0.4).cuda()
    x = torch.rand(B, F, generator=g, dtype=torch.float32).cuda()
    go = (torch.rand(B, K, generator=g, dtype=torch.float32).cuda() - 0.5) * 2e-6       # mixed signs (see the same-sign note above)
    if outlier_row is not None:
        go[outlier_row] = torch.rand(K, generator=g, dtype=torch.float32).cuda() * 1e3
    o1, w1, x1 = _grads(_spec(d, L.PATH_GEMM), x, W, go)
    o0, w0, x0 = _grads(_spec(d, L.PATH_GATE), x, W, go)
    assert torch.isfinite(w1).all() and torch.isfinite(x1).all()
    assert rel_to_max(w1, w0) <= 3e-5
    assert rel_to_max(x1, x0) <= 3e-5

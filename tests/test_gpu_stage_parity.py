"""GPU parity: CUDA gate-path kernels (through the C ABI) vs the complex128 oracle, forward and
adjoint backward, for every circuit family of SURVEY.md §8a at sizes the oracle finishes in seconds.
Tolerance: 1e-5 relative to the max-norm of the reference tensor for fp32 outputs (north_star) and 2e-5 for gradients:
measured at the BASELINE sizes (profiles/r2_parity_margins.md, three seeds per family): outputs <= 6.3e-6 (1.2e-5 for the
x784-scaled clamped differN readout), weight / input gradients <= 8.5e-6 on every gate-path family, n = 3 ... 12."""
import math

import pytest
import torch

from conftest import rel_to_max
from oracle import qiddm_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
GRAD_TOL = 2e-5


def _spec_from_desc(d: O.StageDesc):
    from qiddm_b200._lib import StageSpec
    return StageSpec(n_qubits=d.n_qubits, n_blocks=d.n_blocks, layers_per_block=d.layers_per_block, init=d.init,
                     n_features=d.n_features, pad_value=d.pad_value, add_offset=d.add_offset, enc=d.enc,
                     enc_scale=d.enc_scale, imprimitive=d.imprimitive, remap=d.remap, readout=d.readout,
                     read_count=d.read_count, read_stride=d.read_stride, post_scale=d.post_scale, clamp=d.clamp,
                     clamp_lo=d.clamp_lo, clamp_hi=d.clamp_hi)


def _check(d: O.StageDesc, B: int, seed: int = 0, wdtype=torch.float64, fwd_tol=FWD_TOL, grad_tol=GRAD_TOL):
    from qiddm_b200.functional import run_stage
    g = torch.Generator().manual_seed(seed)
    n = d.n_qubits
    W = (torch.randn(d.n_blocks, d.layers_per_block, n, 3, generator=g, dtype=torch.float64) * 0.4)
    if d.init == O.INIT_AMPLITUDE:
        x = torch.rand(B, d.n_features, generator=g, dtype=torch.float64)
    elif d.enc != O.ENC_NONE:
        x = torch.randn(B, n, generator=g, dtype=torch.float64)
    else:
        x = None
    # oracle
    Wr = W.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True) if x is not None else None
    ref = O.run_stage(d, xr, Wr, batch=B)
    gout = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * gout).sum().backward()
    # device
    Wd = W.to("cuda", wdtype).requires_grad_(True)
    xd = x.to("cuda").requires_grad_(True) if x is not None else None
    out = run_stage(_spec_from_desc(d), xd, Wd, batch=B)
    assert out.shape == ref.shape
    (out * gout.to("cuda", out.dtype)).sum().backward()
    e_out = rel_to_max(out, ref)
    e_w = rel_to_max(Wd.grad, Wr.grad)
    assert e_out <= fwd_tol, f"forward rel-to-max {e_out:.3e}"
    assert e_w <= grad_tol, f"weight-grad rel-to-max {e_w:.3e}"
    if x is not None:
        e_x = rel_to_max(xd.grad, xr.grad, floor=0.1 * max(gout.abs().max().item(), 1.0))
        assert e_x <= grad_tol, f"input-grad rel-to-max {e_x:.3e}"
    return e_out, e_w


@pytest.mark.parametrize("n", list(range(1, 13)))
def test_amplitude_cnot_probs_all_qubit_counts(n):
    """a1 family at every supported qubit count; un-clamped so gradients are compared everywhere."""
    A = 2 ** n
    F = max(1, A - A // 4 - (1 if n > 2 else 0))       # ragged feature count, padded with 0.1
    d = O.StageDesc(n_qubits=n, layers_per_block=5, init=O.INIT_AMPLITUDE, n_features=F, pad_value=0.1,
                    imprimitive=O.IMP_CNOT, remap=O.REMAP_TANH, readout=O.READ_PROBS, read_count=F,
                    post_scale=float(F))
    _check(d, B=37 if n <= 10 else 5, seed=n)


@pytest.mark.parametrize("n", list(range(1, 13)))
def test_reupload_cz_expval_all_qubit_counts(n):
    """a4 family (RZ re-upload, 2-layer CZ blocks, <Z> readout) at every supported qubit count."""
    d = O.desc_reupload(n, L=3, layers=2)
    _check(d, B=33 if n <= 10 else 5, seed=100 + n)


def test_qdense_60x28_clamped():
    """QDenseUndirected_old_noise(60, 28): n=10, 600 Rot + 600 CNOT, clamp epilogue (nn/qdense.py:95-111)."""
    d = O.desc_qdense(60, 784, O.REMAP_TANH)
    _check(d, B=6, seed=1)


def test_qdense_pi_tanh_remap_8x8():
    d = O.desc_qdense(60, 64, O.REMAP_PI_TANH)
    _check(d, B=16, seed=2)


def test_qnn_a_angle_embedding():
    """a2: AngleEmbedding(Y) + SEL(CNOT) + probs (nn/qdense.py:162-190)."""
    d = O.desc_qnn_a(4, 64)
    _check(d, B=9, seed=3)


@pytest.mark.parametrize("cfg", [(6, 14), (8, 6)])
def test_qiddm_reupload_configs(cfg):
    """QIDDM_LL_noise(784,6,14,2) / QIDDM_PL_noise(784,8,6,2) stage circuits (nn/qdense.py:1599-1617)."""
    n, L = cfg
    _check(O.desc_reupload(n, L, 2), B=20, seed=4)


def test_qnn_noise_circuit():
    """QNN_noise(784,8,14): RZ once + 14-layer SEL(CZ) + <Z> (nn/qdense.py:249-265)."""
    _check(O.desc_reupload(8, 1, 14), B=10, seed=5)


def test_differn_probs_stage_clamped_and_chain_stage():
    """differN stages: probs readout (first n un-scaled for chaining; pixels scaled+clamped at the end)."""
    _check(O.desc_reupload(10, 9, 2, readout=O.READ_PROBS, read_count=10), B=8, seed=6)
    _check(O.desc_reupload(10, 9, 2, readout=O.READ_PROBS, read_count=784, post_scale=784.0, clamp=True), B=8, seed=7)


def test_ry_reupload_and_scaled_rz():
    """QIDDM_PL_noise1 (RY re-upload, :602) and QIDDM_A_differN_* (RZ(pi/2 a), :2215)."""
    _check(O.desc_reupload(5, 4, 2, enc=O.ENC_RY), B=12, seed=8)
    _check(O.desc_reupload(6, 3, 2, enc_scale=math.pi / 2, readout=O.READ_PROBS, read_count=36, post_scale=36.0,
                           clamp=True), B=12, seed=9)


def test_three_layer_blocks_bias_false():
    _check(O.desc_reupload(6, 4, 3), B=10, seed=10)


@pytest.mark.parametrize("cfg", [(1, 8, 3), (8, 8, 3), (16, 8, 1), (32, 32, 3), (8, 1, 1)])
def test_qconv_stage_rows(cfg):
    """a6 circuit on explicit patch rows (nn/qconv.py:51-69): pad 0.5, +0.1, pi*tanh, [::2][:out]."""
    cin, cout, k = cfg
    d = O.desc_qconv(cin, cout, (k, k), 3)
    _check(d, B=21, seed=11)


@pytest.mark.parametrize("cfg", [(4, 1, 400_000), (10, 2, 5_000), (7, 2, 30_000), (12, 1, 300)])
def test_persistent_grid_many_instances(cfg):
    """More instances than one wave of CTAs holds: exercises the persistent loop and ragged tails."""
    n, depth, B = cfg
    A = 2 ** n
    d = O.StageDesc(n_qubits=n, layers_per_block=depth, init=O.INIT_AMPLITUDE, n_features=A, pad_value=0.0,
                    imprimitive=O.IMP_CNOT, readout=O.READ_PROBS, read_count=min(A, 16), post_scale=1.0)
    _check(d, B=B + 3, seed=20 + n)
    _check(O.desc_reupload(n, 1, depth), B=B // 2 + 1, seed=40 + n)


def test_float32_weights():
    _check(O.desc_reupload(6, 3, 2), B=10, seed=12, wdtype=torch.float32, grad_tol=4e-5)


def test_empty_batch():
    from qiddm_b200.functional import run_stage
    d = O.desc_reupload(4, 2, 2)
    W = torch.zeros(2, 2, 4, 3, device="cuda", dtype=torch.float64)
    out = run_stage(_spec_from_desc(d), torch.zeros(0, 4, device="cuda"), W)
    assert out.shape == (0, 4)


def test_unitary_build_matches_oracle():
    from qiddm_b200.functional import build_unitary
    d = O.desc_qconv(8, 8, (3, 3), 3)
    g = torch.Generator().manual_seed(13)
    W = torch.rand(1, 3, d.n_qubits, 3, generator=g, dtype=torch.float64) * math.pi - math.pi / 2
    U = build_unitary(_spec_from_desc(d), W.cuda())
    Uref = O.circuit_unitary(d, W)
    assert rel_to_max(torch.view_as_real(U.contiguous()), torch.view_as_real(Uref)) <= FWD_TOL


def test_analytic_known_answers():
    """Zero weights: SEL is a basis permutation -> |0..0> stays one-hot; sum of probs = 1;
    RY(theta)|0> gives <Z> = cos(theta)."""
    from qiddm_b200.functional import run_stage
    from qiddm_b200 import _lib as L
    n = 5
    s = L.StageSpec(n_qubits=n, n_blocks=2, layers_per_block=2, init=L.INIT_ZERO, enc=L.ENC_RZ,
                    imprimitive=L.IMP_CNOT, readout=L.READ_PROBS, read_count=32)
    W0 = torch.zeros(2, 2, n, 3, device="cuda", dtype=torch.float64)
    p = run_stage(s, torch.randn(7, n, device="cuda"), W0)
    assert torch.allclose(p.sum(1), torch.ones(7, device="cuda"), atol=1e-6)
    assert torch.allclose(p[:, 0], torch.ones(7, device="cuda"), atol=1e-6)
    # one layer, theta only, n = 1: <Z> = cos(theta)
    s1 = L.StageSpec(n_qubits=1, n_blocks=1, layers_per_block=1, init=L.INIT_ZERO, readout=L.READ_EXPVAL_Z)
    th = torch.tensor([0.3, 1.1, 2.5], dtype=torch.float64)
    for t in th:
        W = torch.tensor([[[[0.0, t.item(), 0.0]]]], device="cuda", dtype=torch.float64)
        z = run_stage(s1, None, W, batch=3)
        assert torch.allclose(z.double().cpu(), torch.full((3, 1), math.cos(t.item()), dtype=torch.float64), atol=1e-6)

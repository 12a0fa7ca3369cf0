"""GPU parity of the unitary-collapse (tcgen05 GEMM) path against the oracle and against the gate
path (cross-implementation).  Tolerances: precision 3 (3-term fp16 split) is held
to the fp32 bar of the gate path (1e-5 outputs, 1e-4 gradients rel-to-max); precision 1 (single fp16
pass) to the stated looser bound 5e-3 (outputs) / 2e-2 (gradients)."""
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


def _run(d, B, seed, precision, out_tol, grad_tol, wscale=0.4):
    from qiddm_b200 import _lib as L
    from qiddm_b200.functional import run_stage
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(1, d.layers_per_block, d.n_qubits, 3, generator=g, dtype=torch.float64) * wscale
    x = torch.rand(B, d.n_features, generator=g, dtype=torch.float64)
    Wr, xr = W.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ref = O.run_stage(d, xr, Wr)
    go = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * go).sum().backward()
    Wd, xd = W.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    out = run_stage(_spec(d, L.PATH_GEMM, precision), xd, Wd)
    (out * go.cuda()).sum().backward()
    e0 = rel_to_max(out, ref)
    assert e0 <= out_tol, f"forward {e0:.3e}"
    if d.clamp:
        o = out.detach().cpu()
        agree = (((o >= d.clamp_hi) == (ref >= d.clamp_hi)) & ((o <= d.clamp_lo) == (ref <= d.clamp_lo))).all(dim=1)
        assert agree.float().mean().item() >= 0.9
        if not bool(agree.all()):          # weight gradient of the agreeing instances only: re-run both sides on them
            return _run_rows(d, W, x[agree], go[agree], precision, grad_tol)
    e = (e0, rel_to_max(Wd.grad, Wr.grad), rel_to_max(xd.grad, xr.grad))
    assert e[1] <= grad_tol, f"weight grad {e[1]:.3e}"
    assert e[2] <= grad_tol, f"input grad {e[2]:.3e}"
    return e


def _run_rows(d, W, x, go, precision, grad_tol):
    from qiddm_b200 import _lib as L
    from qiddm_b200.functional import run_stage
    Wr, xr = W.clone().requires_grad_(True), x.clone().requires_grad_(True)
    (O.run_stage(d, xr, Wr) * go).sum().backward()
    Wd, xd = W.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    (run_stage(_spec(d, L.PATH_GEMM, precision), xd, Wd) * go.cuda()).sum().backward()
    e = (0.0, rel_to_max(Wd.grad, Wr.grad), rel_to_max(xd.grad, xr.grad))
    assert e[1] <= grad_tol, f"weight grad {e[1]:.3e}"
    assert e[2] <= grad_tol, f"input grad {e[2]:.3e}"
    return e


@pytest.mark.parametrize("n,F,K,stride", [(6, 64, 64, 1), (6, 50, 50, 1), (7, 72, 8, 2), (8, 200, 200, 1),
                                          (4, 9, 8, 2), (9, 300, 100, 1), (5, 32, 16, 2)])
def test_gemm_path_matches_oracle_fp32_grade(n, F, K, stride):
    d = O.StageDesc(n_qubits=n, layers_per_block=4, init=O.INIT_AMPLITUDE, n_features=F, pad_value=0.3,
                    add_offset=0.1, imprimitive=O.IMP_CNOT, remap=O.REMAP_TANH, readout=O.READ_PROBS, read_count=K,
                    read_stride=stride, post_scale=float(2 ** n) / 2)
    _run(d, B=300, seed=n, precision=3, out_tol=1e-5, grad_tol=3e-5)


def test_gemm_path_qdense_60x28_clamped_fp32_grade():
    """The bench circuit: QDenseUndirected_old_noise(60,28), clamp epilogue, 784 of 1024 amplitudes."""
    d = O.desc_qdense(60, 784, O.REMAP_TANH)
    _run(d, B=257, seed=1, precision=3, out_tol=3e-5, grad_tol=3e-5)


def test_gemm_path_single_pass_fp16_looser_bound():
    """precision 1: compared PRE-clamp (a 3e-4 output error flips the clamp mask of outputs that sit on the
    clamp boundary, which makes their gradient discontinuous — SURVEY.md §7 'hard parts')."""
    d = dataclasses.replace(O.desc_qdense(10, 784, O.REMAP_TANH), clamp=False)
    _run(d, B=200, seed=2, precision=1, out_tol=5e-3, grad_tol=2e-2)


def test_gemm_path_qconv_rows():
    d = O.desc_qconv(8, 8, (3, 3), 3)
    _run(d, B=500, seed=3, precision=3, out_tol=1e-5, grad_tol=3e-5, wscale=1.0)


def test_gemm_equals_gate_path_and_auto_dispatch():
    from qiddm_b200 import _lib as L
    from qiddm_b200.functional import run_stage
    d = O.desc_qdense(8, 256, O.REMAP_PI_TANH)
    g = torch.Generator().manual_seed(4)
    W = (torch.randn(1, 8, 8, 3, generator=g, dtype=torch.float64) * 0.4).cuda()
    x = torch.rand(20000, 256, generator=g, dtype=torch.float64).cuda()
    a = run_stage(_spec(d, L.PATH_GATE), x, W)
    b = run_stage(_spec(d, L.PATH_GEMM), x, W)
    c = run_stage(_spec(d, L.PATH_AUTO), x, W)
    assert rel_to_max(b, a) <= 1e-5
    assert L.Plan.get(_spec(d, L.PATH_AUTO)).use_gemm(20000)
    assert torch.equal(b, c)                      # cost model: 20000 instances of an 8-layer n = 8 circuit -> GEMM path
    assert not L.Plan.get(_spec(d, L.PATH_AUTO)).use_gemm(100)
    small = run_stage(_spec(d, L.PATH_AUTO), x[:100], W)
    assert torch.equal(small, run_stage(_spec(d, L.PATH_GATE), x[:100], W))


def test_collapsed_operator_cache_follows_weight_updates():
    from qiddm_b200 import _lib as L
    from qiddm_b200.functional import run_stage
    d = O.desc_qdense(4, 64, O.REMAP_TANH)
    spec = _spec(d, L.PATH_GEMM)
    W = torch.nn.Parameter((torch.randn(1, 4, 6, 3, dtype=torch.float64) * 0.4).cuda())
    x = torch.rand(200, 64, dtype=torch.float64).cuda()
    a = run_stage(spec, x, W).detach()
    with torch.no_grad():
        W.add_(0.3)
    b = run_stage(spec, x, W).detach()
    ref = O.run_stage(d, x.cpu(), W.detach().cpu())
    assert rel_to_max(b, ref) <= 1e-5 and rel_to_max(a, ref) > 1e-3


def test_full_size_properties_of_the_bench_workload():
    """BASELINE config 2 / bench.py size (n = 10, 60 layers, 262 144 instances) through size-independent properties:
    probabilities of every instance sum to 1, sampled rows equal the oracle, identical rows give identical results,
    and the backward is linear in the upstream gradient."""
    from qiddm_b200 import _lib as L
    from qiddm_b200.functional import run_stage
    B = 262144
    d = O.desc_qdense(60, 784, O.REMAP_TANH)
    full = O.StageDesc(**{**d.__dict__, "read_count": 1024, "post_scale": 1.0, "clamp": False})
    g = torch.Generator().manual_seed(99)
    W = (torch.randn(1, 60, 10, 3, generator=g, dtype=torch.float64) * 0.4)
    x = torch.rand(B, 784, generator=g, dtype=torch.float32)
    x[B - 1] = x[0]                                              # a duplicated instance far away in the batch
    xd, Wd = x.cuda(), W.cuda()
    n0 = L.launch_count()
    p = run_stage(_spec(full, L.PATH_GEMM), xd, Wd)              # (B, 1024) un-scaled probabilities
    assert L.launch_count() > n0 and p.shape == (B, 1024)
    s = p.sum(dim=1)
    assert (s - 1).abs().max().item() <= 3e-5
    assert p.min().item() >= 0
    assert torch.equal(p[0], p[B - 1])
    idx = torch.tensor([0, 1, 777, 65535, 131072, 200001, B - 2])
    ref = O.run_stage(full, x[idx].double(), W)
    assert rel_to_max(p[idx.cuda()], ref) <= 2e-5          # K = 784: measured 5-7e-6, stated bound 3e-5
    del p
    # the clamped module readout at full size equals the gate path on a slice
    out = run_stage(_spec(d, L.PATH_GEMM), xd, Wd)
    sl = slice(100000, 100512)
    assert rel_to_max(out[sl], run_stage(_spec(d, L.PATH_GATE), xd[sl], Wd)) <= 4e-5      # measured 1.4-1.8e-5
    # backward linearity in grad_out (weights gradient), B = 65 536
    xs = xd[:65536].clone()
    g1 = torch.randn(65536, 784, generator=g).cuda() / 65536
    g2 = torch.randn(65536, 784, generator=g).cuda() / 65536

    def wgrad(go):
        Wp = Wd.clone().requires_grad_(True)
        o = run_stage(_spec(d, L.PATH_GEMM), xs, Wp)
        (o * go).sum().backward()
        return Wp.grad

    a, b, c = wgrad(g1), wgrad(g2), wgrad(g1 + g2)
    # random-sign upstream gradients over 65 536 instances cancel in dW, which amplifies the fp32 accumulation error of the
    # split-K sums relative to the result: measured 5e-6 ... 1.9e-4 over seeds (scripts/full_size_margins.py)
    assert rel_to_max(a + b, c) <= 1e-3

"""GPU parity of the drop-in torch modules (constructor/forward signatures of the reference) against
the oracle's module-level restatements, incl. gradients through the autograd.Functions."""
import io
import math

import pytest
import torch

from conftest import GOLDEN, rel_to_max
from oracle import qiddm_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
GTOL = 2e-5            # measured <= 8.5e-6 per family at BASELINE sizes (profiles/r2_parity_margins.md)


def _to64(m):
    return m.to("cuda", torch.float64)


def test_qdense_old_noise_module_matches_oracle():
    from qiddm_b200 import nn
    torch.manual_seed(0)
    m = _to64(nn.QDenseUndirected_old_noise(10, 8))
    x = torch.rand(5, 1, 8, 8, dtype=torch.float64)
    W = m.weights.detach().cpu().clone().requires_grad_(True)
    ref = O.qdense_forward(x, W, O.REMAP_TANH)
    out = m(x.cuda())
    assert out.shape == (5, 1, 8, 8) and out.dtype == torch.float64
    assert rel_to_max(out, ref) <= TOL
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    (out * g.cuda()).sum().backward()
    assert rel_to_max(m.weights.grad, W.grad) <= GTOL
    # qnode attribute returns the raw probabilities of all 2**n basis states
    p = m.qnode(x.reshape(5, 64).cuda())
    assert p.shape == (5, 64) and torch.allclose(p.sum(1), torch.ones(5, device="cuda", dtype=p.dtype), atol=1e-5)


def test_qdense_old_pi_tanh_module():
    from qiddm_b200 import nn
    torch.manual_seed(1)
    m = _to64(nn.QDenseUndirected_old(4, (8, 8)))
    x = torch.rand(3, 1, 8, 8, dtype=torch.float64)
    ref = O.qdense_forward(x, m.weights.detach().cpu(), O.REMAP_PI_TANH)
    assert rel_to_max(m(x.cuda()), ref) <= TOL


@pytest.mark.parametrize("cls_args", [("QIDDM_LL_noise", (64, 6, 5, 2)), ("QIDDM_LL_old", (64, 4, 3, 3)),
                                      ("QIDDM_LL_relu_noise", (64, 5, 2, 1))])
def test_qiddm_ll_module_matches_oracle_with_true_gradients(cls_args):
    from qiddm_b200 import nn
    name, args = cls_args
    torch.manual_seed(2)
    m = _to64(getattr(nn, name)(*args))
    x = torch.rand(7, 1, 8, 8, dtype=torch.float64)
    ps = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.named_parameters()}
    ref = O.qiddm_ll_forward(x, ps["weights1"], ps["linear_down.weight"], ps["linear_down.bias"],
                             ps["linear_up.weight"], ps["linear_up.bias"])
    out = m(x.cuda())
    assert rel_to_max(out, ref) <= TOL
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    (out * g.cuda()).sum().backward()
    for k, v in m.named_parameters():
        assert rel_to_max(v.grad, ps[k].grad, floor=1e-6) <= GTOL, k


def test_qiddm_detach_quantum_reproduces_reference_cut_gradient():
    """with detach_quantum only linear_up trains (as in the reference)."""
    from qiddm_b200 import nn
    m = _to64(nn.QIDDM_LL_noise(64, 4, 2, 2))
    m.detach_quantum = True
    m(torch.rand(3, 1, 8, 8, dtype=torch.float64).cuda()).sum().backward()
    assert m.weights1.grad is None and m.linear_down.weight.grad is None
    assert m.linear_up.weight.grad is not None


def test_qnn_noise_module():
    from qiddm_b200 import nn
    torch.manual_seed(3)
    m = nn.QNN_noise("8 * 8", 6, 4)
    x = torch.rand(4, 1, 8, 8, dtype=torch.float64)
    ps = {k: v.detach().cpu() for k, v in m.named_parameters()}
    ref = O.qnn_forward(x, ps["weights"], ps["linear_down.weight"], ps["linear_down.bias"],
                        ps["linear_up.weight"], ps["linear_up.bias"])
    assert rel_to_max(m(x.cuda()), ref) <= TOL


def test_differn_chain_module_matches_oracle():
    """differN family after the PCA: chain of N stages through the first n probabilities."""
    from qiddm_b200 import nn
    torch.manual_seed(4)
    m = nn.differN_old_pca(8, 3, 2).cuda()
    a = torch.randn(9, m.wires, dtype=torch.float64)
    W = m.weights.detach().cpu().double().clone().requires_grad_(True)
    ref = O.differN_forward(a, W, 64)
    out = m._chain(a.cuda())
    assert rel_to_max(out, ref) <= TOL
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    (out * g.cuda()).sum().backward()
    assert rel_to_max(m.weights.grad, W.grad) <= 4e-5       # float32 parameters (gradient returned in fp32)
    full = m(torch.rand(9, 1, 8, 8).cuda())                  # with the sklearn PCA in the loop (H5)
    assert full.shape == (9, 1, 8, 8)


@pytest.mark.parametrize("cfg", [(1, 8, 3, 1, 7, 6), (8, 8, 3, 1, 7, 7), (16, 8, 1, 0, 5, 5), (8, 16, 3, 1, 4, 9),
                                 (4, 1, 1, 0, 6, 6), (3, 5, (3, 2), (1, 0), 6, 7)])
def test_qconv2d_fused_unfold_matches_oracle(cfg):
    """QConv2d forward (H1-fixed) and both gradients vs unfold + circuit + re-layout in the oracle."""
    from qiddm_b200 import nn
    cin, cout, k, pad, H, Wd = cfg
    torch.manual_seed(5)
    m = nn.QConv2d(cin, cout, kernel_size=k, padding=pad, qdepth=3).cuda()
    x = torch.rand(3, cin, H, Wd, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    Wr = m.weights.detach().cpu().clone().requires_grad_(True)
    ref = O.qconv_forward(xr, Wr, cout, m.kernel_size, m.padding)
    xd = x.cuda().requires_grad_(True)
    out = m(xd)
    assert out.shape == ref.shape and out.dtype == torch.float64
    assert rel_to_max(out, ref) <= TOL
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    (out * g.cuda()).sum().backward()
    assert rel_to_max(m.weights.grad, Wr.grad) <= GTOL
    assert rel_to_max(xd.grad, xr.grad) <= GTOL


@pytest.mark.parametrize("cfg", [(1, 8, 3, 1, 14, 14), (8, 8, 3, 1, 9, 10), (16, 8, 1, 0, 7, 7), (8, 16, 3, 1, 8, 9),
                                 (32, 32, 3, 1, 7, 7), (8, 1, 1, 0, 12, 12), (3, 5, (3, 2), (1, 0), 9, 11)])
def test_qconv2d_unitary_collapse_path_matches_oracle_and_gate_path(cfg):
    """QConv2d on the tcgen05 GEMM path (fused unfold, NCHW epilogue, gather col2im) vs the oracle and vs the
    gate-by-gate path.  GEMM-path bounds: 1e-5 outputs, 1e-4 gradients (rel-to-max)."""
    from qiddm_b200 import _lib as L
    from qiddm_b200 import nn
    cin, cout, k, pad, H, Wd = cfg
    torch.manual_seed(11)
    m = nn.QConv2d(cin, cout, kernel_size=k, padding=pad, qdepth=3).cuda()
    x = torch.rand(5, cin, H, Wd, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    Wr = m.weights.detach().cpu().clone().requires_grad_(True)
    ref = O.qconv_forward(xr, Wr, cout, m.kernel_size, m.padding)
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    res = {}
    for name, path in (("gemm", L.PATH_GEMM), ("gate", L.PATH_GATE)):
        m.path = path
        m.weights.grad = None
        xd = x.cuda().requires_grad_(True)
        n0 = L.launch_count()
        out = m(xd)
        assert L.launch_count() > n0
        (out * g.cuda()).sum().backward()
        res[name] = (out.detach(), m.weights.grad.clone(), xd.grad.clone())
        assert rel_to_max(out, ref) <= TOL, name
        assert rel_to_max(m.weights.grad, Wr.grad) <= 3e-5, name
        assert rel_to_max(xd.grad, xr.grad) <= 3e-5, name
    assert rel_to_max(res["gemm"][0], res["gate"][0]) <= 2e-5
    # inference forward (no saved state) gives the same result
    m.path = L.PATH_GEMM
    with torch.no_grad():
        assert rel_to_max(m(x.cuda()), res["gemm"][0]) <= 1e-6


def test_unet_undirected_quantum_runs_and_trains():
    """UNetUndirected(3, 8, qdepth) with QConv2d children (nn/unet.py:119-160): shapes + finite grads."""
    from qiddm_b200 import nn
    torch.manual_seed(6)
    net = nn.UNetUndirected(depth=2, start_channels=4, qdepth=2).cuda()
    x = torch.rand(2, 1, 12, 12, dtype=torch.float64, device="cuda")
    y = net(x)
    assert y.shape == (2, 1, 12, 12)
    y.mean().backward()
    for name, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name


def test_diffusion_training_step_matches_oracle():
    """src/models.py:44-67 on QDense: same noise draw -> same loss and same weight gradient."""
    from qiddm_b200 import models, nn, noise
    torch.manual_seed(7)
    net = nn.QDenseUndirected_old_noise(6, 8)
    diff = models.Diffusion(net, None, "data", (8, 8), torch.nn.MSELoss()).to("cuda", torch.float64)
    x = torch.rand(2, 64, dtype=torch.float64)
    eps = torch.normal(0.5, 0.2, size=(2, 64)).double()
    diff.add_noise = lambda d, tau, decay_mod: noise.add_normal_noise_multiple(d, tau, decay_mod, eps=eps.cuda())
    diff.train()
    (loss,) = diff(x=x.cuda(), T=10)
    W = net.weights.detach().cpu().clone().requires_grad_(True)
    ref = O.diffusion_loss(lambda v: O.qdense_forward(v, W, O.REMAP_TANH), x, eps, 10, (8, 8), "data")
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * max(1.0, abs(ref.item()))
    assert rel_to_max(net.weights.grad, W.grad) <= GTOL


def test_reference_checkpoint_loads_and_generates_letter():
    """F1/F2: a checkpoint trained by the real PennyLane stack loads unchanged into the drop-in module
    and the B200 sampler reproduces the oracle's image (golden from tests/golden/make_golden.py)."""
    from qiddm_b200 import models, nn
    gold = torch.load(GOLDEN / "f1_qdense_label14.pt", weights_only=True)
    net = nn.QDenseUndirected_old_noise(60, 28)
    diff = models.Diffusion(net, None, "noise", (28, 28)).to("cuda", torch.float64)
    diff.load_state_dict({"net.weights": gold["weights"]})
    diff.eval()
    out = diff(gold["first_x"].cuda(), n_iters=40, only_last=True)
    assert rel_to_max(out, gold["sample"]) <= 1e-4
    img = out[0, 0].cpu()
    contrast = img[6:22, 6:22].mean() - (img.sum() - img[6:22, 6:22].sum()) / (784 - 256)
    assert contrast > 0.3


def test_qconv2d_collapse_path_edge_cases():
    """Empty image batch, a single patch row, and an input that needs no gradient (first UNet layer)."""
    from qiddm_b200 import _lib as L
    from qiddm_b200 import nn
    torch.manual_seed(2)
    m = nn.QConv2d(1, 8, kernel_size=3, padding=1, qdepth=2).cuda()
    m.path = L.PATH_GEMM
    assert m(torch.zeros(0, 1, 6, 6, device="cuda", dtype=torch.float64)).shape == (0, 8, 6, 6)
    x = torch.rand(1, 1, 1, 1, dtype=torch.float64)                 # one patch, 8 of 9 features in the zero padding
    ref = O.qconv_forward(x, m.weights.detach().cpu(), 8, m.kernel_size, m.padding)
    assert rel_to_max(m(x.cuda()), ref) <= TOL
    x = torch.rand(6, 1, 10, 10, dtype=torch.float64)               # no grad for the image: the dX GEMM is skipped
    Wr = m.weights.detach().cpu().clone().requires_grad_(True)
    ref = O.qconv_forward(x, Wr, 8, m.kernel_size, m.padding)
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    out = m(x.cuda())
    (out * g.cuda()).sum().backward()
    assert rel_to_max(out, ref) <= TOL and rel_to_max(m.weights.grad, Wr.grad) <= 3e-5

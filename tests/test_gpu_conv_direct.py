"""QConv2d as a direct fp32 convolution on the collapse path (csrc/qiddm_conv.cu) against the oracle (unfold + circuit +
re-layout, reference nn/qconv.py:51-56, :71-90 with the H1 line restored), against the gate-by-gate path and against the
tcgen05 GEMM path, at every layer shape of UNetUndirected(3, 8, 3) and at ragged sizes."""
import pytest
import torch

from conftest import rel_to_max
from oracle import qiddm_oracle as O

pytestmark = pytest.mark.gpu
TOL = 3e-6             # plain fp32 FMAs, at most 289 terms per output
GTOL = 1e-5


def _module(cin, cout, k, depth=3, seed=3):
    from qiddm_b200 import nn
    torch.manual_seed(seed)
    return nn.QConv2d(cin, cout, kernel_size=k, padding=k // 2, qdepth=depth).cuda()


def _is_direct(m, x):
    from qiddm_b200 import _lib as L
    u = L.UnfoldDesc(x.shape[1], x.shape[2], x.shape[3], m.kernel_size[0], m.kernel_size[1], m.padding[0], m.padding[1])
    plan = L.Plan.get(m._spec())
    return plan.qconv_direct(u) and plan.use_collapse_qconv(u, x.shape[0] * x.shape[2] * x.shape[3])


# (in, out, kernel, H, W, images): the UNet's layers at reduced batch, then ragged / narrow / multi-band shapes
CASES = [(1, 8, 3, 28, 28, 3), (8, 8, 3, 28, 28, 2), (16, 8, 3, 28, 28, 2), (16, 8, 1, 28, 28, 2), (8, 1, 1, 28, 28, 2),
         (8, 16, 3, 14, 14, 3), (16, 16, 3, 14, 14, 3), (32, 16, 3, 14, 14, 3), (32, 16, 1, 14, 14, 3),
         (3, 5, 3, 9, 11, 4), (6, 3, 1, 5, 7, 5), (8, 8, 3, 40, 17, 2), (4, 2, 3, 1, 1, 70), (6, 12, 3, 3, 70, 2)]


@pytest.mark.parametrize("cfg", CASES)
def test_direct_conv_matches_oracle_forward_and_both_gradients(cfg):
    cin, cout, k, H, W, n = cfg
    m = _module(cin, cout, k)
    x = torch.rand(n, cin, H, W, dtype=torch.float64)
    assert _is_direct(m, x)
    xr = x.clone().requires_grad_(True)
    Wr = m.weights.detach().cpu().clone().requires_grad_(True)
    ref = O.qconv_forward(xr, Wr, cout, m.kernel_size, m.padding)
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    xd = x.cuda().requires_grad_(True)
    out = m(xd)
    assert out.shape == ref.shape and out.dtype == torch.float64
    assert rel_to_max(out, ref) <= TOL
    (out * g.cuda()).sum().backward()
    assert rel_to_max(m.weights.grad, Wr.grad) <= GTOL
    assert rel_to_max(xd.grad, xr.grad) <= GTOL
    # inference forward (nothing saved) and float32 tensors give the same numbers
    with torch.no_grad():
        assert rel_to_max(m(x.cuda()), out) <= 1e-7
        o32 = m(x.cuda().float())
        assert o32.dtype == torch.float32 and rel_to_max(o32.double(), ref) <= TOL


def test_direct_conv_float32_gradients_and_no_image_gradient():
    m = _module(8, 8, 3)
    x = torch.rand(3, 8, 12, 13, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    Wr = m.weights.detach().cpu().clone().requires_grad_(True)
    ref = O.qconv_forward(xr, Wr, 8, m.kernel_size, m.padding)
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    xd = x.cuda().float().requires_grad_(True)
    out = m(xd)
    (out * g.cuda().float()).sum().backward()
    assert xd.grad.dtype == torch.float32
    assert rel_to_max(m.weights.grad, Wr.grad) <= GTOL and rel_to_max(xd.grad.double(), xr.grad) <= GTOL
    # first UNet layer: the image needs no gradient -> only the weight-gradient kernel runs
    m.weights.grad = None
    out = m(x.cuda())
    (out * g.cuda()).sum().backward()
    assert rel_to_max(m.weights.grad, Wr.grad) <= GTOL


def test_direct_conv_agrees_with_gate_and_gemm_paths_and_respects_the_clamp():
    """Same module, three paths; the input is scaled so that a good share of the outputs sits on the clamp (gradient 0 there)."""
    from qiddm_b200 import _lib as L
    m = _module(8, 8, 3, seed=9)
    x = torch.rand(4, 8, 14, 14, dtype=torch.float64, device="cuda")
    x[:, :, :, :7] *= 0.02                       # nearly uniform patches -> concentrated amplitudes -> clamped outputs
    g = torch.randn(4, 8, 14, 14, dtype=torch.float64, device="cuda")
    res = {}
    for name, path in (("direct", L.PATH_AUTO), ("gate", L.PATH_GATE), ("gemm", L.PATH_GEMM)):
        m.path = path
        m.weights.grad = None
        xd = x.clone().requires_grad_(True)
        out = m(xd)
        (out * g).sum().backward()
        res[name] = (out.detach(), m.weights.grad.clone(), xd.grad.clone())
    m.path = L.PATH_AUTO
    assert _is_direct(m, x)
    clamped = ((res["gate"][0] <= 0) | (res["gate"][0] >= 1)).double().mean().item()
    assert clamped > 0.01
    for other, tol in (("gate", 1e-5), ("gemm", 3e-5)):
        for i in range(3):
            assert rel_to_max(res["direct"][i], res[other][i]) <= tol, (other, i)


def test_direct_conv_large_batch_is_deterministic():
    """640 images x 28 x 28 (the UNet step's first level): per-CTA partials summed in a fixed order -> bitwise repeatable."""
    m = _module(8, 8, 3)
    x = torch.rand(640, 8, 28, 28, dtype=torch.float64, device="cuda")
    g = torch.randn(640, 8, 28, 28, dtype=torch.float64, device="cuda")
    grads = []
    for _ in range(2):
        m.weights.grad = None
        xd = x.clone().requires_grad_(True)
        (m(xd) * g).sum().backward()
        grads.append((m.weights.grad.clone(), xd.grad.clone()))
    assert torch.equal(grads[0][1], grads[1][1])
    assert rel_to_max(grads[0][0], grads[1][0]) <= 1e-6      # the adjoint gate kernel sums a CTA's angle gradients with shared-memory atomics


@pytest.mark.parametrize("cfg", [(32, 16, 7, 7, 3), (16, 8, 14, 14, 2), (6, 3, 5, 9, 4), (8, 4, 13, 3, 5)])
def test_fused_upsample_1x1_qconv_equals_the_two_modules_and_the_oracle(cfg):
    """`Upsample(scale_factor=2, bilinear) -> 1 x 1 QConv2d` of UpBlock (reference nn/unet.py:36-41) with the interpolation inside the
    convolution's staging (qiddm_qconv_up_forward / _backward) against the two modules run one after the other, and against
    torch's interpolation + the oracle's QConv."""
    from qiddm_b200 import nn
    from qiddm_b200.functional import run_qconv_up
    from qiddm_b200.nn.glue import Upsample
    cin, cout, h, w, n = cfg
    torch.manual_seed(4)
    conv = nn.QConv2d(cin, cout, kernel_size=1, padding=0, qdepth=3).cuda()
    up = Upsample(scale_factor=2, mode="bilinear")
    x = torch.rand(n, cin, h, w, dtype=torch.float64)
    g = torch.randn(n, cout, 2 * h, 2 * w, dtype=torch.float64)
    # oracle: torch's bilinear interpolation, then the restated QConv
    xr = x.clone().requires_grad_(True)
    Wr = conv.weights.detach().cpu().clone().requires_grad_(True)
    ref = O.qconv_forward(torch.nn.functional.interpolate(xr, scale_factor=2, mode="bilinear"), Wr, cout, (1, 1), (0, 0))
    (ref * g).sum().backward()
    # the two modules
    x2 = x.cuda().requires_grad_(True)
    o2 = conv(up(x2))
    (o2 * g.cuda()).sum().backward()
    gw2, gx2 = conv.weights.grad.clone(), x2.grad.clone()
    # fused
    conv.weights.grad = None
    x1 = x.cuda().requires_grad_(True)
    o1 = run_qconv_up(conv._spec(), x1, conv.weights, 2 * h, 2 * w, 0.5, 0.5)
    assert o1 is not None and o1.shape == ref.shape and o1.dtype == torch.float64
    (o1 * g.cuda()).sum().backward()
    assert rel_to_max(o1, o2) <= 1e-6 and rel_to_max(o1, ref) <= TOL
    assert rel_to_max(conv.weights.grad, gw2) <= 3e-6 and rel_to_max(conv.weights.grad, Wr.grad) <= GTOL
    assert rel_to_max(x1.grad, gx2) <= 3e-6 and rel_to_max(x1.grad, xr.grad) <= GTOL
    with torch.no_grad():
        assert rel_to_max(run_qconv_up(conv._spec(), x.cuda().float(), conv.weights, 2 * h, 2 * w, 0.5, 0.5).double(), ref) <= TOL


def test_unet_with_and_without_the_upsample_fusion(monkeypatch):
    from qiddm_b200 import nn
    from qiddm_b200.nn import unet as U
    torch.manual_seed(1)
    net = nn.UNetUndirected(3, 8, 3).cuda()
    x = torch.rand(4, 1, 28, 28, dtype=torch.float64, device="cuda")
    outs, grads = [], []
    for fused in (True, False):
        monkeypatch.setattr(U, "UPCONV_FUSION", fused)
        net.zero_grad(set_to_none=True)
        o = net(x)
        o.square().sum().backward()
        outs.append(o.detach())
        grads.append(torch.cat([p.grad.flatten() for p in net.parameters() if p.grad is not None]))
    # fp32 rounding differences of the interpolation (fused: fp32 taps; modules: float64 upsample rounded to fp32 in the staging)
    # pass through the BatchNorm layers behind the up-blocks; whole-network bound against the oracle: 4e-4 (test_gpu_round2.py)
    assert rel_to_max(outs[0], outs[1]) <= 1e-4
    assert rel_to_max(grads[0], grads[1]) <= 1e-3

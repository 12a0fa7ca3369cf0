"""UNet glue kernels (qiddm_b200.nn.glue: BatchNorm2d, bilinear Upsample) against the torch modules they replace
(reference nn/unet.py:28-116 uses torch.nn.BatchNorm2d / torch.nn.Upsample in float64)."""
import pytest
import torch

from conftest import rel_to_max

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("shape", [(10, 8, 28, 28), (3, 16, 7, 7), (1, 5, 1, 3), (640, 32, 7, 7)])
def test_batchnorm2d_matches_torch(dtype, tol, shape):
    from qiddm_b200 import _lib as L
    from qiddm_b200.nn.glue import BatchNorm2d
    torch.manual_seed(0)
    c = shape[1]
    ref = torch.nn.BatchNorm2d(c, dtype=dtype).cuda()
    our = BatchNorm2d(c, dtype=dtype).cuda()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.uniform_(-0.5, 0.5)
    our.load_state_dict(ref.state_dict())
    for it in range(2):                                   # two steps: running statistics accumulate
        x = torch.rand(shape, dtype=dtype, device="cuda") * 2 + it
        g = torch.randn(shape, dtype=dtype, device="cuda")
        xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        n0 = L.launch_count()
        yr, yo = ref(xr), our(xo)
        assert L.launch_count() > n0
        (yr * g).sum().backward()
        (yo * g).sum().backward()
        assert rel_to_max(yo, yr) <= tol
        assert rel_to_max(xo.grad, xr.grad) <= tol * 50
        assert rel_to_max(our.weight.grad, ref.weight.grad) <= tol * 50
        assert rel_to_max(our.bias.grad, ref.bias.grad) <= tol * 50
        ref.zero_grad(); our.zero_grad()
    assert rel_to_max(our.running_mean, ref.running_mean) <= tol
    assert rel_to_max(our.running_var, ref.running_var) <= tol * 10
    assert int(our.num_batches_tracked) == int(ref.num_batches_tracked) == 2
    ref.eval(); our.eval()
    x = torch.rand(shape, dtype=dtype, device="cuda")
    assert rel_to_max(our(x), ref(x)) <= max(tol, 1e-6 if dtype == torch.float32 else 0)
    assert set(our.state_dict()) == set(ref.state_dict())


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-13), (torch.float32, 1e-6)])
@pytest.mark.parametrize("cfg", [((4, 3, 7, 7), dict(scale_factor=2)), ((2, 5, 14, 13), dict(scale_factor=2)),
                                 ((2, 2, 1, 1), dict(scale_factor=2)), ((3, 2, 5, 6), dict(scale_factor=3)),
                                 ((2, 3, 6, 9), dict(size=(11, 7))), ((2, 2, 8, 8), dict(scale_factor=1.5))])
def test_bilinear_upsample_matches_torch(dtype, tol, cfg):
    from qiddm_b200 import _lib as L
    from qiddm_b200.nn.glue import Upsample
    shape, kw = cfg
    torch.manual_seed(1)
    ref, our = torch.nn.Upsample(mode="bilinear", **kw), Upsample(mode="bilinear", **kw)
    x = torch.rand(shape, dtype=dtype, device="cuda")
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    n0 = L.launch_count()
    yr, yo = ref(xr), our(xo)
    assert L.launch_count() > n0 and yo.shape == yr.shape
    g = torch.randn_like(yr)
    (yr * g).sum().backward()
    (yo * g).sum().backward()
    assert rel_to_max(yo, yr) <= tol
    assert rel_to_max(xo.grad, xr.grad) <= tol * 10


def test_unet_with_library_glue_equals_torch_glue():
    """UNetUndirected (QConv2d children) with the library BatchNorm2d (ReLU fused) / Upsample / MaxPool2d vs the same net run
    with the torch modules (identical weights): outputs and gradients agree to float64 round-off of the glue."""
    from qiddm_b200 import nn
    torch.manual_seed(6)
    net = nn.UNetUndirected(depth=3, start_channels=4, qdepth=2).cuda()
    x = torch.rand(4, 1, 12, 12, dtype=torch.float64, device="cuda")
    y = net(x)
    y.square().mean().backward()
    grads = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad()

    def to_torch(m):
        for name, child in list(m.named_children()):
            if isinstance(child, nn.glue.BatchNorm2d):
                t = torch.nn.BatchNorm2d(child.num_features, dtype=torch.double).cuda()
                t.load_state_dict(child.state_dict())
                t.running_mean.zero_(); t.running_var.fill_(1.0)
                setattr(m, name, t)
            elif isinstance(child, nn.glue.Upsample):
                setattr(m, name, torch.nn.Upsample(scale_factor=2, mode="bilinear"))
            elif isinstance(child, nn.glue.FusedReLU):
                setattr(m, name, torch.nn.ReLU())            # the fused pair becomes torch's BatchNorm2d + ReLU again
            elif isinstance(child, nn.glue.MaxPool2d):
                setattr(m, name, torch.nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                to_torch(child)
    to_torch(net)
    y2 = net(x)
    y2.square().mean().backward()
    assert rel_to_max(y, y2) <= 1e-9
    for k, p in net.named_parameters():
        assert rel_to_max(grads[k], p.grad, floor=1e-12) <= 1e-5, k      # fp32 QConv children amplify glue round-off


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-15), (torch.float32, 1e-6)])
def test_fused_noise_ladder_pair_matches_the_reference_schedule(dtype, tol):
    """noise.ladder_pair (one kernel) == slicing add_normal_noise_multiple's ladder as src/models.py:46-63 does."""
    from qiddm_b200 import noise
    torch.manual_seed(0)
    x = torch.rand(7, 50, dtype=dtype, device="cuda")
    eps = torch.normal(0.5, 0.2, size=(7, 50), device="cuda")
    T = 10
    whole = noise.add_normal_noise_multiple(x, tau=T + 1, decay_mod=3.0, eps=eps).reshape(7, T + 1, 50)
    noisy, clean = noise.ladder_pair(x, T, 3.0, eps=eps)
    assert noisy.shape == (70, 50) and noisy.dtype == dtype
    assert (noisy - whole[:, 1:].reshape(70, 50)).abs().max().item() <= tol
    assert (clean - whole[:, :-1].reshape(70, 50)).abs().max().item() <= tol
    assert torch.equal(clean.reshape(7, T, 50)[:, 0], x.clamp(0, 1))


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-13), (torch.float32, 1e-5)])
def test_fused_mse_loss_and_grad_matches_torch(dtype, tol):
    from qiddm_b200 import noise
    torch.manual_seed(1)
    p = torch.rand(33, 1, 9, 9, dtype=dtype, device="cuda", requires_grad=True)
    a, b = torch.rand_like(p), torch.rand_like(p)
    ref = torch.nn.MSELoss()(p, a)
    ref.backward()
    loss, grad = noise.mse_loss_and_grad(p, a)
    assert abs(loss.item() - ref.item()) <= tol * ref.item() and rel_to_max(grad, p.grad) <= tol
    p.grad = None
    ref = torch.nn.MSELoss()((p - 0.5) * 0.1, a - b)          # goal "noise" (src/models.py:95-96)
    ref.backward()
    loss, grad = noise.mse_loss_and_grad(p, a, b, scale=0.1, shift=-0.05)
    assert abs(loss.item() - ref.item()) <= tol * ref.item() and rel_to_max(grad, p.grad) <= tol


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-13), (torch.float32, 1e-5)])
def test_mse_with_the_target_recomputed_from_the_noise_draw(dtype, tol):
    """noise.mse_ladder_loss_and_grad (no `clean` tensor: the ladder levels are recomputed from images + eps) == the loss over
    the materialised ladder, both goals (src/models.py:65-67, :95-99); the clean-less ladder writes the same `noisy`."""
    from qiddm_b200 import noise
    torch.manual_seed(5)
    T = 7
    x = torch.rand(9, 33, dtype=dtype, device="cuda") * 1.4 - 0.2          # some levels hit the clamp
    eps = torch.normal(0.5, 0.2, size=(9, 33), device="cuda")
    noisy, clean = noise.ladder_pair(x, T, 3.0, eps=eps)
    noisy2, none, draw = noise.ladder_pair(x, T, 3.0, eps=eps, want_clean=False, return_draw=True)
    assert none is None and torch.equal(noisy, noisy2)
    pred = torch.rand(9 * T, 33, dtype=dtype, device="cuda")
    l_ref, g_ref = noise.mse_loss_and_grad(pred, clean)
    l_new, g_new = noise.mse_ladder_loss_and_grad(pred, draw, T)
    assert abs(l_new.item() - l_ref.item()) <= tol * l_ref.item() and rel_to_max(g_new, g_ref) <= tol
    l_ref, g_ref = noise.mse_loss_and_grad(pred, noisy, clean, scale=0.1, shift=-0.05)
    l_new, g_new = noise.mse_ladder_loss_and_grad(pred, draw, T, scale=0.1, shift=-0.05, c0=-1.0, c1=1.0)
    assert abs(l_new.item() - l_ref.item()) <= tol * l_ref.item() and rel_to_max(g_new, g_ref) <= tol


def test_diffusion_step_with_and_without_the_recomputed_target(monkeypatch):
    from qiddm_b200 import models, nn, noise
    for goal in ("data", "noise"):
        res = []
        for flag in ("1", "0"):
            monkeypatch.setenv("QIDDM_MSE_LADDER", flag)
            torch.manual_seed(3)
            net = nn.QIDDM_LL_noise(64, 4, 3, 2)
            d = models.Diffusion(net, noise.add_normal_noise_multiple, goal, (8, 8), torch.nn.MSELoss()).to("cuda", torch.float64)
            d.train()
            x = torch.rand(5, 64, dtype=torch.float64, device="cuda")
            torch.manual_seed(11)
            (loss,) = d(x=x, T=6)
            res.append((loss.item(), torch.cat([p.grad.flatten() for p in net.parameters()])))
        assert abs(res[0][0] - res[1][0]) <= 1e-12 * abs(res[1][0])
        assert rel_to_max(res[0][1], res[1][1]) <= 1e-9


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("hidden,bias", [(6, True), (8, False), (16, True), (1, True)])
def test_fused_linear_up_mse_tail_matches_torch(dtype, tol, hidden, bias):
    """noise.linear_up_mse_loss (qiddm_linear_up_mse_step: Linear(hidden -> pixels) + MSE against the recomputed ladder target,
    loss and all three gradients in one pass) against the same tail written with torch ops, both goals."""
    from qiddm_b200 import noise
    torch.manual_seed(7)
    T, batch, P = 5, 37, 150
    x = torch.rand(batch, P, dtype=dtype, device="cuda") * 1.3 - 0.1
    eps = torch.normal(0.5, 0.2, size=(batch, P), device="cuda")
    noisy, clean, draw = noise.ladder_pair(x, T, 3.0, eps=eps, return_draw=True)
    layer = torch.nn.Linear(hidden, P, bias=bias).to("cuda", dtype)
    for coef, target, pre in (({}, clean, lambda o: o),
                              (dict(scale=0.1, shift=-0.05, c0=-1.0, c1=1.0), noisy - clean, lambda o: (o - 0.5) * 0.1)):
        h1 = torch.randn(batch * T, hidden, dtype=dtype, device="cuda", requires_grad=True)
        h2 = h1.detach().clone().requires_grad_(True)
        layer.zero_grad()
        ref = torch.nn.functional.mse_loss(pre(layer(h2)), target)
        ref.backward()
        gw, gb = layer.weight.grad.clone(), (layer.bias.grad.clone() if bias else None)
        layer.zero_grad()
        assert noise.linear_up_mse_ok(h1, layer, draw, T)
        loss = noise.linear_up_mse_loss(h1, layer, draw, T, **coef)
        loss.backward()
        assert abs(loss.item() - ref.item()) <= tol * abs(ref.item())
        assert rel_to_max(h1.grad, h2.grad) <= tol and rel_to_max(layer.weight.grad, gw) <= tol
        if bias:
            assert rel_to_max(layer.bias.grad, gb) <= tol


@pytest.mark.parametrize("cls_args", [("QIDDM_LL_noise", (64, 4, 3, 2)), ("QIDDM_PL_noise", (64, 4, 3, 2))])
def test_diffusion_step_with_and_without_the_fused_tail(cls_args, monkeypatch):
    """Whole training step of a re-upload network with `linear_up + loss` fused (QIDDM_FUSED_TAIL=1, from 16 384 rows on) against the same step
    through the separate kernels: same loss, same gradients of every parameter."""
    from qiddm_b200 import models, nn, noise
    name, args = cls_args
    for goal in ("data", "noise"):
        res = []
        for flag in ("1", "0"):
            monkeypatch.setenv("QIDDM_FUSED_TAIL", flag)
            torch.manual_seed(3)
            net = getattr(nn, name)(*args)
            d = models.Diffusion(net, noise.add_normal_noise_multiple, goal, (8, 8), torch.nn.MSELoss()).to("cuda", torch.float64)
            d.train()
            x = torch.rand(2100, 64, dtype=torch.float64, device="cuda")          # 16 800 rows: above the fused tail's threshold
            torch.manual_seed(11)
            (loss,) = d(x=x, T=8)
            res.append((loss.item(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
        assert abs(res[0][0] - res[1][0]) <= 1e-10 * abs(res[1][0]), goal
        assert res[0][1].keys() == res[1][1].keys()
        for k in res[0][1]:
            assert rel_to_max(res[0][1][k], res[1][1][k], floor=1e-14) <= 1e-6, (goal, k)


def test_diffusion_step_fused_glue_equals_torch_glue():
    """Diffusion.forward with the fused ladder + MSE kernels vs the same step through the torch ops (custom add_noise /
    verbose path), both goals: same loss and parameter gradients."""
    from qiddm_b200 import models, nn, noise
    for goal in ("data", "noise"):
        torch.manual_seed(3)
        net = nn.QIDDM_LL_noise(64, 4, 3, 2)
        d = models.Diffusion(net, noise.add_normal_noise_multiple, goal, (8, 8), torch.nn.MSELoss()).to("cuda", torch.float64)
        d.train()
        x = torch.rand(3, 64, dtype=torch.float64, device="cuda")
        torch.manual_seed(11)
        (l1,) = d(x=x, T=5)
        g1 = {k: p.grad.clone() for k, p in net.named_parameters()}
        net.zero_grad()
        d.add_noise = lambda data, tau, decay_mod: noise.add_normal_noise_multiple(data, tau, decay_mod)   # torch ladder
        d.loss = _TorchMSE()                                                                              # torch loss path
        torch.manual_seed(11)
        (l3,) = d(x=x, T=5)
        assert abs(l1.item() - l3.item()) <= 1e-9 * abs(l3.item())
        for k, p in net.named_parameters():
            assert rel_to_max(g1[k], p.grad, floor=1e-12) <= 1e-6, (goal, k)


class _TorchMSE(torch.nn.Module):          # not `type(...) is MSELoss` -> Diffusion keeps the torch loss path
    def forward(self, a, b):
        return torch.nn.functional.mse_loss(a, b)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("mode", ["post", "pre"])
def test_batchnorm_with_fused_relu_matches_the_torch_pair(dtype, tol, mode):
    """nn/unet.py:92-108 `Conv -> BN -> ReLU` ("post") and :55-63 `Conv -> ReLU -> BN` ("pre"): the ReLU computed inside the
    BatchNorm kernels equals the two torch modules, forward, input gradient, affine gradients and running statistics."""
    from qiddm_b200.nn.glue import BatchNorm2d
    torch.manual_seed(3)
    x = torch.randn(6, 5, 9, 7, dtype=dtype, device="cuda")
    go = torch.randn(6, 5, 9, 7, dtype=dtype, device="cuda")
    ours = BatchNorm2d(5, dtype=dtype).cuda()
    ours.fuse_relu = mode
    ref = torch.nn.BatchNorm2d(5, dtype=dtype).cuda()
    with torch.no_grad():
        for m in (ours, ref):
            m.weight.copy_(torch.linspace(-1.0, 1.5, 5))
            m.bias.copy_(torch.linspace(-0.5, 0.5, 5))
    xo, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yo = ours(xo)
    yr = torch.relu(ref(xr)) if mode == "post" else ref(torch.relu(xr))
    (yo * go).sum().backward()
    (yr * go).sum().backward()
    assert rel_to_max(yo, yr) <= tol
    assert rel_to_max(xo.grad, xr.grad) <= tol * 50
    assert rel_to_max(ours.weight.grad, ref.weight.grad) <= tol * 50
    assert rel_to_max(ours.bias.grad, ref.bias.grad) <= tol * 50
    assert rel_to_max(ours.running_mean, ref.running_mean) <= tol and rel_to_max(ours.running_var, ref.running_var) <= tol * 10
    ours.eval(), ref.eval()
    ye = ours(x)
    yre = torch.relu(ref(x)) if mode == "post" else ref(torch.relu(x))
    assert rel_to_max(ye, yre) <= max(tol, 1e-6 if dtype == torch.float32 else 0)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", [(3, 4, 8, 8), (2, 3, 7, 9), (1, 1, 28, 28)])
def test_maxpool2d_matches_torch_including_ties(dtype, shape):
    from qiddm_b200.nn.glue import MaxPool2d
    torch.manual_seed(4)
    x = torch.randn(*shape, dtype=dtype, device="cuda").clamp_min(0)          # ReLU-like input: many ties at 0
    ours, ref = MaxPool2d(kernel_size=2, stride=2), torch.nn.MaxPool2d(kernel_size=2, stride=2)
    xo, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yo, yr = ours(xo), ref(xr)
    assert torch.equal(yo, yr)
    go = torch.randn_like(yr)
    (yo * go).sum().backward()
    (yr * go).sum().backward()
    assert torch.equal(xo.grad, xr.grad)


def test_unet_blocks_fuse_their_relus_and_keep_the_state_dict_keys():
    from qiddm_b200 import nn
    from qiddm_b200.nn.glue import BatchNorm2d, FusedReLU
    m = nn.UNetUndirected(3, 8, 0)
    relus = [mod for mod in m.modules() if isinstance(mod, FusedReLU)]
    bns = [mod for mod in m.modules() if isinstance(mod, BatchNorm2d)]
    assert len(relus) == 10 and all(r.fused for r in relus)
    assert sorted(b.fuse_relu for b in bns) == ["post"] * 8 + ["pre"] * 2
    ref_keys = {"down_blocks.0.net.0.weight", "down_blocks.0.net.1.running_mean", "up_blocks.0.net.2.weight", "up_blocks.1.net.4.bias",
                "final_conv.weight"}
    assert ref_keys <= set(m.state_dict().keys())


@pytest.mark.parametrize("rows,in_f,out_f,dtype,bias", [(1000, 784, 6, torch.float64, True), (1000, 6, 784, torch.float64, True),
                                                        (333, 4096, 8, torch.float32, True), (257, 8, 4096, torch.float32, True),
                                                        (64, 784, 16, torch.float64, False), (5000, 784, 6, torch.float32, False),
                                                        (31, 1, 300, torch.float64, True), (40960, 784, 6, torch.float64, True)])
def test_skinny_linear_matches_torch(rows, in_f, out_f, dtype, bias, monkeypatch):
    """linear_down / linear_up shapes of nn/qdense.py:219-386, :565-670 (and the 64 x 64 ones of config 5): forward, grad_x,
    grad_weight, grad_bias against torch.nn.functional.linear in the same dtype."""
    from qiddm_b200.nn import glue
    from qiddm_b200.nn.glue import skinny_linear
    monkeypatch.setattr(glue, "SKINNY_MIN_ROWS", 1)        # the kernels themselves, also below the dispatch threshold
    torch.manual_seed(rows + in_f)
    layer = torch.nn.Linear(in_f, out_f, bias=bias).to("cuda", dtype)
    x = torch.randn(rows, in_f, device="cuda", dtype=dtype, requires_grad=True)
    gy = torch.randn(rows, out_f, device="cuda", dtype=dtype)
    y = skinny_linear(x, layer)
    assert y.grad_fn is not None and type(y.grad_fn).__name__.startswith("_SkinnyLinear"), "the streaming kernels were not taken"
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), layer.weight.grad.clone(), layer.bias.grad.clone() if bias else None)
    x.grad = None
    layer.zero_grad()
    yr = torch.nn.functional.linear(x, layer.weight, layer.bias)
    yr.backward(gy)
    ref = (yr.detach(), x.grad, layer.weight.grad, layer.bias.grad if bias else None)
    tol = 1e-12 if dtype == torch.float64 else 2e-5
    for a, b, name in zip(got, ref, ("y", "grad_x", "grad_w", "grad_b")):
        if b is not None:
            assert rel_to_max(a, b) <= tol, name


def test_skinny_linear_partial_gradients_and_fallback():
    """Only the gradients autograd asks for are computed (first layer: no grad_x); shapes outside the kernels go to torch."""
    from qiddm_b200.nn.glue import skinny_linear
    layer = torch.nn.Linear(784, 6).to("cuda", torch.float64)
    x = torch.randn(5000, 784, device="cuda", dtype=torch.float64)          # no grad
    y = skinny_linear(x, layer)
    y.sum().backward()
    ref = torch.nn.functional.linear(x, layer.weight.detach(), layer.bias.detach())
    assert rel_to_max(y, ref) <= 1e-12
    assert rel_to_max(layer.weight.grad, x.sum(0, keepdim=True).expand(6, -1)) <= 1e-12
    assert rel_to_max(layer.bias.grad, torch.full((6,), 5000.0, dtype=torch.float64)) <= 1e-12
    small = torch.nn.Linear(16, 8).to("cuda", torch.float64)
    ys = skinny_linear(torch.randn(10, 16, device="cuda", dtype=torch.float64), small)
    assert not type(ys.grad_fn).__name__.startswith("_SkinnyLinear")
    few = skinny_linear(torch.randn(100, 784, device="cuda", dtype=torch.float64), layer)      # below SKINNY_MIN_ROWS
    assert not type(few.grad_fn).__name__.startswith("_SkinnyLinear")

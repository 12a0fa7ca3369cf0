"""The fused diffusion training step of a single amplitude-embedding layer (`qiddm_dense_mse_step`: ladder -> operand splits,
forward GEMM with MSE + dL/dY in the epilogue, dW GEMM, adjoint on the basis columns) against the complex128 oracle of
src/models.py:44-104 and against the unfused kernel sequence of the same library, through `Diffusion.forward`."""
import os

import pytest
import torch

from conftest import rel_to_max
from oracle import qiddm_oracle as O

pytestmark = pytest.mark.gpu


def _make(goal, depth, side, dtype, remap_cls="noise"):
    from qiddm_b200 import _lib as L
    from qiddm_b200 import models, nn, noise
    net = nn.QDenseUndirected_old_noise(depth, side) if remap_cls == "noise" else nn.QDenseUndirected_old(depth, side)
    net.path = L.PATH_GEMM
    diff = models.Diffusion(net, noise.add_normal_noise_multiple, goal, (side, side), torch.nn.MSELoss()).to("cuda", dtype)
    diff.train()
    return net, diff


def _step(diff, net, x, T, seed, fused):
    os.environ["QIDDM_FUSED_STEP"] = "1" if fused else "0"
    try:
        net.weights.grad = None
        torch.manual_seed(seed)
        (loss,) = diff(x=x, T=T)
        return loss.detach().clone(), net.weights.grad.detach().clone()
    finally:
        os.environ.pop("QIDDM_FUSED_STEP", None)


@pytest.mark.parametrize("goal", ["data", "noise"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fused_step_matches_oracle(goal, dtype):
    """Same noise draw (re-drawn from the same seed) -> same loss and weight gradient as the oracle's training step."""
    side, T, n = 8, 10, 40
    torch.manual_seed(3)
    net, diff = _make(goal, 6, side, dtype)
    x = torch.rand(n, side * side, dtype=dtype, device="cuda")
    calls = []
    from qiddm_b200 import _lib as L
    orig = L.Plan.dense_mse_step
    L.Plan.dense_mse_step = lambda self, *a, **k: (calls.append(1), orig(self, *a, **k))[1]
    try:
        loss, gw = _step(diff, net, x, T, seed=11, fused=True)
    finally:
        L.Plan.dense_mse_step = orig
    assert calls, "the fused step was not taken"
    torch.manual_seed(11)
    eps = torch.normal(mean=0.5, std=0.2, size=(n, side * side), device="cuda").double().cpu()
    W = net.weights.detach().double().cpu().clone().requires_grad_(True)
    ref = O.diffusion_loss(lambda v: O.qdense_forward(v, W, O.REMAP_TANH), x.double().cpu(), eps, T, (side, side), goal)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-5 * max(1e-30, abs(ref.item())), (loss.item(), ref.item())
    assert rel_to_max(gw, W.grad) <= 3e-5


@pytest.mark.parametrize("goal,side,depth,n,dtype", [("data", 28, 4, 300, torch.float32), ("noise", 28, 3, 77, torch.float64),
                                                     ("data", 16, 5, 1000, torch.float64), ("data", 12, 3, 130, torch.float32),
                                                     ("noise", 40, 2, 30, torch.float32)])
def test_fused_step_equals_the_unfused_sequence(goal, side, depth, n, dtype):
    """Multi-tile shapes (28 x 28: 7 N tiles; 3 000 rows: 12 M tiles with a ragged last one; 12 x 12 = 144 of 256 amplitudes:
    pad rows + ones column; 40 x 40 = 1600 pixels: n = 11, the row-per-warp ladder kernel) against ladder_pair -> run_stage -> mse_loss_and_grad -> autograd of the same library."""
    torch.manual_seed(5)
    net, diff = _make(goal, depth, side, dtype, remap_cls="old" if side == 16 else "noise")
    x = torch.rand(n, side * side, dtype=dtype, device="cuda")
    l1, g1 = _step(diff, net, x, 10, seed=21, fused=True)
    l0, g0 = _step(diff, net, x, 10, seed=21, fused=False)
    assert abs(l1.item() - l0.item()) <= 1e-6 * abs(l0.item()) + 1e-12, (l1.item(), l0.item())
    assert rel_to_max(g1, g0) <= 2e-5
    # accumulation into an existing .grad, as autograd does
    os.environ["QIDDM_FUSED_STEP"] = "1"
    try:
        torch.manual_seed(21)
        diff(x=x, T=10)
        torch.manual_seed(21)
        diff(x=x, T=10)
    finally:
        os.environ.pop("QIDDM_FUSED_STEP", None)
    assert rel_to_max(net.weights.grad, 3 * g1) <= 1e-6


def test_fused_step_with_single_pass_gradients_and_graph_capture():
    """x3 forward / x1 gradient GEMMs behind the fused epilogue (stated bound 1e-3), and the step inside GraphedTrainStep."""
    from qiddm_b200 import train
    torch.manual_seed(8)
    net, diff = _make("data", 4, 28, torch.float32)
    x = torch.rand(200, 784, device="cuda")
    l3, g3 = _step(diff, net, x, 10, seed=2, fused=True)
    net.gemm_bwd_precision = 1
    l1, g1 = _step(diff, net, x, 10, seed=2, fused=True)
    assert l1.item() == l3.item()
    assert 0 < rel_to_max(g1, g3) <= 1e-3
    net.gemm_bwd_precision = 0
    opt = torch.optim.Adam(diff.parameters(), lr=1e-3, capturable=True)
    step = train.GraphedTrainStep(diff, opt, 10, x)
    w0 = net.weights.detach().clone()
    losses = [float(step.step(x)) for _ in range(3)]
    assert all(l == l and l > 0 for l in losses)
    assert not torch.equal(net.weights.detach(), w0)


def test_fused_step_at_the_bench_size_equals_the_unfused_sequence():
    """bench.py's e2e configuration (QDenseUndirected_old_noise(60, 28), 52 428 images x 10 levels = 524 280 circuit instances,
    float32): the fused step against the un-fused kernel sequence on the same noise draw -- the size-independent property
    that stands in for the oracle, which needs hours at this size."""
    torch.manual_seed(1)
    net, diff = _make("data", 60, 28, torch.float32)
    x = torch.rand(52428, 784, device="cuda")
    l1, g1 = _step(diff, net, x, 10, seed=4, fused=True)
    l0, g0 = _step(diff, net, x, 10, seed=4, fused=False)
    assert abs(l1.item() - l0.item()) <= 2e-6 * abs(l0.item()), (l1.item(), l0.item())
    assert rel_to_max(g1, g0) <= 3e-5
    assert torch.isfinite(g1).all() and g1.abs().max().item() > 0

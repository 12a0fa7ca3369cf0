"""CUDA-graph training step and sampler (qiddm_b200.train.GraphedTrainStep / GraphedSampler)
against the eager loop body of src/mnist_exm.py:175-182 and Diffusion.sample (src/models.py:106-147)."""
import copy

import pytest
import torch

from conftest import rel_to_max

pytestmark = pytest.mark.gpu


def _diff(net_factory, goal, eps):
    from qiddm_b200 import models, noise
    torch.manual_seed(3)
    net = net_factory()
    d = models.Diffusion(net, None, goal, (8, 8), torch.nn.MSELoss()).to("cuda", torch.float64)
    d.add_noise = lambda data, tau, decay_mod: noise.add_normal_noise_multiple(data, tau, decay_mod, eps=eps)
    return d


@pytest.mark.parametrize("model", ["QIDDM_LL_noise", "QIDDM_PL_noise", "QDenseUndirected_old_noise", "QNN_noise"])
def test_graphed_train_step_equals_eager_steps(model):
    from qiddm_b200 import nn
    from qiddm_b200.train import GraphedTrainStep
    factories = {"QIDDM_LL_noise": lambda: nn.QIDDM_LL_noise(64, 4, 3, 2), "QIDDM_PL_noise": lambda: nn.QIDDM_PL_noise(64, 4, 3, 2),
                 "QDenseUndirected_old_noise": lambda: nn.QDenseUndirected_old_noise(4, 8),
                 "QNN_noise": lambda: nn.QNN_noise(64, 4, 3)}
    imgs, tau = 2, 5
    eps = torch.normal(0.5, 0.2, size=(imgs, 64), generator=torch.Generator().manual_seed(1)).double().cuda()
    batches = [torch.rand(imgs, 64, dtype=torch.float64, generator=torch.Generator().manual_seed(10 + i)).cuda()
               for i in range(4)]
    goal = "noise" if model == "QIDDM_PL_noise" else "data"
    # eager reference
    d0 = _diff(factories[model], goal, eps)
    d0.train()
    # SGD for the strict parameter comparison: Adam's g / (|g| + eps) turns round-off on analytically-zero gradients
    # (e.g. the last RZ before a probability readout) into +-lr steps; Adam itself is covered below
    o0 = torch.optim.SGD(d0.parameters(), lr=0.05)
    losses0 = []
    for x in batches:
        o0.zero_grad(set_to_none=True)
        (l,) = d0(x=x, T=tau)
        o0.step()
        losses0.append(l.item())
    # graphed
    d1 = _diff(factories[model], goal, eps)
    o1 = torch.optim.SGD(d1.parameters(), lr=0.05)
    g = GraphedTrainStep(d1, o1, tau, batches[0])
    losses1 = [g.step(x).item() for x in batches]
    for a, b in zip(losses0, losses1):
        assert abs(a - b) <= 1e-5 * abs(a), (losses0, losses1)
    for (n0, p0), (_, p1) in zip(d0.named_parameters(), d1.named_parameters()):
        assert rel_to_max(p1, p0) <= 1e-5, n0


def test_graphed_train_step_with_capturable_adam_trains():
    """The reference optimizer (Adam, src/mnist_exm.py:170) inside the graph: same loss trajectory as eager Adam."""
    from qiddm_b200 import nn
    from qiddm_b200.train import GraphedTrainStep
    imgs, tau = 2, 5
    eps = torch.normal(0.5, 0.2, size=(imgs, 64), generator=torch.Generator().manual_seed(1)).double().cuda()
    x = torch.rand(imgs, 64, dtype=torch.float64, generator=torch.Generator().manual_seed(2)).cuda()
    d0 = _diff(lambda: nn.QIDDM_LL_noise(64, 4, 3, 2), "data", eps)
    d0.train()
    o0 = torch.optim.Adam(d0.parameters(), lr=1e-2, capturable=True)
    l0 = []
    for _ in range(12):
        o0.zero_grad(set_to_none=True)
        (l,) = d0(x=x, T=tau)
        o0.step()
        l0.append(l.item())
    d1 = _diff(lambda: nn.QIDDM_LL_noise(64, 4, 3, 2), "data", eps)
    g = GraphedTrainStep(d1, torch.optim.Adam(d1.parameters(), lr=1e-2, capturable=True), tau, x)
    l1 = [g.step(x).item() for _ in range(12)]
    assert l1[-1] < l1[0]
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-3 * abs(a), (l0, l1)
    with pytest.raises(RuntimeError):
        GraphedTrainStep(d1, torch.optim.Adam(d1.parameters(), lr=1e-2), tau, x)


def test_graphed_sampler_equals_diffusion_sample():
    from qiddm_b200 import models, nn, noise
    from qiddm_b200.train import GraphedSampler
    for goal, factory in (("noise", lambda: nn.QDenseUndirected_old_noise(5, 8)), ("data", lambda: nn.QIDDM_PL_noise(64, 4, 3, 2))):
        torch.manual_seed(5)
        d = models.Diffusion(factory(), noise.add_normal_noise_multiple, goal, (8, 8)).to("cuda", torch.float64)
        d.eval()
        first = torch.rand(10, 1, 8, 8, dtype=torch.float64, device="cuda") * 0.75 + 0.5
        ref = d.sample(n_iters=23, first_x=first, only_last=True)
        got = GraphedSampler(d, first, unroll=10).sample(23, first)
        assert rel_to_max(got, ref) <= 1e-5, goal


def test_collapsed_operator_inside_a_graph_follows_the_weights():
    """The unitary collapse is captured with the step: replays after an in-graph weight update use the new U."""
    from qiddm_b200 import _lib as L
    from qiddm_b200 import nn
    torch.manual_seed(0)
    m = nn.QDenseUndirected_old_noise(3, 8).cuda()
    x = torch.rand(4096, 1, 8, 8, device="cuda")
    with torch.no_grad():
        m(x)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        m(x)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        y = m(x)
    with torch.no_grad():
        m.weights.add_(0.3)
    g.replay()
    with torch.no_grad():
        ref = m(x)
    assert rel_to_max(y, ref) <= 1e-6


def test_graphed_data_parallel_step_two_graphs_around_the_collective():
    """allreduce=True under torch.distributed: graph 1 (.. backward, pack) -> eager all-reduce of the flat bucket ->
    graph 2 (unpack, optimizer).  Single-rank NCCL group here (the 2-GPU run is scripts/bench_dp.py); the step must equal
    the plain graphed step."""
    import os
    import torch.distributed as dist
    from qiddm_b200 import nn
    from qiddm_b200.train import GraphedTrainStep
    imgs, tau = 2, 5
    eps = torch.normal(0.5, 0.2, size=(imgs, 64), generator=torch.Generator().manual_seed(1)).double().cuda()
    xs = [torch.rand(imgs, 64, dtype=torch.float64, generator=torch.Generator().manual_seed(20 + i)).cuda() for i in range(3)]
    d0 = _diff(lambda: nn.QIDDM_LL_noise(64, 4, 3, 2), "data", eps)
    g0 = GraphedTrainStep(d0, torch.optim.SGD(d0.parameters(), lr=0.05), tau, xs[0])
    l0 = [g0.step(x).item() for x in xs]
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        d1 = _diff(lambda: nn.QIDDM_LL_noise(64, 4, 3, 2), "data", eps)
        g1 = GraphedTrainStep(d1, torch.optim.SGD(d1.parameters(), lr=0.05), tau, xs[0], allreduce=True)
        assert g1.graph_tail is not None
        l1 = [g1.step(x).item() for x in xs]
    finally:
        dist.destroy_process_group()
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-6 * abs(a)
    for (n0, p0), (_, p1) in zip(d0.named_parameters(), d1.named_parameters()):
        assert rel_to_max(p1, p0) <= 1e-6, n0

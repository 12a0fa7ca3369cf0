"""World-size-2 gloo test of the data-parallel step: sharded step with the flat
gradient all-reduce == single-process step on the whole batch (same noise draw), bit-for-bit up to
reduction order.  Runs on CPU with a classical UNet (the N>1 host logic is independent of the kernels)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(seed):
    from qiddm_b200 import models, nn
    torch.manual_seed(seed)
    net = nn.UNetUndirected(depth=2, start_channels=4, qdepth=0)
    # BatchNorm statistics are rank-local by design (no SyncBN); use eval-mode BN so both runs agree exactly
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    return models.Diffusion(net, None, "data", (8, 8), torch.nn.MSELoss()).double()


def _noise_fn(eps_all, offset_holder):
    from qiddm_b200 import noise

    def f(d, tau, decay_mod):
        n = d.shape[0]
        e = eps_all[offset_holder[0]:offset_holder[0] + n]
        return noise.add_normal_noise_multiple(d, tau, decay_mod, eps=e)
    return f


def _worker(rank, world, port, x, eps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qiddm_b200.train import DataParallelTrainer, shard_batch
    diff = _make(0)
    for m in diff.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    start = sum(len(shard_batch(x, r, world)) for r in range(rank))
    diff.add_noise = _noise_fn(eps, [start])
    opt = torch.optim.SGD(diff.parameters(), lr=0.1)
    tr = DataParallelTrainer(diff, opt, tau=4)
    tr.broadcast_parameters()
    # keep BN in eval inside the trainer's diff.train()
    orig_train = diff.train

    def train_keep_bn(mode=True):
        orig_train(mode)
        for m in diff.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
        return diff
    diff.train = train_keep_bn
    tr.step(x)
    if rank == 0:
        torch.save({k: v.detach().clone() for k, v in diff.state_dict().items()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dp_step_equals_single_process_step():
    torch.manual_seed(1)
    x = torch.rand(5, 64, dtype=torch.float64)           # ragged: 3 + 2 images
    eps = torch.rand(5, 64, dtype=torch.float64)
    # single process reference
    from qiddm_b200.train import DataParallelTrainer
    ref = _make(0)
    ref.add_noise = _noise_fn(eps, [0])
    orig = ref.train

    def train_keep_bn(mode=True):
        orig(mode)
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
        return ref
    ref.train = train_keep_bn
    DataParallelTrainer(ref, torch.optim.SGD(ref.parameters(), lr=0.1), tau=4).step(x)
    want = ref.state_dict()

    import tempfile
    ctx = mp.get_context("spawn")
    port = _free_port()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "rank0.pt")
        procs = [ctx.Process(target=_worker, args=(r, 2, port, x, eps, path)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=180)
            assert p.exitcode == 0
        got = torch.load(path, weights_only=True)
    for k, v in want.items():
        assert torch.allclose(got[k], v, atol=1e-12, rtol=0), k


def test_flat_bucket_roundtrip():
    from qiddm_b200.train import FlatGradBucket, allreduce_gradients
    lin = torch.nn.Linear(3, 2)
    lin(torch.ones(1, 3)).sum().backward()
    g0 = [p.grad.clone() for p in lin.parameters()]
    b = FlatGradBucket(lin.parameters())
    allreduce_gradients(b)                  # no process group: identity
    assert b.flat.numel() == 8
    for p, g in zip(lin.parameters(), g0):
        assert torch.equal(p.grad, g)


def test_flat_bucket_gradients_are_views_of_the_bucket():
    """After `zero()` every p.grad is a view of the flat buffer: autograd accumulates in place, nothing is packed or unpacked,
    and parameters of two dtypes get one buffer each."""
    from qiddm_b200.train import FlatGradBucket
    lin64, lin32 = torch.nn.Linear(3, 2).double(), torch.nn.Linear(2, 2)
    params = list(lin64.parameters()) + list(lin32.parameters())
    b = FlatGradBucket(params)
    b.zero()
    assert len(b.flats) == 2 and b.flats[0].dtype == torch.float64 and b.flats[1].dtype == torch.float32
    ptrs = [p.grad.data_ptr() for p in params]
    for _ in range(2):                                   # the second backward ACCUMULATES into the same storage
        lin32(lin64(torch.ones(1, 3, dtype=torch.float64)).float()).sum().backward()
    assert [p.grad.data_ptr() for p in params] == ptrs
    assert torch.equal(b.flats[0][:6].view(2, 3), lin64.weight.grad) and lin64.weight.grad.abs().sum() > 0
    g = lin32.bias.grad.clone()
    b.pack(), b.unpack()                                 # both free
    assert torch.equal(lin32.bias.grad, g) and lin32.bias.grad.data_ptr() == ptrs[3]
    b.zero()
    assert all(float(f.abs().sum()) == 0 for f in b.flats) and [p.grad.data_ptr() for p in params] == ptrs

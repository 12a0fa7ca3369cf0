"""Data-parallel training step for `Diffusion` (the reference is single-process,
src/mnist_exm.py:148-203).

One process per GPU (torchrun).  Each rank takes whole images, so the tau-ladder, the per-batch PCA (H5)
and BatchNorm statistics stay rank-local exactly as in the reference.  `Diffusion.forward` calls
`.backward()` itself (src/models.py:67), so DDP wrappers do not fit: after it returns, all gradients are
packed into ONE flat bucket (circuit weights <= 2.2 k + classical <= 30 k floats), all-reduced once
(NCCL over NVLink on GPUs, gloo on CPU for tests) and averaged.  No collective touches the data path."""
from __future__ import annotations

import copy
import os
import warnings
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous image shard of this rank (images are independent units; ragged tails go to the low ranks)."""
    n = x.shape[0]
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return x[start:start + base + (1 if rank < rem else 0)]


class FlatGradBucket:
    """All gradients of `params` in ONE flat buffer per (dtype, device): every `p.grad` is a VIEW into it (as DDP's
    `gradient_as_bucket_view`), so autograd accumulates straight into the bucket, "packing" and "unpacking" move nothing,
    zeroing is one memset per buffer and the all-reduce is one collective per buffer (one in practice: the QIDDM modules are
    float64 after `.to(dtype=double)`; a float32 circuit-weight tensor next to float64 linears gives two)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.flats: List[torch.Tensor] = []
        self._views: List[torch.Tensor] = []

    def _ensure(self):
        if self._views:
            return
        groups = {}
        for p in self.params:
            groups.setdefault((p.dtype, p.device), []).append(p)
        view_of = {}
        for (dtype, device), ps in groups.items():
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=dtype, device=device)
            self.flats.append(flat)
            off = 0
            for p in ps:
                view_of[id(p)] = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self._views = [view_of[id(p)] for p in self.params]

    @property
    def flat(self) -> torch.Tensor:
        """The flat buffer (the first one when parameters of several dtypes are present; see `flats`)."""
        self._ensure()
        return self.flats[0]

    def zero(self):
        """`opt.zero_grad()` for the bucket: one memset per buffer; every `p.grad` is (re)attached to its view."""
        self._ensure()
        for f in self.flats:
            f.zero_()
        for p, v in zip(self.params, self._views):
            if p.grad is not v:
                p.grad = v

    def pack(self):
        """Make the buffers hold the current gradients.  Free when the gradients already are the views (after `zero()`);
        gradients autograd created on its own (first step after `zero_grad(set_to_none=True)`) are copied in once."""
        self._ensure()
        for p, v in zip(self.params, self._views):
            if p.grad is v:
                continue
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
            p.grad = v
        return self.flats[0]

    def unpack(self):
        """The gradients are views of the buffers: nothing to move."""
        for p, v in zip(self.params, self._views):
            if p.grad is not v:
                p.grad = v


def allreduce_gradients(bucket: FlatGradBucket, weights: Optional[float] = None) -> None:
    """Sum the flat bucket over ranks and divide by the world size (or weight by local/global samples)."""
    bucket.pack()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for flat in bucket.flats:
            if weights is not None:
                flat.mul_(weights)
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.div_(dist.get_world_size())
    bucket.unpack()


class DevicePrefetcher:
    """Double-buffered host -> device input pipeline (the reference's `batch[0].to(dev)`, src/mnist_exm.py:178,
    taken off the critical path): batch i+1 is copied from pinned host memory on a side stream while batch i is
    being simulated.  `next(host_batch)` returns the device copy of the batch submitted by the PREVIOUS call
    (None on the first call) and starts the copy of `host_batch`."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0
        self.pending = None

    def next(self, host_batch: Optional[torch.Tensor]):
        cur = torch.cuda.current_stream(self.device)
        out = None
        if self.pending is not None:
            j = self.pending
            cur.wait_event(self.ready[j])          # the compute stream consumes buffer j from here on
            out = self.bufs[j]
        if host_batch is not None:
            j = self.i & 1
            self.i += 1
            if self.bufs[j] is None or self.bufs[j].shape != host_batch.shape or self.bufs[j].dtype != host_batch.dtype:
                self.bufs[j] = torch.empty(host_batch.shape, dtype=host_batch.dtype, device=self.device)
            else:
                self.stream.wait_event(self.free[j])   # the step that last read buffer j has been enqueued and finished
            with torch.cuda.stream(self.stream):
                self.bufs[j].copy_(host_batch, non_blocking=True)
                self.ready[j].record(self.stream)
            self.pending = j
        else:
            self.pending = None
        return out

    def release(self, batch: torch.Tensor) -> None:
        """Call after the work that reads `batch` has been enqueued on the current stream."""
        for j in range(2):
            if self.bufs[j] is batch:
                self.free[j].record(torch.cuda.current_stream(self.device))


class DataParallelTrainer:
    """opt.zero_grad(); diff(x=x_local, T=tau); all-reduce; opt.step()  — the loop body of
    src/mnist_exm.py:175-182 with the batch sharded over ranks."""

    def __init__(self, diff: torch.nn.Module, optimizer: torch.optim.Optimizer, tau: int):
        self.diff, self.opt, self.tau = diff, optimizer, tau
        self.bucket = FlatGradBucket(diff.parameters())
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0

    def broadcast_parameters(self, src: int = 0) -> None:
        # through `p.detach()`, not `p.data`: the detached alias shares the version counter, so the in-place write
        # invalidates version-keyed caches (the collapsed operator of the GEMM path)
        if self.world > 1:
            for p in self.diff.parameters():
                dist.broadcast(p.detach(), src)
            for b in self.diff.buffers():
                dist.broadcast(b.detach(), src)

    def step(self, x_global: torch.Tensor, already_sharded: bool = False) -> torch.Tensor:
        x = x_global if already_sharded else shard_batch(x_global, self.rank, self.world)
        self.diff.train()
        self.bucket.zero()                        # gradients accumulate straight into the flat bucket
        n_global = x_global.shape[0] if not already_sharded else None
        if x.shape[0] > 0:
            (loss,) = self.diff(x=x, T=self.tau)
        else:
            loss = torch.zeros((), device=next(self.diff.parameters()).device)
        # the local loss is a mean over local samples: weight by the shard size so the result equals the
        # single-process mean over the global batch
        w = (x.shape[0] / n_global) if n_global else None
        allreduce_gradients(self.bucket, weights=w)
        self.opt.step()
        return loss.detach()


class GraphedTrainStep:
    """The loop body of src/mnist_exm.py:175-182 (`opt.zero_grad(); diff(x=x, T=tau); opt.step()`) captured ONCE in a
    CUDA graph and replayed per batch: at the reference batch sizes (batch 1, tau 10 -> 10 circuit
    instances) the step is ~100 dependent kernel launches, so launch latency, not arithmetic, bounds it.

    * inputs are copied into a static device buffer; the noise ladder's RNG draw, the net, the MSE, the backward
      (incl. the adjoint gate kernels / the unitary-collapse GEMMs) and Adam (`capturable=True`) live inside the graph;
      with `allreduce=True` under torch.distributed the step is two graphs around ONE eager NCCL all-reduce of the flat
      gradient bucket, which the gradients are views of (no pack / unpack kernels);
    * models with a host round trip in forward (sklearn PCA) cannot be captured: use `pca_on_device`.
    """

    def __init__(self, diff: torch.nn.Module, optimizer: torch.optim.Optimizer, tau: int, example_x: torch.Tensor,
                 warmup: int = 3, allreduce: bool = False):
        if not example_x.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (no CPU path)")
        for gparam in optimizer.param_groups:
            if "capturable" in gparam and not gparam["capturable"]:
                raise RuntimeError("the optimizer must be built with capturable=True to live inside a CUDA graph")
        self.diff, self.opt, self.tau = diff, optimizer, tau
        self.x = example_x.clone()
        self.bucket = FlatGradBucket(diff.parameters()) if allreduce else None
        self.graph = torch.cuda.CUDAGraph()
        diff.train()
        # warm-up steps (lazy initialisation of cuBLAS / allocator pools / optimizer state) must not count as training:
        # parameters, buffers and optimizer state are restored afterwards, so the first replay is the first step
        model_state = copy.deepcopy(diff.state_dict())
        opt_state = copy.deepcopy(optimizer.state_dict())
        side = torch.cuda.Stream(example_x.device)
        side.wait_stream(torch.cuda.current_stream(example_x.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._body()
        torch.cuda.current_stream(example_x.device).wait_stream(side)
        torch.cuda.synchronize(example_x.device)
        had_state = len(opt_state["state"]) > 0
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")                # QConv2d warns when a CHECKPOINT is loaded; this is our own snapshot
            diff.load_state_dict(model_state)              # in place: the parameter tensors stay the same objects
            if had_state:
                optimizer.load_state_dict(opt_state)
            else:                                          # fresh optimizer: zero the lazily created moments / step
                for st in optimizer.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
        self.opt.zero_grad(set_to_none=True)
        self.world = dist.get_world_size() if (allreduce and dist.is_available() and dist.is_initialized()) else 1
        if self.bucket is None or not (dist.is_available() and dist.is_initialized()):
            with torch.cuda.graph(self.graph):
                self.loss = self._body()
            self.graph_tail = None
        else:
            # data parallel: the collective stays OUTSIDE the captures (graph 1: zero the bucket .. backward into it; eager
            # NCCL all-reduce of the bucket; graph 2: average + optimizer) -- two replays and one NCCL call per step.
            # Capturing the all-reduce into a single graph was measured on 2 x B200 (round 2): 2.17 vs 2.15 ms (4096 images),
            # 0.336 vs 0.347 ms (1 image), 6.10 vs 6.01 ms (UNet) -- no gain, and the process group's teardown then hung.
            with torch.cuda.graph(self.graph):
                self.bucket.zero()
                (loss,) = self.diff(x=self.x, T=self.tau)
                self.loss = loss.detach()
                self.bucket.pack()
            self.graph_tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_tail, pool=self.graph.pool()):
                for f in self.bucket.flats:
                    f.div_(self.world)
                self.opt.step()
        self._params = [p for p in diff.parameters()]

    def _body(self):
        if self.bucket is not None:
            self.bucket.zero()
        else:
            self.opt.zero_grad(set_to_none=True)
        (loss,) = self.diff(x=self.x, T=self.tau)
        if self.bucket is not None:
            allreduce_gradients(self.bucket)
        self.opt.step()
        return loss.detach()

    def step(self, x: torch.Tensor) -> torch.Tensor:
        """Copies `x` (same shape as the example; host or device) into the static buffer and replays the graph(s).
        Returns the (device-resident) loss of this step."""
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        if self.graph_tail is not None:
            for f in self.bucket.flats:
                dist.all_reduce(f, op=dist.ReduceOp.SUM)
            self.graph_tail.replay()
        for p in self._params:        # the replay updated the weights in place: invalidate version-keyed caches
            torch.autograd.graph.increment_version(p)
        return self.loss


class GraphedSampler:
    """`Diffusion.sample` (src/models.py:106-147) as replays of one captured block of `unroll` fixed-point
    iterations on a static buffer: the 1000-iteration sampler of config 4 is 1000 dependent tiny forward passes."""

    def __init__(self, diff: torch.nn.Module, example_x: torch.Tensor, unroll: int = 10, noise_factor: float = 1.0):
        self.diff, self.unroll, self.noise_factor = diff, unroll, noise_factor
        self.x = example_x.clone()
        self.graph = torch.cuda.CUDAGraph()
        diff.eval()
        side = torch.cuda.Stream(example_x.device)
        side.wait_stream(torch.cuda.current_stream(example_x.device))
        with torch.cuda.stream(side), torch.no_grad():
            self._iterate(self.x.clone(), 2)
        torch.cuda.current_stream(example_x.device).wait_stream(side)
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.x.copy_(self._iterate(self.x, unroll))

    def _iterate(self, x, k):
        for _ in range(k):
            pred = self.diff.net(x)
            if self.diff.prediction_goal == "data":
                x = pred
            else:
                x = torch.clamp(x - (pred - 0.5) * 0.1 * self.noise_factor, 0, 1)
        return x

    def sample(self, n_iters: int, first_x: torch.Tensor) -> torch.Tensor:
        """Last iterate after `n_iters` iterations (the `only_last=True` result of Diffusion.sample)."""
        self.x.copy_(first_x)
        blocks, rest = divmod(n_iters, self.unroll)
        for _ in range(blocks):
            self.graph.replay()
        x = self.x
        if rest:
            with torch.no_grad():
                x = self._iterate(x.clone(), rest)
        return x.clone()

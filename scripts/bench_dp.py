#!/usr/bin/env python
"""Data-parallel training throughput of one QIDDM model at N GPUs (BASELINE.json configs 3-5: "8xB200 data-parallel"):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \\
      scripts/bench_dp.py --model unet --images 64 [--graph]
Each rank trains on its own `--images` images per step (weak scaling; tau = 10); gradients go through ONE flat-bucket
all-reduce per step (NCCL over NVLink), inside the CUDA graph when --graph.  Time = CUDA events, max over ranks."""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from qiddm_b200 import models, noise
from qiddm_b200 import nn as qnn
from qiddm_b200.train import DataParallelTrainer, GraphedTrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="unet", choices=["unet", "unet64", "qiddm_ll", "qiddm_pl", "qdense"])
ap.add_argument("--images", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
side = 64 if a.model == "unet64" else 28
net = {"unet": lambda: qnn.UNetUndirected(3, 8, 3), "unet64": lambda: qnn.UNetUndirected(3, 8, 3),
       "qiddm_ll": lambda: qnn.QIDDM_LL_noise(784, 6, 14, 2), "qiddm_pl": lambda: qnn.QIDDM_PL_noise(784, 8, 6, 2),
       "qdense": lambda: qnn.QDenseUndirected_old_noise(60, 28)}[a.model]()
if a.model == "qiddm_pl":
    net.pca_group = 10          # one PCA per image's tau-ladder (reference batch-1 semantics; whole-batch PCA is not capturable)
diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (side, side), torch.nn.MSELoss()).to(dev, torch.float64)
opt = torch.optim.Adam(diff.parameters(), lr=1e-3, capturable=a.graph)
trainer = DataParallelTrainer(diff, opt, tau=10)
trainer.broadcast_parameters()
torch.manual_seed(100 + rank)
x = torch.rand(a.images, side * side, device=dev, dtype=torch.float64)
if a.graph:
    gs = GraphedTrainStep(diff, opt, 10, x, allreduce=world > 1)
    step = lambda: gs.step(x)
else:
    step = lambda: trainer.step(x, already_sharded=True)
for _ in range(3):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = step()
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# replicas must stay identical: compare a parameter checksum across ranks
chk = torch.stack([p.detach().double().sum() for p in diff.parameters()]).sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
if world > 1:
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"what": "dp_train", "model": a.model, "n_gpus": world, "images_per_gpu": a.images, "cuda_graph": a.graph,
                      "ms_per_step": round(ms.item(), 3), "train_samples_per_s": round(a.images * world / ms.item() * 1e3, 1),
                      "loss": float(loss), "replicas_in_sync": bool((hi - lo).abs().item() <= 1e-9 * max(1.0, abs(hi.item())))}))
if world > 1:
    dist.destroy_process_group()

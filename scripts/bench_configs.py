#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 1-5 on ONE B200 (bench.py stays the headline line).

  python scripts/bench_configs.py --what sweep,config1,config3,config4,config5 [--quick]

Every record is one JSON line on stdout: CUDA-event time on the current stream after warm-up, with
the per-kernel library timers (qiddm_timing_collect) beside it.  `--cpu` adds a bounded CPU timing of the
oracle port of the same step (test infrastructure used as the reported CPU baseline, like bench.py's
cpu_baseline leg)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from qiddm_b200 import _lib as L  # noqa: E402
from qiddm_b200 import models, noise  # noqa: E402
from qiddm_b200 import nn as qnn  # noqa: E402

DEV = torch.device("cuda", 0)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, warmup=3, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    L.timing_enable(True)
    L.timing_collect()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = L.launch_count()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    kinds = L.timing_collect()
    L.timing_enable(False)
    ms = a.elapsed_time(b) / iters
    ker = {k: round(v["ms"] / iters, 4) for k, v in kinds.items() if v["launches"]}
    return ms, ker, (L.launch_count() - n0) // iters


# ----------------------------------------------------------------------------------------------
# config 2: QDense / re-upload stage sweep over qubit count x batch
# ----------------------------------------------------------------------------------------------
def sweep(quick):
    ns = [4, 6, 8, 10, 12] if quick else list(range(4, 13))
    for n in ns:
        A = 1 << n
        F = 784 if n == 10 else A
        for fam, depth in (("qdense", 2), ("qdense", 10), ("qdense", 60), ("reupload", 6)):
            if fam == "qdense":
                base = dict(n_qubits=n, layers_per_block=depth, init=L.INIT_AMPLITUDE, n_features=F, pad_value=0.1,
                            imprimitive=L.IMP_CNOT, remap=L.REMAP_TANH, readout=L.READ_PROBS, read_count=F,
                            post_scale=float(F), clamp=True)
                w = torch.randn(depth, n, 3, device=DEV, dtype=torch.float64) * 0.4
                n_rot, n_in = depth * n, F
                paths = [("gate", L.PATH_GATE), ("gemm", L.PATH_GEMM)]
            else:
                base = dict(n_qubits=n, n_blocks=depth, layers_per_block=2, init=L.INIT_ZERO, enc=L.ENC_RZ,
                            imprimitive=L.IMP_CZ, readout=L.READ_EXPVAL_Z)
                w = torch.randn(depth, 2, n, 3, device=DEV, dtype=torch.float64) * 0.4
                n_rot, n_in = depth * 2 * n, n
                paths = [("gate", L.PATH_GATE)]
            batches = [1 << 10, 1 << 14, 1 << 18] if quick else [1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20]
            for B in batches:
                if B * A * 8 > (8 << 30) or B * n_rot * A > 3e12:
                    continue
                for pname, pid in paths:
                    spec = L.StageSpec(path=pid, **base)
                    plan = L.Plan.get(spec)
                    if pname == "gemm" and (not plan.gemm_supported() or B < 2 * A):
                        continue
                    x = torch.rand(B, n_in, device=DEV) if fam == "qdense" else torch.randn(B, n_in, device=DEV)
                    go = torch.randn(B, spec.n_out, device=DEV) / B

                    if pname == "gemm":
                        def fwd():
                            w.add_(0.0)
                            plan.gemm_forward(x, w)

                        def both():
                            w.add_(0.0)
                            _, saved = plan.gemm_forward(x, w, save=True)
                            plan.gemm_backward(x, w, go, saved=saved)
                    else:
                        def fwd():
                            plan.forward(x, w)

                        def both():       # training pair: psi_final travels from the forward to the adjoint launch
                            _, st = plan.forward(x, w, save_state=True)
                            plan.backward(x, w, go, state=st)
                    it = 3 if B * n_rot * A > 2e11 else 5
                    f_ms, _, _ = timed(fwd, 2, it)
                    t_ms, ker, _ = timed(both, 2, it)
                    flop = 14.0 * A * n_rot
                    emit(what="sweep", family=fam, n=n, depth=depth, n_rot=n_rot, batch=B, path=pname,
                         fwd_ms=round(f_ms, 4), fwd_bwd_ms=round(t_ms, 4), fwd_evals_per_s=round(B / f_ms * 1e3),
                         fwd_bwd_evals_per_s=round(B / t_ms * 1e3),
                         gate_alg_tflops_fwd=round(B * flop / f_ms / 1e9, 2) if pname == "gate" else None,
                         alg_hbm_gbs_fwd_bwd=round(B * 4 * (2 * n_in + 2 * spec.n_out) / t_ms / 1e6, 1), kernels=ker)
                    del x, go
                torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------
# training-step helpers
# ----------------------------------------------------------------------------------------------
def train_step_rate(net, shape, imgs, tau, lr, goal="data", dtype=torch.float64, iters=5, warmup=3, graphed=False):
    diff = models.Diffusion(net, noise.add_normal_noise_multiple, goal, shape, torch.nn.MSELoss()).to(DEV, dtype)
    diff.train()
    opt = torch.optim.Adam(diff.parameters(), lr=lr, capturable=graphed)
    x = torch.rand(imgs, shape[0] * shape[1], device=DEV, dtype=dtype)
    if graphed:
        from qiddm_b200.train import GraphedTrainStep
        gs = GraphedTrainStep(diff, opt, tau, x)

        def step():
            gs.step(x)
    else:
        def step():
            opt.zero_grad(set_to_none=True)
            diff(x=x, T=tau)
            opt.step()

    ms, ker, launches = timed(step, warmup, iters)
    return diff, ms, ker, launches


def cpu_step_seconds(fn, reps=2):
    torch.set_num_threads(os.cpu_count() or 1)
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps


def config1(cpu):
    """src/mnist_exm.py defaults: QIDDM_LL_noise(784,6,14,2), QNN_noise(784,8,14); batch 1 image, tau 10."""
    from oracle import qiddm_oracle as O
    for name, args, lr in (("QIDDM_LL_noise", (784, 6, 14, 2), 0.0255), ("QNN_noise", (784, 8, 14), 0.01011)):
        for imgs, graphed, dt in ((1, False, torch.float64), (1, True, torch.float64), (8, False, torch.float64),
                                  (8, True, torch.float64), (512, False, torch.float64), (512, True, torch.float64),
                                  (8192, False, torch.float64), (8192, True, torch.float64), (8192, True, torch.float32)):
            torch.manual_seed(42)
            net = getattr(qnn, name)(*args)
            stages = args[3] if name == "QIDDM_LL_noise" else 1
            _, ms, ker, launches = train_step_rate(net, (28, 28), imgs, 10, lr, iters=20 if imgs <= 8 else 5,
                                                   graphed=graphed, dtype=dt)
            rec = dict(what="config1", model=f"{name}{args}", images_per_step=imgs, tau=10, cuda_graph=graphed,
                       module_dtype=str(dt).replace("torch.", ""), ms_per_step=round(ms, 4),
                       train_samples_per_s=round(imgs / ms * 1e3, 1),
                       circuit_evals_per_s=round(imgs * 10 * stages / ms * 1e3), lib_launches_per_step=launches, kernels=ker)
            if cpu and imgs == 1 and not graphed:
                ps = {k: v.detach().cpu().double().clone().requires_grad_(True) for k, v in net.named_parameters()}
                data = torch.rand(1, 784, dtype=torch.float64)
                eps = torch.normal(0.5, 0.2, size=(1, 784)).double()
                if name == "QIDDM_LL_noise":
                    f = lambda v: O.qiddm_ll_forward(v, ps["weights1"], ps["linear_down.weight"], ps["linear_down.bias"],
                                                     ps["linear_up.weight"], ps["linear_up.bias"])
                else:
                    f = lambda v: O.qnn_forward(v, ps["weights"], ps["linear_down.weight"], ps["linear_down.bias"],
                                                ps["linear_up.weight"], ps["linear_up.bias"])

                def cstep():
                    for p in ps.values():
                        p.grad = None
                    O.diffusion_loss(f, data, eps, 10, (28, 28), "data").backward()
                s = cpu_step_seconds(cstep, 3)
                rec["cpu_oracle_ms_per_step"] = round(s * 1e3, 2)
                rec["cpu_cores"] = os.cpu_count()
            emit(**rec)


def unet_rate(tag, side, imgs_list, qdepth=3, simple=False, graphed=False, dtype=torch.float64):
    for imgs in imgs_list:
        torch.manual_seed(0)
        net = (qnn.UNetUndirectedS if simple else qnn.UNetUndirected)(3, 8, qdepth)
        try:
            _, ms, ker, launches = train_step_rate(net, (side, side), imgs, 10, 1e-3, iters=3, warmup=2, graphed=graphed,
                                                   dtype=dtype)
        except torch.OutOfMemoryError:
            emit(what=tag, images_per_step=imgs, error="oom")
            continue
        patches = 5782 if side == 28 else None
        emit(what=tag, model=f"{'UNetUndirectedS' if simple else 'UNetUndirected'}(3,8,{qdepth}) {side}x{side}",
             images_per_step=imgs, tau=10, cuda_graph=graphed, module_dtype=str(dtype).replace("torch.", ""),
             ms_per_step=round(ms, 3), train_samples_per_s=round(imgs / ms * 1e3, 1),
             circuit_evals_per_s=(round(imgs * 10 * patches / ms * 1e3) if patches else None),
             lib_launches_per_step=launches, kernels=ker)
        del net
        torch.cuda.empty_cache()


def config4(quick):
    """src/emnist_exm.py: QIDDM_PL_noise(784,8,6,2) training + Diffusion.sample(n_iters=1000) on 10 images."""
    torch.manual_seed(0)
    net = qnn.QIDDM_PL_noise(784, 8, 6, 2)
    from qiddm_b200.nn import qdense as qd
    from qiddm_b200.train import GraphedSampler
    for imgs, graphed, group in ((1, False, None), (1, True, None), (8, True, None), (64, False, None), (64, True, 10),
                                 (1024, True, 10), (8192, True, 10)):
        net.pca_group = group
        diff, ms, ker, launches = train_step_rate(net, (28, 28), imgs, 10, 0.0255, goal="noise", iters=10, graphed=graphed)
        emit(what="config4_train", model="QIDDM_PL_noise(784,8,6,2)", images_per_step=imgs, tau=10, cuda_graph=graphed,
             ms_per_step=round(ms, 3), train_samples_per_s=round(imgs / ms * 1e3, 1), lib_launches_per_step=launches,
             kernels=ker, pca=("device, one PCA per image (pca_group = tau)" if group else
                               "device, one PCA over the batch" if qd.PCA_ON_DEVICE else "host sklearn"))
    diff.eval()
    n_iters = 100 if quick else 1000
    for nimg, graphed, group in ((10, False, None), (10, True, None), (64, True, None), (4096, False, None),
                                 (4000, True, 10), (40000, True, 10)):
        net.pca_group = group
        first = torch.rand(nimg, 1, 28, 28, device=DEV, dtype=torch.float64) * 0.75 + 0.5
        gs = GraphedSampler(diff, first, unroll=10) if graphed else None
        torch.cuda.synchronize()
        t = time.perf_counter()
        if graphed:
            gs.sample(n_iters, first)
        else:
            diff.sample(n_iters=n_iters, first_x=first, only_last=True)
        torch.cuda.synchronize()
        s = time.perf_counter() - t
        emit(what="config4_sample", model="QIDDM_PL_noise(784,8,6,2)", images=nimg, n_iters=n_iters, cuda_graph=graphed,
             pca_group=group, seconds=round(s, 3), sampler_iters_per_s=round(n_iters / s, 1),
             image_iterations_per_s=round(nimg * n_iters / s), circuit_evals_per_s=round(nimg * 2 * n_iters / s))


def config5(quick):
    torch.manual_seed(0)
    for imgs in (8, 64) if quick else (8, 64, 512):
        net = qnn.QDenseUndirected_old(60, 64)
        _, ms, ker, launches = train_step_rate(net, (64, 64), imgs, 10, 1e-3, iters=3, warmup=2)
        emit(what="config5_qdense64", model="QDenseUndirected_old(60,64) n=12", images_per_step=imgs, tau=10,
             ms_per_step=round(ms, 3), train_samples_per_s=round(imgs / ms * 1e3, 1),
             circuit_evals_per_s=round(imgs * 10 / ms * 1e3), lib_launches_per_step=launches, kernels=ker)
    net = qnn.QIDDM_PL_noise(4096, 8, 6, 2)
    _, ms, ker, launches = train_step_rate(net, (64, 64), 64, 10, 1e-3, goal="noise", iters=3)
    emit(what="config5_pl64", model="QIDDM_PL_noise(4096,8,6,2)", images_per_step=64, tau=10, ms_per_step=round(ms, 3),
         train_samples_per_s=round(64 / ms * 1e3, 1), lib_launches_per_step=launches, kernels=ker)
    unet_rate("config5_unet64", 64, (8,) if quick else (8, 32))
    unet_rate("config5_unet64", 64, (8,) if quick else (8, 32), graphed=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="config1,config3,config4,config5,sweep")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(DEV)
    emit(what="env", gpu=torch.cuda.get_device_name(0), torch=torch.__version__, cores=os.cpu_count())
    for w in a.what.split(","):
        if w == "sweep":
            sweep(a.quick)
        elif w == "config1":
            config1(a.cpu)
        elif w == "config3":
            unet_rate("config3_unet28", 28, (8, 64) if a.quick else (1, 8, 64, 256))
            unet_rate("config3_unet28", 28, (8, 64) if a.quick else (1, 8, 64, 256), graphed=True)
            unet_rate("config3_unet28_simple", 28, (64,), simple=True)
        elif w == "config4":
            config4(a.quick)
        elif w == "config5":
            config5(a.quick)


if __name__ == "__main__":
    main()

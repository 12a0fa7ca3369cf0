#!/usr/bin/env python
"""bench.py — QDense quantum-layer fwd+bwd throughput on B200 (BASELINE.json configs[1]).

One "step" = one forward + adjoint backward of `QDenseUndirected_old_noise(60, 28)` (n = 10 qubits,
600 Rot + 600 CNOT, MNIST-shaped 28x28 inputs; nn/qdense.py:71-125) over one batch of B synthetic
circuit instances per GPU.  metric = circuit evals/s (one eval = one state-vector simulation of one
instance, fwd+bwd), whole job over all ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL): instances are sharded over ranks (weak
scaling, no data-path collective); only the 1 800-float circuit-weight gradient is all-reduced per step.
`--impl reference` times the CPU oracle port of the reference path on the host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

QDEPTH, SIDE = 60, 28
PIXELS = SIDE * SIDE
NQ = 10
METRIC = "circuit_evals_per_sec_fwd_bwd"
UNIT = "circuit-evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=524288, help="circuit instances per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--path", default="auto", choices=["auto", "gate", "gemm"],
                    help="auto = library dispatch (unitary-collapse tcgen05 GEMM when batch >= 2 * 2^n)")
    ap.add_argument("--precision", type=int, default=3, choices=[1, 3],
                    help="GEMM path: 3 = fp32-grade 3-term fp16 split (default), 1 = single fp16 pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


def workload_config(batch, n_gpus):
    return {"workload": f"QDenseUndirected_old_noise({QDEPTH},{SIDE}) fwd+bwd, n={NQ} qubits, 600 Rot + 600 CNOT, "
                        f"synthetic MNIST-shaped 28x28", "instances_per_gpu": batch, "global_instances": batch * n_gpus,
            "parallelism": f"dp{n_gpus} (instances sharded, weight-grad all-reduce only)",
            "step": "device-resident step = forward + backward for BOTH the input and the weight gradients; the e2e "
                    "Diffusion step (first layer: the noisy images need no gradient) skips the dX GEMM, as the reference does",
            "l2": "inputs+grads per step (>=2x%.0f MB) exceed the 126 MB L2" % (batch * PIXELS * 4 / 1e6)}


# ----------------------------------------------------------------------------------------------
# CPU oracle legs (the only places bench.py may execute oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_oracle_rate(target_seconds: float, steps: int = 1, warmup: int = 0):
    """Times the complex128 oracle port of the reference path (fwd + autograd bwd) on the host cores."""
    import torch
    from oracle import qiddm_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(42)
    W = (torch.randn(QDEPTH, NQ, 3, generator=g, dtype=torch.float64) * 0.4).requires_grad_(True)

    def run(b):
        x = torch.rand(b, 1, SIDE, SIDE, generator=g, dtype=torch.float64)
        go = torch.randn(b, 1, SIDE, SIDE, generator=g, dtype=torch.float64)
        t = time.perf_counter()
        out = O.qdense_forward(x, W, O.REMAP_TANH)
        (out * go).sum().backward()
        W.grad = None
        return time.perf_counter() - t

    run(4)                                   # warm the thread pool / allocator
    t_probe = run(16)
    per = t_probe / 16
    per_step = max(target_seconds / max(steps + warmup, 1), 0.5)
    b = int(min(4096, max(16, per_step / per)))
    for _ in range(4):                       # small batches under-use the cores: refine the estimate at the sample's own size
        t_b = run(b)
        if t_b >= 0.6 * per_step or b >= 4096:
            break
        b = int(min(4096, max(b + 1, b * per_step / t_b)))
    for _ in range(warmup):
        run(b)
    times = [run(b) for _ in range(max(steps, 1))]
    t = sum(times) / len(times)
    cb = {"value": b / t, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{b} instances/step x {len(times)} step(s) of the same circuit, complex128 torch oracle "
                    f"(per-gate ops on a (B,2^n) tensor + autograd, mirrors default.qubit.torch), {t:.2f} s/step"}
    try:    # for scale: the forward alone on the gate-by-gate C restatement (lightning.qubit-like), OpenMP over the instances
        from oracle import c_oracle as C
        d = O.desc_qdense(QDEPTH, PIXELS, O.REMAP_TANH)
        xc = torch.rand(32 * cores, PIXELS, generator=g, dtype=torch.float64)
        C.run_stage(d, xc[:cores], W.detach()[None], threads=cores)
        t0 = time.perf_counter()
        C.run_stage(d, xc, W.detach()[None], threads=cores)
        cb["c_forward_only"] = {"value": xc.shape[0] / (time.perf_counter() - t0), "unit": "circuit-evals/s (forward only)",
                                "cores": cores, "sample": f"{xc.shape[0]} instances, oracle/statevec_oracle.c"}
        goc = torch.randn(xc.shape[0], PIXELS, generator=g, dtype=torch.float64)
        t0 = time.perf_counter()
        C.stage_grads(d, xc, W.detach()[None], goc, threads=cores)
        cb["c_fwd_bwd"] = {"value": xc.shape[0] / (time.perf_counter() - t0), "unit": UNIT, "cores": cores,
                           "sample": f"{xc.shape[0]} instances, forward + adjoint-method backward in C (OpenMP over the "
                                     f"instances): what an optimised CPU simulator reaches; the reference's path is the torch one"}
    except Exception as e:
        cb["c_forward_only"] = {"unavailable": str(e)[:120]}
    return cb, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, t = cpu_oracle_rate(args.cpu_seconds * 2, steps=max(1, min(args.steps, 5)), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.batch, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "PennyLane/Lightning are not installable here (no network); this is the oracle port of the "
                    "reference's default.qubit.torch path on the host cores"}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = sorted(float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import qiddm_b200
    from qiddm_b200 import nn as qnn
    from qiddm_b200._lib import Plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)

    import dataclasses
    from qiddm_b200 import _lib as L
    from qiddm_b200 import models, noise
    torch.manual_seed(42 + rank)
    net = qnn.QDenseUndirected_old_noise(QDEPTH, SIDE).to(dev, torch.float64)
    if world > 1:
        dist.broadcast(net.weights.data, 0)
    path_id = {"auto": L.PATH_AUTO, "gate": L.PATH_GATE, "gemm": L.PATH_GEMM}[args.path]
    spec = dataclasses.replace(net._spec(), path=path_id, gemm_precision=args.precision)
    net._spec = lambda: spec                                   # the module API (e2e) uses the same dispatch
    plan = Plan.get(spec)
    use_gemm = plan.use_gemm(B)
    x = torch.rand(B, PIXELS, device=dev, dtype=torch.float32)               # resident in HBM
    go = torch.randn(B, PIXELS, device=dev, dtype=torch.float32) / (B * PIXELS)
    w = net.weights.detach()

    def step_device():
        w.add_(0.0)            # bumps the version counter like an optimizer step: the collapse is redone
        if use_gemm:
            out, saved = plan.gemm_forward(x, w, save=True)      # training forward keeps Y + operand splits
            gi, gw = plan.gemm_backward(x, w, go, need_grad_in=True, need_grad_w=True, saved=saved)
        else:
            out = plan.forward(x, w)
            gi, gw = plan.backward(x, w, go, need_grad_in=True, need_grad_w=True)
        if world > 1:
            dist.all_reduce(gw)
        return out, gi, gw

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        step_device()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # --- timed region: exactly K steps, CUDA events on the launching (current) stream
    L.timing_enable(True)
    L.timing_collect()
    launches0 = qiddm_b200.launch_count()
    sync_all()
    t_wall0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(K):
        step_device()
    e1.record()
    sync_all()
    t_wall1 = time.perf_counter()
    launches = qiddm_b200.launch_count() - launches0
    kinds = L.timing_collect()
    L.timing_enable(False)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_step = ms / K
    value = B * world / (ms_step * 1e-3)

    # --- e2e: the QIDDM training step through the public module API (src/models.py:44-67): pinned host
    # images -> H2D -> noise ladder (tau = 10 instances per image) -> QDense net -> MSE -> backward ->
    # loss + circuit-weight gradient back on the host.  B instances/step = B/10 images/step.
    TAU = 10
    imgs = max(1, B // TAU)
    diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (SIDE, SIDE), torch.nn.MSELoss()).to(dev)
    diff.train()
    xh = torch.rand(imgs, PIXELS, dtype=torch.float32).pin_memory()
    res_h = torch.empty(1 + net.weights.numel(), dtype=torch.float64).pin_memory()

    from qiddm_b200.train import DevicePrefetcher
    pre = DevicePrefetcher(dev)
    pre.next(xh)                                     # prime: the first batch is in flight

    def step_e2e():
        xd = pre.next(xh)                            # this step's device batch; starts the H2D copy of the next one
        net.weights.grad = None
        with torch.no_grad():
            net.weights.add_(0.0)
        (loss,) = diff(x=xd, T=TAU)
        pre.release(xd)
        g = net.weights.grad
        if world > 1:
            dist.all_reduce(g)
        res_h.copy_(torch.cat([loss.detach().reshape(1).double(), g.reshape(-1).double()]), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        step_e2e()
    sync_all()
    ke = max(3, min(K, 5))
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(ke):
        step_e2e()
    e1.record()
    sync_all()
    ms_e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e = t.item()
    e2e_val = imgs * TAU * world / (ms_e / ke * 1e-3)
    # clocks sampled over both timed regions (device-resident steps and the end-to-end steps)
    clocks = sampler.stop(t_wall0, time.perf_counter()) if rank == 0 else None
    if clocks is not None and "sw_power_cap" in (clocks.get("reasons") or []):
        # nvidia-smi's polled clocks.sm does not resolve the cap: ncu reports 1.17-1.43 GHz inside the GEMM launches
        clocks["note"] = "power cap active: ncu shows 1.17-1.43 GHz SM clock inside the GEMM launches (profiles/r1_gemm_pair_summary.md)"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel (largest share of the timed region, measured live)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    fp32_peak = 148 * 128 * 2 * (float(peaks.get("sm_max_mhz", 1965.0)) * 1e6) / 1e12
    MAIN = ("gate_forward", "gate_backward", "gemm")
    shares = {k: v["ms"] / ms for k, v in kinds.items() if v["launches"] and k in MAIN}
    dom = max(shares, key=shares.get)
    kd = kinds[dom]
    per_launch_ms = kd["ms"] / kd["launches"]
    if dom == "gemm":
        achieved = kd["work"] / (kd["ms"] * 1e-3) / 1e12
        roofline = {"kernel": "gemm_pair_kernel (tcgen05.mma.cta_group::2 kind::f16 + TMA, fused |Y|^2 readout epilogue)",
                    "bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": achieved / tc_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the step's three GEMM launches, from
                    # the committed `ncu --set full` capture of this command (profiles/r1_gemm_pair_ncu_full_raw.csv; 8.3-11.2e9
                    # across the captures of the round: the dW launch's L2 re-reads vary with the box, profiles/r1_gemm_pair_summary.md)
                    "traffic": 11.2e9 if (B == 524288 and args.precision == 3) else None,
                    "traffic_unit": "bytes/launch (algorithmic operand + result bytes: 6.04e9)",
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback"),
                    "ms_per_launch": per_launch_ms, "launches": kd["launches"],
                    "algorithmic_flops": "2*M*N*K per GEMM (single pass); precision %d executes %dx that"
                                         % (args.precision, args.precision),
                    "executed_tflops": achieved * args.precision, "share_of_step": shares[dom]}
    else:
        alg_bytes = B * 4 * (PIXELS * 3) * K                 # x in, grad_out in, grad_in out (fp32) per launch
        achieved = alg_bytes / (kd["ms"] * 1e-3) / 1e9
        tf = kd["work"] / (kd["ms"] * 1e-3) / 1e12
        roofline = {"kernel": f"gate_kernel<{NQ},*> ({dom}, gate path)", "bound": "hbm", "achieved": achieved,
                    "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "ms_per_launch": per_launch_ms,
                    "launches": kd["launches"], "share_of_step": shares[dom],
                    "note": "the gate path is FP32-FMA/shared-memory bound, not HBM bound (SURVEY.md 8d)",
                    "fp32": {"achieved_tflops": tf, "peak_tflops_nominal": fp32_peak, "frac": tf / fp32_peak}}
    roofline["kernel_shares_of_step"] = shares
    roofline["kernel_ms_per_step"] = {k: round(v["ms"] / K, 4) for k, v in kinds.items() if v["launches"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 (tcgen05 fp16 x%d split, fp32 accumulate)" % args.precision) if use_gemm else "f32",
            "data": "synthetic", "config": workload_config(B, world), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": imgs * PIXELS * 4,
                    "d2h_bytes_per_step": res_h.numel() * 8, "ms_per_step": ms_e / ke,
                    "api": "qiddm_b200.models.Diffusion(QDenseUndirected_old_noise(60,28)).forward(x, T=10): "
                           f"{imgs} pinned host images/step -> {imgs * TAU} circuit instances, loss+grad to host; "
                           "H2D double-buffered on a copy stream (qiddm_b200.train.DevicePrefetcher), one copy per step"},
            "path": ("gemm_x%d" % args.precision) if use_gemm else "gate",
            "gpu_launches": int(launches), "roofline": roofline}

    if not args.no_extras:
        line["extras"] = extras(dev)
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_oracle_rate(args.cpu_seconds)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extras(dev):
    """Secondary numbers (not the headline): forward-only rate, and QIDDM train samples/s on an
    MNIST-shaped QIDDM_LL_noise(784,6,14,2) training step (src/mnist_exm.py:46, tau = 10)."""
    import torch
    from qiddm_b200 import models, nn as qnn, noise
    from qiddm_b200._lib import Plan
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    net = qnn.QDenseUndirected_old_noise(QDEPTH, SIDE).to(dev, torch.float64)
    plan = Plan.get(net._spec())
    B = 65536
    x = torch.rand(B, PIXELS, device=dev)
    for _ in range(2):
        plan.forward(x, net.weights.detach())
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(5):
        plan.forward(x, net.weights.detach())
    b.record()
    torch.cuda.synchronize()
    out["qdense_forward_only_evals_per_s"] = B * 5 / (a.elapsed_time(b) * 1e-3)
    # QIDDM train samples/s
    torch.manual_seed(0)
    qn = qnn.QIDDM_LL_noise(784, 6, 14, 2)
    diff = models.Diffusion(qn, noise.add_normal_noise_multiple, "data", (28, 28), torch.nn.MSELoss()).to(dev, torch.float64)
    diff.train()
    opt = torch.optim.Adam(diff.parameters(), lr=0.0255)
    imgs = 4096
    data = torch.rand(imgs, 784, device=dev, dtype=torch.float64)

    def step():
        opt.zero_grad(set_to_none=True)
        diff(x=data, T=10)
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(5):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out["qiddm_ll_train_samples_per_s"] = imgs / (ms * 1e-3)
    out["qiddm_ll_train_circuit_evals_per_s"] = imgs * 10 * 2 / (ms * 1e-3)
    out["qiddm_ll_config"] = "QIDDM_LL_noise(784,6,14,2), 4096 images/step, tau=10, Adam, float64 module I/O"
    del diff, opt, data

    # the other BASELINE configs through CUDA-graph steps (qiddm_b200.train.GraphedTrainStep), float64 modules, tau = 10
    from qiddm_b200.train import GraphedTrainStep

    def graphed_rate(net, imgs, side, goal, iters=10):
        d = models.Diffusion(net, noise.add_normal_noise_multiple, goal, (side, side), torch.nn.MSELoss()).to(dev, torch.float64)
        o = torch.optim.Adam(d.parameters(), lr=1e-3, capturable=True)
        xs = torch.rand(imgs, side * side, device=dev, dtype=torch.float64)
        g = GraphedTrainStep(d, o, 10, xs)
        for _ in range(2):
            g.step(xs)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(iters):
            g.step(xs)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    torch.manual_seed(0)
    ms = graphed_rate(qnn.QIDDM_LL_noise(784, 6, 14, 2), 1, 28, "data", iters=50)
    out["config1_reference_batch_ms_per_step_graphed"] = ms           # src/mnist_exm.py defaults: 1 image x tau 10
    ms = graphed_rate(qnn.UNetUndirected(3, 8, 3), 64, 28, "data", iters=5)
    out["config3_qconv_unet_train_samples_per_s"] = 64 / (ms * 1e-3)
    out["config3_qconv_unet_circuit_evals_per_s"] = 64 * 10 * 5782 / (ms * 1e-3)
    pl = qnn.QIDDM_PL_noise(784, 8, 6, 2)
    pl.pca_group = 10                                                  # one PCA per image's tau-ladder (reference batch-1 semantics)
    ms = graphed_rate(pl, 1024, 28, "noise", iters=10)
    out["config4_qiddm_pl_train_samples_per_s"] = 1024 / (ms * 1e-3)
    out["configs_note"] = ("graphed steps, one B200: config 1 = QIDDM_LL_noise(784,6,14,2) at the reference batch (1 image); "
                           "config 3 = UNetUndirected(3,8,3), 64 images/step, QConv on the tcgen05 path; config 4 = "
                           "QIDDM_PL_noise(784,8,6,2), 1024 images/step, per-image on-device PCA; more in profiles/r1_configs.md")
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

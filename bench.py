#!/usr/bin/env python
"""bench.py — QDense quantum-layer fwd+bwd throughput on B200 (BASELINE.json configs[1]).

One "step" = one forward + adjoint backward of `QDenseUndirected_old_noise(60, 28)` (n = 10 qubits,
600 Rot + 600 CNOT, MNIST-shaped 28x28 inputs; nn/qdense.py:71-125) over one batch of B synthetic
circuit instances per GPU.  metric = circuit evals/s (one eval = one state-vector simulation of one
instance, fwd+bwd), whole job over all ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL): instances are sharded over ranks (weak
scaling, no data-path collective); only the 1 800-float circuit-weight gradient is all-reduced per step.
`--impl reference` times the CPU oracle port of the reference path on the host cores (rank 0 only): W warm-up + K timed
steps, each a FIXED bounded sample (`config.sample_instances`) of the same workload.

Besides the headline (`value`, `e2e`, `roofline` of the tcgen05 GEMM) the line carries `secondary`: the second half of
BASELINE.json's metric, QIDDM train samples/s (configs 1 and 4) as CUDA-graph training steps, data-parallel over the ranks
with the flat-bucket all-reduce, with the gate kernels' FP32 roofline (peak measured in-run by qiddm_probe_fp32_fma).
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
    ap.add_argument("--cpu-instances", type=int, default=1024,
                    help="circuit instances per CPU sample step (cpu_baseline: best of 3; --impl reference: every step)")
    ap.add_argument("--path", default="auto", choices=["auto", "gate", "gemm"],
                    help="auto = library dispatch (unitary-collapse tcgen05 GEMM when batch >= 2 * 2^n)")
    ap.add_argument("--precision", type=int, default=3, choices=[1, 3],
                    help="GEMM path: 3 = fp32-grade 3-term fp16 split (default), 1 = single fp16 pass")
    ap.add_argument("--bwd-precision", type=int, default=0, choices=[0, 1, 3],
                    help="GEMM path, dX / dW GEMMs: 0 = same as --precision; 1 behind --precision 3 = x3 forward, x1 gradients")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--secondary", default="", help="comma list of the secondary configs to run (default: config1,config4,config3,config5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


def workload_config(batch, n_gpus):
    return {"workload": f"QDenseUndirected_old_noise({QDEPTH},{SIDE}) fwd+bwd, n={NQ} qubits, 600 Rot + 600 CNOT, "
                        f"synthetic MNIST-shaped 28x28", "instances_per_gpu": batch, "global_instances": batch * n_gpus,
            "work_per_step": "value: forward + input gradient + weight gradient per instance; e2e and the CPU arms "
                             "(cpu_baseline, --impl reference): forward + weight gradient (the layer is the first one of the "
                             "Diffusion step, its input needs no gradient -- src/models.py:64-67)",
            "parallelism": f"dp{n_gpus} (instances sharded, weight-grad all-reduce only)",
            "step": "device-resident step (`value`, what `roofline` explains) = the layer's generic forward + backward for BOTH "
                    "the input and the weight gradients given an upstream gradient tensor; the e2e Diffusion training step (first "
                    "layer: the noisy images need no gradient, so no dX GEMM, as in the reference) runs as the FUSED step "
                    "qiddm_dense_mse_step: noise ladder -> fp16 operand splits in one pass, forward GEMM whose epilogue forms the "
                    "MSE loss and dL/dY on the accumulator tile (no out / Y / grad_out arrays, no separate MSE and dL/dY passes), "
                    "dW GEMM, adjoint sweep on the basis columns (QIDDM_FUSED_STEP=0: the unfused kernel sequence)",
            "l2": "inputs+grads per step (>=2x%.0f MB) exceed the 126 MB L2" % (batch * PIXELS * 4 / 1e6)}


# ----------------------------------------------------------------------------------------------
# CPU oracle legs (the only places bench.py may execute oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_sample_size(requested: int, steps_total: int) -> int:
    """Fixed sample per CPU step: `requested` (1024) instances, reduced only when W + K steps of it would not end within a few
    minutes (~2 s per 1024 instances on 16 cores)."""
    b = requested
    if steps_total > 40:
        b = max(128, (requested * 40 // steps_total) // 64 * 64)
    return b


def cpu_oracle_rate(sample_instances: int, steps: int = 3, warmup: int = 1, best: bool = True):
    """Times the complex128 oracle port of the reference path (forward + autograd weight gradient) on the host cores:
    `warmup` + `steps` steps of `sample_instances` instances each.  best=True -> rate of the fastest step (cpu_baseline),
    else of the mean step (--impl reference, whose ms_per_step x steps must match the wall clock)."""
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

    b = sample_instances
    run(8)                                   # warm the thread pool / allocator
    for _ in range(warmup):
        run(b)
    times = [run(b) for _ in range(max(steps, 1))]
    t = min(times) if best else sum(times) / len(times)
    cb = {"value": b / t, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{b} instances/step x {len(times)} timed step(s) (+{warmup} warm-up) of the same circuit, forward + weight "
                    f"gradient, complex128 torch oracle (per-gate ops on a (B,2^n) tensor + autograd, mirrors "
                    f"default.qubit.torch); {'best' if best else 'mean'} step {t:.2f} s, all steps "
                    f"{[round(v, 2) for v in times]} s",
          "sample_instances": b, "steps_timed": len(times)}
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
    K, Wm = max(args.steps, 1), max(args.warmup, 0)
    b = cpu_sample_size(args.cpu_instances, K + Wm)
    cb, t = cpu_oracle_rate(b, steps=K, warmup=Wm, best=False)
    cfg = workload_config(args.batch, args.gpus)
    cfg["sample_instances"] = b
    cfg["sample_note"] = (f"each CPU step simulates a bounded sample of {b} instances of the workload (the rate is per "
                          f"instance); instances_per_gpu / global_instances describe the GPU arm's step")
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": Wm, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "cpu_baseline": cb,
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
        dist.broadcast(net.weights.detach(), 0)      # detach() shares the version counter (cache invalidation)
    path_id = {"auto": L.PATH_AUTO, "gate": L.PATH_GATE, "gemm": L.PATH_GEMM}[args.path]
    spec = dataclasses.replace(net._spec(), path=path_id, gemm_precision=args.precision,
                               gemm_bwd_precision=args.bwd_precision)
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
    # the same e2e step with the fusion switched off (QIDDM_FUSED_STEP is read per call): the un-fused kernel sequence, for scale
    e2e_unfused = None
    if use_gemm and not args.no_extras and os.environ.get("QIDDM_FUSED_STEP", "1") != "0":
        os.environ["QIDDM_FUSED_STEP"] = "0"
        try:
            for _ in range(2):
                step_e2e()
            sync_all()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(3):
                step_e2e()
            e1.record()
            sync_all()
            ms_u = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms_u], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_u = t.item()
            e2e_unfused = {"value": imgs * TAU * world / (ms_u / 3 * 1e-3), "unit": UNIT, "ms_per_step": ms_u / 3,
                           "what": "the same Diffusion.forward step as `e2e` with QIDDM_FUSED_STEP=0 (ladder, prep_x, forward GEMM "
                                   "with out + Y stores, MSE pass, G pass, dW GEMM, adjoint)"}
        finally:
            os.environ.pop("QIDDM_FUSED_STEP", None)
    # per-kernel times of the e2e step (library timers, two extra steps outside the timed region)
    L.timing_enable(True)
    L.timing_collect()
    e2e_launches0 = qiddm_b200.launch_count()
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    e2e_launches = (qiddm_b200.launch_count() - e2e_launches0) // 2
    e2e_kinds = {k: round(v["ms"] / 2, 4) for k, v in L.timing_collect().items() if v["launches"]}
    L.timing_enable(False)
    # clocks sampled over both timed regions (device-resident steps and the end-to-end steps)
    clocks = sampler.stop(t_wall0, time.perf_counter()) if rank == 0 else None
    if clocks is not None and "sw_power_cap" in (clocks.get("reasons") or []):
        # nvidia-smi's polled clocks.sm does not resolve the cap: ncu reports 1.27-1.45 GHz inside the GEMM launches
        clocks["note"] = "power cap active: ncu shows 1.27-1.45 GHz SM clock inside the GEMM launches (profiles/r2b_summary.md)"

    # --- sustained rate: the same device-resident step back to back for >= 3 s (the "sustained" tensor peak is a 4 s loop)
    sustained = None
    if world == 1 and not args.no_extras:
        n_s = max(int(3000.0 / ms_step) + 1, K)
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(n_s):
            step_device()
        e1.record()
        torch.cuda.synchronize()
        ms_s = e0.elapsed_time(e1)
        sustained = {"value": B * n_s / (ms_s * 1e-3), "unit": UNIT, "steps": n_s, "seconds": ms_s * 1e-3,
                     "ms_per_step": ms_s / n_s}
    del x, go

    # --- the second half of BASELINE.json's metric: QIDDM train samples/s, data-parallel over the ranks (all ranks take part)
    sec = None
    if not args.no_secondary:
        sec = secondary_train(dev, world, rank, max(5, min(K, 10)), only=args.secondary.split(",") if args.secondary else None)

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
    bwdp = args.bwd_precision or args.precision
    if dom == "gemm":
        achieved = kd["work"] / (kd["ms"] * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this command
        # (profiles/r2_gemm_traffic.json, written by scripts/ncu_summary.py from the .ncu-rep): not measured in this run
        traffic, traffic_src = None, "no ncu capture committed for this configuration"
        tj = ROOT / "profiles" / "r2_gemm_traffic.json"
        if tj.exists():
            try:
                td = json.loads(tj.read_text())
                key = f"B{B}_p{args.precision}{bwdp}"
                if key in td:
                    traffic = float(td[key]["mean_bytes_per_launch"])
                    traffic_src = f"ncu capture {td[key]['source']} (per-launch bytes of forward / dX / dW: {td[key]['per_launch']})"
            except Exception as exc:      # a malformed side file must not kill the bench line
                traffic_src = f"profiles/r2_gemm_traffic.json unreadable: {exc}"
        exec_mult = (args.precision + 2 * bwdp) / 3.0
        roofline = {"kernel": "gemm_pair_kernel (tcgen05.mma.cta_group::2 kind::f16 + TMA, fused |Y|^2 readout epilogue)",
                    "bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": achieved / tc_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the step's three GEMM launches, from
                    # the committed `ncu --set full` capture of this command (profiles/r2b_gemm_pair_ncu_full_raw.csv via
                    # profiles/r2_gemm_traffic.json; not measured in this run -- `traffic_source` says which capture)
                    "traffic": traffic, "traffic_source": traffic_src,
                    "traffic_unit": "DRAM bytes/launch, mean of the step's GEMM launches; the algorithm's own bytes per launch "
                                    "(fp32 x / grad_out in, out / grad_in out, shared by the step's 3 launches): %.3g"
                                    % (4.0 * B * PIXELS * 4 / 3),
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback"),
                    "ms_per_launch": per_launch_ms, "launches": kd["launches"],
                    "algorithmic_flops": "2*M*N*K per GEMM (single pass); forward executes %dx, dX / dW %dx that"
                                         % (args.precision, bwdp),
                    "executed_tflops": achieved * exec_mult, "share_of_step": shares[dom]}
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
            "dtype": ("f32 (tcgen05 fp16 operands: x%d split forward, x%d split gradient GEMMs; fp32 accumulate)"
                      % (args.precision, bwdp)) if use_gemm else "f32",
            "data": "synthetic", "config": workload_config(B, world), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": imgs * PIXELS * 4,
                    "d2h_bytes_per_step": res_h.numel() * 8, "ms_per_step": ms_e / ke,
                    "api": "qiddm_b200.models.Diffusion(QDenseUndirected_old_noise(60,28)).forward(x, T=10): "
                           f"{imgs} pinned host images/step -> {imgs * TAU} circuit instances, loss+grad to host; "
                           "H2D double-buffered on a copy stream (qiddm_b200.train.DevicePrefetcher), one copy per step; "
                           + ("fused step (qiddm_dense_mse_step)" if os.environ.get("QIDDM_FUSED_STEP", "1") != "0" and use_gemm
                              else "unfused kernel sequence"),
                    "kernel_ms_per_step": e2e_kinds, "gpu_launches_per_step": int(e2e_launches)},
            "path": ("gemm_x%d_bwd_x%d" % (args.precision, bwdp)) if use_gemm else "gate",
            "gpu_launches": int(launches), "roofline": roofline}

    if sec is not None:
        line["secondary"] = sec
    if not args.no_extras:
        line["extras"] = extras(dev) if world == 1 else {}
        if sustained is not None:
            line["extras"]["sustained_3s"] = sustained
        if e2e_unfused is not None:
            line["extras"]["e2e_unfused"] = e2e_unfused
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_oracle_rate(cpu_sample_size(args.cpu_instances, 4), steps=3, warmup=1, best=True)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def secondary_train(dev, world, rank, steps, only=None):
    """QIDDM train samples/s (BASELINE.json metric, second half; configs 1 and 4) as the loop body of
    src/mnist_exm.py:175-182 captured in CUDA graphs (qiddm_b200.train.GraphedTrainStep), weak-scaled over the ranks:
    every rank trains on its own images, ONE flat-bucket NCCL all-reduce of all gradients per step.  The gate kernels'
    roofline is FP32-FMA: algorithmic flops (14 * 2^n per Rot, adjoint = 3x forward) over the kernels'
    own CUDA-event time from an eager pass, against the FP32 peak measured here by qiddm_probe_fp32_fma."""
    import torch
    import torch.distributed as dist
    from qiddm_b200 import _lib as L
    from qiddm_b200 import models, noise
    from qiddm_b200 import nn as qnn
    from qiddm_b200.train import DataParallelTrainer, GraphedTrainStep

    ev = lambda: torch.cuda.Event(enable_timing=True)
    fp32_peak = L.fp32_fma_peak_tflops(dev) if rank == 0 else None                       # burst: the harder denominator
    fp32_sustained = L.fp32_fma_peak_tflops(dev, sustained=True) if rank == 0 else None
    # (key, model, constructor, images per GPU and step, goal, pca_group, lr, side, reference)
    cfgs = [("config1", "QIDDM_LL_noise(784,6,14,2)", lambda: qnn.QIDDM_LL_noise(784, 6, 14, 2), 4096, "data", None, 0.0255, SIDE,
             "src/mnist_exm.py:46,139"),
            ("config4", "QIDDM_PL_noise(784,8,6,2)", lambda: qnn.QIDDM_PL_noise(784, 8, 6, 2), 1024, "noise", 10, 0.01, SIDE,
             "src/emnist_exm.py:45"),
            ("config3", "UNetUndirected(3,8,3) (QConv-UNet, 5 782 circuits per image-forward)", lambda: qnn.UNetUndirected(3, 8, 3),
             64, "data", None, 1e-3, SIDE, "src/fashion_exm.py, nn/unet.py:119-160"),
            ("config5", "QIDDM_PL_noise(4096,8,6,2) on 64x64", lambda: qnn.QIDDM_PL_noise(4096, 8, 6, 2), 256, "noise", 10, 0.01, 64,
             "src/bloodmnist.py:47")]
    if only:
        cfgs = [c for c in cfgs if c[0] in only]
    out = {"metric": "qiddm_train_samples_per_sec", "unit": "train-samples/s", "n_gpus": world, "tau": 10,
           "scaling": "weak", "fp32_peak_tflops_measured": fp32_peak, "fp32_sustained_tflops_measured": fp32_sustained,
           "what": "whole-job images/s through Diffusion training steps (noise ladder -> net -> MSE -> adjoint backward -> "
                   "flat-bucket all-reduce -> Adam), CUDA-graph replays, float64 module I/O, fp32 simulation"}
    for key, name, make, imgs, goal, pca_group, lr, side, src in cfgs:
        torch.manual_seed(0)
        net = make()
        if pca_group:
            net.pca_group = pca_group        # one PCA per image's tau-ladder = the reference's batch-1 semantics
        diff = models.Diffusion(net, noise.add_normal_noise_multiple, goal, (side, side), torch.nn.MSELoss()).to(dev, torch.float64)
        opt = torch.optim.Adam(diff.parameters(), lr=lr, capturable=True)
        trainer = DataParallelTrainer(diff, opt, tau=10)
        trainer.broadcast_parameters()
        torch.manual_seed(100 + rank)
        x = torch.rand(imgs, side * side, device=dev, dtype=torch.float64)
        trainer.step(x, already_sharded=True)                    # eager warm-up
        L.timing_enable(True)
        L.timing_collect()
        n0 = L.launch_count()
        a, b = ev(), ev()
        a.record()
        for _ in range(2):
            trainer.step(x, already_sharded=True)
        b.record()
        torch.cuda.synchronize()
        kinds = L.timing_collect()
        L.timing_enable(False)
        launches = (L.launch_count() - n0) // 2
        eager_ms = a.elapsed_time(b) / 2
        gs = GraphedTrainStep(diff, opt, 10, x, allreduce=world > 1)
        for _ in range(3):
            gs.step(x)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(steps):
            gs.step(x)
        b.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        gf, gb = kinds["gate_forward"], kinds["gate_backward"]
        roof = None
        if rank == 0 and gf["launches"] and gb["launches"]:
            tf_f = gf["work"] / (gf["ms"] * 1e-3) / 1e12
            tf_b = gb["work"] / (gb["ms"] * 1e-3) / 1e12
            tf_all = (gf["work"] + gb["work"]) / ((gf["ms"] + gb["ms"]) * 1e-3) / 1e12
            roof = {"kernel": "gate_kernel (forward + adjoint backward)", "bound": "fp32", "unit": "TFLOP/s",
                    "achieved": tf_all, "peak": fp32_peak, "frac": tf_all / fp32_peak if fp32_peak else None,
                    "peak_source": "qiddm_probe_fp32_fma, this run: burst rate of packed FFMA2 chains (best of five 0.7 ms launches); "
                                   "a 50 ms launch of the same loop is power-capped at fp32_sustained_tflops_measured",
                    "forward": {"achieved": tf_f, "frac": tf_f / fp32_peak, "ms_per_step": gf["ms"] / 2},
                    "backward": {"achieved": tf_b, "frac": tf_b / fp32_peak, "ms_per_step": gb["ms"] / 2,
                                 "note": "algorithmic flops of the adjoint counted as 3x the forward (un-apply on psi, apply-dagger "
                                         "on lambda, inner products); psi_final comes from the forward launch, no recomputation"},
                    "share_of_eager_step": (gf["ms"] + gb["ms"]) / 2 / eager_ms}
        cf, cb = kinds["conv_forward"], kinds["conv_backward"]
        if rank == 0 and cf["launches"] and cb["launches"] and cf["ms"] + cb["ms"] > gf["ms"] + gb["ms"]:
            # QConv networks: the direct-convolution kernels (csrc/qiddm_conv.cu) carry the step, not the gate kernels
            tf_f = cf["work"] / (cf["ms"] * 1e-3) / 1e12
            tf_b = cb["work"] / (cb["ms"] * 1e-3) / 1e12
            tf_all = (cf["work"] + cb["work"]) / ((cf["ms"] + cb["ms"]) * 1e-3) / 1e12
            roof = {"kernel": "conv_fwd_kernel / conv_grad + conv_bwd_data + conv_bwd_w kernels (QConv2d as a direct fp32 convolution)",
                    "bound": "fp32", "unit": "TFLOP/s", "achieved": tf_all, "peak": fp32_peak,
                    "frac": tf_all / fp32_peak if fp32_peak else None,
                    "peak_source": "qiddm_probe_fp32_fma, this run (burst rate of packed FFMA2 chains)",
                    "algorithmic_flops": "2 F N per patch and pass (F = C k k features, N = 2 out_channels rows of U); backward = image "
                                         "gradient + weight gradient (the first layer has no image gradient)",
                    "forward": {"achieved": tf_f, "frac": tf_f / fp32_peak, "ms_per_step": cf["ms"] / 2},
                    "backward": {"achieved": tf_b, "frac": tf_b / fp32_peak, "ms_per_step": cb["ms"] / 2},
                    "share_of_eager_step": (cf["ms"] + cb["ms"]) / 2 / eager_ms}
        kshares = {k: round(v["ms"] / 2 / eager_ms, 4) for k, v in kinds.items() if v["launches"] and v["ms"] / 2 / eager_ms >= 0.01}
        evals_per_image = {"config3": 10 * 5782}.get(key, 10 * 2)         # tau x (circuits per image-forward)
        out[key] = {"model": name, "reference": src, "goal": goal, "images_per_gpu": imgs, "value": imgs * world / (ms * 1e-3),
                    "ms_per_step": ms, "steps": steps, "circuit_evals_per_s": imgs * world * evals_per_image / (ms * 1e-3),
                    "kernel_shares_of_eager_step": kshares,
                    "eager_ms_per_step": eager_ms, "gpu_launches_per_step": int(launches), "roofline": roof,
                    "allreduce": "one flat fp64 bucket of all gradients per step (NCCL; the gradients are views of the bucket), "
                                 "between two graph replays" if world > 1 else None}
        del gs, trainer, opt, diff, net, x
        torch.cuda.empty_cache()
    return out


def extras(dev):
    """Side numbers (not the headline): forward-only rate of the bench layer and the config-1 step at the reference batch size."""
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
    out["configs_note"] = ("config 1 = QIDDM_LL_noise(784,6,14,2) at the REFERENCE batch (1 image x tau 10, src/mnist_exm.py:144), one graph "
                           "replay per step; the throughput-sized configs 1 / 3 / 4 / 5 are in `secondary`")
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

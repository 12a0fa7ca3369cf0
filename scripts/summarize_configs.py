#!/usr/bin/env python
"""Turns the JSON lines of scripts/bench_configs.py into the markdown tables kept under profiles/.
  python scripts/summarize_configs.py gpurun_out/configs_full.jsonl > profiles/r1_configs.md"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
env = next((r for r in rows if r["what"] == "env"), {})
print("# BASELINE.json configs 1-5 and the config-2 sweep on one B200 (`scripts/bench_configs.py`)\n")
print(f"GPU: {env.get('gpu')}, torch {env.get('torch')}, host cores {env.get('cores')}.  CUDA-event times after warm-up; "
      "`graph` = the step / sampler replayed from a CUDA graph (`qiddm_b200.train`).  The headline line stays `bench.py`.\n")


def fmt(v, nd=3):
    if v is None:
        return "—"
    if isinstance(v, float):
        return f"{v:,.{nd}f}"
    return f"{v:,}" if isinstance(v, int) else str(v)


def table(title, what, cols):
    sel = [r for r in rows if r["what"] == what]
    if not sel:
        return
    print(f"## {title}\n")
    print("| " + " | ".join(c[0] for c in cols) + " |")
    print("|" + "---|" * len(cols))
    for r in sel:
        print("| " + " | ".join(fmt(c[1](r)) for c in cols) + " |")
    print()


step_cols = [("model", lambda r: r.get("model")), ("images/step", lambda r: r.get("images_per_step")),
             ("graph", lambda r: "yes" if r.get("cuda_graph") else "no"), ("dtype", lambda r: r.get("module_dtype", "float64")),
             ("ms/step", lambda r: r.get("ms_per_step")),
             ("train samples/s", lambda r: r.get("train_samples_per_s")),
             ("circuit evals/s", lambda r: r.get("circuit_evals_per_s")),
             ("gate fwd+bwd ms", lambda r: round(r.get("kernels", {}).get("gate_forward", 0) + r.get("kernels", {}).get("gate_backward", 0), 3) or None),
             ("GEMM ms", lambda r: r.get("kernels", {}).get("gemm")),
             ("CPU oracle ms/step", lambda r: r.get("cpu_oracle_ms_per_step"))]
table("Config 1 — `src/mnist_exm.py` training step (tau = 10, Adam, float64 module I/O)", "config1", step_cols)
table("Config 3 — `UNetUndirected(3,8,3)` QConv-UNet training step, 28x28 (5 782 circuits per image-forward)", "config3_unet28", step_cols[:-1])
table("Config 3 — `UNetUndirectedS(3,8,3)`", "config3_unet28_simple", step_cols[:-1])
table("Config 4 — `QIDDM_PL_noise(784,8,6,2)` training step (goal noise, PCA re-fit per forward)", "config4_train",
      step_cols[:6] + [("pca", lambda r: r.get("pca"))])
table("Config 4 — `Diffusion.sample` fixed-point sampler", "config4_sample",
      [("model", lambda r: r["model"]), ("images", lambda r: r["images"]), ("iterations", lambda r: r["n_iters"]),
       ("graph", lambda r: "yes" if r.get("cuda_graph") else "no"), ("pca group", lambda r: r.get("pca_group")),
       ("seconds", lambda r: r["seconds"]),
       ("iterations/s", lambda r: r["sampler_iters_per_s"]), ("circuit evals/s", lambda r: r["circuit_evals_per_s"])])
table("Config 5 — 64x64: `QDenseUndirected_old(60,64)` (n = 12)", "config5_qdense64", step_cols[:-1])
table("Config 5 — 64x64: `QIDDM_PL_noise(4096,8,6,2)`", "config5_pl64", step_cols[:6])
table("Config 5 — 64x64: `UNetUndirected(3,8,3)`", "config5_unet64", step_cols[:-1])

sw = [r for r in rows if r["what"] == "sweep"]
if sw:
    print("## Config 2 — stage sweep (forward + adjoint backward of ONE stage, fp32 on device)\n")
    print("`qdense` = AmplitudeEmbedding + SEL(CNOT, depth layers) + probs (a1); `reupload` = 6 blocks of RZ re-upload + "
          "2-layer SEL(CZ) + <Z> (a4).  `gate TF/s` = algorithmic forward rate (14 flop per Rot and amplitude) of the gate "
          "kernel; nominal fp32 peak 148 SM x 128 FMA x 2 x 1.965 GHz = 74.4 TFLOP/s.  `alg GB/s` = 4(2 n_in + 2 n_out) bytes "
          "per instance over the fwd+bwd time (HBM peak 6 548 GB/s).\n")
    print("| family | n | depth | batch | path | fwd ms | fwd+bwd ms | fwd+bwd evals/s | gate TF/s (fwd) | alg GB/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in sw:
        print(f"| {r['family']} | {r['n']} | {r['depth']} | {r['batch']:,} | {r['path']} | {r['fwd_ms']:.3f} | "
              f"{r['fwd_bwd_ms']:.3f} | {r['fwd_bwd_evals_per_s']:,} | {fmt(r['gate_alg_tflops_fwd'], 1)} | {r['alg_hbm_gbs_fwd_bwd']:,.0f} |")
    print()
    # winners per (family, n, depth, batch)
    print("### Faster path per point (what `Plan.use_gemm`'s cost model has to reproduce)\n")
    best = {}
    for r in sw:
        if r["family"] != "qdense":
            continue
        k = (r["n"], r["depth"], r["batch"])
        if k not in best or r["fwd_bwd_ms"] < best[k][1]:
            best[k] = (r["path"], r["fwd_bwd_ms"])
    ns = sorted({k[0] for k in best})
    print("| n | depth | " + " | ".join(f"B={b:,}" for b in sorted({k[2] for k in best})) + " |")
    print("|---|---|" + "---|" * len({k[2] for k in best}))
    for n in ns:
        for d in sorted({k[1] for k in best if k[0] == n}):
            cells = []
            for b in sorted({k[2] for k in best}):
                cells.append(best.get((n, d, b), ("—", 0))[0])
            print(f"| {n} | {d} | " + " | ".join(cells) + " |")

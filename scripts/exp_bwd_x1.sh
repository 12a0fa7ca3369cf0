#!/bin/bash
# Round-2 experiment: single-pass gradient GEMMs behind the fp32-grade forward.
#   gpurun --timeout 400 -- 'bash scripts/exp_bwd_x1.sh'
# Prints the measured output / gradient errors against the oracle with the knob on, then a same-box A/B of the bench step.
set -u
mkdir -p gpurun_out
echo "== accuracy, QIDDM_GEMM_BWD_X1=1 (expected: outputs unchanged, gradients ~3e-4 rel-to-max) =="
QIDDM_GEMM_BWD_X1=1 python scripts/diag_gemm_accuracy.py 2>&1 | tee gpurun_out/bwd_x1_accuracy.txt
for v in 0 1 0 1; do
    QIDDM_GEMM_BWD_X1=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bwd_x1_$v.json 2>/dev/null
    python - "$v" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bwd_x1_{sys.argv[1]}.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernel_ms_per_step"]
print(f"BWD_X1={sys.argv[1]}: {d['value'] / 1e6:.2f} M evals/s, {d['ms_per_step']:.2f} ms/step, e2e {d['e2e']['value'] / 1e6:.2f} M; "
      f"gemm fwd {k['gemm_forward']} dX {k['gemm_dx']} dW {k['gemm_dw']} grad_y {k['grad_y']}")
PY
done

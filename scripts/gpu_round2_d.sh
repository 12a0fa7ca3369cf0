#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-secondary"
$CMD > gpurun_out/d_plain.json 2> gpurun_out/d_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 9 -c 3 -f -o gpurun_out/r2_gemm_pair $CMD > gpurun_out/d_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_gemm_launches.csv $CMD > gpurun_out/d_ncu_list.log 2>&1
echo "ncu list rc=$?"
ls -la gpurun_out/*.ncu-rep

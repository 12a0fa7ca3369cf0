#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time timeout 600 python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err ) 2> gpurun_out/i_bench.time; echo "bench rc=$?"; cat gpurun_out/i_bench.time
timeout 900 python scripts/bench_configs.py --what sweep > gpurun_out/i_sweep.jsonl 2> gpurun_out/i_sweep.err; echo "sweep rc=$?"; wc -l gpurun_out/i_sweep.jsonl
CMD1="python scripts/run_step.py qiddm_ll 4096 3"
CMD2="python scripts/run_step.py qiddm_pl 1024 3"
$CMD1 > gpurun_out/i_plain1.log 2>&1 && $CMD2 > gpurun_out/i_plain2.log 2>&1 && \
ncu --set full --clock-control none -k regex:gate_kernel -s 4 -c 4 -f -o gpurun_out/r2_gate_ll $CMD1 > gpurun_out/i_ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none -k regex:gate_kernel -s 4 -c 4 -f -o gpurun_out/r2_gate_pl $CMD2 > gpurun_out/i_ncu2.log 2>&1
echo "ncu2 rc=$?"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_glue.py tests/test_gpu_modules.py tests/test_gpu_graphs.py tests/test_gpu_fused_step.py -x -q > gpurun_out/c13_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/c13_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 5 --warmup 3 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo "bench rc=$?"
QIDDM_MSE_LADDER=0 timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 5 --warmup 3 > gpurun_out/c13_bench_off.json 2> gpurun_out/c13_bench_off.err; echo "bench off rc=$?"

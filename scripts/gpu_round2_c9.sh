#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_direct.py tests/test_gpu_round2.py tests/test_gpu_graphs.py -x -q > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/c9_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-extras --secondary config3 --steps 3 --warmup 3 > gpurun_out/c9_bench3.json 2> gpurun_out/c9_bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/c9_bench3.err
QIDDM_UPCONV_FUSION=0 timeout 300 python bench.py --no-cpu-baseline --no-extras --secondary config3 --steps 3 --warmup 3 > gpurun_out/c9_bench3_off.json 2> gpurun_out/c9_bench3_off.err; echo "bench off rc=$?"

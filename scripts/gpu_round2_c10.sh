#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_direct.py -x -q > gpurun_out/c10_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c10_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-extras --secondary config3 --steps 3 --warmup 3 > gpurun_out/c10_bench3.json 2> gpurun_out/c10_bench3.err; echo "bench rc=$?"
timeout 300 python scripts/profile_unet.py 64 > gpurun_out/c10_prof_unet_on.txt 2>&1; echo "prof on rc=$?"

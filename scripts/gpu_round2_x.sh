#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_glue.py tests/test_gpu_modules.py tests/test_gpu_graphs.py tests/test_gpu_round2.py -x -q > gpurun_out/x_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/x_tests.log
python scripts/profile_step.py qiddm_ll 4096 2>&1 | cut -c1-105,150- | head -30 | tee gpurun_out/x_prof_ll.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --secondary config1,config4,config5 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py -x -q -k "tail or ladder or mse or diffusion" > gpurun_out/c16_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/c16_tests.log | cut -c1-200
QIDDM_FUSED_TAIL=1 timeout 300 python scripts/profile_step.py qiddm_ll 4096 > gpurun_out/c16_prof_ll.txt 2>&1; echo "prof rc=$?"
QIDDM_FUSED_TAIL=1 timeout 300 python bench.py --no-cpu-baseline --no-extras --secondary config1,config4,config5 --steps 5 --warmup 3 > gpurun_out/c16_bench_tail.json 2> gpurun_out/c16_bench_tail.err; echo "bench rc=$?"

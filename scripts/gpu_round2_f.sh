#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_step.py -x -q -k "unfused" > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/f_tests.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_step.py -x -q > gpurun_out/u_tests.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/u_tests.log
python scripts/fused_step_probe.py 2>&1 | tee gpurun_out/u_probe.log

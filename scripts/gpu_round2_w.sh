#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/w_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/w_tests.log
python scripts/fused_step_probe.py 2>&1 | tee gpurun_out/w_probe.log

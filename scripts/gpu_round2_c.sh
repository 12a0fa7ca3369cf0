#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
grep -n "unet end-to-end\|passed\|failed\|^FAILED\|^E  " gpurun_out/c_pytest.log | head -40
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"

#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
timeout 900 python scripts/measure_parity_margins.py > gpurun_out/b_margins.log 2>&1; echo "margins rc=$?"
timeout 300 python scripts/measure_tc_bias.py > gpurun_out/b_bias.log 2>&1; echo "bias rc=$?"; cat gpurun_out/b_bias.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py -m gpu -q > gpurun_out/k_pytest.log 2>&1; tail -2 gpurun_out/k_pytest.log
CMD="python scripts/run_step.py unet 64 4"
$CMD > gpurun_out/k_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_unet_launches.csv $CMD > gpurun_out/k_ncu.log 2>&1
echo "ncu rc=$?"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
bash scripts/gemm_knob_sweep.sh "" "QIDDM_GEMM_L2_HINTS=1" "QIDDM_GEMM_L2_HINTS=2" "QIDDM_GEMM_L2_HINTS=3" "QIDDM_GEMM_L2_HINTS=7" "QIDDM_GEMM_STAGES=2" "" 2>&1 | tee gpurun_out/e_sweep.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/e_pytest.log 2>&1; tail -3 gpurun_out/e_pytest.log

#!/bin/bash
# dual-N dW GEMM + provisional G scale: parity first, then the knob comparison
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_big_batch.py tests/test_gpu_gemm_path.py -x -q > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/q_tests.log
bash scripts/gemm_knob_sweep.sh "" "QIDDM_GEMM_DW_BK=64" "QIDDM_GEMM_DW_DUAL=0" "QIDDM_GEMM_GSAMPLE=1" "" 2>&1 | tee gpurun_out/q_sweep.log

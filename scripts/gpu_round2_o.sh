#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
bash scripts/gemm_knob_sweep.sh "" "QIDDM_GEMM_SKIP_LOADS=1" "QIDDM_GEMM_SKIP_LOADS=2" "QIDDM_GEMM_SKIP_LOADS=3" "" 2>&1 | tee gpurun_out/o_sweep.log

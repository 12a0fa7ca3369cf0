#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
bash scripts/gemm_knob_sweep.sh "" "QIDDM_GEMM_DW_SPLITS=74" "QIDDM_GEMM_DW_SPLITS=148" "" 2>&1 | tee gpurun_out/s2_sweep.log
for sp in 0 16 32 64; do QIDDM_GEMM_DW_SPLITS=$sp python scripts/dual_check.py 2>&1 | sed "s/^/splits=$sp /"; done | tee gpurun_out/s2_err.log

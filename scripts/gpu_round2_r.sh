#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(python scripts/dual_check.py; QIDDM_GEMM_DW_DUAL=0 python scripts/dual_check.py; QIDDM_GEMM_DW_BK=64 python scripts/dual_check.py) 2>&1 | tee gpurun_out/r_dual.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
for bn in 0 224 208 176 160 128 112 256; do TAG="BN=$bn" QIDDM_GEMM_BN=$bn python scripts/gemm_fwd_probe.py; done
TAG="TMA_EPI=0" QIDDM_GEMM_TMA_EPI=0 python scripts/gemm_fwd_probe.py
TAG="TMA_EPI=0 NOSTORE" QIDDM_GEMM_TMA_EPI=0 QIDDM_GEMM_NOSTORE=1 python scripts/gemm_fwd_probe.py
} 2>&1 | tee gpurun_out/f_probe.log
bash scripts/gemm_knob_sweep.sh "" "QIDDM_GEMM_VARTAIL=0" "" "QIDDM_GEMM_VARTAIL=0" 2>&1 | tee gpurun_out/f_sweep.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; tail -3 gpurun_out/f_pytest.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python scripts/fused_step_probe.py 2>&1 | tee gpurun_out/v_probe.log
CMD="python scripts/fused_step_once.py 52428 4"
$CMD > gpurun_out/v_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 4 -c 2 -f -o gpurun_out/r2_fused_fwd $CMD > gpurun_out/v_ncu.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep

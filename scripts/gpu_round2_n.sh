#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python scripts/run_step.py unet 64 2"
$CMD > gpurun_out/n_plain.log 2>&1 && ncu --set full --clock-control none -k regex:'prep_xt_unfold|fold_rows|grad_y_kernel|g_bound' -s 36 -c 36 -f -o gpurun_out/r2_unet_aux $CMD > gpurun_out/n_ncu.log 2>&1
echo "ncu rc=$?"

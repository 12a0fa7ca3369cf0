#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python scripts/run_step.py qiddm_ll 4096 2"
$CMD > gpurun_out/c18_plain.log 2>&1 || exit 1
timeout 240 ncu --set full --clock-control none -k regex:'tail_|mse_ladder|noise_ladder' --kernel-name-base demangled -s 6 -c 6 -f -o /tmp/r2c_tail $CMD > gpurun_out/c18_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2c_tail.ncu-rep --page raw --csv > gpurun_out/r2c_tail_raw.csv 2>/dev/null
ls -la gpurun_out/r2c_tail_raw.csv

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python scripts/run_step.py unet 64 2"
$CMD > gpurun_out/g4_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'gate_kernel<\(int\)4|gate_kernel<4' --kernel-name-base demangled -c 4 -f -o gpurun_out/r2b_gate_n4 $CMD > gpurun_out/g4_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/g4_ncu.log
ls -la gpurun_out/r2b_gate_n4.ncu-rep

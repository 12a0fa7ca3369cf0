#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests/test_gpu_glue.py tests/test_gpu_conv_direct.py -x -q > gpurun_out/c2_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/c2_tests.log
python bench.py --no-cpu-baseline --no-extras --secondary config3 --steps 3 --warmup 3 > gpurun_out/c2_bench3.json 2> gpurun_out/c2_bench3.err; echo "bench rc=$?"
CMD="python scripts/run_step.py unet 64 2"
$CMD > gpurun_out/c2_plain.log 2>&1 || exit 1
# second step: fwd L2 (8->8) = 44; backward of the last 8->8 layer: grad 57, data 58, w 59; 16->8: data 61, w 62
ncu --set full --clock-control none --import-source on -k regex:'conv_(fwd|bwd|grad)' --kernel-name-base demangled -s 44 -c 1 -f -o /tmp/r2c_conv_fwd $CMD > gpurun_out/c2_ncu1.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'conv_(fwd|bwd|grad)' --kernel-name-base demangled -s 57 -c 6 -f -o /tmp/r2c_conv_bwd $CMD > gpurun_out/c2_ncu2.log 2>&1; echo "ncu rc=$?"
for f in fwd bwd; do
  ncu -i /tmp/r2c_conv_$f.ncu-rep --page raw --csv > gpurun_out/r2c_conv_${f}_raw.csv 2>/dev/null
  ncu -i /tmp/r2c_conv_$f.ncu-rep --page source --csv > gpurun_out/r2c_conv_${f}_source.csv 2>/dev/null
done

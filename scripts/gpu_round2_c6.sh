#!/bin/bash
# evidence of the direct QConv path: smoke, full GPU suite, default bench + reference arm, ncu launch list of the UNet step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c6_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c6_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c6_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c6_tests.log
python bench.py > gpurun_out/c6_bench_default.json 2> gpurun_out/c6_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c6_bench_reference.json 2> gpurun_out/c6_bench_reference.err; echo "ref rc=$?"
CMD="python scripts/run_step.py unet 64 3"
$CMD > gpurun_out/c6_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2c_unet_launches.csv $CMD > gpurun_out/c6_ncu1.log 2>&1; echo "launch list rc=$?"

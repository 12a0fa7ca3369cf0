#!/bin/bash
# final evidence of the round: default bench line, reference arm, ncu launch list + full captures of the GEMM launches of the
# device-resident step and of the fused e2e step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python bench.py > gpurun_out/z_bench_default.json 2> gpurun_out/z_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-secondary"
$CMD > gpurun_out/z_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_gemm_launches.csv $CMD > gpurun_out/z_ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 9 -c 3 -f -o gpurun_out/r2b_gemm_pair $CMD > gpurun_out/z_ncu2.log 2>&1; echo "ncu gemm rc=$?"
CMD2="python scripts/fused_step_once.py 52428 4"
$CMD2 > gpurun_out/z_plain2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'gemm_pair_kernel|ladder_prep' -s 3 -c 3 -f -o gpurun_out/r2b_fused_step $CMD2 > gpurun_out/z_ncu3.log 2>&1; echo "ncu fused rc=$?"
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
(python scripts/gscale_check.py; QIDDM_GEMM_GSAMPLE=1 python scripts/gscale_check.py) 2>&1 | tee gpurun_out/s_gscale.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-secondary"
for v in 64d 0; do
  if [ $v == 64d ]; then export QIDDM_GEMM_DW_BK=64; unset QIDDM_GEMM_DW_DUAL; else export QIDDM_GEMM_DW_DUAL=0; fi
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum --clock-control none -k regex:gemm_pair_kernel -s 9 -c 3 --csv --log-file gpurun_out/s_ncu_dw_$v.csv $CMD > gpurun_out/s_ncu_$v.log 2>&1
  echo "ncu $v rc=$?"
done

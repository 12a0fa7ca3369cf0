#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py -x -q -k skinny 2>&1 | tail -5
python scripts/linear_probe.py 2>&1 | tee gpurun_out/y_probe.log
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_read.sum --clock-control none -k regex:'narrow_out|outer_kernel|wide_out_kernel' -s 9 -c 5 --csv --log-file gpurun_out/y_ncu_raw.csv python scripts/linear_probe.py > gpurun_out/y_ncu.log 2>&1
grep -v "^==" gpurun_out/y_ncu_raw.csv | cut -d, -f5,13- | head -20

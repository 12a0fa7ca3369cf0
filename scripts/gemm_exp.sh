#!/bin/bash
# GEMM-path experiments (env knobs of qiddm_gemm.cu)
run() { echo "== $* $EXTRA"; env "$@" timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; print({x:k[x] for x in ('gemm_forward','gemm_dx','gemm_dw')}, 'ms/step', round(d['ms_per_step'],3))"; }
run QIDDM_GEMM_BK32=1 QIDDM_GEMM_TMA_EPI=1
run QIDDM_GEMM_BK32=1 QIDDM_GEMM_TMA_EPI=0
run QIDDM_GEMM_BK32=0 QIDDM_GEMM_TMA_EPI=0
run QIDDM_GEMM_BK32=0 QIDDM_GEMM_TMA_EPI=1
run QIDDM_GEMM_BK32=0 QIDDM_GEMM_NOSTORE=1 QIDDM_GEMM_TMA_EPI=0
EXTRA="--precision 1" run QIDDM_GEMM_BK32=0 QIDDM_GEMM_TMA_EPI=1

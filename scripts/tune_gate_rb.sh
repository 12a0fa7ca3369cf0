#!/bin/bash
# Times the gate-path forward/backward kernels for each register-tile width (B200).
for rb in 3 4 5; do
  QIDDM_RB_FWD=$rb QIDDM_RB_BWD=$rb timeout 300 python bench.py --path gate --batch 16384 --steps 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; print('n=10 RB=$rb', 'fwd ms', k.get('gate_forward'), 'bwd ms', k.get('gate_backward'), 'evals/s', round(d['value']))"
done

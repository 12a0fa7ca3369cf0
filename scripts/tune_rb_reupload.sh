#!/bin/bash
# Register-tile width sweep of the gate kernels on the re-upload family (B200): QIDDM_RB_FWD / QIDDM_RB_BWD.
for n in 6 7 8 9 10; do for rb in 3 4 5; do
  QIDDM_RB_FWD=$rb QIDDM_RB_BWD=$rb timeout 120 python scripts/bench_stage.py --family reupload --n $n --L 6 --batch 131072 --iters 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n=$n RB=$rb fwd', d['fwd_ms'], 'bwd', d['bwd_ms'])"
done; done

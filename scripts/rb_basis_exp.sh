for f in 5 4 3; do for b in 5 4 3; do
QIDDM_RB_FWD=$f QIDDM_RB_BWD=$b timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; print('RBF=$f RBB=$b', 'gate fwd', k.get('gate_forward'), 'bwd', k.get('gate_backward'), 'step', round(d['ms_per_step'],3))"
done; done

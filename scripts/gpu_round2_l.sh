#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/l_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; tail -3 gpurun_out/l_pytest.log; grep -n "^FAILED\|^E  " gpurun_out/l_pytest.log | head
timeout 300 python bench.py --no-cpu-baseline --no-extras --secondary config3 --steps 5 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/l_bench.json').read().strip().splitlines()[-1])
c=d['secondary']['config3']; print('config3', round(c['value']), 'samples/s', round(c['ms_per_step'],3), 'eager', round(c['eager_ms_per_step'],3), 'launches', c['gpu_launches_per_step'])
print('step', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])
PY

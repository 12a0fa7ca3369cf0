#!/bin/bash
# direct-convolution QConv: memcheck, parity tests, UNet step with and without it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python scripts/conv_direct_check.py > gpurun_out/c1_plain.log 2>&1; echo "plain rc=$?"; tail -5 gpurun_out/c1_plain.log
timeout 600 compute-sanitizer --tool memcheck python scripts/conv_direct_check.py > gpurun_out/c1_san.log 2>&1; echo "san rc=$?"; tail -4 gpurun_out/c1_san.log
timeout 900 python -m pytest tests/test_gpu_conv_direct.py -x -q > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/c1_tests.log
QIDDM_QCONV_DIRECT=0 python scripts/profile_unet.py 64 > gpurun_out/c1_prof_unet_off.txt 2>&1; echo "prof off rc=$?"
python scripts/profile_unet.py 64 > gpurun_out/c1_prof_unet_on.txt 2>&1; echo "prof on rc=$?"
python bench.py --no-cpu-baseline --no-extras --secondary config3 --steps 3 --warmup 3 > gpurun_out/c1_bench3.json 2> gpurun_out/c1_bench3.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/c1_bench3.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('secondary',{}).get('config3'),indent=0)[:1500])
P

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-extras --secondary config1,config4 --steps 5 --warmup 3 > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; echo "bench rc=$?"

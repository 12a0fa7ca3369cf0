#!/bin/bash
# re-entry baseline: the GPU test suite and the default bench line on the restored tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/p_tests.log
timeout 600 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?"
head -c 600 gpurun_out/p_bench.json

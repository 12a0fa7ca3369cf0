#!/bin/bash
# full GPU suite + default bench line with the direct QConv path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c3_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/c3_tests.log
python bench.py > gpurun_out/c3_bench_default.json 2> gpurun_out/c3_bench_default.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c3_smoke.log

#!/bin/bash
# final state of the round: smoke, full GPU suite, default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c11_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c11_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c11_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c11_tests.log
python bench.py > gpurun_out/c11_bench_default.json 2> gpurun_out/c11_bench_default.err; echo "bench rc=$?"

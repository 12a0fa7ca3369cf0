#!/bin/bash
# round-2 GPU call A: tests, parity margins, bench lines
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 900 python scripts/measure_parity_margins.py > gpurun_out/a_margins.log 2>&1; echo "margins rc=$?"
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --bwd-precision 1 --no-cpu-baseline --no-extras --no-secondary > gpurun_out/a_bench_bwd1.json 2> gpurun_out/a_bench_bwd1.err; echo "bench bwd1 rc=$?"
timeout 300 python bench.py --precision 1 --no-cpu-baseline --no-extras --no-secondary > gpurun_out/a_bench_p1.json 2> gpurun_out/a_bench_p1.err; echo "bench p1 rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "bench ref rc=$?"
tail -c 1500 gpurun_out/a_bench.json

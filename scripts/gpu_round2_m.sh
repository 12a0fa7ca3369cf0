#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-8}
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err ) 2> gpurun_out/m_time.txt; echo "rc=$?"; cat gpurun_out/m_time.txt | tail -4
tail -c 300 gpurun_out/m_bench_n$N.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/m_bench_ref_n$N.json 2> gpurun_out/m_bench_ref_n$N.err ) 2>> gpurun_out/m_time.txt; echo "ref rc=$?"

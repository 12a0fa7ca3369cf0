#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "n8 rc=$?"; tail -c 300 gpurun_out/n8_bench.err

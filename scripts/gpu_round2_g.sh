#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
{
TAG="default" python scripts/gemm_fwd_probe.py
TAG="STCS" QIDDM_GEMM_L2_HINTS=8 python scripts/gemm_fwd_probe.py
TAG="STCS+Wlast" QIDDM_GEMM_L2_HINTS=10 python scripts/gemm_fwd_probe.py
TAG="default" python scripts/gemm_fwd_probe.py
TAG="STCS" QIDDM_GEMM_L2_HINTS=8 python scripts/gemm_fwd_probe.py
} 2>&1 | tee gpurun_out/g3_probe.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/g3_bench_n2.json 2> gpurun_out/g3_bench_n2.err; echo "n2 rc=$?"; tail -c 600 gpurun_out/g3_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/g3_bench_ref_n2.json 2> gpurun_out/g3_bench_ref_n2.err; echo "ref n2 rc=$?"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/bench_dp.py --model qiddm_ll --images 4096 --graph > gpurun_out/h_dp_ll.json 2> gpurun_out/h_dp_ll.err; echo "dp ll rc=$?"; cat gpurun_out/h_dp_ll.json
QIDDM_GRAPH_ALLREDUCE=1 timeout 300 $TR --master-port 29512 scripts/bench_dp.py --model qiddm_ll --images 4096 --graph > gpurun_out/h_dp_ll_cap.json 2> gpurun_out/h_dp_ll_cap.err; echo "dp ll captured rc=$?"; cat gpurun_out/h_dp_ll_cap.json; tail -c 400 gpurun_out/h_dp_ll_cap.err
timeout 300 $TR --master-port 29513 scripts/bench_dp.py --model qiddm_ll --images 1 --graph --steps 200 > gpurun_out/h_dp_ll_b1.json 2>/dev/null; cat gpurun_out/h_dp_ll_b1.json
QIDDM_GRAPH_ALLREDUCE=1 timeout 300 $TR --master-port 29514 scripts/bench_dp.py --model qiddm_ll --images 1 --graph --steps 200 > gpurun_out/h_dp_ll_b1_cap.json 2>/dev/null; cat gpurun_out/h_dp_ll_b1_cap.json
timeout 300 $TR --master-port 29515 scripts/bench_dp.py --model unet --images 64 --graph > gpurun_out/h_dp_unet.json 2>/dev/null; cat gpurun_out/h_dp_unet.json
QIDDM_GRAPH_ALLREDUCE=1 timeout 300 $TR --master-port 29516 scripts/bench_dp.py --model unet --images 64 --graph > gpurun_out/h_dp_unet_cap.json 2>/dev/null; cat gpurun_out/h_dp_unet_cap.json

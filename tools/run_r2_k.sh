#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_comm_gpu.py -m gpu -x -q 2>&1 | grep -E "passed|failed|Error|assert|E  " | head -20 | tee gpurun_out/r2k_pytest.log
timeout 300 $TR --nproc-per-node 2 --master-port 29561 tools/comm_bench.py 2>&1 | grep -E "peer|nccl" | tee gpurun_out/r2k_comm2.log
timeout 600 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2k_n2_peer.json 2> gpurun_out/r2k_n2_peer.err; tail -c 300 gpurun_out/r2k_n2_peer.json

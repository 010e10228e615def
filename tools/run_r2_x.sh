#!/bin/bash
# two-GPU check: peer all-reduce tests, bench lines of config 2 (weak) and config 4 (strong) at N = 2
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_comm_gpu.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2x_pytest.log; cat gpurun_out/r2x_pytest.log
timeout 600 $TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2x_n2.json 2> gpurun_out/r2x_n2.err; tail -c 400 gpurun_out/r2x_n2.json; tail -2 gpurun_out/r2x_n2.err
timeout 600 $TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 --cpu-batch 1 > gpurun_out/r2x_n2_ref.json 2> gpurun_out/r2x_n2_ref.err; tail -c 300 gpurun_out/r2x_n2_ref.json

#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2l_n1.json 2> gpurun_out/r2l_n1.err; tail -c 400 gpurun_out/r2l_n1.json; tail -2 gpurun_out/r2l_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2l_n2_peer.json 2> gpurun_out/r2l_n2_peer.err; tail -c 300 gpurun_out/r2l_n2_peer.json; tail -3 gpurun_out/r2l_n2_peer.err
timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --allreduce nccl > gpurun_out/r2l_n2_nccl.json 2> gpurun_out/r2l_n2_nccl.err; tail -c 300 gpurun_out/r2l_n2_nccl.json; tail -3 gpurun_out/r2l_n2_nccl.err
timeout 600 python bench.py --config 4 --no-cpu-baseline --no-hotpath > gpurun_out/r2l_c4_n1.json 2> gpurun_out/r2l_c4_n1.err; tail -c 300 gpurun_out/r2l_c4_n1.json; tail -3 gpurun_out/r2l_c4_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --config 4 --no-cpu-baseline --no-hotpath > gpurun_out/r2l_c4_n2.json 2> gpurun_out/r2l_c4_n2.err; tail -c 300 gpurun_out/r2l_c4_n2.json; tail -3 gpurun_out/r2l_c4_n2.err

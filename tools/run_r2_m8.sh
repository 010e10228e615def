#!/bin/bash
# 8-GPU runs: config 2 weak scaling (peer all-reduce with 16 / 8 / 32 CTAs per overlapped bucket) and config 4 strong scaling
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 400 $TR --master-port 29541 bench.py --gpus 8 --steps 30 > gpurun_out/r2m8_n8.json 2> gpurun_out/r2m8_n8.err; tail -c 150 gpurun_out/r2m8_n8.json; tail -1 gpurun_out/r2m8_n8.err
timeout 400 $TR --master-port 29542 bench.py --gpus 8 --steps 30 --comm-ctas 8 > gpurun_out/r2m8_n8_c8.json 2> gpurun_out/r2m8_n8_c8.err; tail -c 150 gpurun_out/r2m8_n8_c8.json
timeout 400 $TR --master-port 29543 bench.py --gpus 8 --steps 30 --comm-ctas 32 > gpurun_out/r2m8_n8_c32.json 2> gpurun_out/r2m8_n8_c32.err; tail -c 150 gpurun_out/r2m8_n8_c32.json
timeout 400 $TR --master-port 29544 bench.py --gpus 8 --steps 30 --config 4 > gpurun_out/r2m8_c4_n8.json 2> gpurun_out/r2m8_c4_n8.err; tail -c 150 gpurun_out/r2m8_c4_n8.json

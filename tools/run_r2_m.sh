#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_comm_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2m_pytest.log
for n in 4 8; do
timeout 600 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n > gpurun_out/r2m_n$n.json 2> gpurun_out/r2m_n$n.err; tail -c 200 gpurun_out/r2m_n$n.json; tail -2 gpurun_out/r2m_n$n.err
timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --config 4 --no-cpu-baseline --no-hotpath > gpurun_out/r2m_c4_n$n.json 2> gpurun_out/r2m_c4_n$n.err; tail -c 200 gpurun_out/r2m_c4_n$n.json; tail -2 gpurun_out/r2m_c4_n$n.err
done
timeout 600 $TR --nproc-per-node 8 --master-port 29548 bench.py --gpus 8 --allreduce nccl > gpurun_out/r2m_n8_nccl.json 2> gpurun_out/r2m_n8_nccl.err; tail -c 200 gpurun_out/r2m_n8_nccl.json

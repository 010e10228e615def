#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_comm_gpu.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2k_pytest.log

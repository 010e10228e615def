#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_guard_gpu.py -m gpu -q -x -k "census or uflow_loss" 2>&1 | tail -3
timeout 200 python tools/microbench.py census 2>&1 | grep census_

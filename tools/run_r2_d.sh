#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2d_pytest.log
for v in 1 16 24 32; do echo "census variant $v"; timeout 300 python tools/microbench.py census --census-variant $v 2>&1 | grep census_; done | tee gpurun_out/r2d_census.log

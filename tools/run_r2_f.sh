#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_corr_gpu.py tests/test_ref_cuda_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2f_pytest.log
for v in 0 10 11 12 13 14 15; do echo "corr variant $v"; timeout 300 python tools/microbench.py corr_bwd --variant $v --shapes 16x32x96x128,64x32x96x128,16x32x48x64,16x32x24x32,16x32x12x16,16x64x48x64,16x96x24x32,16x128x12x16,1x128x12x20,1x32x96x160 2>&1 | grep corr_; done | tee gpurun_out/r2f_corr.log

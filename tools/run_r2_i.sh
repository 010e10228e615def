#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_warp_gpu.py tests/test_loss_gpu.py tests/test_model.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2i_pytest.log
for wv in 0 1; do echo "warp variant $wv"; timeout 300 python tools/microbench.py warp --flow smooth --warp-variant $wv --shapes 8x32x96x128,16x32x96x128,64x32x96x128,16x32x48x64,16x32x24x32,8x3x384x512,1x128x12x20 2>&1 | grep warp_; done | tee gpurun_out/r2i_warp.log
timeout 300 python tools/microbench.py warp --flow iid --shapes 16x32x96x128,8x3x384x512 2>&1 | grep warp_ | tee -a gpurun_out/r2i_warp.log

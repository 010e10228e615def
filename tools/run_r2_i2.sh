#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_warp_gpu.py -m gpu -x -q 2>&1 | tail -3
for wv in 1; do echo "warp variant $wv"; timeout 300 python tools/microbench.py warp --flow smooth --warp-variant $wv --shapes 8x32x96x128,16x32x96x128,64x32x96x128,16x32x48x64,8x3x384x512 2>&1 | grep warp_bwd_; done | tee gpurun_out/r2i2_warp.log

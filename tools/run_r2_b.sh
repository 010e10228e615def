#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/lds_probe > gpurun_out/r2b_lds.json 2>&1; cat gpurun_out/r2b_lds.json
for f in iid smooth wild; do timeout 300 python tools/microbench.py warp --flow $f --shapes 8x3x384x512,16x32x96x128 2>&1 | grep warp_; done | tee gpurun_out/r2b_warp.log

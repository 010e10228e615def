#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 1 2 3 4; do echo "warp variant $v"; timeout 200 python tools/microbench.py warp --flow smooth --warp-variant $v --shapes 16x32x96x128,64x32x96x128 2>&1 | grep "warp_bwd_"; done

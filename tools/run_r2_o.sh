#!/bin/bash
set -u
mkdir -p gpurun_out
S="--shapes 16x32x48x64,16x32x24x32,16x32x12x16,8x32x48x64,1x32x96x160,1x64x48x80,1x96x24x40,1x128x12x20,8x32x96x128,2x32x96x128,16x64x48x64,16x96x24x32,16x128x12x16"
timeout 120 python tools/corr_variant_check.py fwd 32 > gpurun_out/r2o_check.log 2>&1; grep -c bit-equal gpurun_out/r2o_check.log; grep -v bit-equal gpurun_out/r2o_check.log
for v in 30 31 32; do echo "variant $v"; timeout 300 python tools/microbench.py corr_fwd $S --variant $v 2>&1 | grep corr_fwd; done > gpurun_out/r2o_corr_small.log 2>&1

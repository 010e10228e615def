#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/corr_variant_check.py bwd 13 14 > gpurun_out/r2w_check.log 2>&1; grep -c bit-equal gpurun_out/r2w_check.log; grep -v bit-equal gpurun_out/r2w_check.log
S="--shapes 8x32x96x128,16x32x96x128,64x32x96x128,16x64x48x64,16x32x48x64,16x32x112x256"
for v in 0 13 14; do echo "variant $v"; timeout 200 python tools/microbench.py corr_bwd $S --variant $v 2>&1 | grep corr_bwd; done > gpurun_out/r2w_corr.log 2>&1; cat gpurun_out/r2w_corr.log

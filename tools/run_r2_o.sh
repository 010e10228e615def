#!/bin/bash
set -u
mkdir -p gpurun_out
S="--shapes 16x32x96x128,64x32x96x128,64x64x48x64,16x32x112x256"
timeout 120 python tools/corr_variant_check.py 35 37 > gpurun_out/r2o_check.log 2>&1; grep -c bit-equal gpurun_out/r2o_check.log; grep -v bit-equal gpurun_out/r2o_check.log
for v in 31 35 36 37; do echo "variant $v"; timeout 200 python tools/microbench.py corr_fwd $S --variant $v 2>&1 | grep corr_fwd; done > gpurun_out/r2o_corr.log 2>&1
cat gpurun_out/r2o_corr.log

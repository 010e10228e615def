#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/microbench.py warp --flow smooth --shapes 16x32x96x128 > gpurun_out/r2h_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:warp_fwd_kernel -s 3 -c 1 -f -o gpurun_out/warp_fwd python tools/microbench.py warp --flow smooth --shapes 16x32x96x128 > gpurun_out/r2h_ncu.log 2>&1
tail -3 gpurun_out/r2h_plain.log; tail -3 gpurun_out/r2h_ncu.log

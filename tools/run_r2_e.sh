#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/microbench.py census --shapes 8x3x384x512 > gpurun_out/r2e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:census_fwd_sym -s 2 -c 2 -f -o gpurun_out/census_sym python tools/microbench.py census --shapes 8x3x384x512 > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_plain.log; tail -5 gpurun_out/r2e_ncu.log

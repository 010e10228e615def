#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 8 12 16 20; do echo "census variant $v"; timeout 300 python tools/microbench.py census --census-variant $v --shapes 8x3x384x512,16x3x320x1024 2>&1 | grep census_; done | tee gpurun_out/r2d2_census.log

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2g_pytest.log
timeout 300 python tools/microbench.py corr_bwd --shapes 16x32x96x128,16x32x48x64,16x32x24x32,16x32x12x16,16x64x48x64,16x96x24x32,16x128x12x16,1x192x6x10,1x128x12x20,1x96x24x40,1x64x48x80,1x32x96x160 2>&1 | grep corr_ | tee gpurun_out/r2g_corr.log
timeout 300 python tools/microbench.py census --shapes 8x3x384x512,16x3x320x1024,32x3x448x1024 2>&1 | grep census_ | tee gpurun_out/r2g_census.log

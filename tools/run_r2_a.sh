#!/bin/bash
# round-2 call A: peaks, GPU tests after the hygiene changes, baseline microbench at the judged shapes
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
./tools/peaks > gpurun_out/r2a_peaks.json 2>&1; cat gpurun_out/r2a_peaks.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
timeout 600 python tools/microbench.py corr --shapes 8x32x96x128,16x32x96x128,64x32x96x128,16x32x48x64,16x32x24x32,16x32x12x16,16x64x48x64,16x96x24x32,16x128x12x16,1x192x6x10,1x128x12x20,1x96x24x40,1x64x48x80,1x32x96x160 --csv gpurun_out/r2a_corr.csv > gpurun_out/r2a_corr.log 2>&1; tail -30 gpurun_out/r2a_corr.log
timeout 600 python tools/microbench.py warp --shapes 8x32x96x128,16x32x96x128,64x32x96x128 --csv gpurun_out/r2a_warp.csv > gpurun_out/r2a_warp.log 2>&1; tail -20 gpurun_out/r2a_warp.log
timeout 300 python tools/microbench.py census --csv gpurun_out/r2a_census.csv > gpurun_out/r2a_census.log 2>&1; tail -5 gpurun_out/r2a_census.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_guard_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2u_guard.log; cat gpurun_out/r2u_guard.log
ncu --set full --clock-control none --import-source on -k regex:probe -s 6 -c 1 -o gpurun_out/corr_mma_probe_r2 -f ./tools/corr_mma_probe > gpurun_out/ncu_mma_probe.log 2>&1; tail -2 gpurun_out/ncu_mma_probe.log

#!/bin/bash
# ncu launch list of ONE eager chairs_uflow train step (per-launch times are cold-cache and serialised: compare shares)
set -u
mkdir -p gpurun_out
python bench.py --profile-step > gpurun_out/plain_step_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step_r2.csv python bench.py --profile-step > gpurun_out/ncu_step_r2.log 2>&1
tail -2 gpurun_out/ncu_step_r2.log; wc -l gpurun_out/launches_step_r2.csv

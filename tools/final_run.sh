#!/bin/bash
# End-of-round evidence run (under gpurun): GPU tests, the bench line, the step launch list, one ncu capture.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_pytest.log
cat gpurun_out/final_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err
tail -c 600 gpurun_out/bench_r1_final2.json
python bench.py --profile-step > gpurun_out/plain_step_final2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step_final2.csv python bench.py --profile-step > gpurun_out/ncu_step_final2.log 2>&1

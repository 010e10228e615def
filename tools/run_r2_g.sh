#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_model.py tests/test_prob_model.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline --no-hotpath > gpurun_out/r2g_config2.json 2> gpurun_out/r2g_config2.err; tail -2 gpurun_out/r2g_config2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_config2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['kernels']['_hot_path_us_per_step'])
for h in d['kernels']['_hot_by_shape'][:14]: print('   ', h['call'], round(h['us_per_launch'],1), h['launches_per_step'], h['bound'], round(h['frac'],3))
P

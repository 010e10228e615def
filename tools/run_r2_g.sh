#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_guard_gpu.py tests/test_model.py tests/test_prob_model.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python tools/microbench.py census 2>&1 | grep census_
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2g_config2.json 2> gpurun_out/r2g_config2.err; tail -2 gpurun_out/r2g_config2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_config2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['kernels']['_hot_path_us_per_step'])
r=d['roofline']; print({k:r[k] for k in ('kernel','bound','frac','avg_launch_us','step_share')})
for h in d['roofline_hotpath']: print('  hot', h['kernel'], h['bound'], round(h['frac'],3), round(h['us'],1))
for h in d['kernels']['_hot_by_shape'][:8]: print('   ', h['call'], round(h['us_per_launch'],1), h['launches_per_step'], h['bound'], round(h['frac'],3))
P

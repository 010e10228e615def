#!/bin/bash
# Round-2 evidence run on one GPU: GPU tests, bench lines of configs 2 (default), 1, 3, 4, 5, kernel sweep, comparators.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2n_pytest.log; cat gpurun_out/r2n_pytest.log
timeout 600 python bench.py > gpurun_out/r2n_config2.json 2> gpurun_out/r2n_config2.err; tail -c 200 gpurun_out/r2n_config2.json; tail -2 gpurun_out/r2n_config2.err
timeout 400 python bench.py --config 1 > gpurun_out/r2n_config1.json 2> gpurun_out/r2n_config1.err; tail -c 200 gpurun_out/r2n_config1.json; tail -2 gpurun_out/r2n_config1.err
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2n_config3.json 2> gpurun_out/r2n_config3.err; tail -c 200 gpurun_out/r2n_config3.json; tail -2 gpurun_out/r2n_config3.err
timeout 600 python bench.py --config 4 --steps 20 > gpurun_out/r2n_config4.json 2> gpurun_out/r2n_config4.err; tail -c 200 gpurun_out/r2n_config4.json; tail -2 gpurun_out/r2n_config4.err
timeout 400 python bench.py --config 5 > gpurun_out/r2n_config5.json 2> gpurun_out/r2n_config5.err; tail -c 200 gpurun_out/r2n_config5.json; tail -2 gpurun_out/r2n_config5.err
timeout 300 python tools/microbench.py compare --csv gpurun_out/r2n_compare.csv > gpurun_out/r2n_compare.log 2>&1; tail -3 gpurun_out/r2n_compare.log | cut -c1-200
timeout 400 python tools/microbench.py all --csv gpurun_out/r2n_microbench.csv > gpurun_out/r2n_microbench.log 2>&1; tail -3 gpurun_out/r2n_microbench.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2n_config2.json 2> gpurun_out/r2n_config2.err; tail -c 200 gpurun_out/r2n_config2.json; tail -2 gpurun_out/r2n_config2.err
timeout 400 python bench.py --config 1 > gpurun_out/r2n_config1.json 2> gpurun_out/r2n_config1.err; tail -c 200 gpurun_out/r2n_config1.json; tail -2 gpurun_out/r2n_config1.err
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2n_config3.json 2> gpurun_out/r2n_config3.err; tail -c 200 gpurun_out/r2n_config3.json; tail -2 gpurun_out/r2n_config3.err
timeout 600 python bench.py --config 4 --steps 20 > gpurun_out/r2n_config4.json 2> gpurun_out/r2n_config4.err; tail -c 200 gpurun_out/r2n_config4.json; tail -2 gpurun_out/r2n_config4.err

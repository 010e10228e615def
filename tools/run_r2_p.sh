#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_corr_gpu.py tests/test_ref_cuda_gpu.py tests/test_pwclite.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2p_pytest.log; cat gpurun_out/r2p_pytest.log
timeout 300 python tools/microbench.py corr_fwd 2>&1 | grep corr_fwd > gpurun_out/r2p_corr.log; cat gpurun_out/r2p_corr.log

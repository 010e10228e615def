#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_corr_gpu.py tests/test_guard_gpu.py tests/test_ref_cuda_gpu.py tests/test_pwclite.py tests/test_model.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2p_pytest.log; cat gpurun_out/r2p_pytest.log
timeout 300 python tools/microbench.py corr_fwd --shapes 16x32x48x64,16x32x24x32,16x32x12x16,8x32x48x64,1x32x96x160,1x64x48x80,8x32x96x128,2x32x96x128,16x64x48x64,16x96x24x32,16x128x12x16,16x32x96x128,64x32x96x128 2>&1 | grep corr_fwd > gpurun_out/r2p_corr.log; cat gpurun_out/r2p_corr.log

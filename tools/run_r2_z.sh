#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_warp_gpu.py tests/test_guard_gpu.py tests/test_model.py tests/test_pwclite.py -m gpu -q -x 2>&1 | tail -5
timeout 200 python tools/microbench.py warp --flow smooth --shapes 8x32x96x128,16x32x96x128,64x32x96x128,16x32x48x64,16x32x24x32,32x32x112x256 2>&1 | grep "warp_bwd_"
timeout 100 python tools/microbench.py warp --flow iid --shapes 16x32x96x128,64x32x96x128 2>&1 | grep "warp_bwd "
timeout 100 python tools/microbench.py warp --flow wild --shapes 16x32x96x128 2>&1 | grep "warp_bwd"
echo "old kernel (variant 9)"
timeout 100 python tools/microbench.py warp --flow smooth --warp-variant 9 --shapes 16x32x96x128,64x32x96x128 2>&1 | grep "warp_bwd_"

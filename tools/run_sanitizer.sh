#!/bin/bash
# compute-sanitizer over the kernel parity tests: memcheck on all of them, racecheck on the kernels that
# synchronise through shared memory (correlation TMA/mbarrier rings, row-scan carries, census strips, stencil).
set -u
mkdir -p gpurun_out
T="tests/test_corr_gpu.py tests/test_warp_gpu.py tests/test_loss_gpu.py tests/test_triag_gpu.py"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_memcheck.log \
    python -m pytest $T -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2_memcheck_pytest.log
tail -3 gpurun_out/r2_memcheck_pytest.log; tail -4 gpurun_out/r2_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r2_racecheck.log \
    python -m pytest tests/test_corr_gpu.py tests/test_triag_gpu.py tests/test_loss_gpu.py -m gpu -q -x -p no:cacheprovider \
    -k "golden or fast_path or both_tiled or cp_async or census or substitution" > gpurun_out/r2_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2_racecheck_pytest.log
tail -3 gpurun_out/r2_racecheck_pytest.log; tail -4 gpurun_out/r2_racecheck.log

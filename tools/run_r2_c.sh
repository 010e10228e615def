#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/insitu_fields.py 2>&1 | grep -v "^   \|^warp call" | tee gpurun_out/r2c_insitu2.log | grep -E "warp|census|range|inside|resize" | head -60

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_round2_golden.py tests/test_model.py tests/test_pwclite.py -m gpu -q -x 2>&1 | tail -3
python - <<'P'
import torch, sys
sys.path.insert(0, '.')
from arflow_b200 import _lib
from tools.microbench import time_graph
lib = _lib.load()
cs = lambda: torch.cuda.current_stream().cuda_stream
for (N, Hi, Wi) in [(32, 192, 256), (32, 96, 128), (32, 48, 64), (64, 160, 512)]:
    def mk(kind):
        def make():
            a = torch.randn(N, Hi, Wi, device="cuda"); b = torch.randn(N, 2 * Hi, 2 * Wi, device="cuda")
            if kind == "fwd":
                return lambda: lib.arf_resize_bilinear_fwd(a.data_ptr(), b.data_ptr(), N, Hi, Wi, 2 * Hi, 2 * Wi, 0.5, 0.5, 2.0, 0, cs())
            return lambda: lib.arf_resize_bilinear_bwd(b.data_ptr(), a.data_ptr(), N, Hi, Wi, 2 * Hi, 2 * Wi, 0.5, 0.5, 2.0, 0, cs())
        return make
    nb = N * Hi * Wi * 5 * 4
    for kind in ("fwd", "bwd"):
        med, _ = time_graph(mk(kind), nb)
        print("resize_x2_%s N%d %dx%d  %.1f us  %.0f GB/s" % (kind, N, Hi, Wi, med * 1e6, nb / med / 1e9))
P

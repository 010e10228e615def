#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2d_pytest.log
for v in 12 16 24; do echo "census variant $v"; timeout 300 python tools/microbench.py census --census-variant $v --shapes 8x3x384x512,16x3x320x1024 2>&1 | grep census_; done | tee gpurun_out/r2d3_census.log
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from arflow_b200 import _lib, uflow_utils as uu
lib = _lib.load()
g = torch.Generator().manual_seed(3)
for kind in ("noise", "similar", "identical"):
    a = torch.rand(4, 3, 384, 512, generator=g)
    b = torch.rand(4, 3, 384, 512, generator=g) if kind == "noise" else (a + 0.01 * torch.randn(a.shape, generator=g)).clamp(0, 1) if kind == "similar" else a.clone()
    m = torch.ones(4, 1, 384, 512)
    out = {}
    for v in (1, 16):
        lib.arf_debug_set(5, v)
        out[v] = (uu.census_loss(a.cuda(), b.cuda(), m.cuda()).item(), uu.census_loss_no_penalty(a.cuda(), b.cuda(), m.cuda())[0].double().cpu())
    # float64 truth from the explicit torch chain
    import oracle.arflow_oracle as orc
    hd = orc.census_loss_no_penalty(a[:1].double(), b[:1].double(), m[:1].double())[0]
    for v in (1, 16):
        e = (out[v][1][:1] - hd).abs()
        print(kind, "variant", v, "loss", out[v][0], "max|h-h64|", e.max().item(), "mean", e.mean().item(), "max|h64|", hd.abs().max().item())
    lo = orc.census_loss(a.double(), b.double(), m.double()).item()
    print(kind, "loss64", lo, "rel err v1 %.2e  v16 %.2e" % (abs(out[1][0] - lo) / abs(lo), abs(out[16][0] - lo) / abs(lo)))
PY

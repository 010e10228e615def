"""Bit-compare corr kernel variants (debug hook 1) with the default kernels on a few shapes.
    python tools/corr_variant_check.py fwd 30 31      python tools/corr_variant_check.py bwd 13 14"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arflow_b200 import _lib
lib = _lib.load()
cs = lambda: torch.cuda.current_stream().cuda_stream
kind = sys.argv[1]
for v in [int(a) for a in sys.argv[2:]]:
    for (B, C, h, w) in [(2, 32, 96, 128), (3, 20, 50, 68), (1, 64, 48, 64), (2, 196, 12, 16), (1, 32, 37, 100)]:
        torch.manual_seed(0)
        f1, f2 = torch.randn(B, C, h, w, device="cuda"), torch.randn(B, C, h, w, device="cuda")
        go = torch.randn(B, 81, h, w, device="cuda")
        res = []
        for var in (0, v):
            lib.arf_debug_set(1, var)
            if kind == "fwd":
                o = torch.full((B, 81, h, w), float("nan"), device="cuda")
                lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), o.data_ptr(), B, C, h, w, 4, 1, 4, 1, 1, cs())
                res.append([o])
            else:
                g1, g2 = torch.full_like(f1, float("nan")), torch.full_like(f2, float("nan"))
                lib.arf_corr_bwd(f1.data_ptr(), f2.data_ptr(), go.data_ptr(), g1.data_ptr(), g2.data_ptr(), B, C, h, w, 4, 1, 4, 1, 1, cs())
                res.append([g1, g2])
            lib.arf_debug_set(1, 0)
        torch.cuda.synchronize()
        ok = all(torch.equal(a, b) for a, b in zip(*res))
        print(kind, "variant", v, (B, C, h, w), "bit-equal" if ok else "DIFF max %.3e" % max(float((a - b).abs().max()) for a, b in zip(*res)))

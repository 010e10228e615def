"""Bit-compare a corr_fwd kernel variant (debug hook 1) with the default kernel on a few shapes."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arflow_b200 import _lib
lib = _lib.load()
cs = lambda: torch.cuda.current_stream().cuda_stream
for v in [int(a) for a in sys.argv[1:]]:
    for (B, C, h, w) in [(2, 32, 96, 128), (3, 20, 50, 68), (1, 64, 48, 64), (2, 196, 12, 16), (1, 32, 37, 100)]:
        torch.manual_seed(0)
        f1, f2 = torch.randn(B, C, h, w, device="cuda"), torch.randn(B, C, h, w, device="cuda")
        o0 = torch.full((B, 81, h, w), float("nan"), device="cuda")
        o1 = torch.full((B, 81, h, w), float("nan"), device="cuda")
        lib.arf_debug_set(1, 0)
        lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), o0.data_ptr(), B, C, h, w, 4, 1, 4, 1, 1, cs())
        lib.arf_debug_set(1, v)
        lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), o1.data_ptr(), B, C, h, w, 4, 1, 4, 1, 1, cs())
        lib.arf_debug_set(1, 0)
        torch.cuda.synchronize()
        print("variant", v, (B, C, h, w), "bit-equal" if torch.equal(o0, o1) else "DIFF max %.3e" % float((o0 - o1).abs().max()))

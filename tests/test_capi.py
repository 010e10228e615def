"""The C-ABI library loads and exports every symbol include/arflow_b200.h declares (no GPU needed)."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "arflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from arflow_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 6
    for n in names:
        assert hasattr(lib, n), "libarflow_b200.so does not export %s" % n
        assert n in _lib.PROTOTYPES, "no ctypes prototype for %s" % n
    assert sorted(_lib.PROTOTYPES) == names, "prototype table and header disagree"
    assert lib.arf_version() >= 100
    assert _lib.error_string(0) == "ok"
    assert "invalid" in _lib.error_string(-1)


def test_out_dims_match_reference_arithmetic(oracle):
    from arflow_b200.correlation import corr_out_dims
    for (H, W, pad, ks, md, s1, s2) in [(9, 11, 4, 1, 4, 1, 1), (24, 32, 20, 3, 20, 1, 2), (16, 20, 3, 1, 3, 2, 1),
                                        (13, 17, 4, 3, 4, 2, 2), (6, 10, 4, 1, 4, 1, 1)]:
        assert corr_out_dims(H, W, pad, ks, md, s1, s2) == oracle.corr_dims(H, W, pad, ks, md, s1, s2)


def test_no_cpu_fallback():
    import torch
    from arflow_b200.correlation import Correlation
    from arflow_b200.warp_utils import flow_warp
    x = torch.randn(1, 4, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)(x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        flow_warp(x, torch.zeros(1, 2, 8, 8))

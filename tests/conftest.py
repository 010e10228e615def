import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity bars of BASELINE.json's north_star (fp32): values 1e-5, gradients 1e-4, relative to the
# largest magnitude of the reference tensor (SURVEY §7 "hard parts": relative error is ill-defined
# where |value| ~ 0).
RTOL_VALUE = 1e-5
RTOL_GRAD = 1e-4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_err(a, ref):
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    scale = ref.abs().max().item()
    if scale == 0.0:
        scale = 1.0
    return (a - ref).abs().max().item() / scale


def elementwise_err(a, ref):
    """max over elements of |a - ref| / (|ref| + rms(ref)): a relative bound with an absolute floor, so that an error in a
    small-magnitude region is not hidden behind the tensor's largest value."""
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    rms = ref.pow(2).mean().sqrt().item() if ref.numel() else 0.0
    if rms == 0.0:
        rms = 1.0
    return ((a - ref).abs() / (ref.abs() + rms)).max().item() if ref.numel() else 0.0


def assert_close(a, ref, rtol, what="", elementwise=True):
    """Two bars: max|a-ref| / max|ref| <= rtol, and, element by element, |a-ref| <= 4 rtol (|ref| + rms(ref))."""
    e = rel_err(a, ref)
    assert e <= rtol, "%s: max|a-ref|/max|ref| = %.3e > %.1e" % (what, e, rtol)
    if elementwise:
        w = elementwise_err(a, ref)
        assert w <= 4 * rtol, "%s: element-wise |a-ref|/(|ref|+rms) = %.3e > %.1e" % (what, w, 4 * rtol)


@pytest.fixture(scope="session")
def oracle():
    import oracle.arflow_oracle as o
    return o

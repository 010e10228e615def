"""Round-2 fixtures of the unmodified reference (tests/golden/make_golden.py round2): normalize_features,
align_corners=True flow up-sampling, a larger inverse-diagonal case.  CPU tests pin the oracle, GPU tests the kernels."""
import pytest
import torch

from conftest import RTOL_GRAD, RTOL_VALUE, assert_close, load_golden

FLAGS = {"all": (True, True, True, True), "perimg": (True, True, True, False), "perch": (True, False, False, True)}
_W = lambda shape: torch.randn(shape, generator=torch.Generator().manual_seed(1234))


def _norm_check(fn, g, dev, tight):
    for tag, flags in FLAGS.items():
        a = g["in0"].to(dev).double().requires_grad_(True) if tight else g["in0"].to(dev).requires_grad_(True)
        b = g["in1"].to(dev).double().requires_grad_(True) if tight else g["in1"].to(dev).requires_grad_(True)
        o1, o2 = fn([a, b], *flags)
        gen = torch.Generator().manual_seed(1234)        # make_golden.both: one generator across the outputs
        loss = 0
        for o in (o1, o2):
            loss = loss + (o * torch.randn(o.shape, generator=gen).to(o)).sum()
        ga, gb = torch.autograd.grad(loss, [a, b])
        assert_close(o1, g[tag + "_out0_f64"], 1e-12 if tight else RTOL_VALUE, tag + " out0")
        assert_close(o2, g[tag + "_out1_f64"], 1e-12 if tight else RTOL_VALUE, tag + " out1")
        assert_close(ga, g[tag + "_grad0_f64"], 1e-10 if tight else RTOL_GRAD, tag + " grad0")
        assert_close(gb, g[tag + "_grad1_f64"], 1e-10 if tight else RTOL_GRAD, tag + " grad1")


def test_oracle_normalize_features():
    from oracle.cpu_nets import normalize_features
    _norm_check(normalize_features, load_golden("normalize_features"), "cpu", True)


def test_oracle_inverse_diagonal_12x17(oracle):
    g = load_golden("invdiag_12x17")
    assert_close(oracle.inverse_diagonal(g["in0"].double(), g["in1"].double(), g["in2"].double()), g["out0_f64"], 1e-12)


@pytest.mark.gpu
def test_normalize_features_golden():
    """N1 (uflow_model.py:8-50) against the reference's own outputs; `all` runs on arf_featnorm_fwd/bwd."""
    from arflow_b200.uflow_model import normalize_features
    _norm_check(normalize_features, load_golden("normalize_features"), "cuda", False)


@pytest.mark.gpu
def test_inverse_diagonal_12x17_golden():
    from arflow_b200 import triag_solve as ts
    g = load_golden("invdiag_12x17")
    a, b, c = (g["in%d" % i].cuda().contiguous() for i in range(3))
    assert_close(ts.inverse_diagonal(a, b, c), g["out0_f64"], RTOL_VALUE)


@pytest.mark.gpu
@pytest.mark.parametrize("s", [2, 4])
def test_resize_align_corners_golden(s):
    """F.interpolate(flow * s, scale_factor=s, bilinear, align_corners=True) (pwclite.py:178-179, 203)."""
    from arflow_b200.uflow_utils import interpolate_align_corners
    g = load_golden("resize_align_corners")
    x = g["in0"].cuda().requires_grad_(True)
    out = interpolate_align_corners(x, s, mul=float(s))
    (gx,) = torch.autograd.grad((out * _W(out.shape).cuda()).sum(), [x])
    assert_close(out, g["up%d_out0_f64" % s], RTOL_VALUE)
    assert_close(gx, g["up%d_grad0_f64" % s], RTOL_GRAD)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 2, 7, 5), (2, 1, 2, 2), (1, 2, 24, 40)])
def test_resize_align_corners_vs_aten(shape):
    from arflow_b200.uflow_utils import interpolate_align_corners
    x = torch.randn(shape, generator=torch.Generator().manual_seed(sum(shape)))
    ref_in = x.double().requires_grad_(True)
    ref = torch.nn.functional.interpolate(ref_in * 2, scale_factor=2, mode="bilinear", align_corners=True)
    w = _W(ref.shape)
    (gref,) = torch.autograd.grad((ref * w.double()).sum(), [ref_in])
    xc = x.cuda().requires_grad_(True)
    out = interpolate_align_corners(xc, 2, mul=2.0)
    (gx,) = torch.autograd.grad((out * w.cuda()).sum(), [xc])
    assert_close(out, ref, RTOL_VALUE)
    assert_close(gx, gref, RTOL_GRAD)

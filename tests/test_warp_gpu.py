"""Parity of the warp kernels (flow_warp / resample) with the golden fixtures and the oracle."""
import pytest
import torch

from conftest import RTOL_GRAD, RTOL_VALUE, assert_close, load_golden

pytestmark = pytest.mark.gpu


def _check(g, out, gx, gf):
    """The parity target is the reference's fp32 grid_sample path (north_star): its normalise ->
    un-normalise round trip perturbs coordinates by ~1e-5 px in fp32, so fp32 results sit ~1e-5*|dI/dx|
    away from the float64 truth; the kernels reproduce that arithmetic and are held to 1e-5 against
    the fp32 reference, and to a looser sanity bound against float64."""
    assert_close(out, g["out0_f32"], RTOL_VALUE, "warp vs fp32 reference")
    assert_close(gx, g["grad0_f32"], RTOL_GRAD, "grad x vs fp32 reference")
    assert_close(gf, g["grad1_f32"], RTOL_GRAD, "grad flow vs fp32 reference")
    assert_close(out, g["out0_f64"], 2e-4, "warp vs float64")
    assert_close(gf, g["grad1_f64"], 1e-3, "grad flow vs float64")


def _run_flow_warp(x, f, **kw):
    from arflow_b200.warp_utils import flow_warp
    x = x.cuda().requires_grad_(True)
    f = f.cuda().requires_grad_(True)
    out = flow_warp(x, f, **kw)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1234)).cuda()
    gx, gf = torch.autograd.grad((out * w).sum(), [x, f])
    return out, gx, gf


@pytest.mark.parametrize("pad", ["zeros", "border", "reflection"])
@pytest.mark.parametrize("align", [True, False])
def test_flow_warp_golden(pad, align):
    g = load_golden("flow_warp_%s_%d" % (pad, align))
    out, gx, gf = _run_flow_warp(g["in0"], g["in1"], pad=pad, align_corners=align)
    _check(g, out, gx, gf)


def test_flow_warp_nearest_othersize_golden():
    from arflow_b200.warp_utils import flow_warp
    g = load_golden("flow_warp_nearest")
    assert_close(flow_warp(g["in0"].cuda(), g["in1"].cuda(), mode="nearest"), g["out0_f32"], RTOL_VALUE)
    g = load_golden("flow_warp_othersize")
    out, gx, gf = _run_flow_warp(g["in0"], g["in1"])
    _check(g, out, gx, gf)


def test_resample_golden():
    from arflow_b200.uflow_utils import flow_to_warp, resample
    g = load_golden("resample")
    x = g["in0"].cuda().requires_grad_(True)
    f = g["in1"].cuda().requires_grad_(True)
    out = resample(x, flow_to_warp(f))
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1234)).cuda()
    gx, gf = torch.autograd.grad((out * w).sum(), [x, f])
    _check(g, out, gx, gf)


@pytest.mark.parametrize("shape", [(2, 32, 48, 64), (1, 3, 96, 128), (2, 196, 6, 8), (1, 5, 31, 45)])
def test_flow_warp_vs_oracle(oracle, shape):
    B, C, H, W = shape
    gen = torch.Generator().manual_seed(H * W + C)
    x = torch.randn(shape, generator=gen)
    f = torch.randn(B, 2, H, W, generator=gen) * 2.0
    xd, fd = x.clone().requires_grad_(True), f.clone().requires_grad_(True)   # fp32 oracle, see _check
    ref = oracle.warp(xd, fd, kind="flow")
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1234))
    rx, rf = torch.autograd.grad((ref * w).sum(), [xd, fd])
    out, gx, gf = _run_flow_warp(x, f)
    assert_close(out, ref, RTOL_VALUE)
    assert_close(gx, rx, RTOL_GRAD)
    assert_close(gf, rf, RTOL_GRAD)


def test_identity_and_integer_shift_full_size():
    """Properties at the config-2 image shape: zero flow is the identity; an integer flow is a shift."""
    from arflow_b200.warp_utils import flow_warp
    x = torch.rand(8, 3, 384, 512, generator=torch.Generator().manual_seed(3)).cuda()
    z = torch.zeros(8, 2, 384, 512, device="cuda")
    assert_close(flow_warp(x, z), x, 2e-4)
    f = z.clone()
    f[:, 0] = 3.0
    f[:, 1] = -2.0
    y = flow_warp(x, f)
    assert_close(y[:, :, 2:, :-3], x[:, :, :-2, 3:], 2e-4)
    assert float(y[:, :, :1].abs().max()) < 1e-3

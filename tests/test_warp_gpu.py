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


# --------------------------------------------------------------------------- hard cases
@pytest.mark.parametrize("shape,sigma,shift,pad,align", [
    ((2, 32, 48, 64), 2.0, 0.0, "zeros", True),
    ((1, 3, 96, 128), 2.0, 0.0, "zeros", True),
    ((1, 7, 33, 50), 2.0, 0.0, "zeros", True),         # ragged: W % 4 != 0, partial last warp
    ((1, 20, 40, 72), 30.0, 0.0, "zeros", True),       # wild flow: neighbouring lanes sample far-apart places
    ((2, 9, 64, 96), 0.7, 37.5, "zeros", True),        # large uniform shift, partly off-image
    ((1, 9, 64, 96), 0.7, -200.0, "zeros", True),      # every tap off-image
    ((1, 12, 48, 64), 6.0, 0.0, "border", True),
    ((1, 12, 48, 64), 6.0, 0.0, "reflection", False),
    ((1, 40, 100, 200), 1.0, 3.0, "zeros", False),
    ((3, 5, 9, 20), 1.0, 0.0, "zeros", True),          # tiny image: several channel groups per block
    ((1, 64, 12, 20), 0.3, 0.0, "zeros", True),        # smooth flow: east / west taps of neighbouring lanes merge
])
def test_warp_gradients_vs_oracle_hard_cases(oracle, shape, sigma, shift, pad, align):
    """Forward, source gradient (merged reds) and flow gradient (in-block channel-group sum) on ragged, wild, shifted and
    off-image fields, every padding mode."""
    B, C, H, W = shape
    gen = torch.Generator().manual_seed(H * W + C)
    x = torch.randn(shape, generator=gen)
    f = torch.randn(B, 2, H, W, generator=gen) * sigma + shift
    xd, fd = x.clone().requires_grad_(True), f.clone().requires_grad_(True)
    ref = oracle.warp(xd, fd, kind="flow", pad=pad, align_corners=align)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1234))
    rx, rf = torch.autograd.grad((ref * w).sum(), [xd, fd])
    out, gx, gf = _run_flow_warp(x, f, pad=pad, align_corners=align)
    assert_close(out, ref, RTOL_VALUE)
    assert_close(gx, rx, RTOL_GRAD)
    assert_close(gf, rf, RTOL_GRAD)
    # flow gradient only (source detached) takes a different instantiation
    from arflow_b200.warp_utils import flow_warp
    fc = f.cuda().requires_grad_(True)
    (gf2,) = torch.autograd.grad((flow_warp(x.cuda(), fc, pad=pad, align_corners=align) * w.cuda()).sum(), [fc])
    assert_close(gf2, rf, RTOL_GRAD)


def test_warp_benchmark_size_properties():
    """64x32x96x128 (config 5) on a smooth flow: the flow gradient is deterministic run to run, and the source gradient
    conserves mass: sum(gx) == sum over pixels of gy * (sum of in-range tap weights)."""
    from arflow_b200.warp_utils import flow_warp
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(64, 32, 96, 128, generator=gen).cuda().requires_grad_(True)
    f = torch.nn.functional.interpolate(torch.randn(64, 2, 12, 16, generator=gen) * 3, scale_factor=8,
                                        mode="bilinear").cuda().requires_grad_(True)
    w = torch.randn(64, 32, 96, 128, generator=gen).cuda()
    gx1, gf1 = torch.autograd.grad((flow_warp(x, f) * w).sum(), [x, f])
    gx2, gf2 = torch.autograd.grad((flow_warp(x, f) * w).sum(), [x, f])
    assert torch.equal(gf1, gf2)
    assert_close(gx1, gx2, 2e-6)
    ones = torch.ones_like(x)
    wsum = flow_warp(ones, f.detach())                 # per pixel: sum of the in-range tap weights
    assert_close(gx1.double().sum(), (w.double() * wsum.double()).sum(), 1e-6)


def test_flow_grad_only_full_size(oracle):
    """Image warp of the loss (source detached): flow gradient only at 384x512, against the oracle (B = 2)."""
    from arflow_b200.warp_utils import flow_warp
    gen = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 384, 512, generator=gen)
    f = torch.nn.functional.interpolate(torch.randn(2, 2, 24, 32, generator=gen) * 6, scale_factor=16, mode="bilinear")
    w = torch.randn(2, 3, 384, 512, generator=gen)
    fc = f.cuda().requires_grad_(True)
    (g_dev,) = torch.autograd.grad((flow_warp(x.cuda(), fc) * w.cuda()).sum(), [fc])
    fd = f.clone().requires_grad_(True)
    (g_ref,) = torch.autograd.grad((oracle.warp(x, fd, kind="flow") * w).sum(), [fd])
    assert_close(g_dev, g_ref, RTOL_GRAD)


def test_resample_flow_is_resample_of_flow_to_warp():
    """The callers' fused form (grid added in-kernel) is bit-identical to the reference's two-step form."""
    from arflow_b200.uflow_utils import flow_to_warp, mask_invalid, mask_invalid_flow, resample, resample_flow
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(2, 32, 24, 40, generator=gen).cuda().requires_grad_(True)
    f = (torch.randn(2, 2, 24, 40, generator=gen) * 4).cuda().requires_grad_(True)
    w = torch.randn(2, 32, 24, 40, generator=gen).cuda()
    a = resample(x, flow_to_warp(f))
    ga = torch.autograd.grad((a * w).sum(), [x, f])
    b = resample_flow(x, f)
    gb = torch.autograd.grad((b * w).sum(), [x, f])
    assert torch.equal(a, b)
    assert_close(gb[0], ga[0], 2e-6)
    assert torch.equal(gb[1], ga[1])
    assert torch.equal(mask_invalid_flow(f), mask_invalid(flow_to_warp(f)))

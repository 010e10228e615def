"""Pins the oracle (oracle/) to outputs of the unmodified reference (tests/golden/*.npz).  CPU only."""
import pytest
import torch

from conftest import assert_close, load_golden

TIGHT = 1e-12   # float64 oracle vs float64 reference


@pytest.mark.parametrize("name", ["corr_b2c5_9x11", "corr_b1c32_12x16", "corr_b1c3_6x40"])
def test_correlation(oracle, name):
    g = load_golden(name)
    f1, f2 = g["in0"], g["in1"]
    assert_close(oracle.corr_fwd_c(f1, f2), g["out0_f64"], TIGHT, "C oracle fwd")
    assert_close(oracle.cost_volume(f1.double(), f2.double()), g["out0_f64"], TIGHT, "torch oracle fwd")
    assert_close(oracle.cost_volume(f1, f2), g["out0_f32"], 1e-6, "torch oracle fwd f32")
    # gradients of sum(out * w), w drawn as in make_golden.both
    w = torch.randn(g["out0_f64"].shape, generator=torch.Generator().manual_seed(1234)).double()
    g1, g2 = oracle.corr_bwd_c(f1, f2, w)
    assert_close(g1, g["grad0_f64"], TIGHT, "C oracle grad f1")
    assert_close(g2, g["grad1_f64"], TIGHT, "C oracle grad f2")


def _warp_grads(oracle, x, f, **kw):
    x = x.double().requires_grad_(True)
    f = f.double().requires_grad_(True)
    out = oracle.warp(x, f, **kw)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1234)).double()
    gx, gf = torch.autograd.grad((out * w).sum(), [x, f])
    return out, gx, gf


@pytest.mark.parametrize("pad", ["zeros", "border", "reflection"])
@pytest.mark.parametrize("align", [True, False])
def test_flow_warp(oracle, pad, align):
    g = load_golden("flow_warp_%s_%d" % (pad, align))
    out, gx, gf = _warp_grads(oracle, g["in0"], g["in1"], kind="flow", pad=pad, align_corners=align)
    assert_close(out, g["out0_f64"], TIGHT, "warp")
    assert_close(gx, g["grad0_f64"], TIGHT, "grad x")
    assert_close(gf, g["grad1_f64"], 1e-10, "grad flow")


def test_flow_warp_nearest_and_othersize(oracle):
    g = load_golden("flow_warp_nearest")
    assert_close(oracle.warp(g["in0"].double(), g["in1"].double(), mode="nearest"), g["out0_f64"], TIGHT)
    g = load_golden("flow_warp_othersize")
    out, gx, gf = _warp_grads(oracle, g["in0"], g["in1"], kind="flow")
    assert_close(out, g["out0_f64"], TIGHT)
    assert_close(gx, g["grad0_f64"], TIGHT)
    assert_close(gf, g["grad1_f64"], 1e-10)


def test_resample(oracle):
    g = load_golden("resample")
    x, f = g["in0"], g["in1"]
    xd = x.double().requires_grad_(True)
    fd = f.double().requires_grad_(True)
    out = oracle.warp(xd, oracle.flow_to_warp(fd), kind="coords")
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1234)).double()
    gx, gf = torch.autograd.grad((out * w).sum(), [xd, fd])
    assert_close(out, g["out0_f64"], TIGHT)
    assert_close(gx, g["grad0_f64"], TIGHT)
    assert_close(gf, g["grad1_f64"], 1e-10)

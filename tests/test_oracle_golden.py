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


def _w(shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(1234)).double()


def _grads(fn, inputs, wrt):
    """d sum(out*w) / d inputs[wrt], w drawn exactly like tests/golden/make_golden.py::both."""
    ins = [i.double().clone().requires_grad_(k in wrt) for k, i in enumerate(inputs)]
    out = fn(*ins)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    g = torch.Generator().manual_seed(1234)
    loss = 0
    for o in outs:
        if o.requires_grad:
            loss = loss + (o * torch.randn(o.shape, generator=g).double()).sum()
    return outs, torch.autograd.grad(loss, [ins[k] for k in wrt])


def test_masks(oracle):
    g = load_golden("masks")
    flow, flow_b = g["in0"].double(), g["in1"].double()
    assert torch.equal(oracle.mask_invalid(oracle.flow_to_warp(flow)), g["invalid_out0_f64"])
    assert_close(oracle.range_map(flow), g["range_out0_f64"], TIGHT)
    assert_close(oracle.range_map(flow), g["range_wu_out0_f64"], TIGHT)
    assert_close(oracle.splat_count(oracle.flow_to_warp(flow)), g["corrmap_out0_f64"], TIGHT)
    assert torch.equal(oracle.occu_mask_backward(flow, 0.2), g["occ_bw_out0_f64"])
    assert_close(oracle.occu_mask_backward(flow, 0.0), g["occ_bw0_out0_f64"], 1e-7)  # reference returns .float()
    assert torch.equal(oracle.border_mask(flow), g["border_out0_f64"])
    assert torch.equal(oracle.occu_mask_bidirection(flow * 0.3, flow_b * 0.3), g["occ_bi_out0_f64"])


def test_resize(oracle):
    g = load_golden("resize")
    for s in (2, 4):
        for is_flow in (0, 1):
            (out,), (gi,) = _grads(lambda a: oracle.resize_bilinear(a, float(s), bool(is_flow)), [g["in0"]], (0,))
            assert_close(out, g["up%d_%d_out0_f64" % (s, is_flow)], TIGHT)
            assert_close(gi, g["up%d_%d_grad0_f64" % (s, is_flow)], TIGHT)
            (out,), (gi,) = _grads(lambda a: oracle.resize_bilinear(a, 1.0 / s, bool(is_flow)), [g["in1"]], (0,))
            assert_close(out, g["down%d_%d_out0_f64" % (s, is_flow)], TIGHT)
            assert_close(gi, g["down%d_%d_grad0_f64" % (s, is_flow)], TIGHT)


def test_census(oracle):
    g = load_golden("census")
    a, b, m = g["in0"], g["in1"], g["in2"]
    (loss,), (ga, gb) = _grads(lambda x, y, mm: oracle.census_loss(x, y, mm), [a, b, m], (0, 1))
    assert_close(loss, g["loss_out0_f64"], TIGHT)
    assert_close(ga, g["loss_grad0_f64"], 1e-10)
    assert_close(gb, g["loss_grad1_f64"], 1e-10)
    (h, w), (ga, gb) = _grads(lambda x, y, mm: oracle.census_loss_no_penalty(x, y, mm), [a, b, m], (0, 1))
    assert_close(h, g["nopen_out0_f64"], TIGHT)
    assert_close(w, g["nopen_out1_f64"], TIGHT)
    assert_close(gb, g["nopen_grad1_f64"], 1e-10)
    (d,), (gb,) = _grads(lambda x, y: oracle.ternary_loss(x, y, 1)[0], [a, b], (1,))
    assert_close(d, g["tern1_out0_f64"], TIGHT)
    assert_close(gb, g["tern1_grad1_f64"], 1e-10)
    assert_close(oracle.ternary_loss(a.double(), b.double(), 3, True)[0], g["tern3s_out0_f64"], TIGHT)
    assert torch.equal(oracle.ternary_loss(a.double(), b.double(), 2)[1], g["tern2mask_out0_f64"])


def test_smooth_blocks(oracle):
    g = load_golden("smooth_blocks")
    flo, image = g["in0"], g["in1"]
    for key, fn in (("s1abs", lambda f, i: oracle.smooth_grad_1st(f, i, 10.0)),
                    ("s1uf", lambda f, i: oracle.smooth_grad_1st(f, i, 10.0, "uflow")),
                    ("s2", lambda f, i: oracle.smooth_grad_2nd(f, i, 10.0))):
        (out,), (gf,) = _grads(fn, [flo, image], (0,))
        assert_close(out, g[key + "_out0_f64"], TIGHT, key)
        assert_close(gf, g[key + "_grad0_f64"], 1e-10, key)


@pytest.mark.parametrize("order", [1, 2])
def test_uflow_loss(oracle, order):
    g = load_golden("uflow_loss_order%d" % order)
    ins = [g["in0"], g["in1"], g["in2"], g["in3"]]
    outs, (g0, g2) = _grads(lambda o0, o1, o2, t: oracle.uflow_loss([o0, o1, o2], t, smooth_order=order), ins, (0, 2))
    for k in range(5):
        assert_close(outs[k], g["out%d_f64" % k], 1e-11, "output %d" % k)
    assert_close(g0, g["grad0_f64"], 1e-9)
    assert_close(g2, g["grad2_f64"], 1e-9)


def test_triangular(oracle):
    g = load_golden("triag")
    for k in (1, 3):
        A, X = g["A%d" % k], g["X%d" % k]
        for key, tr in (("mv", False), ("mvT", True)):
            (out,), (gA, gX) = _grads(lambda a, x: oracle.stencil_mv(a, x, k, tr), [A, X], (0, 1))
            assert_close(out, g["%s%d_out0_f64" % (key, k)], TIGHT, key)
            assert_close(gA, g["%s%d_grad0_f64" % (key, k)], TIGHT, key + " dA")
            assert_close(gX, g["%s%d_grad1_f64" % (key, k)], TIGHT, key + " dX")
    a, b, c, d, x = (g["in%d" % i] for i in range(5))
    assert_close(oracle.substitution(a, b, c, d, x), g["fsub_out0_f64"], 1e-11)
    assert_close(oracle.substitution(a, b, c, d, x, upper=True), g["bsub_out0_f64"], 1e-11)
    a5, b5, c5 = a[:, :, :4, :5], b[:, :, :4, :4], c[:, :, :3, :5]
    assert_close(oracle.inverse_diagonal(a5, b5, c5), g["invdiag_out0_f64"], 1e-11)
    # L (L^-1 x) = x ties the solver to the product
    y = oracle.substitution(a, b, c, d, x)
    from_taps = torch.zeros(2, 8, 6, 7, dtype=torch.float64)
    for ch in range(2):
        from_taps[:, 0 + ch] = a[:, ch]
        from_taps[:, 2 + ch, :, :-1] = b[:, ch]
        from_taps[:, 4 + ch, :-1, :] = c[:, ch]
        from_taps[:, 6 + ch, :-1, :-1] = d[:, ch]
    assert_close(oracle.stencil_mv(from_taps, y, 1), x.double(), 1e-10)
    assert_close(oracle.stencil_mv(from_taps, y, 1), g["mv4_out0_f64"] * 0 + x.double(), 1e-10)


def test_ssim_and_resampler(oracle):
    g = load_golden("ssim_resampler")
    a, b, m = g["in0"], g["in1"], g["in2"]
    outs, (ga, gb) = _grads(lambda x, y, mm: tuple(oracle.ssim_loss(x, y, mm)[0]) + (oracle.ssim_loss(x, y, mm)[1],),
                            [a, b, m], (0, 1))
    for k in range(3):
        assert_close(outs[k], g["ssimloss_out%d_f64" % k], TIGHT)
    assert_close(ga, g["ssimloss_grad0_f64"], 1e-10)
    assert_close(gb, g["ssimloss_grad1_f64"], 1e-10)
    (o,), (ga, gb) = _grads(lambda x, y: oracle.ssim_valid(x, y, 1), [a, b], (0, 1))
    assert_close(o, g["ssim1_out0_f64"], TIGHT)
    assert_close(ga, g["ssim1_grad0_f64"], 1e-10)
    assert_close(oracle.ssim_valid(a.double(), b.double(), 2), g["ssim2_out0_f64"], TIGHT)
    (o,), (gd, gw) = _grads(lambda d, w: oracle.resampler(d, w[..., 0], w[..., 1]), [g["data"], g["warp"]], (0, 1))
    assert_close(o, g["resampler_out0_f64"], TIGHT)
    assert_close(gd, g["resampler_grad0_f64"], TIGHT)
    assert_close(gw, g["resampler_grad1_f64"], 1e-10)

"""Parity of masks, resize, census, smoothness and the whole UFlowLoss with the golden fixtures (B200)."""
import types

import pytest
import torch

from conftest import RTOL_GRAD, RTOL_VALUE, assert_close, load_golden

pytestmark = pytest.mark.gpu


def _grads(fn, inputs, wrt):
    ins = [i.cuda().clone().requires_grad_(k in wrt) for k, i in enumerate(inputs)]
    out = fn(*ins)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    g = torch.Generator().manual_seed(1234)
    loss = 0
    for o in outs:
        w = torch.randn(o.shape, generator=g)
        if o.requires_grad:
            loss = loss + (o * w.cuda()).sum()
    return outs, torch.autograd.grad(loss, [ins[k] for k in wrt])


def test_masks_golden():
    from arflow_b200 import uflow_utils as uu
    from arflow_b200 import warp_utils as wu
    g = load_golden("masks")
    flow, flow_b = g["in0"].cuda(), g["in1"].cuda()
    assert torch.equal(uu.mask_invalid(uu.flow_to_warp(flow)).cpu(), g["invalid_out0_f32"])
    assert_close(uu.compute_range_map(flow), g["range_out0_f64"], RTOL_VALUE)
    assert_close(wu.compute_range_map(flow), g["range_out0_f64"], RTOL_VALUE)
    assert_close(wu.get_corresponding_map(uu.flow_to_warp(flow)), g["corrmap_out0_f64"], RTOL_VALUE)
    assert torch.equal(wu.get_occu_mask_backward(flow, th=0.2).cpu(), g["occ_bw_out0_f32"])
    assert_close(wu.get_occu_mask_backward(flow, th=0.0), g["occ_bw0_out0_f64"], RTOL_VALUE)
    assert torch.equal(wu.border_mask(flow).cpu(), g["border_out0_f32"])
    assert torch.equal(wu.get_occu_mask_bidirection(flow * 0.3, flow_b * 0.3).cpu(), g["occ_bi_out0_f32"])


@pytest.mark.parametrize("shape", [(2, 24, 32), (3, 17, 20), (1, 9, 13), (2, 96, 128)])
def test_inside_masks_vs_torch(shape):
    """mask_invalid / border_mask for widths that take the four-pixels-per-thread kernel (W % 4 == 0) and the scalar one:
    bit-equal to the reference's comparisons on coordinates built by the same single fp32 add."""
    from arflow_b200 import uflow_utils as uu
    from arflow_b200 import warp_utils as wu
    B, H, W = shape
    gen = torch.Generator().manual_seed(B * H + W)
    flow = (torch.randn(B, 2, H, W, generator=gen) * 6).cuda()
    flow[:, :, 0, 0] = 0.0                      # a coordinate exactly on the border
    flow[:, 0, -1, -1] = 0.0
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
    x, y = xx + flow[:, 0], yy + flow[:, 1]
    ref0 = ((x >= 0) & (x <= W - 1) & (y >= 0) & (y <= H - 1)).float().unsqueeze(1)
    ref1 = ((x > 0) & (x < W - 1) & (y > 0) & (y < H - 1)).float().unsqueeze(1)
    assert torch.equal(uu.mask_invalid_flow(flow), ref0)
    assert torch.equal(wu.border_mask(flow), ref1)


def test_resize_golden():
    from arflow_b200 import uflow_utils as uu
    g = load_golden("resize")
    for s in (2, 4):
        for is_flow in (0, 1):
            (out,), (gi,) = _grads(lambda a: uu.upsample(a, bool(is_flow), scale_factor=float(s)), [g["in0"]], (0,))
            assert_close(out, g["up%d_%d_out0_f64" % (s, is_flow)], RTOL_VALUE)
            assert_close(gi, g["up%d_%d_grad0_f64" % (s, is_flow)], RTOL_GRAD)
            (out,), (gi,) = _grads(lambda a: uu.downsample(a, bool(is_flow), scale_factor=float(s)), [g["in1"]], (0,))
            assert_close(out, g["down%d_%d_out0_f64" % (s, is_flow)], RTOL_VALUE)
            assert_close(gi, g["down%d_%d_grad0_f64" % (s, is_flow)], RTOL_GRAD)


@pytest.mark.parametrize("shape", [(2, 2, 7, 5), (3, 2, 2, 2), (1, 3, 24, 32), (4, 2, 96, 128), (1, 2, 33, 130)])
def test_upsample_x2_fast_path_vs_aten_fp64(shape):
    """The exact x2 geometry (every flow up-sampling of the networks) runs on its own kernels; they must agree with
    ATen's float64 bilinear interpolation, values and gradient, at odd sizes, at the 2 x 2 minimum and at full size."""
    from arflow_b200 import uflow_utils as uu
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=gen)
    w = torch.randn(shape[0], shape[1], 2 * shape[2], 2 * shape[3], generator=gen)
    for is_flow in (False, True):
        xc = x.cuda().requires_grad_(True)
        out = uu.upsample(xc, is_flow, scale_factor=2.0)
        (gi,) = torch.autograd.grad((out * w.cuda()).sum(), [xc])
        xd = x.double().requires_grad_(True)
        ref = torch.nn.functional.interpolate(xd, scale_factor=2.0, mode="bilinear", align_corners=False) * (2.0 if is_flow else 1.0)
        (gr,) = torch.autograd.grad((ref * w.double()).sum(), [xd])
        assert_close(out, ref, RTOL_VALUE, "x2 value")
        assert_close(gi, gr, RTOL_GRAD, "x2 gradient")


def test_census_golden():
    from arflow_b200 import loss_blocks as lb
    from arflow_b200 import uflow_utils as uu
    g = load_golden("census")
    a, b, m = g["in0"], g["in1"], g["in2"]
    (loss,), (ga, gb) = _grads(lambda x, y, mm: uu.census_loss(x, y, mm), [a, b, m], (0, 1))
    assert_close(loss, g["loss_out0_f64"], RTOL_VALUE, "census_loss")
    assert_close(ga, g["loss_grad0_f64"], RTOL_GRAD, "d census_loss / d image_a")
    assert_close(gb, g["loss_grad1_f64"], RTOL_GRAD, "d census_loss / d image_b")
    (h, w), (ga, gb) = _grads(lambda x, y, mm: uu.census_loss_no_penalty(x, y, mm), [a, b, m], (0, 1))
    assert_close(h, g["nopen_out0_f64"], RTOL_VALUE, "hamming")
    assert_close(w, g["nopen_out1_f64"], RTOL_VALUE, "weight")
    assert_close(ga, g["nopen_grad0_f64"], RTOL_GRAD)
    assert_close(gb, g["nopen_grad1_f64"], RTOL_GRAD)
    (d,), (gb,) = _grads(lambda x, y: lb.TernaryLoss(x, y, max_distance=1)[0], [a, b], (1,))
    assert_close(d, g["tern1_out0_f64"], RTOL_VALUE)
    assert_close(gb, g["tern1_grad1_f64"], RTOL_GRAD)
    d3, _ = lb.TernaryLoss(a.cuda(), b.cuda(), max_distance=3, sum_dist=True)
    assert_close(d3, g["tern3s_out0_f64"], RTOL_VALUE)
    assert torch.equal(lb.TernaryLoss(a.cuda(), b.cuda(), max_distance=2)[1].cpu(), g["tern2mask_out0_f32"])
    # the explicit transform helper agrees with the fused kernel
    ham = uu.soft_hamming(uu.census_transform(a.cuda(), 7), uu.census_transform(b.cuda(), 7))
    assert_close(ham, g["nopen_out0_f64"], RTOL_VALUE)


@pytest.mark.parametrize("strip", [8, 16, 24])
def test_census_pair_symmetric_kernels(strip):
    """The pair-symmetric strip kernels (every unordered pixel pair evaluated once) against the golden fixture and,
    on ragged shapes, against the per-pixel kernels: Hamming map, fused loss and both image gradients."""
    from arflow_b200 import _lib
    from arflow_b200 import loss_blocks as lb
    from arflow_b200 import uflow_utils as uu
    lib = _lib.load()
    try:
        lib.arf_debug_set(5, strip)
        g = load_golden("census")
        a, b, m = g["in0"], g["in1"], g["in2"]
        (loss,), (ga, gb) = _grads(lambda x, y, mm: uu.census_loss(x, y, mm), [a, b, m], (0, 1))
        assert_close(loss, g["loss_out0_f64"], RTOL_VALUE, "census_loss")
        assert_close(ga, g["loss_grad0_f64"], RTOL_GRAD, "d census_loss / d image_a")
        assert_close(gb, g["loss_grad1_f64"], RTOL_GRAD, "d census_loss / d image_b")
        (h, w), (ga, gb) = _grads(lambda x, y, mm: uu.census_loss_no_penalty(x, y, mm), [a, b, m], (0, 1))
        assert_close(h, g["nopen_out0_f64"], RTOL_VALUE, "hamming")
        assert_close(ga, g["nopen_grad0_f64"], RTOL_GRAD)
        assert_close(gb, g["nopen_grad1_f64"], RTOL_GRAD)
        (d,), (gb,) = _grads(lambda x, y: lb.TernaryLoss(x, y, max_distance=1)[0], [a, b], (1,))
        assert_close(d, g["tern1_out0_f64"], RTOL_VALUE)
        assert_close(gb, g["tern1_grad1_f64"], RTOL_GRAD)
        gen = torch.Generator().manual_seed(5)
        for (B, H, W) in [(2, 37, 71), (1, 16, 32), (3, 100, 130), (1, 9, 200), (2, 64, 58), (1, 65, 59)]:
            for md in (1, 2, 3):
                x, y = torch.rand(B, 3, H, W, generator=gen), torch.rand(B, 3, H, W, generator=gen)
                mm = (torch.rand(B, 1, H, W, generator=gen) > 0.3).float()
                res = {}
                for variant in (1, strip):
                    lib.arf_debug_set(5, variant)
                    if md == 3:
                        fn = lambda p, r, k: uu.census_loss(p, r, k)
                        res[variant] = _grads(fn, [x, y, mm], (0, 1))
                    else:
                        fn = lambda p, r: lb.TernaryLoss(p, r, max_distance=md)[0]
                        res[variant] = _grads(fn, [x, y], (0, 1))
                (o_ref,), g_ref = res[1]
                (o_sym,), g_sym = res[strip]
                assert_close(o_sym, o_ref.cpu(), RTOL_VALUE, "value %s md %d" % ((B, H, W), md))
                for u, v in zip(g_sym, g_ref):
                    assert_close(u, v.cpu(), RTOL_GRAD, "grad %s md %d" % ((B, H, W), md))
    finally:
        lib.arf_debug_set(5, 0)


def test_smooth_blocks_golden():
    from arflow_b200 import loss_blocks as lb
    g = load_golden("smooth_blocks")
    flo, image = g["in0"], g["in1"]
    for key, fn in (("s1abs", lambda f, i: lb.smooth_grad_1st(f, i, 10.0)),
                    ("s1uf", lambda f, i: lb.smooth_grad_1st(f, i, 10.0, penalty="uflow")),
                    ("s2", lambda f, i: lb.smooth_grad_2nd(f, i, 10.0))):
        (out,), (gf,) = _grads(fn, [flo, image], (0,))
        assert_close(out, g[key + "_out0_f64"], RTOL_VALUE, key)
        assert_close(gf, g[key + "_grad0_f64"], RTOL_GRAD, key)


@pytest.mark.parametrize("order", [1, 2])
def test_uflow_loss_golden(order):
    from arflow_b200.uflow_loss import UFlowLoss
    g = load_golden("uflow_loss_order%d" % order)
    cfg = types.SimpleNamespace(edge_constant=150, w_smooth=4.0, w_census=1.0, with_bk=True, smooth_order=order)
    ins = [g["in0"], g["in1"], g["in2"], g["in3"]]
    outs, (g0, g2) = _grads(lambda o0, o1, o2, t: UFlowLoss(cfg)([o0, o1, o2], t), ins, (0, 2))
    # the parity target of the warp inside the loss is the fp32 grid_sample path (see test_warp_gpu._check)
    for k in range(4):
        assert_close(outs[k], g["out%d_f32" % k], RTOL_VALUE, "output %d" % k)
    assert_close(outs[4], g["out4_f32"], 1e-4, "mask1")
    assert_close(g0, g["grad0_f32"], RTOL_GRAD, "d/d output[0]")
    assert_close(g2, g["grad2_f32"], RTOL_GRAD, "d/d output[2]")
    assert_close(outs[0], g["out0_f64"], 1e-4, "total vs float64")


def test_uflow_loss_full_size_vs_oracle(oracle):
    """Config-2 shape (B reduced to 2 so the oracle finishes in seconds): loss values and gradients."""
    from arflow_b200.uflow_loss import UFlowLoss
    gen = torch.Generator().manual_seed(21)
    B, H, W = 2, 384, 512
    o0 = torch.randn(B, 4, H, W, generator=gen) * 3
    o1 = torch.randn(B, 4, H // 2, W // 2, generator=gen)
    o2 = torch.randn(B, 4, H // 4, W // 4, generator=gen)
    t = torch.rand(B, 6, H, W, generator=gen)
    cfg = types.SimpleNamespace(edge_constant=150, w_smooth=4.0, w_census=1.0, with_bk=True, smooth_order=1)
    a0, a2 = o0.cuda().requires_grad_(True), o2.cuda().requires_grad_(True)
    out = UFlowLoss(cfg)([a0, o1.cuda(), a2], t.cuda())
    out[0].backward()
    r0, r2 = o0.clone().requires_grad_(True), o2.clone().requires_grad_(True)
    ref = oracle.uflow_loss([r0, o1, r2], t)
    ref[0].backward()
    for k in range(4):
        assert_close(out[k], ref[k], RTOL_VALUE, "output %d" % k)
    # the oracle runs in float32 here too: through a bilinear warp of white-noise images its own element-wise noise in
    # the near-zero regions of this gradient (values span 1e-16 .. 1e-6) is of the order of the element-wise bar
    assert_close(a0.grad, r0.grad, RTOL_GRAD, "d/d output[0]", elementwise=False)
    assert_close(a2.grad, r2.grad, RTOL_GRAD, "d/d output[2]", elementwise=False)


def test_ssim_and_resampler_golden():
    from arflow_b200 import loss_blocks as lb
    from arflow_b200 import uflow_resampler as ur
    from arflow_b200 import uflow_utils as uu
    g = load_golden("ssim_resampler")
    a, b, m = g["in0"], g["in1"], g["in2"]
    outs, (ga, gb) = _grads(lambda x, y, mm: tuple(uu.ssim_loss(x, y, mm)[0]) + (uu.ssim_loss(x, y, mm)[1],),
                            [a, b, m], (0, 1))
    for k in range(3):
        assert_close(outs[k], g["ssimloss_out%d_f64" % k], RTOL_VALUE, "ssim_loss out %d" % k)
    assert_close(ga, g["ssimloss_grad0_f64"], RTOL_GRAD)
    assert_close(gb, g["ssimloss_grad1_f64"], RTOL_GRAD)
    (o,), (ga, gb) = _grads(lambda x, y: lb.SSIM(x, y, md=1), [a, b], (0, 1))
    assert_close(o, g["ssim1_out0_f64"], RTOL_VALUE)
    assert_close(ga, g["ssim1_grad0_f64"], RTOL_GRAD)
    assert_close(gb, g["ssim1_grad1_f64"], RTOL_GRAD)
    assert_close(lb.SSIM(a.cuda(), b.cuda(), md=2), g["ssim2_out0_f64"], RTOL_VALUE)
    (o,), (gd, gw) = _grads(lambda d, w: ur.resampler(d, w), [g["data"], g["warp"]], (0, 1))
    assert_close(o, g["resampler_out0_f64"], RTOL_VALUE)
    assert_close(gd, g["resampler_grad0_f64"], RTOL_GRAD)
    assert_close(gw, g["resampler_grad1_f64"], RTOL_GRAD)


@pytest.mark.parametrize("small", [True, False])
def test_census_loss_groups_equals_separate_calls(small):
    """census_loss_groups on stacked inputs (one launch for both slices) == one census_loss per batch slice: the per-pixel
    arithmetic is the same, only the order of the partial sums of the normaliser differs (double accumulation)."""
    from arflow_b200 import uflow_utils as uu
    gen = torch.Generator().manual_seed(9)
    B, H, W = (2, 40, 72) if small else (4, 128, 160)     # per-pixel kernels / pair-symmetric strip kernels
    a, b = torch.rand(2 * B, 3, H, W, generator=gen), torch.rand(2 * B, 3, H, W, generator=gen)
    m = (torch.rand(2 * B, 1, H, W, generator=gen) > 0.3).float()
    ac, bc = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    l = uu.census_loss_groups(ac, bc, m.cuda(), 2)
    ga, gb = torch.autograd.grad(l[0] * 1.5 + l[1] * 0.5, [ac, bc])
    for g, wgt in ((0, 1.5), (1, 0.5)):
        sl = slice(g * B, (g + 1) * B)
        a1, b1 = a[sl].cuda().requires_grad_(True), b[sl].cuda().requires_grad_(True)
        l1 = uu.census_loss(a1, b1, m[sl].cuda())
        g1a, g1b = torch.autograd.grad(l1 * wgt, [a1, b1])
        assert_close(l[g], l1, 1e-6, "loss of slice %d" % g)
        assert_close(ga[sl], g1a, 1e-6, "d/d image_a")
        assert_close(gb[sl], g1b, 1e-6, "d/d image_b")


@pytest.mark.parametrize("C", [3, 8, 12, 40])
def test_resampler_both_backward_kernels_vs_oracle(oracle, C):
    """arf_resampler_bwd owns a sample point by a thread (C <= 8) or by a warp (wider data); both against the float64
    oracle (uflow_resampler.py:155-241), coordinates partly outside the image."""
    from arflow_b200 import uflow_resampler as ur
    gen = torch.Generator().manual_seed(C)
    data = torch.randn(2, 9, 11, C, generator=gen)
    warp = torch.rand(2, 5, 7, 2, generator=gen) * torch.tensor([13.0, 11.0]) - 1.5
    w = torch.randn(2, 5, 7, C, generator=gen)
    d64, w64 = data.double().requires_grad_(True), warp.double().requires_grad_(True)
    ref = oracle.resampler(d64, w64[..., 0], w64[..., 1])
    rd, rw = torch.autograd.grad((ref * w.double()).sum(), [d64, w64])
    dc, wc = data.cuda().requires_grad_(True), warp.cuda().requires_grad_(True)
    out = ur.resampler(dc, wc)
    gd, gw = torch.autograd.grad((out * w.cuda()).sum(), [dc, wc])
    assert_close(out, ref, RTOL_VALUE, "resampler")
    assert_close(gd, rd, RTOL_GRAD, "d / d data")
    assert_close(gw, rw, RTOL_GRAD, "d / d warp")


@pytest.mark.parametrize("name", ["elbo_sparse", "elbo_diag"])
def test_uflow_elbo_loss_golden(name):
    """UFlowElboLoss (non-diagonal stencil covariance, and diagonal with closed-form smoothness, out-of-frame
    and occlusion penalties) against the reference run with the same injected noise."""
    import json
    import numpy as np
    import os
    from conftest import GOLDEN
    from arflow_b200.uflow_elbo_loss import UFlowElboLoss
    g = load_golden(name) if False else None
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        cfg = types.SimpleNamespace(**json.loads(str(z["cfg"])))
        t = {k: torch.from_numpy(z[k]) for k in z.files if k != "cfg"}
    eps = [t["eps0"].cuda(), t["eps1"].cuda()]

    def run(f, b, i1, i2):
        loss = UFlowElboLoss(cfg)
        it = iter(eps)
        loss._normal = lambda size, like: next(it)
        out = loss({"flows_fw": [None, None, f], "flows_bw": [None, None, b]}, i1, i2)
        o4 = out[4] if torch.is_tensor(out[4]) else torch.zeros((), device=f.device)
        return out[:4] + (o4,)

    outs, (gf, gb) = _grads(run, [t["in0"], t["in1"], t["in2"], t["in3"]], (0, 1))
    for k in range(5):
        assert_close(outs[k], t["out%d_f32" % k], 2e-5, "%s output %d" % (name, k))
    # The reference's own fp32 gradient sits 2.5e-4 (sparse) / 8.6e-5 (diag) away from its float64 gradient on
    # these inputs (x4 upsample -> coordinate round trip -> census amplifies fp32 rounding), so 1e-4 against
    # either is below the reference's noise floor here; the bar is 2x that floor, against both.
    tol = 5e-4 if name == "elbo_sparse" else 2e-4
    assert_close(gf, t["grad0_f32"], tol, "d/d flows_fw[2] vs fp32 reference")
    assert_close(gb, t["grad1_f32"], tol, "d/d flows_bw[2] vs fp32 reference")
    assert_close(gf, t["grad0_f64"], tol, "d/d flows_fw[2] vs float64 reference")
    assert_close(gb, t["grad1_f64"], tol, "d/d flows_bw[2] vs float64 reference")

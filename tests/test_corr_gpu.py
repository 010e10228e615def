"""Parity of the correlation kernels with the oracle / golden fixtures (B200, through the C-ABI)."""
import pytest
import torch

from conftest import RTOL_GRAD, RTOL_VALUE, assert_close, load_golden

pytestmark = pytest.mark.gpu


def _run(f1, f2, w, **kw):
    from arflow_b200.correlation import Correlation
    f1 = f1.cuda().requires_grad_(True)
    f2 = f2.cuda().requires_grad_(True)
    out = Correlation(**kw)(f1, f2)
    g1, g2 = torch.autograd.grad((out * w.cuda().float()).sum(), [f1, f2])
    return out, g1, g2


@pytest.mark.parametrize("name", ["corr_b2c5_9x11", "corr_b1c32_12x16", "corr_b1c3_6x40"])
def test_golden(name):
    g = load_golden(name)
    w = torch.randn(g["out0_f64"].shape, generator=torch.Generator().manual_seed(1234))
    out, g1, g2 = _run(g["in0"], g["in1"], w, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)
    assert_close(out, g["out0_f64"], RTOL_VALUE, "cost volume")
    assert_close(g1, g["grad0_f64"], RTOL_GRAD, "grad f1")
    assert_close(g2, g["grad1_f64"], RTOL_GRAD, "grad f2")


@pytest.mark.parametrize("shape", [(2, 32, 24, 32), (1, 196, 6, 8), (3, 7, 33, 65), (1, 64, 48, 64), (2, 96, 12, 20)])
def test_fast_path_vs_oracle(oracle, shape):
    gen = torch.Generator().manual_seed(sum(shape))
    f1, f2 = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen)
    ref = oracle.corr_fwd_c(f1, f2)
    w = torch.randn(ref.shape, generator=gen)
    r1, r2 = oracle.corr_bwd_c(f1, f2, w)
    out, g1, g2 = _run(f1, f2, w, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)
    assert_close(out, ref, RTOL_VALUE, "cost volume")
    assert_close(g1, r1, RTOL_GRAD, "grad f1")
    assert_close(g2, r2, RTOL_GRAD, "grad f2")


@pytest.mark.parametrize("geom", [(3, 1, 3, 1, 1), (4, 3, 4, 1, 1), (6, 3, 4, 2, 2), (20, 1, 20, 1, 2), (2, 1, 4, 1, 1),
                                  (4, 1, 4, 2, 1), (5, 5, 3, 3, 1)])
def test_general_geometry_vs_oracle(oracle, geom):
    pad, ks, md, s1, s2 = geom
    gen = torch.Generator().manual_seed(pad * 100 + ks * 10 + md)
    shape = (2, 6, 26, 30) if md < 20 else (1, 4, 30, 34)
    f1, f2 = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen)
    ref = oracle.corr_fwd_c(f1, f2, pad, ks, md, s1, s2)
    w = torch.randn(ref.shape, generator=gen)
    r1, r2 = oracle.corr_bwd_c(f1, f2, w, pad, ks, md, s1, s2)
    out, g1, g2 = _run(f1, f2, w, pad_size=pad, kernel_size=ks, max_displacement=md, stride1=s1, stride2=s2)
    assert_close(out, ref, RTOL_VALUE, "cost volume")
    assert_close(g1, r1, RTOL_GRAD, "grad f1")
    assert_close(g2, r2, RTOL_GRAD, "grad f2")


def test_native_style_constructor_and_cost_volume_api(oracle):
    from arflow_b200.correlation import Correlation, compute_cost_volume
    gen = torch.Generator().manual_seed(7)
    f1, f2 = torch.randn(2, 32, 20, 28, generator=gen), torch.randn(2, 32, 20, 28, generator=gen)
    ref = oracle.cost_volume(f1.double(), f2.double())
    a = Correlation(max_displacement=4, kernel_size=1, stride1=1, stride2=1, corr_multiply=1)(f1.cuda(), f2.cuda())
    b = compute_cost_volume(f1.cuda(), f2.cuda(), 4)
    assert_close(a, ref, RTOL_VALUE)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        compute_cost_volume(f1.cuda()[:, :, :4], f2.cuda()[:, :, :4], 4)


def test_linearity_at_full_size():
    """Size-independent property at the config-2 L1 shape: corr(a*f1, f2) == a*corr(f1, f2), and the
    zero-displacement plane equals mean_c(f1*f2)."""
    from arflow_b200.correlation import compute_cost_volume
    gen = torch.Generator().manual_seed(11)
    f1 = torch.randn(8, 32, 96, 128, generator=gen).cuda()
    f2 = torch.randn(8, 32, 96, 128, generator=gen).cuda()
    cv = compute_cost_volume(f1, f2, 4)
    assert_close(compute_cost_volume(2.0 * f1, f2, 4), 2.0 * cv, 1e-6)
    assert_close(cv[:, 40], (f1 * f2).mean(1), RTOL_VALUE)
    # shifting f2 by one pixel moves the planes: plane(dy,dx) of shifted == plane(dy,dx+1) of original (interior)
    f2s = torch.roll(f2, shifts=1, dims=3)
    cvs = compute_cost_volume(f1, f2s, 4)
    assert_close(cvs[:, 41, :, 8:-8], cv[:, 40, :, 8:-8], RTOL_VALUE)


@pytest.mark.parametrize("shape", [(2, 32, 24, 32), (1, 20, 17, 44), (2, 8, 40, 36)])
def test_cp_async_staging_path_matches_tma_path(oracle, shape):
    """W % 4 == 0 shapes run the TMA producer; the debug hook forces the cp.async producer that serves
    every other width.  Both must give the same cost volume."""
    from arflow_b200 import _lib
    from arflow_b200.correlation import compute_cost_volume
    gen = torch.Generator().manual_seed(5)
    f1, f2 = torch.randn(shape, generator=gen).cuda(), torch.randn(shape, generator=gen).cuda()
    a = compute_cost_volume(f1, f2, 4)
    _lib.call("arf_debug_set", 0, 1)
    try:
        b = compute_cost_volume(f1, f2, 4)
    finally:
        _lib.call("arf_debug_set", 0, 0)
    assert torch.equal(a, b)
    assert_close(a, oracle.corr_fwd_c(f1.cpu(), f2.cpu()), RTOL_VALUE)



@pytest.mark.parametrize("shape", [(2, 32, 24, 32), (3, 20, 50, 68), (1, 7, 37, 100), (2, 196, 12, 16), (1, 64, 13, 36)])
def test_both_tiled_forward_kernels_vs_oracle(oracle, shape):
    """arf_corr_fwd routes between three tiled launches by a cost model (32x12 or 32x4 tiles with lane pairs sharing
    64-bit operands / 32x8 column-thread tiles); the debug hook forces each in turn.  Ragged heights, widths that leave partial
    tiles and channel counts that leave partial stages: both must match the C oracle and each other bit for bit (same
    FMA order per accumulator)."""
    from arflow_b200 import _lib
    from arflow_b200.correlation import compute_cost_volume
    gen = torch.Generator().manual_seed(sum(shape))
    f1, f2 = torch.randn(shape, generator=gen).cuda(), torch.randn(shape, generator=gen).cuda()
    ref = oracle.corr_fwd_c(f1.cpu(), f2.cpu())
    outs = []
    for variant in (30, 31, 32):
        _lib.call("arf_debug_set", 1, variant)
        try:
            outs.append(compute_cost_volume(f1, f2, 4))
        finally:
            _lib.call("arf_debug_set", 1, 0)
        assert_close(outs[-1], ref, RTOL_VALUE, "variant %d" % variant)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("shape", [(2, 32, 24, 32), (3, 20, 50, 68), (1, 64, 13, 36), (2, 40, 12, 16)])
def test_all_tiled_backward_item_shapes_vs_oracle(oracle, shape):
    """arf_corr_bwd picks one of six work-item shapes by a cost model; the debug hook forces each.  All must match the C
    oracle and each other bit for bit (the per-element order of the fused multiply-adds does not depend on the item)."""
    from arflow_b200 import _lib
    gen = torch.Generator().manual_seed(sum(shape))
    f1, f2 = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen)
    w = torch.randn(shape[0], 81, shape[2], shape[3], generator=gen)
    r1, r2 = oracle.corr_bwd_c(f1, f2, w)
    res = []
    for variant in (10, 11, 12, 13, 14, 15):
        _lib.call("arf_debug_set", 1, variant)
        try:
            _, g1, g2 = _run(f1, f2, w, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)
        finally:
            _lib.call("arf_debug_set", 1, 0)
        assert_close(g1, r1, RTOL_GRAD, "grad f1, item shape %d" % variant)
        assert_close(g2, r2, RTOL_GRAD, "grad f2, item shape %d" % variant)
        res.append((g1, g2))
    for g1, g2 in res[1:]:
        assert torch.equal(g1, res[0][0]) and torch.equal(g2, res[0][1])


# Full BASELINE sizes against the C oracle (double accumulation): config 2's finest level, and batch 16 at the
# channel counts / pyramid levels of the PWC-Lite family (models/pwclite.py:113: C = 64, 96, 128, 192 at 1/8 .. 1/64
# of 384x512) plus config 1's own level shapes.
@pytest.mark.parametrize("shape", [(8, 32, 96, 128), (16, 64, 48, 64), (16, 96, 24, 32), (16, 128, 12, 16),
                                   (16, 192, 6, 8), (1, 192, 6, 10), (1, 128, 12, 20), (1, 96, 24, 40),
                                   (1, 64, 48, 80), (1, 32, 96, 160)])
def test_full_size_vs_oracle(oracle, shape):
    gen = torch.Generator().manual_seed(sum(shape))
    f1, f2 = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen)
    ref = oracle.corr_fwd_c(f1, f2)
    w = torch.randn(ref.shape, generator=gen)
    r1, r2 = oracle.corr_bwd_c(f1, f2, w)
    out, g1, g2 = _run(f1, f2, w, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)
    assert_close(out, ref, RTOL_VALUE, "cost volume")
    assert_close(g1, r1, RTOL_GRAD, "grad f1")
    assert_close(g2, r2, RTOL_GRAD, "grad f2")
    # second, element-wise check with an absolute floor: |a - ref| <= rtol * (|ref| + rms(ref))
    rms = ref.double().pow(2).mean().sqrt().item()
    err = (out.detach().double().cpu() - ref.double()).abs()
    assert bool((err <= 1e-5 * (ref.double().abs() + rms)).all()), "element-wise: max %.3e" % err.max().item()

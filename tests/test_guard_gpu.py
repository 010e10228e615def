"""Own out-of-bounds check of the hot-path kernels (compute-sanitizer is closed on this GPU pool, see
profiles/r2_sanitizer_closed.txt): every OUTPUT of a C-ABI call lives inside a larger buffer whose guard bands hold a
sentinel, inputs sit at the very end of their allocations' used range with NaN guard bands after them (an out-of-range
READ that reaches a result shows up as NaN), shapes are ragged so partial tiles, partial TMA boxes and the cp.async
staging paths all run.  After the call the guards must be untouched and the results finite."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096          # floats on either side
SENT = -12345.5


class Arena:
    def __init__(self):
        self.bufs = []

    def out(self, *shape):
        n = 1
        for s in shape:
            n *= s
        pad = (-n) % 4                                    # keep the carved tensor 16-byte aligned (TMA paths)
        raw = torch.full((GUARD + n + pad + GUARD,), SENT, device="cuda")
        self.bufs.append((raw, n))
        return raw[GUARD:GUARD + n].view(*shape)

    def inp(self, t):
        n = t.numel()
        raw = torch.full((GUARD + n + ((-n) % 4) + GUARD,), float("nan"), device="cuda")
        raw[GUARD:GUARD + n] = t.flatten().cuda()
        return raw[GUARD:GUARD + n].view(*t.shape)

    def check(self, what):
        torch.cuda.synchronize()
        for raw, n in self.bufs:
            assert bool((raw[:GUARD] == SENT).all()), what + ": wrote before an output"
            assert bool((raw[GUARD + n:] == SENT).all()), what + ": wrote past an output"
            assert bool(torch.isfinite(raw[GUARD:GUARD + n]).all()), what + ": non-finite result (out-of-range read?)"
            assert bool((raw[GUARD:GUARD + n] != SENT).any()), what + ": output never written"


def _cs():
    return torch.cuda.current_stream().cuda_stream


CORR_SHAPES = [(2, 32, 24, 32), (1, 20, 13, 36), (3, 7, 33, 65), (1, 192, 6, 10), (2, 40, 50, 68), (1, 9, 5, 7), (16, 32, 12, 16)]


@pytest.mark.parametrize("shape", CORR_SHAPES)
@pytest.mark.parametrize("variant", [0, 30, 31, 32, 10, 11, 12, 13, 14, 15])   # 30-32: forward launches, 10-15: backward item shapes
def test_correlation_stays_in_bounds(shape, variant):
    from arflow_b200 import _lib
    lib = _lib.load()
    B, C, H, W = shape
    g = torch.Generator().manual_seed(1)
    ar = Arena()
    f1, f2 = ar.inp(torch.randn(shape, generator=g)), ar.inp(torch.randn(shape, generator=g))
    go = ar.inp(torch.randn(B, 81, H, W, generator=g))
    out, g1, g2 = ar.out(B, 81, H, W), ar.out(*shape), ar.out(*shape)
    lib.arf_debug_set(1, variant)
    try:
        assert lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), out.data_ptr(), B, C, H, W, 4, 1, 4, 1, 1, _cs()) == 0
        assert lib.arf_corr_bwd(f1.data_ptr(), f2.data_ptr(), go.data_ptr(), g1.data_ptr(), g2.data_ptr(), B, C, H, W,
                                4, 1, 4, 1, 1, _cs()) == 0
    finally:
        lib.arf_debug_set(1, 0)
    ar.check("correlation %s variant %d" % (shape, variant))


@pytest.mark.parametrize("geom", [(3, 1, 3, 1, 1), (6, 3, 4, 2, 2), (20, 1, 20, 1, 2), (4, 3, 4, 1, 1)])
def test_literal_correlation_stays_in_bounds(geom):
    from arflow_b200 import _lib
    import ctypes
    lib = _lib.load()
    pad, ks, md, s1, s2 = geom
    B, C, H, W = 2, 6, 26, 30
    d2, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.arf_corr_out_dims(H, W, pad, ks, md, s1, s2, ctypes.byref(d2), ctypes.byref(oh), ctypes.byref(ow)) == 0
    g = torch.Generator().manual_seed(2)
    ar = Arena()
    f1, f2 = ar.inp(torch.randn(B, C, H, W, generator=g)), ar.inp(torch.randn(B, C, H, W, generator=g))
    go = ar.inp(torch.randn(B, d2.value, oh.value, ow.value, generator=g))
    out = ar.out(B, d2.value, oh.value, ow.value)
    assert lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), out.data_ptr(), B, C, H, W, pad, ks, md, s1, s2, _cs()) == 0
    if s1 == 1:      # for stride1 > 1 the reference's own backward writes out of bounds; ours is only defined at s1 == 1
        g1, g2 = ar.out(B, C, H, W), ar.out(B, C, H, W)
        assert lib.arf_corr_bwd(f1.data_ptr(), f2.data_ptr(), go.data_ptr(), g1.data_ptr(), g2.data_ptr(), B, C, H, W,
                                pad, ks, md, s1, s2, _cs()) == 0
    ar.check("literal correlation %s" % (geom,))


@pytest.mark.parametrize("shape", [(2, 32, 24, 32), (1, 3, 37, 53), (3, 5, 9, 130), (1, 32, 96, 128)])
@pytest.mark.parametrize("scale", [1.0, 40.0])
def test_warp_stays_in_bounds(shape, scale):
    """scale 40: most taps fall outside the source - every tap predicate is exercised."""
    from arflow_b200 import _lib
    lib = _lib.load()
    B, C, H, W = shape
    g = torch.Generator().manual_seed(3)
    ar = Arena()
    x, gy = ar.inp(torch.randn(shape, generator=g)), ar.inp(torch.randn(shape, generator=g))
    fl = ar.inp(torch.randn(B, 2, H, W, generator=g) * scale)
    y, gx, gf = ar.out(*shape), ar.out(*shape), ar.out(B, 2, H, W)
    a = (B, C, H, W, H, W, float(W - 1), float(H - 1), 0, 0, 0, 1)
    assert lib.arf_warp_fwd(x.data_ptr(), fl.data_ptr(), y.data_ptr(), *a, _cs()) == 0
    assert lib.arf_warp_bwd(x.data_ptr(), fl.data_ptr(), gy.data_ptr(), gx.data_ptr(), gf.data_ptr(), *a, _cs()) == 0
    ar.check("warp %s scale %g" % (shape, scale))


@pytest.mark.parametrize("shape", [(1, 40, 70), (2, 64, 96), (1, 33, 58), (2, 96, 200)])
def test_census_stays_in_bounds(shape):
    from arflow_b200 import _lib
    lib = _lib.load()
    B, H, W = shape
    g = torch.Generator().manual_seed(4)
    ar = Arena()
    a, b = ar.inp(torch.rand(B, 3, H, W, generator=g)), ar.inp(torch.rand(B, 3, H, W, generator=g))
    m = ar.inp(torch.rand(B, 1, H, W, generator=g))
    npart = lib.arf_census_num_partials(B, H, W)
    ham, part, sums = ar.out(B, 1, H, W), ar.out(2 * npart), ar.out(3)
    ga, gb = ar.out(B, 3, H, W), ar.out(B, 3, H, W)
    gl = torch.ones(1, device="cuda")
    assert lib.arf_census_fwd(a.data_ptr(), b.data_ptr(), m.data_ptr(), ham.data_ptr(), part.data_ptr(), sums.data_ptr(),
                              B, H, W, 7, 1.0, 0.01, 0.4, _cs()) == 0
    torch.cuda.synchronize()
    assert lib.arf_census_bwd(a.data_ptr(), b.data_ptr(), None, ham.data_ptr(), m.data_ptr(), sums.data_ptr(),
                              gl.data_ptr(), ga.data_ptr(), gb.data_ptr(), B, H, W, 7, 1.0, 0.01, 0.4, _cs()) == 0
    torch.cuda.synchronize()
    # the partial-sum workspace is sized for the largest kernel family; only its used prefix is written
    ar.bufs = [(r, n) for (r, n) in ar.bufs if n != 2 * npart]
    ar.check("census %s" % (shape,))


@pytest.mark.parametrize("shape", [(3, 37, 53, 1), (2, 24, 130, 3), (2, 112, 256, 3)])
@pytest.mark.parametrize("transposed", [0, 1])
def test_stencil_product_stays_in_bounds(shape, transposed):
    from arflow_b200 import _lib
    lib = _lib.load()
    N, M, Nn, k = shape
    g = torch.Generator().manual_seed(5)
    taps = (k + 1) ** 2
    ar = Arena()
    A, X = ar.inp(torch.randn(N, 2 * taps, M, Nn, generator=g)), ar.inp(torch.randn(N, 2, M, Nn, generator=g))
    gY = ar.inp(torch.randn(N, 2, M, Nn, generator=g))
    Y, dA, dX = ar.out(N, 2, M, Nn), ar.out(N, 2 * taps, M, Nn), ar.out(N, 2, M, Nn)
    assert lib.arf_stencil_mv_fwd(A.data_ptr(), X.data_ptr(), Y.data_ptr(), N, M, Nn, k, transposed, _cs()) == 0
    assert lib.arf_stencil_mv_bwd(A.data_ptr(), X.data_ptr(), gY.data_ptr(), dA.data_ptr(), dX.data_ptr(), N, M, Nn, k,
                                  transposed, _cs()) == 0
    ar.check("stencil %s" % (shape,))


@pytest.mark.parametrize("shape", [(5, 37, 53), (3, 9, 1030), (2, 112, 256), (1, 1, 7)])
@pytest.mark.parametrize("upper", [0, 1])
def test_substitution_stays_in_bounds(shape, upper):
    """row-scan kernel (N <= 1024) and the wavefront fall-back (N = 1030), with and without the diagonal neighbour."""
    from arflow_b200 import _lib
    lib = _lib.load()
    S, M, Nn = shape
    g = torch.Generator().manual_seed(6)
    ar = Arena()
    A = ar.inp(torch.rand(S, M, Nn, generator=g) + 1.5)
    Bm = ar.inp(torch.randn(S, M, max(Nn - 1, 1), generator=g)[:, :, :Nn - 1].contiguous() * 0.3) if Nn > 1 else None
    Cm = ar.inp(torch.randn(S, max(M - 1, 1), Nn, generator=g)[:, :M - 1].contiguous() * 0.3) if M > 1 else None
    Dm = ar.inp(torch.randn(S, M - 1, Nn - 1, generator=g) * 0.3) if (M > 1 and Nn > 1) else None
    X = ar.inp(torch.randn(S, M, Nn, generator=g))
    Y = ar.out(S, M, Nn)
    ptr = lambda t: t.data_ptr() if t is not None else X.data_ptr()     # never dereferenced when the extent is empty
    assert lib.arf_trisolve(A.data_ptr(), ptr(Bm), ptr(Cm), Dm.data_ptr() if Dm is not None else None, X.data_ptr(),
                            Y.data_ptr(), S, M, Nn, upper, _cs()) == 0
    ar.check("substitution %s upper=%d" % (shape, upper))


@pytest.mark.parametrize("wild", [False, True])
def test_resampler_stays_in_bounds(wild):
    """NHWC resampler, interleaved (x, y) coordinates (element stride 2); wild = coordinates far outside the image,
    +-inf and NaN among them: such taps must contribute nothing and write nothing."""
    from arflow_b200 import _lib
    lib = _lib.load()
    B, H, W, C = 2, 13, 21, 5
    g = torch.Generator().manual_seed(7)
    ar = Arena()
    data = ar.inp(torch.randn(B, H, W, C, generator=g))
    xy = torch.rand(B, H, W, 2, generator=g) * torch.tensor([W - 1.0, H - 1.0])
    if wild:
        xy = (xy - 5.0) * 40.0
        xy.view(-1)[::7] = float("inf")
        xy.view(-1)[3::11] = float("-inf")
        xy.view(-1)[5::13] = float("nan")
        xy.view(-1)[1::17] = 3.0e9
    wxy = ar.inp(xy)
    go = ar.inp(torch.randn(B, H * W, C, generator=g))
    out, gd, gw = ar.out(B, H * W, C), ar.out(B, H, W, C), ar.out(B, H, W, 2)
    wp, gp = wxy.data_ptr(), gw.data_ptr()
    assert lib.arf_resampler_fwd(data.data_ptr(), wp, wp + 4, 2, out.data_ptr(), B, H, W, C, H * W, _cs()) == 0
    assert lib.arf_resampler_bwd(data.data_ptr(), wp, wp + 4, 2, go.data_ptr(), gd.data_ptr(), gp, gp + 4, 2, B, H, W, C,
                                 H * W, _cs()) == 0
    if wild:      # NaN / inf coordinates give NaN interpolation weights: only the guard bands are checked
        torch.cuda.synchronize()
        for raw, n in ar.bufs:
            assert bool((raw[:GUARD] == SENT).all()) and bool((raw[GUARD + n + ((-n) % 4):] == SENT).all())
    else:
        ar.check("resampler")

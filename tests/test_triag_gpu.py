"""Parity of the stencil-triangular kernels with the golden fixtures / oracle (B200)."""
import pytest
import torch

from conftest import RTOL_GRAD, RTOL_VALUE, assert_close, load_golden

pytestmark = pytest.mark.gpu


def _grads(fn, inputs, wrt):
    ins = [i.cuda().clone().requires_grad_(k in wrt) for k, i in enumerate(inputs)]
    out = fn(*ins)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1234)).cuda()
    return out, torch.autograd.grad((out * w).sum(), [ins[k] for k in wrt])


@pytest.mark.parametrize("k", [1, 3])
def test_stencil_products_golden(k):
    from arflow_b200 import triag_solve as ts
    g = load_golden("triag")
    A, X = g["A%d" % k], g["X%d" % k]
    for key, fn in (("mv", ts.matrix_vector_product_general), ("mvT", ts.matrix_vector_product_T_general)):
        out, (gA, gX) = _grads(lambda a, x: fn(a, x, k=k), [A, X], (0, 1))
        assert_close(out, g["%s%d_out0_f64" % (key, k)], RTOL_VALUE, key)
        assert_close(gA, g["%s%d_grad0_f64" % (key, k)], RTOL_GRAD, key + " dA")
        assert_close(gX, g["%s%d_grad1_f64" % (key, k)], RTOL_GRAD, key + " dX")


def test_substitution_and_inverse_diagonal_golden():
    from arflow_b200 import triag_solve as ts
    g = load_golden("triag")
    a, b, c, d, x = (g["in%d" % i].cuda() for i in range(5))
    assert_close(ts.forward_substitution(a, b, c, d, x), g["fsub_out0_f64"], RTOL_VALUE)
    assert_close(ts.backward_substitution(a, b, c, d, x), g["bsub_out0_f64"], RTOL_VALUE)
    assert_close(ts.matrix_vector_product(a, b, c, d, x), g["mv4_out0_f64"], RTOL_VALUE)
    assert_close(ts.matrix_vector_product_T(a, b, c, d, x), g["mv4T_out0_f64"], RTOL_VALUE)
    a5, b5, c5 = a[:, :, :4, :5].contiguous(), b[:, :, :4, :4].contiguous(), c[:, :, :3, :5].contiguous()
    assert_close(ts.inverse_diagonal(a5, b5, c5), g["invdiag_out0_f64"], RTOL_VALUE)
    with pytest.raises(RuntimeError, match="contiguous"):
        ts.forward_substitution(a, b, c, d, x.new_empty(2, 2, 7, 6).transpose(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        ts.forward_substitution(a.cpu(), b, c, d, x)


def test_substitution_autograd_vs_oracle(oracle):
    """ForwardSubst / BackwardSubst: dX is the transposed solve, dA..dD the outer products (triag_solve.py:163-202)."""
    from arflow_b200 import triag_solve as ts
    gen = torch.Generator().manual_seed(3)
    K, L, M, N = 2, 2, 37, 53
    A = 1.0 + torch.rand(K, L, M, N, generator=gen)
    B = 0.4 * torch.randn(K, L, M, N - 1, generator=gen)
    C = 0.4 * torch.randn(K, L, M - 1, N, generator=gen)
    D = 0.2 * torch.randn(K, L, M - 1, N - 1, generator=gen)
    X = torch.randn(K, L, M, N, generator=gen)
    w = torch.randn(K, L, M, N, generator=gen)
    for cls, upper in ((ts.ForwardSubst, False), (ts.BackwardSubst, True)):
        ins = [t.cuda().requires_grad_(True) for t in (A, B, C, D, X)]
        Y = cls.apply(*ins)
        (Y * w.cuda()).sum().backward()
        ref = oracle.substitution(A, B, C, D, X, upper=upper)
        assert_close(Y, ref, RTOL_VALUE, "solve")
        dX = oracle.substitution(A, B, C, D, w, upper=not upper)
        assert_close(ins[4].grad, dX, RTOL_GRAD, "dX")
        assert_close(ins[0].grad, -dX * ref, RTOL_GRAD, "dA")


def test_solve_inverts_product_at_config3_size():
    """Size-independent property at the config-3 level-2 shape (32 x 2 systems of 112 x 256):
    forward_substitution(L, L x) == x and backward_substitution(L^T, L^T x) == x."""
    from arflow_b200 import triag_solve as ts
    gen = torch.Generator().manual_seed(9)
    K, L, M, N = 32, 2, 112, 256
    A = (1.0 + torch.rand(K, L, M, N, generator=gen)).cuda()
    B = (0.3 * torch.randn(K, L, M, N - 1, generator=gen)).cuda()
    C = (0.3 * torch.randn(K, L, M - 1, N, generator=gen)).cuda()
    D = (0.2 * torch.randn(K, L, M - 1, N - 1, generator=gen)).cuda()
    X = torch.randn(K, L, M, N, generator=gen).cuda()
    assert_close(ts.forward_substitution(A, B, C, D, ts.matrix_vector_product(A, B, C, D, X)), X, 1e-4)
    assert_close(ts.backward_substitution(A, B, C, D, ts.matrix_vector_product_T(A, B, C, D, X)), X, 1e-4)
    # the general-k product agrees with the four-array form for k = 1
    taps = torch.zeros(K, 8, M, N, device="cuda")
    # one system per (batch, channel): rearrange (K,L,...) -> batch K, channels (u,v) = L
    taps[:, 0:2] = A
    taps[:, 2:4, :, :-1] = B
    taps[:, 4:6, :-1, :] = C
    taps[:, 6:8, :-1, :-1] = D
    assert_close(ts.matrix_vector_product_general(taps, X, k=1), ts.matrix_vector_product(A, B, C, D, X), RTOL_VALUE)


@pytest.mark.parametrize("M,N", [(1, 1), (1, 40), (9, 1), (5, 31), (7, 33), (20, 300), (3, 1024), (70, 1000), (12, 1030)])
@pytest.mark.parametrize("with_d", [True, False])
def test_row_scan_and_wavefront_solves_agree_with_oracle(oracle, M, N, with_d):
    """arf_trisolve picks the row-scan kernel for N <= 1024 and the anti-diagonal wavefront above that
    (arf_debug_set key 4 = 1 forces the wavefront).  Both follow triag_solve.py:76-115; ragged widths
    cover the warp seams, single rows / columns the degenerate recurrences, 70 x 1000 the reuse of the carry slots
    (every 64 rows) with 32 warps in the pipeline."""
    from arflow_b200 import _lib, triag_solve as ts
    gen = torch.Generator().manual_seed(M * 1000 + N)
    K, L = 2, 2
    A = 1.0 + torch.rand(K, L, M, N, generator=gen)
    B = 0.4 * torch.randn(K, L, M, max(N - 1, 0), generator=gen)
    C = 0.4 * torch.randn(K, L, max(M - 1, 0), N, generator=gen)
    D = 0.2 * torch.randn(K, L, max(M - 1, 0), max(N - 1, 0), generator=gen) if with_d else None
    X = torch.randn(K, L, M, N, generator=gen)
    a, b, c, x = (t.cuda() for t in (A, B, C, X))
    d = D.cuda() if with_d else None

    def solve(upper):
        if with_d:
            return (ts.backward_substitution if upper else ts.forward_substitution)(a, b, c, d, x)
        y = torch.empty_like(x)                      # the C-ABI takes D = NULL (three-array systems)
        _lib.call("arf_trisolve", a.data_ptr(), b.data_ptr(), c.data_ptr(), None, x.data_ptr(), y.data_ptr(),
                  K * L, M, N, int(upper), _lib.stream_ptr())
        return y

    for upper in (False, True):
        ref = oracle.substitution(A, B, C, D, X, upper=upper)
        outs = []
        for variant in (0, 1):
            _lib.load().arf_debug_set(4, variant)
            try:
                outs.append(solve(upper))
            finally:
                _lib.load().arf_debug_set(4, 0)
            assert_close(outs[-1], ref, RTOL_VALUE, "upper=%s variant=%d" % (upper, variant))
        assert_close(outs[0], outs[1], RTOL_VALUE, "scan vs wavefront")

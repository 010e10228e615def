"""Drop-in for utils/triag_solve.py of deu439/ARFlow: stencil-triangular products and solves."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


class _StencilMV(Function):
    @staticmethod
    def forward(ctx, A, X, k, transposed):
        A, X = A.contiguous(), X.contiguous()
        N, CA, H, W = A.shape
        if X.shape != (N, 2, H, W) or CA != 2 * (k + 1) ** 2:
            raise ValueError("stencil product: expected A (N,2(k+1)^2,H,W) and X (N,2,H,W)")
        with torch.cuda.device_of(A):
            Y = torch.empty_like(X)
            _lib.call("arf_stencil_mv_fwd", _lib.dev_ptr(A, "A"), _lib.dev_ptr(X, "X"), _lib.dev_ptr(Y), N, H, W, k,
                      int(transposed), _lib.stream_ptr())
        ctx.save_for_backward(A, X)
        ctx.cfg = (k, int(transposed))
        return Y

    @staticmethod
    def backward(ctx, gY):
        A, X = ctx.saved_tensors
        k, transposed = ctx.cfg
        N, _, H, W = A.shape
        gY = gY.contiguous()
        with torch.cuda.device_of(A):
            dA = torch.empty_like(A) if ctx.needs_input_grad[0] else None
            dX = torch.empty_like(X) if ctx.needs_input_grad[1] else None
            _lib.call("arf_stencil_mv_bwd", _lib.dev_ptr(A), _lib.dev_ptr(X), _lib.dev_ptr(gY, "grad"),
                      _lib.dev_ptr(dA, allow_none=True), _lib.dev_ptr(dX, allow_none=True), N, H, W, k, transposed,
                      _lib.stream_ptr())
        return dA, dX, None, None


def matrix_vector_product_general(A, X, k=1):
    """triag_solve.py:29-43 — y = L x for the (k+1)^2-tap lower-triangular stencil matrix."""
    return _StencilMV.apply(A, X, k, False)


def matrix_vector_product_T_general(A, X, k=1):
    """triag_solve.py:59-73 — y = L^T x."""
    return _StencilMV.apply(A, X, k, True)


def matrix_vector_product(A, B, C, D, X):
    """triag_solve.py:20-28 (|D|C| / |B|A| stencil), channels handled independently."""
    B_Y = torch.nn.functional.pad(B * X[:, :, :, 0:-1], (1, 0))
    C_Y = torch.nn.functional.pad(C * X[:, :, 0:-1, :], (0, 0, 1, 0))
    D_Y = torch.nn.functional.pad(D * X[:, :, 0:-1, 0:-1], (1, 0, 1, 0))
    return A * X + B_Y + C_Y + D_Y


def matrix_vector_product_T(A, B, C, D, X):
    """triag_solve.py:51-56."""
    B_Y = torch.nn.functional.pad(B * X[:, :, :, 1:], (0, 1))
    C_Y = torch.nn.functional.pad(C * X[:, :, 1:, :], (0, 0, 0, 1))
    D_Y = torch.nn.functional.pad(D * X[:, :, 1:, 1:], (0, 1, 0, 1))
    return A * X + B_Y + C_Y + D_Y


def _solve(A, B, C, D, X, upper):
    for name, t in (("A", A), ("B", B), ("C", C), ("D", D), ("X", X)):
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor" % name)       # triag_solve.cpp:8
        if not t.is_contiguous():
            raise RuntimeError("%s must be contiguous" % name)          # triag_solve.cpp:9
    K, L, M, N = A.shape
    if B.shape != (K, L, M, N - 1) or C.shape != (K, L, M - 1, N) or D.shape != (K, L, M - 1, N - 1) or X.shape != A.shape:
        raise ValueError("substitution: inconsistent A/B/C/D/X shapes")
    with torch.cuda.device_of(A):
        Y = torch.empty_like(X)
        _lib.call("arf_trisolve", _lib.dev_ptr(A, "A"), _lib.dev_ptr(B, "B"), _lib.dev_ptr(C, "C"), _lib.dev_ptr(D, "D"),
                  _lib.dev_ptr(X, "X"), _lib.dev_ptr(Y), K * L, M, N, int(upper), _lib.stream_ptr())
    return Y


def forward_substitution(A, B, C, D, X):
    """Solves L y = x (triag_solve.py:76-94; triag_solve_cuda.forward_substitution)."""
    return _solve(A, B, C, D, X, False)


def backward_substitution(A, B, C, D, X):
    """Solves L^T y = x (triag_solve.py:97-115; triag_solve_cuda.backward_substitution)."""
    return _solve(A, B, C, D, X, True)


class ForwardSubst(Function):
    """triag_solve.py:163-181."""

    @staticmethod
    def forward(ctx, A, B, C, D, X):
        Y = forward_substitution(A, B, C, D, X)
        ctx.save_for_backward(A, B, C, D, Y)
        return Y

    @staticmethod
    @once_differentiable
    def backward(ctx, dY):
        A, B, C, D, Y = ctx.saved_tensors
        dX = backward_substitution(A, B, C, D, dY.contiguous())
        dA = -dX * Y
        dB = -dX[:, :, :, 1:] * Y[:, :, :, :-1]
        dC = -dX[:, :, 1:, :] * Y[:, :, :-1, :]
        dD = -dX[:, :, 1:, 1:] * Y[:, :, :-1, :-1]
        return dA, dB, dC, dD, dX


class BackwardSubst(Function):
    """triag_solve.py:184-202."""

    @staticmethod
    def forward(ctx, A, B, C, D, X):
        Y = backward_substitution(A, B, C, D, X)
        ctx.save_for_backward(A, B, C, D, Y)
        return Y

    @staticmethod
    @once_differentiable
    def backward(ctx, dY):
        A, B, C, D, Y = ctx.saved_tensors
        dX = forward_substitution(A, B, C, D, dY.contiguous())
        dA = -dX * Y
        dB = -dX[:, :, :, :-1] * Y[:, :, :, 1:]
        dC = -dX[:, :, :-1, :] * Y[:, :, 1:, :]
        dD = -dX[:, :, :-1, :-1] * Y[:, :, 1:, 1:]
        return dA, dB, dC, dD, dX


def inverse_diagonal(A, B, C):
    """triag_solve_cuda.inverse_diagonal (triag_solve.cpp:38-45): diag((L L^T)^-1), L from A, B, C."""
    for name, t in (("A", A), ("B", B), ("C", C)):
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor" % name)
        if not t.is_contiguous():
            raise RuntimeError("%s must be contiguous" % name)
    K, L, M, N = A.shape
    with torch.cuda.device_of(A):
        H = torch.empty_like(A)
        _lib.call("arf_inv_diag", _lib.dev_ptr(A, "A"), _lib.dev_ptr(B, "B"), _lib.dev_ptr(C, "C"), _lib.dev_ptr(H),
                  K * L, M, N, _lib.stream_ptr())
    return H


def marginal_variances(A, B, C):
    """triag_solve.py:205-232 — same quantity as inverse_diagonal."""
    return inverse_diagonal(A, B, C)

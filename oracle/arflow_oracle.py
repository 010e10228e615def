"""ORACLE — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under arflow_b200/ does.  Every function restates one reference function
(deu439/ARFlow, cited as file:line) with explicit index arithmetic on CPU tensors, generic in dtype:
run it in float64 for ground truth, in float32 for "reference precision".

Parity pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, produced by importing /root/reference in
the build container (tests/golden/make_golden.py, fixtures committed under tests/golden/*.npz) and
checked by tests/test_oracle_golden.py.
"""
import ctypes
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_clib = None


def _c():
    global _clib
    if _clib is None:
        if not os.path.exists(_CLIB_PATH):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        _clib = ctypes.CDLL(_CLIB_PATH)
    return _clib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------- correlation ----
def corr_dims(H, W, pad, ks, md, s1, s2):
    """correlation_cuda.cc:25-34."""
    d2, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = _c().oracle_corr_dims(H, W, pad, ks, md, s1, s2, ctypes.byref(d2), ctypes.byref(oh), ctypes.byref(ow))
    if rc:
        raise ValueError("bad correlation geometry")
    return d2.value, oh.value, ow.value


def corr_fwd_c(f1, f2, pad=4, ks=1, md=4, s1=1, s2=1):
    """Literal C restatement of correlation_cuda_kernel.cu:41-114, double accumulation."""
    a1 = np.ascontiguousarray(f1.detach().double().numpy())
    a2 = np.ascontiguousarray(f2.detach().double().numpy())
    B, C, H, W = a1.shape
    d2, oh, ow = corr_dims(H, W, pad, ks, md, s1, s2)
    out = np.zeros((B, d2, oh, ow), dtype=np.float64)
    rc = _c().oracle_corr_fwd(_dp(a1), _dp(a2), _dp(out), B, C, H, W, pad, ks, md, s1, s2)
    assert rc == 0
    return torch.from_numpy(out)


def corr_bwd_c(f1, f2, gout, pad=4, ks=1, md=4, s1=1, s2=1):
    """Literal C restatement of correlation_cuda_kernel.cu:116-300 (both input gradients)."""
    a1 = np.ascontiguousarray(f1.detach().double().numpy())
    a2 = np.ascontiguousarray(f2.detach().double().numpy())
    go = np.ascontiguousarray(gout.detach().double().numpy())
    B, C, H, W = a1.shape
    g1 = np.zeros_like(a1)
    g2 = np.zeros_like(a2)
    rc = _c().oracle_corr_bwd(_dp(a1), _dp(a2), _dp(go), _dp(g1), _dp(g2), B, C, H, W, pad, ks, md, s1, s2)
    assert rc == 0
    return torch.from_numpy(g1), torch.from_numpy(g2)


def cost_volume(f1, f2, md=4):
    """correlation_native.py:13-23 == uflow_model.py:53-92 (pad=md, ks=1, strides 1), vectorised:
    one unfold of the zero-padded f2 instead of 81 slice-multiply-mean passes.  Differentiable, so
    it also is the CPU baseline for forward+backward."""
    B, C, H, W = f1.shape
    D = 2 * md + 1
    f2p = F.pad(f2, (md, md, md, md))
    win = f2p.unfold(2, H, 1).unfold(3, W, 1)          # B, C, D, D, H, W   (dy, dx leading)
    return (f1[:, :, None, None] * win).mean(dim=1).reshape(B, D * D, H, W)


# --------------------------------------------------------------------------- warp -----------
def _unnormalize(g, size, align):
    # ATen grid_sampler_unnormalize
    if align:
        return (g + 1) / 2 * (size - 1)
    return ((g + 1) * size - 1) / 2


def _reflect(c, twice_low, twice_high):
    if twice_low == twice_high:
        return torch.zeros_like(c)
    mn = twice_low / 2.0
    span = (twice_high - twice_low) / 2.0
    c = (c - mn).abs()
    extra = torch.fmod(c, span)
    flips = torch.floor(c / span)
    return torch.where(flips % 2 == 0, extra + mn, span - extra + mn)


def _source_index(p, n1, size, pad, align):
    g = 2.0 * p / n1 - 1.0
    c = _unnormalize(g, size, align)
    if pad == "border":
        c = c.clamp(0, size - 1)
    elif pad == "reflection":
        c = _reflect(c, 0, 2 * (size - 1)) if align else _reflect(c, -1, 2 * size - 1)
        c = c.clamp(0, size - 1)
    return c


def _gather2d(x, yi, xi):
    """x: (B,C,H,W); yi,xi: (B,h,w) long -> (B,C,h,w), zeros where the index is outside."""
    B, C, H, W = x.shape
    ok = (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
    lin = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).view(B, 1, -1).expand(B, C, -1)
    v = x.reshape(B, C, H * W).gather(2, lin).view(B, C, *yi.shape[1:])
    return v * ok.unsqueeze(1).to(x.dtype)


def warp(x, field, kind="flow", pad="zeros", mode="bilinear", align_corners=True, nW1=None, nH1=None):
    """flow_warp (warp_utils.py:83-90; kind='flow', divisors W-1, H-1 of the flow) and
    resample (uflow_utils.py:53-77; kind='coords', divisors max(W-1,1), max(H-1,1) of the source),
    with grid_sample written out as explicit gathers (differentiable through torch indexing)."""
    B, C, Hs, Ws = x.shape
    _, _, Ho, Wo = field.shape
    if nW1 is None:
        nW1, nH1 = ((Wo - 1, Ho - 1) if kind == "flow" else (max(Ws - 1, 1), max(Hs - 1, 1)))
    px, py = field[:, 0], field[:, 1]
    if kind == "flow":
        jj = torch.arange(Wo, dtype=field.dtype).view(1, 1, Wo)
        ii = torch.arange(Ho, dtype=field.dtype).view(1, Ho, 1)
        px, py = jj + px, ii + py
    X = _source_index(px, nW1, Ws, pad, align_corners)
    Y = _source_index(py, nH1, Hs, pad, align_corners)
    if mode == "nearest":
        return _gather2d(x, torch.round(Y).long(), torch.round(X).long())
    x0, y0 = torch.floor(X), torch.floor(Y)
    x1, y1 = x0 + 1, y0 + 1
    w_nw = ((x1 - X) * (y1 - Y)).unsqueeze(1)
    w_ne = ((X - x0) * (y1 - Y)).unsqueeze(1)
    w_sw = ((x1 - X) * (Y - y0)).unsqueeze(1)
    w_se = ((X - x0) * (Y - y0)).unsqueeze(1)
    x0l, x1l, y0l, y1l = x0.long(), x1.long(), y0.long(), y1.long()
    return (_gather2d(x, y0l, x0l) * w_nw + _gather2d(x, y0l, x1l) * w_ne +
            _gather2d(x, y1l, x0l) * w_sw + _gather2d(x, y1l, x1l) * w_se)


def flow_to_warp(flow):
    """uflow_utils.py:6-32."""
    B, _, H, W = flow.shape
    jj = torch.arange(W, dtype=flow.dtype).view(1, 1, 1, W).expand(B, 1, H, W)
    ii = torch.arange(H, dtype=flow.dtype).view(1, 1, H, 1).expand(B, 1, H, W)
    return torch.cat([jj, ii], 1) + flow

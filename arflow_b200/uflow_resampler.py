"""Drop-in for utils/uflow_resampler.py of deu439/ARFlow (the TF `resampler` port, NHWC)."""
import torch

from . import _lib


class _ResamplerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, warp_x, warp_y):
        data = data.contiguous()
        warp_x, warp_y = warp_x.contiguous(), warp_y.contiguous()
        B, H, W, C = data.shape
        P = warp_x.numel() // B
        with torch.cuda.device_of(data):
            out = torch.empty(tuple(warp_x.shape) + (C,), dtype=data.dtype, device=data.device)
            _lib.call("arf_resampler_fwd", _lib.dev_ptr(data, "data"), _lib.dev_ptr(warp_x, "warp_x"),
                      _lib.dev_ptr(warp_y, "warp_y"), 1, _lib.dev_ptr(out), B, H, W, C, P, _lib.stream_ptr())
        ctx.save_for_backward(data, warp_x, warp_y)
        return out

    @staticmethod
    def backward(ctx, gout):
        data, warp_x, warp_y = ctx.saved_tensors
        B, H, W, C = data.shape
        P = warp_x.numel() // B
        gout = gout.contiguous()
        with torch.cuda.device_of(data):
            gd = torch.empty_like(data) if ctx.needs_input_grad[0] else None
            gx = torch.empty_like(warp_x) if ctx.needs_input_grad[1] else None
            gy = torch.empty_like(warp_y) if ctx.needs_input_grad[2] else None
            _lib.call("arf_resampler_bwd", _lib.dev_ptr(data), _lib.dev_ptr(warp_x), _lib.dev_ptr(warp_y), 1,
                      _lib.dev_ptr(gout, "grad"), _lib.dev_ptr(gd, allow_none=True), _lib.dev_ptr(gx, allow_none=True),
                      _lib.dev_ptr(gy, allow_none=True), 1, B, H, W, C, P, _lib.stream_ptr())
        return gd, gx, gy


def resampler_with_unstacked_warp(data, warp_x, warp_y, safe=True):
    """uflow_resampler.py:155-241.  data (B,H,W,C); warp_x, warp_y (B, ...) -> (B, ..., C).
    Taps outside the image contribute zero (the reference's safe=True; with safe=False it raises on such
    coordinates, here they are handled the same safe way)."""
    assert warp_x.size() == warp_y.size(), "warp_x and warp_y incompatible!"
    assert warp_x.shape[0] == data.shape[0], "warp_x and data have incompatible first dimension (batch size)"
    return _ResamplerFunction.apply(data, warp_x, warp_y)


def resampler(data, warp):
    """uflow_resampler.py:137-152 — warp (..., 2) holds (x, y)."""
    warp_x, warp_y = torch.unbind(warp, dim=-1)
    return resampler_with_unstacked_warp(data, warp_x, warp_y)

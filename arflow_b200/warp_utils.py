"""Drop-in for utils/warp_utils.py of deu439/ARFlow (flow_warp and the mask helpers)."""
import torch

from . import _lib

_PAD = {"zeros": 0, "border": 1, "reflection": 2}
_INTERP = {"bilinear": 0, "nearest": 1}
FIELD_FLOW, FIELD_COORDS = 0, 1


class _WarpFunction(torch.autograd.Function):
    """y = grid_sample(x, normalise(field)) with the reference's fp32 coordinate round trip."""

    @staticmethod
    def forward(ctx, x, field, nW1, nH1, field_kind, interp, pad_mode, align):
        if x.dim() != 4 or field.dim() != 4 or field.shape[1] != 2 or x.shape[0] != field.shape[0]:
            raise ValueError("warp: expected x (B,C,H,W) and a (B,2,h,w) field")
        x = x.contiguous()
        field = field.contiguous()
        B, C, Hs, Ws = x.shape
        Ho, Wo = field.shape[2:]
        args = (B, C, Hs, Ws, Ho, Wo, float(nW1), float(nH1), field_kind, interp, pad_mode, int(align))
        with torch.cuda.device_of(x):
            y = torch.empty((B, C, Ho, Wo), dtype=x.dtype, device=x.device)
            _lib.call("arf_warp_fwd", _lib.dev_ptr(x, "x"), _lib.dev_ptr(field, "flow"), _lib.dev_ptr(y),
                      *args, _lib.stream_ptr())
        ctx.save_for_backward(x, field)
        ctx.args = args
        return y

    @staticmethod
    def backward(ctx, gy):
        x, field = ctx.saved_tensors
        gy = gy.contiguous()
        need_x, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        with torch.cuda.device_of(x):
            gx = torch.empty_like(x) if need_x else None
            gf = torch.empty_like(field) if need_f else None
            _lib.call("arf_warp_bwd", _lib.dev_ptr(x), _lib.dev_ptr(field), _lib.dev_ptr(gy, "grad"),
                      _lib.dev_ptr(gx, allow_none=True), _lib.dev_ptr(gf, allow_none=True),
                      *ctx.args, _lib.stream_ptr())
        return gx, gf, None, None, None, None, None, None


def _modes(pad, mode):
    if pad not in _PAD:
        raise ValueError("nn.functional.grid_sample(): expected padding_mode to be 'zeros', 'border', "
                         "or 'reflection', but got: '%s'" % pad)
    if mode not in _INTERP:
        if mode == "bicubic":
            raise NotImplementedError("arflow_b200 warp: mode='bicubic' is not implemented")
        raise ValueError("nn.functional.grid_sample(): expected mode to be 'bilinear', 'nearest' or "
                         "'bicubic', but got: '%s'" % mode)
    return _PAD[pad], _INTERP[mode]


def flow_warp(x, flow, pad='zeros', mode='bilinear', align_corners=True):
    """utils/warp_utils.py:83-90 — sample x at (j + u, i + v).

    The grid is normalised with the flow's own (W-1, H-1) (norm_grid, :16-23) and un-normalised
    with the source size by grid_sample; W == 1 or H == 1 divides by zero exactly as there.
    """
    pad_mode, interp = _modes(pad, mode)
    _, _, H, W = flow.shape
    return _WarpFunction.apply(x, flow, W - 1, H - 1, FIELD_FLOW, interp, pad_mode, bool(align_corners))

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


# --------------------------------------------------------------------------- grids, masks --
def mesh_grid(B, H, W):
    """warp_utils.py:7-13 (integer pixel grid, B x 2 x H x W, CPU like the reference)."""
    x_base = torch.arange(0, W).repeat(B, H, 1)
    y_base = torch.arange(0, H).repeat(B, W, 1).transpose(1, 2)
    return torch.stack([x_base, y_base], 1)


def norm_grid(v_grid):
    """warp_utils.py:16-23."""
    _, _, H, W = v_grid.size()
    v_grid_norm = torch.zeros_like(v_grid)
    v_grid_norm[:, 0, :, :] = 2.0 * v_grid[:, 0, :, :] / (W - 1) - 1.0
    v_grid_norm[:, 1, :, :] = 2.0 * v_grid[:, 1, :, :] / (H - 1) - 1.0
    return v_grid_norm.permute(0, 2, 3, 1)


def _splat(field, kind):
    # differentiable in the field like the reference's scatter_add of bilinear weights (warp_utils.py:60-78, 222-238)
    from .uflow_utils import _RangeMapFunction
    return _RangeMapFunction.apply(field, kind)


def get_corresponding_map(data):
    """warp_utils.py:26-80 — forward splat of unnormalised coordinates (B,2,H,W) -> (B,1,H,W)."""
    return _splat(data, FIELD_COORDS)


def compute_range_map(flow):
    """warp_utils.py:158-239 (identical to uflow_utils.compute_range_map)."""
    assert flow.dim() == 4
    return _splat(flow, FIELD_FLOW)


def _count_to_mask(count, mode, th=0.0):
    with torch.cuda.device_of(count):
        out = torch.empty_like(count)
        _lib.call("arf_count_to_mask", _lib.dev_ptr(count), _lib.dev_ptr(out), count.numel(), mode, float(th),
                  _lib.stream_ptr())
    return out


def get_occu_mask_backward(flow21, th=0.2):
    """warp_utils.py:103-116 — 1 (or close to 1) at occluded pixels."""
    corr_map = _splat(flow21, FIELD_FLOW)  # get_corresponding_map(base_grid + flow21)
    return _count_to_mask(corr_map, 1, th) if th > 0 else _count_to_mask(corr_map, 2)


def get_occu_mask_bidirection(flow12, flow21, scale=0.01, bias=0.5):
    """warp_utils.py:93-100 — forward-backward consistency check."""
    flow21_warped = flow_warp(flow21, flow12, pad='zeros').detach()
    flow12 = flow12.detach().contiguous()
    B, _, H, W = flow12.shape
    with torch.cuda.device_of(flow12):
        occ = torch.empty((B, 1, H, W), dtype=flow12.dtype, device=flow12.device)
        _lib.call("arf_occ_bidir", _lib.dev_ptr(flow12, "flow12"), _lib.dev_ptr(flow21_warped), _lib.dev_ptr(occ),
                  B, H, W, float(scale), float(bias), _lib.stream_ptr())
    return occ


def border_mask(flow):
    """warp_utils.py:119-134 — 1 where the correspondence lies strictly inside the image."""
    flow = flow.detach().contiguous()
    B, _, H, W = flow.shape
    with torch.cuda.device_of(flow):
        mask = torch.empty((B, 1, H, W), dtype=flow.dtype, device=flow.device)
        _lib.call("arf_inside_mask", _lib.dev_ptr(flow, "flow"), _lib.dev_ptr(mask), B, H, W, FIELD_FLOW, 1,
                  _lib.stream_ptr())
    return mask


def flow_to_warp(flow):
    """warp_utils.py:138-155 — the NHWC, (y, x)-ordered variant kept by this module."""
    flow = flow.permute(0, 2, 3, 1).flip(-1)
    _, height, width, _ = flow.size()
    i_grid, j_grid = torch.meshgrid(torch.arange(height, device=flow.device, dtype=flow.dtype),
                                    torch.arange(width, device=flow.device, dtype=flow.dtype), indexing='ij')
    return torch.stack([i_grid, j_grid], dim=2) + flow

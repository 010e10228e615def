"""Drop-in for utils/uflow_utils.py of deu439/ARFlow."""
import torch

from . import _lib
from .warp_utils import FIELD_COORDS, _WarpFunction


def flow_to_warp(flow):
    """uflow_utils.py:6-32 — absolute sampling coordinates (x, y) = pixel grid + flow.
    The grid is built on the flow's device (the reference builds it on the CPU and copies it
    over on every call, :19-22)."""
    B, _, H, W = flow.shape
    jj = torch.arange(W, device=flow.device, dtype=flow.dtype).view(1, 1, 1, W)
    ii = torch.arange(H, device=flow.device, dtype=flow.dtype).view(1, 1, H, 1)
    return torch.cat([flow[:, 0:1] + jj, flow[:, 1:2] + ii], dim=1)


def resample(source, coords):
    """uflow_utils.py:53-77 — bilinear sample of `source` at absolute `coords` (x, y), zeros outside,
    align_corners=True; coordinates go through the same 2*c/max(W-1,1)-1 normalisation round trip."""
    _, _, H, W = source.shape
    return _WarpFunction.apply(source, coords, max(W - 1, 1), max(H - 1, 1), FIELD_COORDS, 0, 0, True)

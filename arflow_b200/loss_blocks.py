"""Drop-in for losses/loss_blocks.py of deu439/ARFlow (the ARFlow-era loss blocks)."""
import torch
import torch.nn.functional as F

from . import _lib
from .uflow_utils import _CensusHammingFunction


def penalty_ddflow(diff, eps=0.01, q=0.4):
    """loss_blocks.py:5-6."""
    return torch.pow((torch.abs(diff) + eps), q)


def penalty_uflow(x):
    """loss_blocks.py:8-9."""
    return torch.sqrt(torch.pow(x, 2.0) + 0.001 ** 2)


def TernaryLoss(im, im_warp, max_distance=1, sum_dist=False):
    """loss_blocks.py:12-62 -> (dist, mask): soft Hamming distance of the ternary/census transforms
    (mean over the patch for ARFlow, sum for UFlow) and the mask with a max_distance border zeroed."""
    patch_size = 2 * max_distance + 1
    scale = 1.0 if sum_dist else 1.0 / (patch_size * patch_size)
    dist = _CensusHammingFunction.apply(im, im_warp, patch_size, scale)
    n, _, h, w = im.size()
    inner = torch.ones(n, 1, h - 2 * max_distance, w - 2 * max_distance, dtype=im.dtype, device=im.device)
    mask = F.pad(inner, [max_distance] * 4)
    return dist, mask


def gradient(data):
    """loss_blocks.py:87-90."""
    D_dy = data[:, :, 1:] - data[:, :, :-1]
    D_dx = data[:, :, :, 1:] - data[:, :, :, :-1]
    return D_dx, D_dy


class _SmoothFunction(torch.autograd.Function):
    """final * (mean_x(w_x * pen(d_x)) + mean_y(w_y * pen(d_y))) — csrc/smooth.cu."""

    @staticmethod
    def forward(ctx, flow, image, order, wstride, woff, penalty, edge, eps2, final_scale):
        if image.requires_grad:
            raise NotImplementedError("arflow_b200 smoothness: the image is data; no gradient w.r.t. it is provided")
        flow, image = flow.contiguous(), image.contiguous()
        B, Cf, H, W = flow.shape
        if Cf != 2 or image.shape[0] != B or image.shape[2:] != flow.shape[2:]:
            raise ValueError("smoothness: expected flow (B,2,H,W) and image (B,C,H,W)")
        Ci = image.shape[1]
        args = (B, Ci, H, W, order, wstride, woff, penalty, float(edge), float(eps2), float(final_scale))
        lib = _lib.load()
        with torch.cuda.device_of(flow):
            out = torch.empty((1,), dtype=flow.dtype, device=flow.device)
            partials = torch.empty((2 * lib.arf_smooth_num_partials(B, H, W),), dtype=flow.dtype, device=flow.device)
            _lib.call("arf_smooth_fwd", _lib.dev_ptr(image, "image"), _lib.dev_ptr(flow, "flow"), _lib.dev_ptr(out),
                      _lib.dev_ptr(partials), *args, _lib.stream_ptr())
        ctx.save_for_backward(flow, image)
        ctx.args = args
        return out[0]

    @staticmethod
    def backward(ctx, gloss):
        flow, image = ctx.saved_tensors
        gloss = gloss.reshape(1).contiguous()
        with torch.cuda.device_of(flow):
            gflow = torch.empty_like(flow)
            _lib.call("arf_smooth_bwd", _lib.dev_ptr(image), _lib.dev_ptr(flow), _lib.dev_ptr(gloss, "grad"),
                      _lib.dev_ptr(gflow), *ctx.args, _lib.stream_ptr())
        return gflow, None, None, None, None, None, None, None, None


def smooth_grad_1st(flo, image, alpha, penalty="abs"):
    """loss_blocks.py:93-109."""
    if penalty == "abs":
        pen = 1
    elif penalty == "uflow":
        pen = 0
    else:
        raise NotImplementedError()
    return _SmoothFunction.apply(flo, image, 1, 1, 0, pen, alpha, 0.001 ** 2, 0.25)


def smooth_grad_2nd(flo, image, alpha):
    """loss_blocks.py:112-124."""
    return _SmoothFunction.apply(flo, image, 2, 1, 1, 1, alpha, 0.0, 0.5)


def SSIM(x, y, md=1):
    """loss_blocks.py:65-84 — valid (unpadded) box filters, output (B,C,H-2md,W-2md)."""
    from .uflow_utils import _SsimFunction
    return _SsimFunction.apply(x, y, 2 * md + 1, True, 1)

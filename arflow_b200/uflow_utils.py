"""Drop-in for utils/uflow_utils.py of deu439/ARFlow."""
import torch

from . import _lib
from .warp_utils import FIELD_COORDS, FIELD_FLOW, _WarpFunction  # noqa: F401


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


def resample_flow(source, flow):
    """resample(source, flow_to_warp(flow)) (uflow_utils.py:53-77 after :6-32) without materialising the coordinate
    grid: the kernel adds the pixel grid to the flow itself — the same single fp32 add, so the result is bit-identical —
    which saves the arange / cat / add kernels and their autograd counterparts on every call."""
    _, _, H, W = source.shape
    return _WarpFunction.apply(source, flow, max(W - 1, 1), max(H - 1, 1), FIELD_FLOW, 0, 0, True)


def mask_invalid_flow(flow):
    """mask_invalid(flow_to_warp(flow)) (uflow_utils.py:35-50) straight from the flow."""
    flow = flow.detach().contiguous()
    B, _, H, W = flow.shape
    with torch.cuda.device_of(flow):
        mask = torch.empty((B, 1, H, W), dtype=flow.dtype, device=flow.device)
        _lib.call("arf_inside_mask", _lib.dev_ptr(flow, "flow"), _lib.dev_ptr(mask), B, H, W, FIELD_FLOW, 0,
                  _lib.stream_ptr())
    return mask


# --------------------------------------------------------------------------- masks ---------
def _new_like(t, shape):
    return torch.empty(shape, dtype=t.dtype, device=t.device)


def mask_invalid(coords):
    """uflow_utils.py:35-50 — 1 where the coordinate lies inside [0,W-1]x[0,H-1]."""
    coords = coords.detach().contiguous()
    B, _, H, W = coords.shape
    with torch.cuda.device_of(coords):
        mask = _new_like(coords, (B, 1, H, W))
        _lib.call("arf_inside_mask", _lib.dev_ptr(coords, "coords"), _lib.dev_ptr(mask), B, H, W, FIELD_COORDS, 0,
                  _lib.stream_ptr())
    return mask


class _RangeMapFunction(torch.autograd.Function):
    """Forward splat count; `kind` 0 = the field is a flow (base grid added), 1 = absolute (x, y) coordinates."""

    @staticmethod
    def forward(ctx, flow, kind=0):
        flow = flow.contiguous()
        B, _, H, W = flow.shape
        with torch.cuda.device_of(flow):
            count = _new_like(flow, (B, 1, H, W))
            _lib.call("arf_range_map", _lib.dev_ptr(flow, "flow"), _lib.dev_ptr(count), B, H, W, kind, _lib.stream_ptr())
        ctx.save_for_backward(flow)
        ctx.kind = kind
        return count

    @staticmethod
    def backward(ctx, gcount):
        (flow,) = ctx.saved_tensors
        B, _, H, W = flow.shape
        gcount = gcount.contiguous()
        with torch.cuda.device_of(flow):
            gflow = torch.empty_like(flow)
            _lib.call("arf_range_map_bwd", _lib.dev_ptr(flow), _lib.dev_ptr(gcount, "grad"), _lib.dev_ptr(gflow),
                      B, H, W, ctx.kind, _lib.stream_ptr())
        return gflow, None


def compute_range_map(flow):
    """uflow_utils.py:80-160 — how often each pixel is hit by the forward splat of `flow` (bilinear weights,
    targets outside the image dropped).  Differentiable w.r.t. the flow like the reference's scatter_add
    (the occlusion penalty of the ELBO loss uses that, uflow_elbo_loss.py:552-559)."""
    assert flow.dim() == 4
    return _RangeMapFunction.apply(flow)


def clamp01(count, mode=0, th=0.0):
    """clamp(count, 0, 1) (mode 0), clamp < th (mode 1), 1 - clamp (mode 2) in one pass."""
    count = count.contiguous()
    with torch.cuda.device_of(count):
        out = torch.empty_like(count)
        _lib.call("arf_count_to_mask", _lib.dev_ptr(count), _lib.dev_ptr(out), count.numel(), mode, float(th),
                  _lib.stream_ptr())
    return out


# --------------------------------------------------------------------------- resize --------
class _ResizeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, Ho, Wo, rh, rw, mul, align=0):
        img = img.contiguous()
        B, C, Hi, Wi = img.shape
        args = (B * C, Hi, Wi, Ho, Wo, float(rh), float(rw), float(mul), int(align))
        with torch.cuda.device_of(img):
            out = _new_like(img, (B, C, Ho, Wo))
            _lib.call("arf_resize_bilinear_fwd", _lib.dev_ptr(img, "img"), _lib.dev_ptr(out), *args, _lib.stream_ptr())
        ctx.args = args
        ctx.in_shape = img.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous()
        with torch.cuda.device_of(gout):
            gin = _new_like(gout, ctx.in_shape)
            _lib.call("arf_resize_bilinear_bwd", _lib.dev_ptr(gout, "grad"), _lib.dev_ptr(gin), *ctx.args,
                      _lib.stream_ptr())
        return gin, None, None, None, None, None, None


def interpolate_align_corners(img, scale_factor, mul=1.0):
    """mul * F.interpolate(img, scale_factor=s, mode='bilinear', align_corners=True): the flow up-sampling of the
    PWC-Lite family (models/pwclite.py:178-179, 203; `mul` folds the `flow * s` in)."""
    import math
    _, _, H, W = img.shape
    Ho, Wo = int(math.floor(H * scale_factor)), int(math.floor(W * scale_factor))
    if Ho < 2 or Wo < 2 or H < 2 or W < 2:
        raise ValueError("interpolate_align_corners: needs at least 2 rows and columns on both sides")
    # ATen area_pixel_compute_scale: float(in - 1) / (out - 1), one fp32 rounding
    import numpy as np
    rh = float(np.float32(H - 1) / np.float32(Ho - 1))
    rw = float(np.float32(W - 1) / np.float32(Wo - 1))
    return _ResizeFunction.apply(img, Ho, Wo, rh, rw, mul, 1)


def _interpolate(img, scale_factor, mul):
    # F.interpolate(scale_factor=s): output size floor(in*s); source step exactly 1/s
    import math
    _, _, H, W = img.shape
    Ho, Wo = int(math.floor(H * scale_factor)), int(math.floor(W * scale_factor))
    return _ResizeFunction.apply(img, Ho, Wo, 1.0 / scale_factor, 1.0 / scale_factor, mul)


def upsample(img, is_flow, scale_factor=2.0):
    """uflow_utils.py:163-182 — bilinear, align_corners=False; flow values scaled with the size."""
    return _interpolate(img, scale_factor, scale_factor if is_flow else 1.0)


def downsample(img, is_flow, scale_factor=2.0):
    """uflow_utils.py:185-204."""
    return _interpolate(img, 1 / scale_factor, 1 / scale_factor if is_flow else 1.0)


# --------------------------------------------------------------------------- small helpers -
def image_grads(image_batch, stride=1):
    """uflow_utils.py:207-210 (slicing only)."""
    image_batch_x = image_batch[:, :, :, stride:] - image_batch[:, :, :, :-stride]
    image_batch_y = image_batch[:, :, stride:] - image_batch[:, :, :-stride]
    return image_batch_x, image_batch_y


def abs_robust_loss(diff, eps=0.01, q=0.4):
    """uflow_utils.py:213-214."""
    return torch.pow((torch.abs(diff) + eps), q)


def robust_l1(x):
    """uflow_utils.py:337-338."""
    return (x + 0.001 ** 2) ** 0.5


def rgb_to_grayscale(image):
    """uflow_utils.py:227-231."""
    grayscale = image[:, 0, :, :] * 0.2989 + image[:, 1, :, :] * 0.5870 + image[:, 2, :, :] * 0.1140
    return grayscale.unsqueeze(1)


def zero_mask_border(mask, patch_size):
    """uflow_utils.py:234-238."""
    p = patch_size // 2
    return torch.nn.functional.pad(mask[:, :, p:-p, p:-p], [p] * 4)


# --------------------------------------------------------------------------- census --------
class _CensusHammingFunction(torch.autograd.Function):
    """soft_hamming(census_transform(a), census_transform(b)) in one pass (uflow_utils.py:241-279)."""

    @staticmethod
    def forward(ctx, im_a, im_b, patch, scale):
        im_a, im_b = im_a.contiguous(), im_b.contiguous()
        if im_a.shape != im_b.shape or im_a.dim() != 4 or im_a.shape[1] != 3:
            raise ValueError("census: expected two (B,3,H,W) images of equal shape")
        B, _, H, W = im_a.shape
        with torch.cuda.device_of(im_a):
            ham = _new_like(im_a, (B, 1, H, W))
            _lib.call("arf_census_fwd", _lib.dev_ptr(im_a, "image_a"), _lib.dev_ptr(im_b, "image_b"), None,
                      _lib.dev_ptr(ham), None, None, B, H, W, patch, float(scale), 0.01, 0.4, _lib.stream_ptr())
        ctx.save_for_backward(im_a, im_b)
        ctx.cfg = (patch, float(scale))
        return ham

    @staticmethod
    def backward(ctx, gham):
        im_a, im_b = ctx.saved_tensors
        patch, scale = ctx.cfg
        B, _, H, W = im_a.shape
        gham = gham.contiguous()
        with torch.cuda.device_of(im_a):
            ga = torch.empty_like(im_a) if ctx.needs_input_grad[0] else None
            gb = torch.empty_like(im_b) if ctx.needs_input_grad[1] else None
            _lib.call("arf_census_bwd", _lib.dev_ptr(im_a), _lib.dev_ptr(im_b), _lib.dev_ptr(gham, "grad"), None, None,
                      None, None, _lib.dev_ptr(ga, allow_none=True), _lib.dev_ptr(gb, allow_none=True),
                      B, H, W, patch, scale, 0.01, 0.4, _lib.stream_ptr())
        return ga, gb, None, None


# Data-parallel training shards the batch, but the reference normalises the masked census mean by the mask sum of
# the WHOLE batch (uflow_utils.py:293).  With a process group registered here, the per-rank denominator is replaced
# by world * (global denominator) so that the gradient average over ranks equals the single-process gradient:
#   mean_r [ W * num_r / (sum_r den_r + 1e-6) ] = sum_r num_r / (sum_r den_r + 1e-6).
# The mask is detached, so this is a forward-only 4-byte all-reduce.  It is a collective inside the step: through NCCL
# it cannot be captured in a CUDA graph on this pool, through the peer-memory kernel (arflow_b200/comm.py) it can
# (train_step.py).  The default (None) keeps the per-rank normaliser.
_census_group = None


def set_census_normaliser_group(group, enabled=True):
    """group: a torch.distributed process group (or dist.group.WORLD); enabled=False / group=None switches back."""
    global _census_group
    _census_group = group if enabled else None


def globalise_census_sums(sums, group):
    """sums = [num, den, num/(den+1e-6)] of this rank -> the same triple with the batch-global normaliser.
    group: a torch.distributed process group, or an arflow_b200.comm.PeerAllReduce of >= 4 floats (capturable)."""
    if hasattr(group, "all_reduce_"):
        world = group.world
        buf = group.buffer
        buf[0:1].copy_(sums[1:2])
        group.all_reduce_(0, 4, average=False)
        den = buf[0:1].clone()
    else:
        import torch.distributed as dist
        world = dist.get_world_size(group)
        den = sums[1:2].clone()
        dist.all_reduce(den, group=group)
    out = sums.clone()
    out[1] = (den[0] + 1e-6) / world - 1e-6          # so that 1 / (out[1] + 1e-6) = world / (global den + 1e-6)
    out[2] = sums[0] / (out[1] + 1e-6)
    return out


class _CensusLossFunction(torch.autograd.Function):
    """census_loss (uflow_utils.py:282-293) fused: transform, soft Hamming, robust penalty, border-zeroed
    mask and the batch-global masked mean, one pass forward and one pass backward."""

    @staticmethod
    def forward(ctx, im_a, im_b, mask, patch, eps, q):
        im_a, im_b = im_a.contiguous(), im_b.contiguous()
        if im_a.shape != im_b.shape or im_a.dim() != 4 or im_a.shape[1] != 3:
            raise ValueError("census_loss: expected two (B,3,H,W) images of equal shape")
        B, _, H, W = im_a.shape
        mask = mask.detach().contiguous()
        if mask.shape != (B, 1, H, W):
            raise ValueError("census_loss: mask must be (B,1,H,W)")
        lib = _lib.load()
        with torch.cuda.device_of(im_a):
            ham = _new_like(im_a, (B, 1, H, W))
            partials = _new_like(im_a, (2 * lib.arf_census_num_partials(B, H, W),))
            sums = _new_like(im_a, (3,))
            _lib.call("arf_census_fwd", _lib.dev_ptr(im_a, "image_a"), _lib.dev_ptr(im_b, "image_b"),
                      _lib.dev_ptr(mask, "mask"), _lib.dev_ptr(ham), _lib.dev_ptr(partials), _lib.dev_ptr(sums),
                      B, H, W, patch, 1.0, float(eps), float(q), _lib.stream_ptr())
        if _census_group is not None:
            sums = globalise_census_sums(sums, _census_group)
        ctx.save_for_backward(im_a, im_b, mask, ham, sums)
        ctx.cfg = (patch, float(eps), float(q))
        return sums[2]

    @staticmethod
    def backward(ctx, gloss):
        im_a, im_b, mask, ham, sums = ctx.saved_tensors
        patch, eps, q = ctx.cfg
        B, _, H, W = im_a.shape
        gloss = gloss.reshape(1).contiguous()
        with torch.cuda.device_of(im_a):
            ga = torch.empty_like(im_a) if ctx.needs_input_grad[0] else None
            gb = torch.empty_like(im_b) if ctx.needs_input_grad[1] else None
            _lib.call("arf_census_bwd", _lib.dev_ptr(im_a), _lib.dev_ptr(im_b), None, _lib.dev_ptr(ham),
                      _lib.dev_ptr(mask), _lib.dev_ptr(sums), _lib.dev_ptr(gloss, "grad"),
                      _lib.dev_ptr(ga, allow_none=True), _lib.dev_ptr(gb, allow_none=True),
                      B, H, W, patch, 1.0, eps, q, _lib.stream_ptr())
        return ga, gb, None, None, None, None


class _CensusLossGroupsFunction(torch.autograd.Function):
    """census_loss for `groups` equally sized batch slices of stacked inputs (UFlowLoss runs its two directions stacked on
    the batch): every slice is its own census_loss - own mask sum, own normaliser (uflow_utils.py:293) - evaluated in ONE
    launch each way (arf_census_*_groups), the backward writing ONE gradient tensor.  Returns a (groups,) tensor."""

    @staticmethod
    def forward(ctx, im_a, im_b, mask, groups, patch, eps, q):
        im_a, im_b = im_a.contiguous(), im_b.contiguous()
        if im_a.shape != im_b.shape or im_a.dim() != 4 or im_a.shape[1] != 3 or im_a.shape[0] % groups != 0:
            raise ValueError("census_loss: expected two (groups*B,3,H,W) images of equal shape")
        N, _, H, W = im_a.shape
        mask = mask.detach().contiguous()
        if mask.shape != (N, 1, H, W):
            raise ValueError("census_loss: mask must be (groups*B,1,H,W)")
        lib = _lib.load()
        with torch.cuda.device_of(im_a):
            ham = _new_like(im_a, (N, 1, H, W))
            partials = _new_like(im_a, (2 * lib.arf_census_num_partials(N, H, W),))
            sums = _new_like(im_a, (3 * groups,))
            _lib.call("arf_census_fwd_groups", _lib.dev_ptr(im_a, "image_a"), _lib.dev_ptr(im_b, "image_b"),
                      _lib.dev_ptr(mask, "mask"), _lib.dev_ptr(ham), _lib.dev_ptr(partials), _lib.dev_ptr(sums),
                      N, H, W, groups, patch, 1.0, float(eps), float(q), _lib.stream_ptr())
        if _census_group is not None:
            sums = torch.cat([globalise_census_sums(sums[3 * g:3 * g + 3], _census_group) for g in range(groups)])
        ctx.save_for_backward(im_a, im_b, mask, ham, sums)
        ctx.cfg = (groups, patch, float(eps), float(q))
        return sums[2::3].clone()

    @staticmethod
    def backward(ctx, gloss):
        im_a, im_b, mask, ham, sums = ctx.saved_tensors
        groups, patch, eps, q = ctx.cfg
        N, _, H, W = im_a.shape
        gloss = gloss.contiguous()
        with torch.cuda.device_of(im_a):
            ga = torch.empty_like(im_a) if ctx.needs_input_grad[0] else None
            gb = torch.empty_like(im_b) if ctx.needs_input_grad[1] else None
            _lib.call("arf_census_bwd_groups", _lib.dev_ptr(im_a), _lib.dev_ptr(im_b), None, _lib.dev_ptr(ham),
                      _lib.dev_ptr(mask), _lib.dev_ptr(sums), _lib.dev_ptr(gloss, "grad"),
                      _lib.dev_ptr(ga, allow_none=True), _lib.dev_ptr(gb, allow_none=True),
                      N, H, W, groups, patch, 1.0, eps, q, _lib.stream_ptr())
        return ga, gb, None, None, None, None, None


def census_loss_groups(image_a, image_b, mask, groups, patch_size=7):
    """census_loss (uflow_utils.py:282-293) of each of `groups` equal batch slices of the stacked inputs -> (groups,)."""
    return _CensusLossGroupsFunction.apply(image_a, image_b, mask, groups, patch_size, 0.01, 0.4)


def census_transform(image, patch_size):
    """uflow_utils.py:241-261 — the explicit patch*patch-channel transform (compatibility helper; the loss
    functions below never materialise it)."""
    intensities = rgb_to_grayscale(image) * 255
    p = patch_size // 2
    neighbors = torch.nn.functional.unfold(intensities, patch_size, padding=p)
    neighbors = neighbors.view(image.shape[0], patch_size * patch_size, image.shape[2], image.shape[3])
    diff = neighbors - intensities
    return diff / torch.sqrt(.81 + torch.square(diff))


def soft_hamming(a, b, thresh=.1):
    """uflow_utils.py:264-279."""
    sq_dist = torch.square(a - b)
    return torch.sum(sq_dist / (thresh + sq_dist), 1, keepdim=True)


def census_loss(image_a, image_b, mask, patch_size=7):
    """uflow_utils.py:282-293."""
    return _CensusLossFunction.apply(image_a, image_b, mask, patch_size, 0.01, 0.4)


def census_loss_no_penalty(image_a, image_b, mask, patch_size=7):
    """uflow_utils.py:296-306 — per-pixel soft Hamming distance and normalised weight map."""
    hamming = _CensusHammingFunction.apply(image_a, image_b, patch_size, 1.0)
    padded_mask = zero_mask_border(mask, patch_size)
    return hamming, padded_mask / (torch.sum(padded_mask.detach()) + 1e-6)


# --------------------------------------------------------------------------- SSIM ----------
class _SsimFunction(torch.autograd.Function):
    """Five box filters + the SSIM ratios in one pass (csrc/ssim.cu)."""

    @staticmethod
    def forward(ctx, x, y, patch, valid, mode):
        x, y = x.contiguous(), y.contiguous()
        if x.shape != y.shape or x.dim() != 4:
            raise ValueError("SSIM: expected two (B,C,H,W) tensors of equal shape")
        B, C, H, W = x.shape
        r = patch // 2
        Ho, Wo = (H - 2 * r, W - 2 * r) if valid else (H, W)
        with torch.cuda.device_of(x):
            o1 = _new_like(x, (B, C, Ho, Wo))
            o2 = _new_like(x, (B, C, Ho, Wo)) if mode == 0 else None
            _lib.call("arf_ssim_fwd", _lib.dev_ptr(x, "x"), _lib.dev_ptr(y, "y"), _lib.dev_ptr(o1),
                      _lib.dev_ptr(o2, allow_none=True), B * C, H, W, patch, int(valid), mode, _lib.stream_ptr())
        ctx.save_for_backward(x, y)
        ctx.cfg = (patch, int(valid), mode, Ho, Wo)
        if mode == 0:
            return o1, o2
        return o1

    @staticmethod
    def backward(ctx, g1, g2=None):
        x, y = ctx.saved_tensors
        patch, valid, mode, Ho, Wo = ctx.cfg
        B, C, H, W = x.shape
        g1 = g1.contiguous()
        g2 = g2.contiguous() if g2 is not None else None
        with torch.cuda.device_of(x):
            coef = _new_like(x, (5, B * C, Ho, Wo))
            gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
            _lib.call("arf_ssim_bwd", _lib.dev_ptr(x), _lib.dev_ptr(y), _lib.dev_ptr(g1, "grad"),
                      _lib.dev_ptr(g2, allow_none=True), _lib.dev_ptr(coef), _lib.dev_ptr(gx, allow_none=True),
                      _lib.dev_ptr(gy, allow_none=True), B * C, H, W, patch, valid, mode, _lib.stream_ptr())
        return gx, gy, None, None, None


def ssim_loss(image_a, image_b, mask, patch_size=7):
    """uflow_utils.py:309-334 -> ([d1_sq, d2_sq], weight)."""
    d1_sq, d2_sq = _SsimFunction.apply(image_a, image_b, patch_size, False, 0)
    padded_mask = zero_mask_border(mask, patch_size=patch_size)
    return [d1_sq, d2_sq], padded_mask / (torch.sum(padded_mask.detach()) + 1e-6)

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
    """correlation_native.py:13-23 == uflow_model.py:53-92 (pad=md, ks=1, strides 1): one plane per
    displacement (dy slow, dx fast), each the channel mean of f1 times the shifted, zero-padded f2.
    Differentiable; this is the CPU baseline for forward+backward (as fast as the reference's own loop —
    an unfold-based single-op version measured 3.5x slower on 8 cores)."""
    B, C, H, W = f1.shape
    D = 2 * md + 1
    f2p = F.pad(f2, (md, md, md, md))
    planes = [(f1 * f2p[:, :, dy:dy + H, dx:dx + W]).mean(1) for dy in range(D) for dx in range(D)]
    return torch.stack(planes, 1)


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


# --------------------------------------------------------------------------- masks ----------
def mask_invalid(coords):
    """uflow_utils.py:35-50."""
    H, W = coords.shape[2:]
    x, y = coords[:, 0:1], coords[:, 1:2]
    return ((x >= 0) & (x <= W - 1) & (y >= 0) & (y <= H - 1)).to(coords.dtype)


def border_mask(flow):
    """warp_utils.py:119-134."""
    H, W = flow.shape[2:]
    c = flow_to_warp(flow)
    x, y = c[:, 0:1], c[:, 1:2]
    return ((x > 0) & (x < W - 1) & (y > 0) & (y < H - 1)).to(flow.dtype)


def splat_count(coords):
    """The scatter shared by compute_range_map (uflow_utils.py:80-160, warp_utils.py:158-239) and
    get_corresponding_map (warp_utils.py:26-80): bilinear weights of every target added to its four
    integer neighbours, neighbours outside the image dropped."""
    B, _, H, W = coords.shape
    x, y = coords[:, 0].reshape(B, -1), coords[:, 1].reshape(B, -1)
    x0, y0 = torch.floor(x), torch.floor(y)
    fx, fy = x - x0, y - y0
    out = torch.zeros(B, H * W, dtype=coords.dtype)
    for dy, wy in ((0, 1 - fy), (1, fy)):
        for dx, wx in ((0, 1 - fx), (1, fx)):
            xi, yi = (x0 + dx).long(), (y0 + dy).long()
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            lin = yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)
            out.scatter_add_(1, lin, wy * wx * ok.to(coords.dtype))
    return out.view(B, 1, H, W)


def range_map(flow):
    return splat_count(flow_to_warp(flow))


def occu_mask_backward(flow21, th=0.2):
    """warp_utils.py:103-116."""
    c = range_map(flow21).clamp(0, 1)
    return (c < th).to(flow21.dtype) if th > 0 else 1 - c


def occu_mask_bidirection(flow12, flow21, scale=0.01, bias=0.5):
    """warp_utils.py:93-100."""
    w = warp(flow21, flow12, kind="flow")
    mag = (flow12 ** 2).sum(1, keepdim=True) + (w ** 2).sum(1, keepdim=True)
    return (((flow12 + w) ** 2).sum(1, keepdim=True) > scale * mag + bias).to(flow12.dtype)


# --------------------------------------------------------------------------- resize ---------
def _lin_taps(n_out, n_in, step, dtype):
    # ATen area_pixel_compute_source_index, align_corners=False, clamp below at 0
    s = (step * (torch.arange(n_out, dtype=dtype) + 0.5) - 0.5).clamp(min=0)
    i0 = s.floor().long().clamp(max=n_in - 1)
    i1 = (i0 + 1).clamp(max=n_in - 1)
    l1 = s - i0.to(dtype)
    return i0, i1, 1 - l1, l1


def resize_bilinear(img, scale_factor, is_flow):
    """upsample / downsample of uflow_utils.py:163-204 (pass scale_factor < 1 to downsample)."""
    B, C, H, W = img.shape
    Ho, Wo = int(math.floor(H * scale_factor)), int(math.floor(W * scale_factor))
    y0, y1, hy0, hy1 = _lin_taps(Ho, H, 1.0 / scale_factor, img.dtype)
    x0, x1, wx0, wx1 = _lin_taps(Wo, W, 1.0 / scale_factor, img.dtype)
    top = img[:, :, y0][:, :, :, x0] * wx0 + img[:, :, y0][:, :, :, x1] * wx1
    bot = img[:, :, y1][:, :, :, x0] * wx0 + img[:, :, y1][:, :, :, x1] * wx1
    out = top * hy0.view(-1, 1) + bot * hy1.view(-1, 1)
    return out * scale_factor if is_flow else out


# --------------------------------------------------------------------------- census ---------
def _gray255(im):
    return ((im[:, 0] * 0.2989 + im[:, 1] * 0.5870 + im[:, 2] * 0.1140) * 255).unsqueeze(1)


def census_hamming(im_a, im_b, patch=7, mean=False):
    """soft_hamming(census_transform(a), census_transform(b)) (uflow_utils.py:241-279), offset by offset
    instead of through a patch*patch-channel identity convolution.  mean=True: TernaryLoss ARFlow variant."""
    r = patch // 2
    ga, gb = _gray255(im_a), _gray255(im_b)
    pa, pb = F.pad(ga, [r] * 4), F.pad(gb, [r] * 4)
    H, W = ga.shape[2:]
    h = torch.zeros_like(ga)
    for dy in range(patch):
        for dx in range(patch):
            da = pa[:, :, dy:dy + H, dx:dx + W] - ga
            db = pb[:, :, dy:dy + H, dx:dx + W] - gb
            ta = da / torch.sqrt(0.81 + da * da)
            tb = db / torch.sqrt(0.81 + db * db)
            sq = (ta - tb) ** 2
            h = h + sq / (0.1 + sq)
    return h / (patch * patch) if mean else h


def zero_border(mask, r):
    out = torch.zeros_like(mask)
    out[:, :, r:-r, r:-r] = mask[:, :, r:-r, r:-r]
    return out


def census_loss(im_a, im_b, mask, patch=7):
    """uflow_utils.py:282-293."""
    h = census_hamming(im_a, im_b, patch)
    pm = zero_border(mask, patch // 2)
    return ((h.abs() + 0.01) ** 0.4 * pm).sum() / (pm.detach().sum() + 1e-6)


def census_loss_no_penalty(im_a, im_b, mask, patch=7):
    """uflow_utils.py:296-306."""
    pm = zero_border(mask, patch // 2)
    return census_hamming(im_a, im_b, patch), pm / (pm.detach().sum() + 1e-6)


def ternary_loss(im, im_warp, max_distance=1, sum_dist=False):
    """loss_blocks.py:12-62 -> (dist, mask)."""
    patch = 2 * max_distance + 1
    dist = census_hamming(im, im_warp, patch, mean=not sum_dist)
    mask = zero_border(torch.ones_like(dist), max_distance)
    return dist, mask


# --------------------------------------------------------------------------- smoothness -----
def _diff(t, dim, stride=1):
    n = t.shape[dim]
    return t.narrow(dim, stride, n - stride) - t.narrow(dim, 0, n - stride)


def smooth_uflow(im_2, flow_2, edge_constant, w_smooth, order=1):
    """One direction of the smoothness block of UFlowLoss.forward (uflow_loss.py:62-102); im_2 is the
    quarter-resolution image."""
    total = 0
    for dim in (3, 2):
        g = _diff(im_2.detach(), dim, stride=order)
        w = torch.exp(-(edge_constant * g).abs().mean(1, keepdim=True))
        d = _diff(flow_2, dim)
        if order == 2:
            d = _diff(d, dim)
        total = total + (w * (d * d + 0.001 ** 2) ** 0.5).mean()
    return w_smooth * total / 2.0


def smooth_grad_1st(flo, image, alpha, penalty="abs"):
    """loss_blocks.py:93-109."""
    total = 0
    for dim in (3, 2):
        w = torch.exp(-_diff(image, dim).abs().mean(1, keepdim=True) * alpha)
        d = _diff(flo, dim)
        p = d.abs() if penalty == "abs" else torch.sqrt(d * d + 0.001 ** 2)
        total = total + (w * p / 2.0).mean() / 2.0
    return total


def smooth_grad_2nd(flo, image, alpha):
    """loss_blocks.py:112-124."""
    total = 0
    for dim in (3, 2):
        w = torch.exp(-_diff(image, dim).abs().mean(1, keepdim=True) * alpha)
        d2 = _diff(_diff(flo, dim), dim)
        n = w.shape[dim]
        total = total + (w.narrow(dim, 1, n - 1) * d2.abs()).mean() / 2.0
    return total


def uflow_loss(output, target, w_census=1.0, w_smooth=4.0, edge_constant=150.0, with_bk=True, smooth_order=1):
    """UFlowLoss.forward (uflow_loss.py:13-109) -> (total, loss_warp, loss_smooth, mean|flow|, mask1)."""
    f12_0, f21_0 = output[0][:, 0:2], output[0][:, 2:4]
    f12_2, f21_2 = output[2][:, 0:2], output[2][:, 2:4]
    im1, im2 = target[:, :3], target[:, 3:]

    def direction(im_a, im_b, f_ab_0, f_ba_2):
        w0 = flow_to_warp(f_ab_0)
        recons = warp(im_b.detach(), w0, kind="coords")
        occ = resize_bilinear(range_map(f_ba_2).clamp(0, 1), 4.0, False)
        mask = (occ * mask_invalid(w0)).detach()
        return census_loss(im_a, recons, mask), mask

    l1, mask1 = direction(im1, im2, f12_0, f21_2)
    loss_warp = w_census * l1
    loss_smooth = smooth_uflow(resize_bilinear(im1, 0.25, False), f12_2, edge_constant, w_smooth, smooth_order)
    if with_bk:
        l2, _ = direction(im2, im1, f21_0, f12_2)
        loss_warp = loss_warp + w_census * l2
        loss_smooth = loss_smooth + smooth_uflow(resize_bilinear(im2, 0.25, False), f21_2, edge_constant, w_smooth,
                                                 smooth_order)
    return loss_warp + loss_smooth, loss_warp, loss_smooth, output[0].abs().mean(), mask1


# --------------------------------------------------------------------------- CPU train step -
class OracleOps:
    """The oracle's hot-path ops under the names the reference's PWCFlow calls them by (uflow_utils.*)."""
    flow_to_warp = staticmethod(flow_to_warp)

    @staticmethod
    def resample(source, coords):
        return warp(source, coords, kind="coords")

    @staticmethod
    def upsample(img, is_flow, scale_factor=2.0):
        return resize_bilinear(img, scale_factor, is_flow)

    @staticmethod
    def compute_cost_volume(f1, f2, max_displacement):
        return cost_volume(f1, f2, max_displacement)


class CpuTrainStep:
    """The reference's UFlow training step restated on the CPU (PWCFlow + UFlowLoss + Adam,
    trainer/uflow_trainer.py:30-73, configs/chairs_uflow.json / kitti_uflow.json): the baseline bench.py times
    beside the B200 numbers.  Network = oracle/cpu_nets.PWCFlowCPU (a restatement of models/uflow_model.py, no
    arflow_b200 code), hot path = this file."""

    def __init__(self, level_dropout=0.1, feature_norm=True, smooth_order=1, lr=1e-4, seed=0):
        from .cpu_nets import PWCFlowCPU
        torch.manual_seed(seed)
        self.model = PWCFlowCPU(level_dropout=level_dropout, feature_norm=feature_norm)
        self.model.train()
        self.smooth_order = smooth_order
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8)

    def __call__(self, img_pair):
        res = self.model(img_pair, with_bk=True)
        flows = [torch.cat([a, b], 1) for a, b in zip(res['flows_fw'], res['flows_bw'])]
        out = uflow_loss(flows, img_pair, w_census=1.0, w_smooth=4.0, edge_constant=150.0, with_bk=True,
                         smooth_order=self.smooth_order)
        self.opt.zero_grad()
        out[0].backward()
        self.opt.step()
        return float(out[0].detach())


class CpuPwcLiteInference:
    """Config 1 on the CPU: PWCLite(upsample=True, n_frames=2, reduce_dense=True) two-view inference through
    `correlation_native` + `flow_warp` (models/pwclite.py:260-283, restated in oracle/cpu_nets.PWCLiteCPU)."""

    def __init__(self, seed=0):
        from .cpu_nets import PWCLiteCPU
        torch.manual_seed(seed)
        self.model = PWCLiteCPU(upsample=True, n_frames=2, reduce_dense=True).eval()

    def __call__(self, img_pair):
        with torch.no_grad():
            return self.model(img_pair, with_bk=False)['flows_fw'][0]


# --------------------------------------------------------------------------- triangular -----
def stencil_mv(A, X, k=1, transposed=False):
    """matrix_vector_product_general / _T_general (triag_solve.py:29-43, 59-73): tap (i,j) of channel ch
    lives in A[:, 2*(i*(k+1)+j)+ch]; L x accumulates A*X of the source pixel into the pixel (i,j) further
    down/right, L^T x reads X from there."""
    H, W = X.shape[2:]
    Y = torch.zeros_like(X)
    for i in range(k + 1):
        for j in range(k + 1):
            t = 2 * (i * (k + 1) + j)
            a = A[:, t:t + 2, :H - i, :W - j]
            if transposed:
                Y[:, :, :H - i, :W - j] = Y[:, :, :H - i, :W - j] + a * X[:, :, i:, j:]
            else:
                Y[:, :, i:, j:] = Y[:, :, i:, j:] + a * X[:, :, :H - i, :W - j]
    return Y


def substitution(A, B, C, D, X, upper=False):
    """forward_substitution / backward_substitution (triag_solve.py:76-115; triag_solve_cuda.cu:7-69),
    vectorised over the (batch, channel) systems, float64 numpy."""
    a, b, c, x = (t.detach().double().numpy() for t in (A, B, C, X))
    d = D.detach().double().numpy() if D is not None else None
    M, N = a.shape[2:]
    y = x.copy()
    rows = range(M - 1, -1, -1) if upper else range(M)
    cols = range(N - 1, -1, -1) if upper else range(N)
    s = 1 if upper else -1          # neighbour direction
    for i in rows:
        for j in cols:
            ii, jj = i + s, j + s
            hi, hj = 0 <= ii < M, 0 <= jj < N
            ci, cj = (i, j) if upper else (i - 1, j - 1)   # coefficient position
            v = y[:, :, i, j]
            if hi:
                v = v - y[:, :, ii, j] * c[:, :, ci, j]
            if hj:
                v = v - y[:, :, i, jj] * b[:, :, i, cj]
            if hi and hj and d is not None:
                v = v - y[:, :, ii, jj] * d[:, :, ci, cj]
            y[:, :, i, j] = v / a[:, :, i, j]
    return torch.from_numpy(y)


def inverse_diagonal(A, B, C):
    """inverse_diagonal (triag_solve_cuda.cu:72-139) == marginal_variances (triag_solve.py:205-218):
    squared norm of every column of L^-1, one unit-vector solve per pixel."""
    M, N = A.shape[2:]
    Hh = torch.zeros_like(A, dtype=torch.float64)
    for i in range(M):
        for j in range(N):
            e = torch.zeros_like(A, dtype=torch.float64)
            e[:, :, i, j] = 1
            y = substitution(A, B, C, None, e)
            Hh[:, :, i, j] = (y * y).sum(dim=(2, 3))
    return Hh


# --------------------------------------------------------------------------- SSIM / resampler
def _box(t, patch, valid):
    """AvgPool2d(patch, 1, 0 | patch//2) with zeros counted (count_include_pad): sum of shifted slices / patch^2."""
    r = patch // 2
    if not valid:
        t = F.pad(t, [r] * 4)
    H, W = t.shape[2] - 2 * r, t.shape[3] - 2 * r
    acc = 0
    for dy in range(patch):
        for dx in range(patch):
            acc = acc + t[:, :, dy:dy + H, dx:dx + W]
    return acc / (patch * patch)


def _ssim_terms(x, y, patch, valid):
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mx, my = _box(x, patch, valid), _box(y, patch, valid)
    sx = _box(x * x, patch, valid) - mx * mx
    sy = _box(y * y, patch, valid) - my * my
    sxy = _box(x * y, patch, valid) - mx * my
    return (2 * mx * my + C1) / (mx * mx + my * my + C1), (2 * sxy + C2) / (sx + sy + C2)


def ssim_loss(image_a, image_b, mask, patch=7):
    """uflow_utils.py:309-334."""
    S1, S2 = _ssim_terms(image_a, image_b, patch, False)
    pm = zero_border(mask, patch // 2)
    return [(1 - S1).clamp(0, 1), (1 - S2).clamp(0, 1)], pm / (pm.detach().sum() + 1e-6)


def ssim_valid(x, y, md=1):
    """SSIM of loss_blocks.py:65-84."""
    S1, S2 = _ssim_terms(x, y, 2 * md + 1, True)
    return ((1 - S1 * S2) / 2).clamp(0, 1)


def resampler(data, warp_x, warp_y):
    """resampler_with_unstacked_warp (uflow_resampler.py:155-241): NHWC data, floor/ceil taps, zero outside."""
    B, H, W, C = data.shape
    shape = warp_x.shape
    wx, wy = warp_x.reshape(B, -1), warp_y.reshape(B, -1)
    x0, y0 = torch.floor(wx), torch.floor(wy)
    fx, fy = (wx - x0).unsqueeze(-1), (wy - y0).unsqueeze(-1)
    x1, y1 = torch.ceil(wx), torch.ceil(wy)
    flat = data.reshape(B, H * W, C)

    def tap(yy, xx):
        ok = ((xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)).unsqueeze(-1).to(data.dtype)
        lin = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long().unsqueeze(-1).expand(B, -1, C)
        return flat.gather(1, lin) * ok

    out = (tap(y0, x0) * (1 - fx) + tap(y0, x1) * fx) * (1 - fy) + (tap(y1, x0) * (1 - fx) + tap(y1, x1) * fx) * fy
    return out.reshape(*shape, C)

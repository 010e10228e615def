"""Convolution + bias + leaky ReLU with a fused epilogue, and the NHWC dense-block plumbing of the PWC decoders.

The convolution itself is cuDNN (aten.convolution / aten.convolution_backward — out of scope).
What is fused is everything the reference spends around it in `nn.Sequential(nn.Conv2d, nn.LeakyReLU)` and
`func.leaky_relu(conv(x))` (models/uflow_model.py:134-135, 427-436): bias add + activation become one in-place
pass over the convolution output, and leaky-ReLU gradient + bias gradient become one pass in the backward
(instead of an elementwise kernel plus a separate full-tensor reduction per layer).  Same fp32 arithmetic:
y = leaky(conv + b);  g = gy * (y > 0 ? 1 : slope);  db = sum g.

Layout.  cuDNN's sm_100 convolution kernels are NHWC kernels; fed NCHW activations they transpose every input and
output themselves, and NHWC activations whose channel count is not a multiple of 4 take a padding copy.  With
`nhwc_concat` the decoder builds its dense-block inputs directly as packed channels-last tensors whose channel
count is a multiple of 8 (zero channels at the end of the first concat, matching zero columns inserted into the
weights by `pad_in_channels`), so the convolutions run without any layout kernel.  Tensors keep their logical
(B, C, H, W) shape; "NHWC" below means `memory_format=torch.channels_last` strides.
"""
import torch
import torch.nn.functional as func

from . import _lib

CL = torch.channels_last


def is_nhwc(t):
    """True for a channels-last dense 4-D tensor that is not also NCHW-contiguous (C == 1 or H*W == 1 are both)."""
    return t.dim() == 4 and t.is_contiguous(memory_format=CL) and not t.is_contiguous()


def _int_padding(conv):
    if isinstance(conv.padding, str):
        if conv.padding == "valid":
            return [0, 0]
        # 'same' (stride 1, odd kernels — the only way the PWC networks use it): symmetric d*(k-1)/2
        if any(s != 1 for s in conv.stride) or any(k % 2 == 0 for k in conv.kernel_size):
            raise NotImplementedError("conv_bias_leaky: padding='same' needs stride 1 and odd kernels")
        return [d * (k - 1) // 2 for d, k in zip(conv.dilation, conv.kernel_size)]
    return list(conv.padding)


class _ConvBiasLeaky(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, slope, real_in=None):
        nhwc = is_nhwc(x)
        y = func.conv2d(x, weight, None, stride, padding, dilation)
        y = y.contiguous(memory_format=CL) if nhwc else y.contiguous()
        B, C, H, W = y.shape
        bptr = _lib.dev_ptr(bias.contiguous(), "bias") if bias is not None else None
        with torch.cuda.device_of(y):
            if nhwc:
                _lib.call("arf_bias_leaky_nhwc_fwd", y.data_ptr(), bptr, B * H * W, C, float(slope), _lib.stream_ptr())
            else:
                _lib.call("arf_bias_leaky_fwd", _lib.dev_ptr(y, "conv output"), bptr, B, C, H * W, float(slope),
                          _lib.stream_ptr())
        ctx.save_for_backward(x, weight, y)
        ctx.cfg = (list(stride), list(padding), list(dilation), float(slope), bias is not None, nhwc)
        ctx.real_in = real_in
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        stride, padding, dilation, slope, has_bias, nhwc = ctx.cfg
        gy = gy.contiguous(memory_format=CL) if nhwc else gy.contiguous()
        B, C, H, W = y.shape
        need_b = has_bias and ctx.needs_input_grad[2]
        lib = _lib.load()
        with torch.cuda.device_of(y):
            g = torch.empty_like(y)     # preserves the memory format
            db = torch.empty(C, dtype=y.dtype, device=y.device) if need_b else None
            if nhwc:
                part = (torch.empty(lib.arf_bias_leaky_nhwc_num_partials(B * H * W, C), dtype=y.dtype, device=y.device)
                        if need_b else None)
                _lib.call("arf_bias_leaky_nhwc_bwd", gy.data_ptr(), y.data_ptr(), g.data_ptr(),
                          part.data_ptr() if need_b else None, db.data_ptr() if need_b else None,
                          B * H * W, C, slope, _lib.stream_ptr())
            else:
                part = (torch.empty(lib.arf_bias_leaky_num_partials(B, C, H * W), dtype=y.dtype, device=y.device)
                        if need_b else None)
                _lib.call("arf_bias_leaky_bwd", _lib.dev_ptr(gy, "grad"), _lib.dev_ptr(y), _lib.dev_ptr(g),
                          _lib.dev_ptr(part, allow_none=True), _lib.dev_ptr(db, allow_none=True),
                          B, C, H * W, slope, _lib.stream_ptr())
        if (nhwc and ctx.real_in == 3 and tuple(weight.shape) == (32, 8, 3, 3) and stride == [2, 2] and padding == [1, 1]
                and dilation == [1, 1] and not ctx.needs_input_grad[0] and x.dtype == torch.float32):
            # first pyramid layer on the zero-padded image: arf_conv3x3s2_first_wgrad instead of cuDNN's 64x64 wgrad tile
            N, _, Hi, Wi = x.shape
            with torch.cuda.device_of(x):
                out = torch.empty(27 * 32, dtype=x.dtype, device=x.device)
                part = torch.empty(lib.arf_conv3x3s2_first_wgrad_workspace(N, Hi, Wi), dtype=x.dtype, device=x.device)
                _lib.call("arf_conv3x3s2_first_wgrad", x.data_ptr(), g.data_ptr(), out.data_ptr(), part.data_ptr(),
                          N, Hi, Wi, 3, 32, _lib.stream_ptr())
                gw = torch.zeros((32, 8, 3, 3), dtype=x.dtype, device=x.device).contiguous(memory_format=CL)
                gw[:, :3] = out.view(3, 3, 3, 32).permute(3, 2, 0, 1)       # [kh][kw][ci][co] -> (co, ci, kh, kw)
            return None, gw, db, None, None, None, None, None
        gx, gw, _ = torch.ops.aten.convolution_backward(
            g, x, weight, None, stride, padding, dilation, False, [0, 0], 1,
            [bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1]), False])
        return gx, gw, db, None, None, None, None, None


def conv_bias_leaky(conv, x, negative_slope, weight=None, bias=None, real_in=None):
    """leaky_relu(conv(x)) for an nn.Conv2d `conv` (groups 1); `weight` / `bias` override conv.weight / conv.bias
    (padded / channels-last copies).  real_in: number of leading input channels that are not zero padding (the
    image's 3 of 8 in the first pyramid layer) - the weight gradient of the rest is returned as zero.  CUDA tensors take the fused path.  A CPU tensor only occurs in the
    `-m "not gpu"` host-logic tests, which run the module tree with injected oracle ops (PWCFlow(ops=...)) to check the
    wiring against the reference's goldens without a GPU; bench.py's CPU arm does not come through here (it runs
    oracle/cpu_nets.py, which imports nothing from this package)."""
    if not x.is_cuda:
        return func.leaky_relu(conv(x), negative_slope=negative_slope)
    if conv.groups != 1 or conv.padding_mode != "zeros":
        raise NotImplementedError("conv_bias_leaky: groups == 1 and zero padding only")
    w = conv.weight if weight is None else weight
    if x.dtype != torch.float32 or w.dtype != torch.float32:
        raise TypeError("conv_bias_leaky: float32 only")
    return _ConvBiasLeaky.apply(x, w, conv.bias if bias is None else bias, tuple(conv.stride),
                                tuple(_int_padding(conv)), tuple(conv.dilation), negative_slope, real_in)


class _FlowOutConv(torch.autograd.Function):
    """Conv2d(Cin, 2, 3, padding=1) on a packed channels-last input -> NCHW output (the flow head of every decoder
    level and of the refinement network), forward and all three gradients by arf_conv3x3_small_* in fp32: one pass
    over the input each way instead of cuDNN's channel-padded 256-wide tiles (67 us forward, 33 + 145 + 12 us backward
    at 16 x 96 x 128, for a 25 MB operand)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        N, Cin, H, W = x.shape
        w = weight.contiguous(memory_format=CL)
        with torch.cuda.device_of(x):
            y = torch.empty((N, 2, H, W), dtype=x.dtype, device=x.device)
            _lib.call("arf_conv3x3_small_fwd", x.data_ptr(), w.data_ptr(), _lib.dev_ptr(bias, "bias", allow_none=True),
                      y.data_ptr(), N, H, W, Cin, 2, _lib.stream_ptr())
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        N, Cin, H, W = x.shape
        lib = _lib.load()
        nw = 2 * 9 * Cin
        with torch.cuda.device_of(x):
            gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None          # channels-last like x
            out = torch.empty(nw + 2, dtype=x.dtype, device=x.device)
            part = torch.empty(lib.arf_conv3x3_small_bwd_workspace(N, H, W, Cin, 2), dtype=x.dtype, device=x.device)
            _lib.call("arf_conv3x3_small_bwd", x.data_ptr(), _lib.dev_ptr(gy, "grad"), w.data_ptr(),
                      gx.data_ptr() if gx is not None else None, out.data_ptr(), part.data_ptr(), N, H, W, Cin, 2,
                      _lib.stream_ptr())
        gw = out[:nw].view(2, 3, 3, Cin).permute(0, 3, 1, 2)          # a channels-last (2, Cin, 3, 3) tensor
        gb = out[nw:] if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb


def _is_flow_out_conv(conv, x, w):
    return (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and is_nhwc(x) and conv.groups == 1
            and tuple(w.shape[2:]) == (3, 3) and w.shape[0] == 2 and w.shape[1] == x.shape[1] and x.shape[1] % 32 == 0
            and tuple(conv.stride) == (1, 1) and tuple(conv.dilation) == (1, 1) and _int_padding(conv) == [1, 1]
            and conv.padding_mode == "zeros" and (conv.bias is None or conv.bias.dtype == torch.float32))


def conv_plain(conv, x, weight=None):
    """conv(x) with an optional replacement weight (padded / channels-last copy); bias and geometry from `conv`.
    The two-channel 3x3 flow heads on channels-last CUDA inputs run arf_conv3x3_small_* and return NCHW (see
    _FlowOutConv)."""
    w = conv.weight if weight is None else weight
    if _is_flow_out_conv(conv, x, w):
        return _FlowOutConv.apply(x, w, conv.bias)
    return func.conv2d(x, w, conv.bias, conv.stride, _int_padding(conv) if x.is_cuda else conv.padding, conv.dilation)


def image_pair_nhwc(pairs, scale=2.0, shift=-1.0):
    """(B, 6, H, W) image pairs -> the stacked batch [first images; second images] (2B, 8, H, W) channels-last with
    value * scale + shift on the three real channels and five zero channels (arf_image_pair_pack): the networks' input
    stage - torch.cat of the two slices, x * 2 - 1, the channels-last pack and its zero tail - in one pass.  No
    gradient (images are data)."""
    B, C2, H, W = pairs.shape
    if C2 != 6 or not pairs.is_cuda or pairs.dtype != torch.float32 or pairs.requires_grad:
        raise ValueError("image_pair_nhwc: a (B, 6, H, W) float32 CUDA tensor without gradient")
    with torch.cuda.device_of(pairs):
        out = torch.empty((2 * B, H, W, 8), dtype=pairs.dtype, device=pairs.device)
        _lib.call("arf_image_pair_pack", out.data_ptr(), _lib.dev_ptr(pairs.contiguous(), "pairs"), B, H * W, 3, 8,
                  float(scale), float(shift), _lib.stream_ptr())
    return out.permute(0, 3, 1, 2)


class _BiasAddNhwc(torch.autograd.Function):
    """y += bias over the channel axis of a dense channels-last tensor, in place (arf_bias_leaky_nhwc_fwd with slope 1).
    ATen adds a channels-last ConvTranspose2d's bias with its generic strided kernel: 63 us for the (16, 32, 96, 128)
    context up-sampling of chairs_uflow, against ~9 us for one coalesced pass."""

    @staticmethod
    def forward(ctx, y, bias):
        N, C, H, W = y.shape
        with torch.cuda.device_of(y):
            _lib.call("arf_bias_leaky_nhwc_fwd", y.data_ptr(), _lib.dev_ptr(bias, "bias"), N * H * W, C, 1.0,
                      _lib.stream_ptr())
        ctx.mark_dirty(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        return gy, gy.sum((0, 2, 3))


def conv_transpose_bias(up, x):
    """up(x) for a ConvTranspose2d on a channels-last CUDA input: cuDNN without the bias, then the fused bias pass."""
    w = up.weight.contiguous(memory_format=CL)
    if up.bias is None or not x.is_cuda:
        return func.conv_transpose2d(x, w, up.bias, up.stride, up.padding, up.output_padding, up.groups, up.dilation)
    y = func.conv_transpose2d(x, w, None, up.stride, up.padding, up.output_padding, up.groups, up.dilation)
    if not is_nhwc(y) or y.dtype != torch.float32:
        return y + up.bias.view(1, -1, 1, 1)
    return _BiasAddNhwc.apply(y, up.bias)


# ----------------------------------------------------------------------------- NHWC dense-block plumbing --------
def round_up(n, m):
    return (n + m - 1) // m * m


_ZEROS = {}


def _zeros_cl(like, shape):
    """A cached all-zero channels-last block (read-only by convention): the padding pieces of pad_weight are the same
    every step, so they are allocated and filled once instead of costing a fill kernel per layer and step."""
    key = (tuple(shape), like.device, like.dtype)
    z = _ZEROS.get(key)
    if z is None:
        z = torch.zeros(shape, dtype=like.dtype, device=like.device)
        if len(shape) == 4:
            z = z.contiguous(memory_format=CL)
        _ZEROS[key] = z
    return z


class _PadWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, in_pads, out_pad):
        import ctypes
        Co, Ci, KH, KW = weight.shape
        pads = [(int(a), int(n)) for a, n in in_pads if n]
        if len(pads) > 4:
            raise ValueError("pad_weight: at most four insertions")
        Cip, Cop = Ci + sum(n for _, n in pads), Co + int(out_pad)
        at = (ctypes.c_int * 4)(*([a for a, _ in pads] + [0] * (4 - len(pads))))
        cnt = (ctypes.c_int * 4)(*([n for _, n in pads] + [0] * (4 - len(pads))))
        with torch.cuda.device_of(weight):
            out = torch.empty((Cop, Cip, KH, KW), dtype=weight.dtype, device=weight.device, memory_format=CL)
            _lib.call("arf_pad_weight", out.data_ptr(), weight.data_ptr(), Co, Ci, KH, KW, Cop, Cip, *weight.stride(),
                      len(pads), at, cnt, 1, _lib.stream_ptr())
        ctx.meta = (weight.shape, weight.stride(), pads, Cop, Cip)
        return out

    @staticmethod
    def backward(ctx, gout):
        import ctypes
        shape, stride, pads, Cop, Cip = ctx.meta
        Co, Ci, KH, KW = shape
        gout = gout.contiguous(memory_format=CL)
        at = (ctypes.c_int * 4)(*([a for a, _ in pads] + [0] * (4 - len(pads))))
        cnt = (ctypes.c_int * 4)(*([n for _, n in pads] + [0] * (4 - len(pads))))
        with torch.cuda.device_of(gout):
            gw = torch.empty_strided(shape, stride, dtype=gout.dtype, device=gout.device)   # the parameter's own layout
            _lib.call("arf_pad_weight", gw.data_ptr(), gout.data_ptr(), Co, Ci, KH, KW, Cop, Cip, *stride,
                      len(pads), at, cnt, 0, _lib.stream_ptr())
        return gw, None, None


def pad_weight(weight, in_pads=(), out_pad=0):
    """Channels-last copy of a (Cout, Cin, kh, kw) weight with zero input channels inserted and zero output channels
    appended.  in_pads: [(position in the ORIGINAL input-channel order, count), ...] in increasing position — the
    counterpart of the zero channels the NHWC dense block carries (the tail of its first concat, the tail of a
    convolution output that was widened to a multiple of 64).  One kernel each way (arf_pad_weight); the gradient comes
    back in the parameter's own layout with the inserted rows and columns dropped."""
    pads = [(a, n) for a, n in in_pads if n]
    if not pads and not out_pad and weight.is_contiguous(memory_format=CL):
        return weight                      # channels-last parameter (train_step.py), nothing to insert: no copy
    if not weight.is_cuda:
        raise RuntimeError("pad_weight: CUDA tensors only (the CPU twin of the networks does not pad)")
    dense = weight.is_contiguous() or weight.is_contiguous(memory_format=CL)
    return _PadWeight.apply(weight if dense else weight.contiguous(), tuple(pads), int(out_pad))


def pad_in_channels(weight, at, n_pad):
    """pad_weight with one insertion."""
    return pad_weight(weight, [(at, n_pad)])


def out_channel_pad(cout):
    """Zero output channels to append so that cuDNN runs its sm_100 forward kernel: measured on B200 / cuDNN 9
    (tools/conv_layer_probe.py, 16 x 408 x 96 x 128 NHWC, 3x3): Cout = 96 -> 1107 us forward (sm80 fallback kernel),
    Cout = 128 -> 281 us, 64 -> 205 us, 32 -> 152 us; the backward costs 656 vs 587 us."""
    return (-cout) % 64 if cout > 64 else 0


class _NhwcConcat(torch.autograd.Function):
    """torch.cat(parts, dim=1) into a packed channels-last tensor with the channel count rounded up to `c_total`
    (zero tail).  Parts may be NHWC or NCHW; each part's gradient comes back in the part's own layout.  slopes[i] != 1
    applies a leaky ReLU to part i on the way in (NCHW parts only) — the cost volume's activation, fused."""

    @staticmethod
    def forward(ctx, c_total, slopes, *parts):
        B, _, H, W = parts[0].shape
        dev, dt = parts[0].device, parts[0].dtype
        out = torch.empty((B, c_total, H, W), dtype=dt, device=dev, memory_format=CL)
        meta, off = [], 0
        with torch.cuda.device_of(out):
            for p, slope in zip(parts, slopes):
                if p.shape[0] != B or p.shape[2:] != (H, W) or p.dtype != torch.float32 or not p.is_cuda:
                    raise ValueError("nhwc_concat: parts must be float32 CUDA tensors of equal batch and spatial size")
                nhwc = is_nhwc(p)
                if not nhwc:
                    p = p.contiguous()
                C = p.shape[1]
                if slope != 1.0:
                    if nhwc or C <= 4:
                        raise NotImplementedError("nhwc_concat: fused activation needs an NCHW part with more than 4 channels")
                    _lib.call("arf_nhwc_pack_act", out.data_ptr(), p.data_ptr(), B, H * W, C, c_total, off, float(slope),
                              _lib.stream_ptr())
                else:
                    _lib.call("arf_nhwc_pack", out.data_ptr(), p.data_ptr(), B, H * W, C, c_total, off, int(nhwc),
                              _lib.stream_ptr())
                meta.append((C, off, nhwc, float(slope)))
                off += C
            if off > c_total:
                raise ValueError("nhwc_concat: parts exceed c_total")
            if off < c_total:
                _lib.call("arf_nhwc_pack", out.data_ptr(), None, B, H * W, c_total - off, c_total, off, 1,
                          _lib.stream_ptr())
        ctx.meta = meta
        ctx.dims = (B, H, W, c_total)
        if any(s != 1.0 for s in slopes):
            ctx.save_for_backward(out)       # the activation mask is read back from the packed output
        return out

    @staticmethod
    def backward(ctx, gout):
        B, H, W, c_total = ctx.dims
        gout = gout.contiguous(memory_format=CL)
        fwd = ctx.saved_tensors[0] if ctx.saved_tensors else None
        grads = [None, None]
        with torch.cuda.device_of(gout):
            for i, (C, off, nhwc, slope) in enumerate(ctx.meta):
                if not ctx.needs_input_grad[i + 2]:
                    grads.append(None)
                    continue
                g = torch.empty((B, C, H, W), dtype=gout.dtype, device=gout.device,
                                memory_format=CL if nhwc else torch.contiguous_format)
                if slope != 1.0:
                    _lib.call("arf_nhwc_unpack_act", g.data_ptr(), gout.data_ptr(), fwd.data_ptr(), B, H * W, C, c_total,
                              off, slope, _lib.stream_ptr())
                else:
                    _lib.call("arf_nhwc_unpack", g.data_ptr(), gout.data_ptr(), B, H * W, C, c_total, off, int(nhwc),
                              _lib.stream_ptr())
                grads.append(g)
        return tuple(grads)


def nhwc_concat(parts, multiple=8):
    """Channel concat of CUDA tensors into a packed channels-last tensor, channels padded up to `multiple`.
    A part may be given as (tensor, negative_slope): leaky ReLU applied on the way in (NCHW parts).
    Returns (tensor, n_real_channels)."""
    tensors = [p[0] if isinstance(p, tuple) else p for p in parts]
    slopes = tuple(float(p[1]) if isinstance(p, tuple) else 1.0 for p in parts)
    n = sum(t.shape[1] for t in tensors)
    return _NhwcConcat.apply(round_up(n, multiple), slopes, *tensors), n


class _DenseBlockNhwc(torch.autograd.Function):
    """The PWC decoder's dense block (models/uflow_model.py:195-198) on packed channels-last tensors, with a manual
    backward:   x_{i+1} = cat([x_i, leaky(conv_i(x_i) + b_i)]),   returns the last layer's output.

    Forward = what conv_bias_leaky + nhwc_concat do layer by layer, except that a layer's activated output is written
    straight into its column slice of the next input.  The backward walks the layers in reverse and
    keeps ONE running gradient G (w.r.t. x_{i+1}): a layer's output gradient is read straight out of G's column
    slice by the fused epilogue backward (no unpack copy), and the part of G that belongs to x_i is accumulated
    into the convolution's input gradient in place (arf_nhwc_unpack_add) — instead of autograd's unpack copy of
    both parts plus a separate add per layer.  All weights arrive padded / channels-last (pad_weight)."""

    @staticmethod
    def forward(ctx, slope, geom, x0, *wb):
        n = len(wb) // 2
        ws, bs = wb[:n], wb[n:]
        xs, widths = [x0], []
        x = x0
        B, _, H, W = x0.shape
        rows = B * H * W
        y = None
        for i in range(n):
            stride, padding, dilation = geom[i]
            y = func.conv2d(x, ws[i], None, stride, padding, dilation).contiguous(memory_format=CL)
            C = y.shape[1]
            widths.append(C)
            bptr = bs[i].contiguous().data_ptr() if bs[i] is not None else None
            with torch.cuda.device_of(y):
                if i + 1 < n:
                    # next input = [x | leaky(conv + b)]: the prefix is copied, the activated output is written straight
                    # into its column slice (it is read from there again in the backward; no separate copy of it exists)
                    cin = x.shape[1]
                    nx = torch.empty((B, cin + C, H, W), dtype=x.dtype, device=x.device, memory_format=CL)
                    _lib.call("arf_nhwc_pack", nx.data_ptr(), x.data_ptr(), B, H * W, cin, cin + C, 0, 1, _lib.stream_ptr())
                    _lib.call("arf_bias_leaky_nhwc_fwd_ld", y.data_ptr(), nx.data_ptr() + 4 * cin, cin + C, bptr, rows, C,
                              float(slope), _lib.stream_ptr())
                    x = nx
                    xs.append(x)
                else:
                    _lib.call("arf_bias_leaky_nhwc_fwd", y.data_ptr(), bptr, rows, C, float(slope), _lib.stream_ptr())
        ctx.save_for_backward(*xs, y, *ws)
        ctx.cfg = (n, float(slope), geom, [b is not None for b in bs], widths)
        return y

    @staticmethod
    def backward(ctx, gy_last):
        n, slope, geom, has_bias, widths = ctx.cfg
        saved = ctx.saved_tensors
        xs, y_last, ws = saved[:n], saved[n], saved[n + 1:2 * n + 1]
        lib = _lib.load()
        B, _, H, W = xs[0].shape
        rows = B * H * W
        gws, gbs = [None] * n, [None] * n
        G = None            # gradient w.r.t. xs[i + 1] (packed, width = xs[i + 1].shape[1])
        with torch.cuda.device_of(xs[0]):
            for i in range(n - 1, -1, -1):
                C, cin = widths[i], xs[i].shape[1]
                if G is None:
                    gy = gy_last.contiguous(memory_format=CL)
                    gy_ptr, gy_ld = gy.data_ptr(), C
                    y_ptr, y_ld = y_last.data_ptr(), C
                else:
                    gy_ptr, gy_ld = G.data_ptr() + 4 * cin, G.shape[1]              # columns [cin, cin + C) of G
                    y_ptr, y_ld = xs[i + 1].data_ptr() + 4 * cin, xs[i + 1].shape[1]   # ... and of the next input
                need_b = has_bias[i] and ctx.needs_input_grad[3 + n + i]
                g = torch.empty((B, C, H, W), dtype=xs[0].dtype, device=xs[0].device, memory_format=CL)
                db = torch.empty(C, dtype=g.dtype, device=g.device) if need_b else None
                part = (torch.empty(lib.arf_bias_leaky_nhwc_num_partials(rows, C), dtype=g.dtype, device=g.device)
                        if need_b else None)
                _lib.call("arf_bias_leaky_nhwc_bwd_ld", gy_ptr, gy_ld, y_ptr, y_ld, g.data_ptr(),
                          part.data_ptr() if need_b else None, db.data_ptr() if need_b else None, rows, C, slope,
                          _lib.stream_ptr())
                stride, padding, dilation = geom[i]
                need_x = i > 0 or ctx.needs_input_grad[2]
                gx, gw, _ = torch.ops.aten.convolution_backward(
                    g, xs[i], ws[i], None, list(stride), list(padding), list(dilation), False, [0, 0], 1,
                    [bool(need_x), bool(ctx.needs_input_grad[3 + i]), False])
                if need_x:
                    gx = gx.contiguous(memory_format=CL)
                    if G is not None:
                        _lib.call("arf_nhwc_unpack_add", gx.data_ptr(), G.data_ptr(), B, H * W, cin, G.shape[1], 0,
                                  _lib.stream_ptr())
                G = gx
                gws[i], gbs[i] = gw, db
        return (None, None, G, *gws, *gbs)


def dense_block_nhwc(x0, convs, weights, biases, negative_slope):
    """x0: packed channels-last input; convs: the nn.Conv2d modules (geometry); weights / biases: their padded,
    channels-last weights and (padded) biases.  Returns the last layer's activated output (channels-last)."""
    geom = tuple((tuple(c.stride), tuple(_int_padding(c)), tuple(c.dilation)) for c in convs)
    return _DenseBlockNhwc.apply(negative_slope, geom, x0, *weights, *biases)


class _ToNchw(torch.autograd.Function):
    """channels-last -> NCHW copy (and NCHW -> channels-last for the gradient) through the tiled transpose of
    arf_nhwc_transpose: the features the NCHW hot-path kernels (warp, cost volume) read.  out[(n + shift) % N] = x[n]."""

    @staticmethod
    def forward(ctx, x, shift):
        B, C, H, W = x.shape
        with torch.cuda.device_of(x):
            out = torch.empty((B, C, H, W), dtype=x.dtype, device=x.device)
            _lib.call("arf_nhwc_transpose", out.data_ptr(), x.data_ptr(), B, H * W, C, 1, int(shift), _lib.stream_ptr())
        ctx.shift = int(shift)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B, C, H, W = g.shape
        with torch.cuda.device_of(g):
            out = torch.empty((B, C, H, W), dtype=g.dtype, device=g.device, memory_format=CL)
            _lib.call("arf_nhwc_transpose", out.data_ptr(), g.data_ptr(), B, H * W, C, 0, ctx.shift, _lib.stream_ptr())
        return out, None


def to_nchw(x, batch_shift=0):
    """NCHW-contiguous copy of a channels-last CUDA float32 tensor with its batch rotated: out[(n + shift) % N] = x[n]
    (shift = N/2 swaps the halves of a stacked two-direction batch).  Coalesced transpose both ways."""
    if not (x.is_cuda and x.dtype == torch.float32 and is_nhwc(x)):
        x = x.contiguous()
        return torch.roll(x, batch_shift, 0) if batch_shift else x
    return _ToNchw.apply(x, batch_shift)

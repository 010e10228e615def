"""Convolution + bias + leaky ReLU with a fused epilogue (arf_bias_leaky_fwd / _bwd).

The convolution itself is cuDNN (aten.convolution / aten.convolution_backward — out of scope).
What is fused is everything the reference spends around it in `nn.Sequential(nn.Conv2d, nn.LeakyReLU)` and
`func.leaky_relu(conv(x))` (models/uflow_model.py:134-135, 427-436): bias add + activation become one in-place
pass over the convolution output, and leaky-ReLU gradient + bias gradient become one pass in the backward
(instead of an elementwise kernel plus a separate full-tensor reduction per layer).  Same fp32 arithmetic:
y = leaky(conv + b);  g = gy * (y > 0 ? 1 : slope);  db = sum g.
"""
import torch
import torch.nn.functional as func

from . import _lib


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
    def forward(ctx, x, weight, bias, stride, padding, dilation, slope):
        y = func.conv2d(x, weight, None, stride, padding, dilation).contiguous()
        B, C, H, W = y.shape
        with torch.cuda.device_of(y):
            _lib.call("arf_bias_leaky_fwd", _lib.dev_ptr(y, "conv output"),
                      _lib.dev_ptr(bias.contiguous(), "bias") if bias is not None else None,
                      B, C, H * W, float(slope), _lib.stream_ptr())
        ctx.save_for_backward(x, weight, y)
        ctx.cfg = (list(stride), list(padding), list(dilation), float(slope), bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        stride, padding, dilation, slope, has_bias = ctx.cfg
        gy = gy.contiguous()
        B, C, H, W = y.shape
        need_b = has_bias and ctx.needs_input_grad[2]
        lib = _lib.load()
        with torch.cuda.device_of(y):
            g = torch.empty_like(y)
            db = torch.empty(C, dtype=y.dtype, device=y.device) if need_b else None
            part = (torch.empty(lib.arf_bias_leaky_num_partials(B, C, H * W), dtype=y.dtype, device=y.device)
                    if need_b else None)
            _lib.call("arf_bias_leaky_bwd", _lib.dev_ptr(gy, "grad"), _lib.dev_ptr(y), _lib.dev_ptr(g),
                      _lib.dev_ptr(part, allow_none=True), _lib.dev_ptr(db, allow_none=True),
                      B, C, H * W, slope, _lib.stream_ptr())
        gx, gw, _ = torch.ops.aten.convolution_backward(
            g, x, weight, None, stride, padding, dilation, False, [0, 0], 1,
            [bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1]), False])
        return gx, gw, db, None, None, None, None


def conv_bias_leaky(conv, x, negative_slope):
    """leaky_relu(conv(x)) for an nn.Conv2d `conv` (groups 1).  CUDA tensors take the fused path; a CPU tensor means
    the caller is the oracle-backed CPU twin of the network (tests, bench.py's cpu_baseline) and gets plain torch."""
    if not x.is_cuda:
        return func.leaky_relu(conv(x), negative_slope=negative_slope)
    if conv.groups != 1 or conv.padding_mode != "zeros":
        raise NotImplementedError("conv_bias_leaky: groups == 1 and zero padding only")
    return _ConvBiasLeaky.apply(x, conv.weight, conv.bias, tuple(conv.stride), tuple(_int_padding(conv)),
                                tuple(conv.dilation), negative_slope)

"""Cost-volume correlation behind the reference's module API.

Drop-in for (paths relative to deu439/ARFlow):
  * models/correlation_package/correlation.py:47-61   Correlation(pad_size, kernel_size,
        max_displacement, stride1, stride2, corr_multiply)   (legacy autograd there; new-style here)
  * models/correlation_native.py:6-23                 Correlation(max_displacement=4, *args, **kwargs)
  * models/uflow_model.py:53-92                       compute_cost_volume(features1, features2, max_displacement)
All three compute the same tensor for the settings the models use; one kernel pair
(`arf_corr_fwd` / `arf_corr_bwd`) sits behind all of them.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


def corr_out_dims(H, W, pad, ks, md, s1, s2):
    d2, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.call("arf_corr_out_dims", H, W, pad, ks, md, s1, s2,
              ctypes.byref(d2), ctypes.byref(oh), ctypes.byref(ow))
    return d2.value, oh.value, ow.value


class CorrelationFunction(torch.autograd.Function):
    """New-style replacement of correlation.py:6-44 (which no longer runs on torch >= 2)."""

    @staticmethod
    def forward(ctx, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2,
                corr_multiply):
        if input1.shape != input2.shape or input1.dim() != 4:
            raise ValueError("Correlation: inputs must be two (B,C,H,W) tensors of equal shape")
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        B, C, H, W = input1.shape
        geom = (pad_size, kernel_size, max_displacement, stride1, stride2)
        d2, oh, ow = corr_out_dims(H, W, *geom)
        with torch.cuda.device_of(input1):
            out = torch.empty((B, d2, oh, ow), dtype=input1.dtype, device=input1.device)
            _lib.call("arf_corr_fwd", _lib.dev_ptr(input1, "input1"), _lib.dev_ptr(input2, "input2"),
                      _lib.dev_ptr(out), B, C, H, W, *geom, _lib.stream_ptr())
        ctx.save_for_backward(input1, input2)
        ctx.geom = geom
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        B, C, H, W = input1.shape
        grad_output = grad_output.contiguous()
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        with torch.cuda.device_of(input1):
            g1 = torch.empty_like(input1) if need1 else None
            g2 = torch.empty_like(input2) if need2 else None
            _lib.call("arf_corr_bwd", _lib.dev_ptr(input1), _lib.dev_ptr(input2),
                      _lib.dev_ptr(grad_output, "grad_output"),
                      _lib.dev_ptr(g1, allow_none=True), _lib.dev_ptr(g2, allow_none=True),
                      B, C, H, W, *ctx.geom, _lib.stream_ptr())
        return g1, g2, None, None, None, None, None, None


class Correlation(nn.Module):
    """Both reference constructors, told apart by the call form:

      * CUDA package, correlation.py:48 - `Correlation(pad_size=0, kernel_size=0, max_displacement=0, stride1=1,
        stride2=2, corr_multiply=1)`: chosen when `pad_size` is passed by keyword or two or more positional arguments
        are given (what every model does, pwclite.py:124-126).  Its defaults are kept, including stride2=2 and the
        unusable kernel_size=0 (division by kernel_size**2 * C), which raises at the first forward here.
      * correlation_native.py:7 - `Correlation(max_displacement=4, *args, **kwargs)`: at most one positional argument
        and no `pad_size`; every other argument is swallowed as there, geometry pad = md, kernel 1, strides 1.

    `corr_multiply` is accepted and ignored, as in the reference kernels.
    """

    _PKG = ("pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply")

    def __init__(self, *args, **kwargs):
        super().__init__()
        unknown = set(kwargs) - set(self._PKG)
        if unknown:
            raise TypeError("Correlation: unexpected arguments %s" % sorted(unknown))
        if len(args) > len(self._PKG):
            raise TypeError("Correlation: at most %d positional arguments" % len(self._PKG))
        if "pad_size" in kwargs or len(args) >= 2:
            cfg = dict(pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1)
            for k, v in zip(self._PKG, args):
                if k in kwargs:
                    raise TypeError("Correlation: got multiple values for argument '%s'" % k)
                cfg[k] = v
            cfg.update(kwargs)
        else:
            md = args[0] if args else kwargs.get("max_displacement", 4)
            if args and "max_displacement" in kwargs:
                raise TypeError("Correlation: got multiple values for argument 'max_displacement'")
            cfg = dict(pad_size=md, kernel_size=1, max_displacement=md, stride1=1, stride2=1,
                       corr_multiply=kwargs.get("corr_multiply", 1))
        self.pad_size = cfg["pad_size"]
        self.kernel_size = cfg["kernel_size"]
        self.max_displacement = cfg["max_displacement"]
        self.stride1 = cfg["stride1"]
        self.stride2 = cfg["stride2"]
        self.corr_multiply = cfg["corr_multiply"]
        self.output_dim = 2 * (self.max_displacement // max(self.stride2, 1)) + 1

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2, self.pad_size, self.kernel_size,
                                         self.max_displacement, self.stride1, self.stride2,
                                         self.corr_multiply)


def compute_cost_volume(features1, features2, max_displacement):
    """uflow_model.py:53-92 (same ValueError for a displacement that does not fit)."""
    _, _, height, _ = features1.shape
    if max_displacement <= 0 or max_displacement >= height:
        raise ValueError(f'Max displacement of {max_displacement} is too large.')
    return CorrelationFunction.apply(features1, features2, max_displacement, 1, max_displacement, 1, 1, 1)

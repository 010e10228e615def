"""Fused bias + leaky-ReLU epilogue around the cuDNN convolutions (arf_bias_leaky_fwd / _bwd) against the plain
nn.Conv2d -> leaky_relu chain the reference uses (models/uflow_model.py:134-135, 427-436)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as func

from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cin,cout,k,stride,padding,dilation,shape", [
    (3, 32, 3, 2, 1, 1, (2, 3, 40, 56)),            # first pyramid layer
    (32, 32, 3, 1, 1, 1, (2, 32, 20, 28)),
    (147, 128, 3, 1, "same", 1, (2, 147, 24, 32)),  # decoder layer
    (34, 128, 3, 1, "same", 4, (1, 34, 33, 47)),    # dilated refinement layer, HW % 4 != 0 -> scalar path
    (16, 8, 1, 1, 0, 1, (3, 16, 9, 10)),
])
def test_conv_bias_leaky_matches_torch(cin, cout, k, stride, padding, dilation, shape):
    from arflow_b200.fused_conv import conv_bias_leaky
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(cin * 7 + cout)
        conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, dilation=dilation).cuda()
        x = torch.randn(shape, device="cuda", requires_grad=True)
        ref = func.leaky_relu(conv(x), negative_slope=0.1)
        w = torch.randn_like(ref)
        rg = torch.autograd.grad((ref * w).sum(), [x, conv.weight, conv.bias])
        out = conv_bias_leaky(conv, x, 0.1)
        og = torch.autograd.grad((out * w).sum(), [x, conv.weight, conv.bias])
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert_close(out, ref, 1e-6, "forward")
    for a, b, name in zip(og, rg, ("grad x", "grad weight", "grad bias")):
        assert_close(a, b, 1e-5, name)


def test_conv_without_bias_and_without_input_grad():
    from arflow_b200.fused_conv import conv_bias_leaky
    conv = nn.Conv2d(8, 16, 3, padding=1, bias=False).cuda()
    x = torch.randn(2, 8, 16, 16, device="cuda")          # no grad needed for the input (first layer of a net)
    out = conv_bias_leaky(conv, x, 0.1)
    ref = func.leaky_relu(conv(x), negative_slope=0.1)
    (gw,) = torch.autograd.grad(out.sum(), [conv.weight])
    (rw,) = torch.autograd.grad(ref.sum(), [conv.weight])
    assert_close(out, ref, 1e-6)
    assert_close(gw, rw, 1e-4)     # TF32 convolutions (torch default) on both sides

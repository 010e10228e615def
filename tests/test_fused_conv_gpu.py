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


# --------------------------------------------------------------------------- channels-last plumbing
@pytest.mark.parametrize("cin,cout,k,stride,padding,dilation,shape", [
    (152, 128, 3, 1, "same", 1, (2, 152, 24, 32)),
    (408, 96, 3, 1, "same", 1, (2, 408, 12, 16)),      # 96 output channels: 24 column groups (not a divisor of 256)
    (40, 128, 3, 1, "same", 8, (1, 40, 33, 47)),
    (3, 32, 3, 2, 1, 1, (2, 3, 40, 56)),
    (32, 30, 3, 1, 1, 1, (2, 32, 9, 11)),              # C % 4 != 0 -> scalar epilogue
])
def test_conv_bias_leaky_channels_last(cin, cout, k, stride, padding, dilation, shape):
    from arflow_b200.fused_conv import CL, conv_bias_leaky, is_nhwc
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(cin + cout)
        conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, dilation=dilation).cuda()
        x = torch.randn(shape, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
        ref = func.leaky_relu(conv(x), negative_slope=0.1)
        w = torch.randn_like(ref)
        rg = torch.autograd.grad((ref * w).sum(), [x, conv.weight, conv.bias])
        out = conv_bias_leaky(conv, x, 0.1, weight=conv.weight.contiguous(memory_format=CL))
        og = torch.autograd.grad((out * w).sum(), [x, conv.weight, conv.bias])
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert is_nhwc(out)
    assert_close(out, ref, 1e-5, "forward")
    for a, b, name in zip(og, rg, ("grad x", "grad weight", "grad bias")):
        assert_close(a, b, 2e-5, name)


@pytest.mark.parametrize("B,H,W", [(2, 24, 32), (3, 7, 9)])
def test_nhwc_concat_matches_torch_cat(B, H, W):
    """Mixed NCHW / channels-last parts, unaligned offsets, zero tail; gradients return in each part's layout."""
    from arflow_b200.fused_conv import CL, is_nhwc, nhwc_concat
    gen = torch.Generator().manual_seed(B * H)
    parts = [torch.randn(B, 32, H, W, generator=gen).cuda().contiguous(memory_format=CL),
             torch.randn(B, 2, H, W, generator=gen).cuda(),
             torch.randn(B, 81, H, W, generator=gen).cuda(),
             torch.randn(B, 32, H, W, generator=gen).cuda().contiguous(memory_format=CL)]
    parts = [p.requires_grad_(True) for p in parts]
    out, n = nhwc_concat(parts)
    assert n == 147 and out.shape == (B, 152, H, W) and is_nhwc(out)
    ref = torch.cat([p.detach() for p in parts], dim=1)
    assert torch.equal(out.detach()[:, :147], ref) and float(out.detach()[:, 147:].abs().max()) == 0.0
    w = torch.randn(out.shape, generator=gen).cuda()
    grads = torch.autograd.grad((out * w).sum(), parts)
    off = 0
    for p, g in zip(parts, grads):
        assert torch.equal(g, w[:, off:off + p.shape[1]])
        assert is_nhwc(g) == is_nhwc(p)
        off += p.shape[1]
    # aligned two-part concat of channels-last tensors (the dense-block step)
    a = torch.randn(B, 152, H, W, generator=gen).cuda().contiguous(memory_format=CL)
    b = torch.randn(B, 128, H, W, generator=gen).cuda().contiguous(memory_format=CL)
    out2, n2 = nhwc_concat([a, b])
    assert n2 == 280 and torch.equal(out2, torch.cat([a, b], dim=1))


def test_pad_in_channels():
    from arflow_b200.fused_conv import pad_in_channels
    w = torch.randn(8, 147, 3, 3, device="cuda", requires_grad=True)
    wp = pad_in_channels(w, 147, 5)
    assert wp.shape == (8, 152, 3, 3) and float(wp.detach()[:, 147:].abs().max()) == 0.0 and torch.equal(wp[:, :147], w)
    w2 = torch.randn(8, 275, 3, 3, device="cuda", requires_grad=True)
    wp2 = pad_in_channels(w2, 147, 5)
    assert torch.equal(wp2[:, :147], w2[:, :147]) and torch.equal(wp2[:, 152:], w2[:, 147:])
    (g,) = torch.autograd.grad(wp2.sum(), [w2])
    assert torch.equal(g, torch.ones_like(w2))


# --------------------------------------------------------------------------- normalize_features (row N1)
@pytest.mark.parametrize("shape", [(2, 32, 24, 32), (3, 32, 6, 8), (1, 5, 7, 9), (2, 32, 48, 64), (2, 32, 32, 33)])
def test_normalize_features_fused_matches_torch_chain(shape):
    """arf_featnorm_* against the reference's chain of torch ops (models/uflow_model.py:8-50), values and gradients: the
    single-launch kernels (n <= 8192 elements per map and sample) and the three-kernel chain."""
    from arflow_b200.uflow_model import normalize_features
    gen = torch.Generator().manual_seed(shape[1] + shape[2])
    f1 = (torch.randn(shape, generator=gen) * 0.7 + 0.3).cuda().requires_grad_(True)
    f2 = (torch.randn(shape, generator=gen) * 1.3 - 0.2).cuda().requires_grad_(True)

    def chain(a, b):      # the reference's own formulation, in float64 for the ground truth
        a, b = a.double(), b.double()
        stats = [torch.var_mean(t, dim=[1, 2, 3], keepdim=True) for t in (a, b)]
        mean = (stats[0][1] + stats[1][1]) / 2
        std = torch.sqrt((stats[0][0] + stats[1][0]) / 2 + 1e-16)
        return (a - mean) / std, (b - mean) / std
    r1, r2 = chain(f1, f2)
    w1, w2 = torch.randn(shape, generator=gen).cuda(), torch.randn(shape, generator=gen).cuda()
    rg = torch.autograd.grad((r1 * w1).sum() + (r2 * w2).sum(), [f1, f2])
    y1, y2 = normalize_features([f1, f2], normalize=True, center=True, moments_across_channels=True,
                                moments_across_images=True)
    og = torch.autograd.grad((y1 * w1).sum() + (y2 * w2).sum(), [f1, f2])
    assert_close(y1, r1, 1e-5, "y1")
    assert_close(y2, r2, 1e-5, "y2")
    assert_close(og[0], rg[0], 1e-4, "grad f1")
    assert_close(og[1], rg[1], 1e-4, "grad f2")
    # only one output used downstream
    (g_only,) = torch.autograd.grad((normalize_features([f1, f2], True, True, True, True)[0] * w1).sum(), [f2])
    (r_only,) = torch.autograd.grad((chain(f1, f2)[0] * w1).sum(), [f2])
    assert_close(g_only, r_only, 1e-4, "grad f2 through y1 only")


def test_dense_block_manual_backward_matches_layerwise_autograd():
    """_DenseBlockNhwc (one running gradient, strided epilogue backward, accumulate-unpack) against the same block
    composed from conv_bias_leaky + nhwc_concat under plain autograd."""
    from arflow_b200.fused_conv import CL, conv_bias_leaky, dense_block_nhwc, nhwc_concat, pad_weight
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(3)
        widths, c0 = [128, 128, 96, 64, 32], 152
        convs, cin = [], c0
        for c in widths:
            convs.append(nn.Conv2d(cin, c, 3, padding="same").cuda())
            cin += c
        x0 = torch.randn(2, c0, 12, 16, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
        params = [p for conv in convs for p in (conv.weight, conv.bias)]

        x = x0
        for i, conv in enumerate(convs):
            y = conv_bias_leaky(conv, x, 0.1, weight=pad_weight(conv.weight))
            if i + 1 < len(convs):
                x, _ = nhwc_concat([x, y])
        w = torch.randn_like(y)
        ref = torch.autograd.grad((y * w).sum(), [x0] + params)

        out = dense_block_nhwc(x0, convs, [pad_weight(c.weight) for c in convs], [c.bias for c in convs], 0.1)
        got = torch.autograd.grad((out * w).sum(), [x0] + params)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert torch.equal(out, y)
    for a, b in zip(got, ref):
        assert_close(a, b, 2e-5)


def test_nhwc_concat_with_fused_leaky_relu():
    """(tensor, slope) parts: leaky ReLU applied inside the pack, its gradient inside the unpack."""
    from arflow_b200.fused_conv import CL, nhwc_concat
    gen = torch.Generator().manual_seed(8)
    a = torch.randn(2, 32, 13, 17, generator=gen).cuda().contiguous(memory_format=CL).requires_grad_(True)
    cv = torch.randn(2, 81, 13, 17, generator=gen).cuda().requires_grad_(True)
    out, n = nhwc_concat([a, (cv, 0.1)])
    ref = torch.cat([a, func.leaky_relu(cv, negative_slope=0.1)], dim=1)
    assert n == 113 and out.shape[1] == 120
    assert torch.equal(out.detach()[:, :113], ref.detach())
    w = torch.randn(out.shape, generator=gen).cuda()
    g = torch.autograd.grad((out * w).sum(), [a, cv])
    r = torch.autograd.grad((ref * w[:, :113]).sum(), [a, cv])
    assert torch.equal(g[0], r[0]) and torch.equal(g[1], r[1])


@pytest.mark.parametrize("shape", [(2, 32, 12, 16), (1, 32, 5, 7)])
def test_conv_transpose_bias_matches_torch(shape):
    """context up-sampling (models/uflow_model.py:283-291): ConvTranspose2d(32, 32, 4, 2, 1) on a channels-last input
    with the bias added by the fused pass; values and the three gradients against nn.ConvTranspose2d itself."""
    from arflow_b200.fused_conv import CL, conv_transpose_bias
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(5)
        up = nn.ConvTranspose2d(32, 32, 4, 2, 1).cuda()
        x = torch.randn(shape, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
        ref = up(x)
        w = torch.randn_like(ref)
        rg = torch.autograd.grad((ref * w).sum(), [x, up.weight, up.bias])
        out = conv_transpose_bias(up, x)
        og = torch.autograd.grad((out * w).sum(), [x, up.weight, up.bias])
        assert_close(out, ref, 1e-5, "conv_transpose + bias")
        for a, b, name in zip(og, rg, ("dx", "dw", "db")):
            assert_close(a, b, 1e-4, name)
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("shape,bias", [((2, 32, 12, 16), True), ((1, 64, 7, 9), True), ((2, 32, 5, 40), False),
                                        ((3, 32, 33, 70), True)])
def test_flow_head_conv_gradients_match_torch(shape, bias):
    """Conv2d(Cin, 2, 3, padding=1) on a channels-last input (models/uflow_model.py:139-143): output, input gradient
    (cuDNN) and the weight / bias gradients of arf_conv3x3_small_wgrad against torch's own fp32 convolution; widths
    that are not multiples of 4 or of the 32-pixel run, one and two channel groups."""
    from arflow_b200.fused_conv import CL, _FlowOutConv, conv_plain
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(shape[1] + shape[3])
        conv = nn.Conv2d(shape[1], 2, 3, padding=1, bias=bias).cuda()
        x = torch.randn(shape, device="cuda").contiguous(memory_format=CL).requires_grad_(True)
        params = [x, conv.weight] + ([conv.bias] if bias else [])
        ref = conv(x)
        w = torch.randn_like(ref)
        rg = torch.autograd.grad((ref * w).sum(), params)
        out = conv_plain(conv, x)
        assert out.is_contiguous() and isinstance(out.grad_fn, _FlowOutConv._backward_cls)
        og = torch.autograd.grad((out * w).sum(), params)
        assert_close(out, ref, 1e-5, "flow head")
        for a, b, name in zip(og, rg, ("dx", "dw", "db")):
            assert_close(a, b, 1e-4, name)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_image_pair_pack_is_cat_scale_shift_and_zero_pad():
    """arf_image_pair_pack == torch.cat([x[:, :3], x[:, 3:]], 0) * 2 - 1 as 8-channel channels-last, bit for bit."""
    from arflow_b200.fused_conv import image_pair_nhwc, is_nhwc
    torch.manual_seed(11)
    x = torch.rand(3, 6, 10, 13, device="cuda")
    out = image_pair_nhwc(x)
    ref = torch.cat([x[:, :3], x[:, 3:]], 0) * 2. - 1.
    assert out.shape == (6, 8, 10, 13) and is_nhwc(out)
    assert torch.equal(out[:, :3], ref) and float(out[:, 3:].abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(2, 3, 20, 28), (1, 3, 15, 21), (3, 3, 8, 70)])
def test_first_pyramid_conv_weight_gradient(shape):
    """Conv2d(3, 32, 3, stride 2, padding 1) + leaky ReLU on the zero-padded 8-channel image
    (models/uflow_model.py:427-436): weight / bias gradients of arf_conv3x3s2_first_wgrad + the fused epilogue against
    torch on the 3-channel image; even and odd sizes, rows longer than one 32-pixel run."""
    from arflow_b200.fused_conv import conv_bias_leaky, nhwc_concat, pad_in_channels
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(shape[2])
        conv = nn.Conv2d(3, 32, 3, stride=2, padding=1).cuda()
        x = torch.randn(shape, device="cuda")
        ref = func.leaky_relu(conv(x), negative_slope=0.1)
        w = torch.randn_like(ref)
        rg = torch.autograd.grad((ref * w).sum(), [conv.weight, conv.bias])
        x8, _ = nhwc_concat([x])
        assert x8.shape[1] == 8
        out = conv_bias_leaky(conv, x8, 0.1, weight=pad_in_channels(conv.weight, 3, 5), real_in=3)
        og = torch.autograd.grad((out * w).sum(), [conv.weight, conv.bias])
        assert_close(out, ref, 1e-5, "first conv")
        assert_close(og[0], rg[0], 1e-4, "dw")
        assert_close(og[1], rg[1], 1e-4, "db")
    finally:
        torch.backends.cudnn.allow_tf32 = old

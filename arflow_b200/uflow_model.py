"""PWCFlow — the UFlow PWC network that calls the hot path (drop-in for models/uflow_model.py).

The module tree and parameter names are the reference's (`_refine_model`, `_flow_layers`,
`_context_up_layers`, `_feature_pyramid_extractor._convs`), so a reference state_dict loads by name.
Convolutions stay on cuDNN (out of scope); what changes is everything between them:
  * warp / cost volume / x2 upsampling go through the arflow_b200 kernels,
  * `forward` can run both flow directions as ONE pass over a 2B batch (`stack_directions=True`):
    the two `forward_2_frames` calls of the reference (uflow_model.py:255-257) are independent and
    every op is per-sample, so the result is identical while every kernel sees twice the work,
  * level dropout is drawn on the device (no `.item()` host sync, uflow_model.py:211-214), one
    Bernoulli per direction and level exactly like the reference's two sequential calls, which keeps
    the step capturable in a CUDA graph.
"""
import torch
import torch.nn as nn
import torch.nn.functional as func

from .fused_conv import (CL, _zeros_cl, conv_transpose_bias, image_pair_nhwc, conv_bias_leaky, conv_plain, dense_block_nhwc, nhwc_concat, out_channel_pad,
                         pad_in_channels, pad_weight, to_nchw)


class _CudaOps:
    """The hot-path ops the network needs; the oracle provides a CPU twin for the baseline leg."""

    def __init__(self):
        from . import correlation, uflow_utils
        self.flow_to_warp = uflow_utils.flow_to_warp
        self.resample = uflow_utils.resample
        self.resample_flow = uflow_utils.resample_flow      # resample(x, flow_to_warp(flow)) in one call
        self.upsample = uflow_utils.upsample
        self.compute_cost_volume = correlation.compute_cost_volume


class _FeatNormFunction(torch.autograd.Function):
    """normalize_features for two maps with all four switches on (arf_featnorm_fwd / _bwd, csrc/normalize.cu)."""

    @staticmethod
    def forward(ctx, f1, f2):
        from . import _lib
        f1, f2 = f1.contiguous(), f2.contiguous()
        B, n = f1.shape[0], f1[0].numel()
        lib = _lib.load()
        with torch.cuda.device_of(f1):
            y1, y2 = torch.empty_like(f1), torch.empty_like(f2)
            stats = torch.empty(B * 4, dtype=f1.dtype, device=f1.device)
            ws = torch.empty(lib.arf_featnorm_workspace(B, n) // 8, dtype=torch.float64, device=f1.device)
            _lib.call("arf_featnorm_fwd", _lib.dev_ptr(f1, "features1"), _lib.dev_ptr(f2, "features2"), _lib.dev_ptr(y1),
                      _lib.dev_ptr(y2), _lib.dev_ptr(stats), ws.data_ptr(), B, n, _lib.stream_ptr())
        ctx.save_for_backward(f1, f2, stats)
        return y1, y2

    @staticmethod
    def backward(ctx, g1, g2):
        from . import _lib
        f1, f2, stats = ctx.saved_tensors
        B, n = f1.shape[0], f1[0].numel()
        g1 = torch.zeros_like(f1) if g1 is None else g1.contiguous()
        g2 = torch.zeros_like(f2) if g2 is None else g2.contiguous()
        lib = _lib.load()
        with torch.cuda.device_of(f1):
            d1 = torch.empty_like(f1) if ctx.needs_input_grad[0] else None
            d2 = torch.empty_like(f2) if ctx.needs_input_grad[1] else None
            coef = torch.empty(B * 2, dtype=f1.dtype, device=f1.device)
            ws = torch.empty(lib.arf_featnorm_workspace(B, n) // 8, dtype=torch.float64, device=f1.device)
            _lib.call("arf_featnorm_bwd", _lib.dev_ptr(f1), _lib.dev_ptr(f2), _lib.dev_ptr(g1, "grad"), _lib.dev_ptr(g2, "grad"),
                      _lib.dev_ptr(stats), _lib.dev_ptr(d1, allow_none=True), _lib.dev_ptr(d2, allow_none=True),
                      _lib.dev_ptr(coef), ws.data_ptr(), B, n, _lib.stream_ptr())
        return d1, d2


def normalize_features(feature_list, normalize, center, moments_across_channels, moments_across_images):
    """uflow_model.py:8-50 — per-sample moments (unbiased variance), optionally shared by the images.
    The setting the networks use (two CUDA maps, all four switches on) runs as one fused reduction + one elementwise
    pass each way; every other combination (and the CPU twin) is the reference's chain of torch ops."""
    if (normalize and center and moments_across_channels and moments_across_images and len(feature_list) == 2
            and feature_list[0].is_cuda and feature_list[0].dtype == torch.float32
            and feature_list[0].shape == feature_list[1].shape):
        return list(_FeatNormFunction.apply(feature_list[0], feature_list[1]))
    dim = [1, 2, 3] if moments_across_channels else [2, 3]
    stats = [torch.var_mean(f, dim=dim, keepdim=True) for f in feature_list]
    variances = [s[0] for s in stats]
    means = [s[1] for s in stats]
    if moments_across_images:
        n = float(len(feature_list))
        mean_all = sum(means) / n
        var_all = sum(variances) / n
        means = [mean_all] * len(means)
        variances = [var_all] * len(variances)
    stds = [torch.sqrt(v + 1e-16) for v in variances]
    if center:
        feature_list = [f - m for f, m in zip(feature_list, means)]
    if normalize:
        feature_list = [f / s for f, s in zip(feature_list, stds)]
    return feature_list


def warp_features(ops, features, flow):
    """ops.resample(features, ops.flow_to_warp(flow)) (uflow_model.py:163-165); one fused call where the op set has it."""
    fused = getattr(ops, "resample_flow", None)
    return fused(features, flow) if fused is not None else ops.resample(features, ops.flow_to_warp(flow))


def compute_cost_volume(features1, features2, max_displacement):
    """uflow_model.py:53-92."""
    from .correlation import compute_cost_volume as _cv
    return _cv(features1, features2, max_displacement)


# ----------------------------------------------------------------------------- channels-last decoder pieces ----
# Shared by PWCFlow and PWCProbFlow (CUDA only).  Same operations in the same order as the reference's decoder; what
# changes is where the bytes live: dense-block inputs are built by `nhwc_concat` as packed NHWC tensors with 8-aligned
# channel counts (zero channels after the first concat of a level, zero weight columns to match), the convolutions
# run on cuDNN's NHWC kernels without layout conversions, and only what the NCHW hot-path kernels touch (features
# for warp / cost volume, the few output channels) is converted.
def _conv_leaky_padded(conv, x, alpha, in_pads):
    """conv + bias + leaky on a channels-last input that carries zero channels at `in_pads`; the output is widened
    with zero channels to a width cuDNN has a fast kernel for (out_channel_pad).  Returns (y, n_zero_out_channels)."""
    op = out_channel_pad(conv.out_channels)
    bias = conv.bias
    if op and bias is not None:
        bias = torch.cat([bias, _zeros_cl(bias, (op,))])
    return conv_bias_leaky(conv, x, alpha, weight=pad_weight(conv.weight, in_pads, op), bias=bias), op


def decoder_level_nhwc(layers, parts, alpha):
    """One pyramid level's dense block + output convolution (models/uflow_model.py:189-205).
    parts: tensors to concatenate (NCHW or channels-last).  Returns (context: channels-last, out: NCHW)."""
    x_in, real = nhwc_concat(parts)
    pads = [(real, x_in.shape[1] - real)]      # (position in the real channel order, zero channels carried there)
    convs = [layer[0] for layer in list(layers)[:-1]]      # layer = Sequential(Conv2d, LeakyReLU)
    weights, biases = [], []
    for i, conv in enumerate(convs):
        # the last layer's output is not concatenated again: keep its width (the reference's final cat is never read)
        op = out_channel_pad(conv.out_channels) if i + 1 < len(convs) else 0
        weights.append(pad_weight(conv.weight, pads, op))
        b = conv.bias
        biases.append(torch.cat([b, _zeros_cl(b, (op,))]) if (op and b is not None) else b)
        real += conv.out_channels
        if op:
            pads = pads + [(real, op)]
    context = dense_block_nhwc(x_in, convs, weights, biases, alpha)
    last = layers[-1]
    out = conv_plain(last, context, weight=last.weight.contiguous(memory_format=CL)).contiguous()
    return context, out


def refine_nhwc(refine_model, context, out, alpha):
    """The dilated refinement stack on cat([context, out]) (models/uflow_model.py:212-216) -> NCHW."""
    x, c0 = nhwc_concat([context, out])
    pads = [(c0, x.shape[1] - c0)]
    refine = list(refine_model)              # conv, LeakyReLU, conv, LeakyReLU, ..., conv
    for conv in refine[:-1:2]:
        x, op = _conv_leaky_padded(conv, x, alpha, pads)
        pads = [(conv.out_channels, op)]
    return conv_plain(refine[-1], x, weight=pad_weight(refine[-1].weight, pads)).contiguous()


def context_up_nhwc(up, context):
    """ConvTranspose2d x2 of the context features, channels-last in and out."""
    return conv_transpose_bias(up, context)


class PWCFeaturePyramid(nn.Module):
    """uflow_model.py:364-470 — five levels of three 3x3 convolutions, the first of each with stride 2."""

    def __init__(self, leaky_relu_alpha=0.1, filters=None, level1_num_layers=3, level1_num_filters=32,
                 level1_num_1x1=0, original_layer_sizes=False, num_levels=5, channel_multiplier=1.,
                 pyramid_resolution='half', num_channels=3):
        super().__init__()
        self._channel_multiplier = channel_multiplier
        if num_levels > 6:
            raise NotImplementedError('Max number of pyramid levels is 6')
        if filters is None:
            if original_layer_sizes:
                filters = ((3, 16), (3, 32), (3, 64), (3, 96), (3, 128), (3, 196))[:num_levels]
            else:
                filters = ((level1_num_layers, level1_num_filters),) + ((3, 32),) * 5
                filters = filters[:num_levels]
        assert filters and all(len(t) == 2 and t[0] > 0 for t in filters)
        self._leaky_relu_alpha = leaky_relu_alpha
        self._level1_num_1x1 = level1_num_1x1
        self._convs = nn.ModuleList()
        c = num_channels
        for level, (num_layers, num_filters) in enumerate(filters):
            group = nn.ModuleList()
            for i in range(num_layers):
                stride = 2 if (i == 0 or (i == 1 and level == 0 and pyramid_resolution == 'quarter')) else 1
                is3 = level > 0 or i < num_layers - level1_num_1x1
                out_c = int(num_filters * channel_multiplier)
                # explicit zero pad + 'valid' conv of the reference == padding=1 for the 3x3 layers
                group.append(nn.Conv2d(c, out_c, kernel_size=(3, 3) if is3 else (1, 1), stride=stride,
                                       padding=1 if is3 else 0))
                c = out_c
            self._convs.append(group)

    def forward(self, x, split_features_by_sample=False, nhwc=False, prepacked=None):
        """nhwc (CUDA only): activations and per-step weight copies are channels-last, so cuDNN's NHWC kernels run
        without layout conversions; the returned features are channels-last tensors.
        prepacked: (tensor, c_img) - the normalised input already as zero-padded channels-last (image_pair_nhwc)."""
        n_pad, c_img = 0, None
        if prepacked is not None:
            x, c_img = prepacked
            nhwc = True
            n_pad = x.shape[1] - c_img
        else:
            x = x * 2. - 1.
            nhwc = nhwc and x.is_cuda
            if nhwc:
                # 3 image channels -> 8 (zeros): cuDNN's aligned NHWC kernels instead of its 3-channel fallback
                c_img = x.shape[1]
                x, _ = nhwc_concat([x])
                n_pad = x.shape[1] - c_img
        features = []
        first = True
        for group in self._convs:
            for conv in group:
                w = None
                if nhwc:
                    w = pad_in_channels(conv.weight, conv.weight.shape[1], n_pad if first else 0)
                x = conv_bias_leaky(conv, x, self._leaky_relu_alpha, weight=w,
                                    real_in=c_img if (first and n_pad) else None)
                first = False
            features.append(x)
        if split_features_by_sample:
            n = len(features[0])
            features = [[f[i:i + 1] for f in features] for i in range(n)]
        return features


class PWCFlow(nn.Module):
    """uflow_model.py:96-362.  cfg needs `level_dropout` and `feature_norm`."""

    def __init__(self, cfg, ops=None, stack_directions=True, nhwc=True):
        super().__init__()
        self._ops = ops if ops is not None else _CudaOps()
        self._stack_directions = stack_directions
        self._nhwc = nhwc      # CUDA only: channels-last conv stacks (see fused_conv.py); results are unchanged
        self._leaky_relu_alpha = 0.1
        self._drop_out_rate = cfg.level_dropout
        self._num_context_up_channels = 32
        self._num_levels = 5
        self._normalize_before_cost_volume = cfg.feature_norm
        self._channel_multiplier = 1
        self._accumulate_flow = True

        self._refine_model = self._build_refinement_model()
        self._flow_layers = self._build_flow_layers()
        self._context_up_layers = nn.ModuleList(
            [nn.ConvTranspose2d(self._num_context_up_channels, self._num_context_up_channels, kernel_size=(4, 4),
                                stride=2, padding=1) for _ in range(self._num_levels)])
        self._feature_pyramid_extractor = PWCFeaturePyramid()

    # ------------------------------------------------------------------ construction
    def _build_flow_layers(self):
        result = nn.ModuleList([None])  # no flow is estimated at level 0
        block_layers = [128, 128, 96, 64, 32]
        for i in range(1, self._num_levels):
            layers = nn.ModuleList()
            c_in = 81 + 32
            if i != self._num_levels - 1:
                c_in += 2 + self._num_context_up_channels
            for c in block_layers:
                layers.append(nn.Sequential(nn.Conv2d(c_in, c, kernel_size=(3, 3), stride=1, padding='same'),
                                            nn.LeakyReLU(negative_slope=self._leaky_relu_alpha)))
                c_in += c
            layers.append(nn.Conv2d(block_layers[-1], 2, kernel_size=(3, 3), padding='same'))
            result.append(layers)
        return result

    def _build_refinement_model(self):
        layers = []
        c_in = 32 + 2
        for c, d in [(128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1)]:
            layers.append(nn.Conv2d(c_in, c, kernel_size=(3, 3), stride=1, padding='same', dilation=d))
            layers.append(nn.LeakyReLU(negative_slope=self._leaky_relu_alpha))
            c_in = c
        layers.append(nn.Conv2d(c_in, 2, kernel_size=(3, 3), stride=1, padding='same'))
        return nn.ModuleList(layers)

    def init_weights(self, xavier=False):
        """uflow_model.py:124-136.  The reference iterates `self.named_modules()` — (name, module) tuples —
        so its isinstance checks never match and the call leaves PyTorch's default (Kaiming-uniform)
        initialisation in place.  That observable behaviour is kept; `xavier=True` applies what the
        reference's code evidently intended (Xavier-uniform weights, zero biases)."""
        if not xavier:
            return
        for layer in self.modules():
            if isinstance(layer, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.xavier_uniform_(layer.weight)
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, 0)

    # ------------------------------------------------------------------ forward
    def _keep(self, like, groups):
        """Level-dropout multiplier: one Bernoulli per direction (`groups`), broadcast over its samples."""
        if not (self.training and self._drop_out_rate > 0):
            return None
        keep = (torch.rand(groups, device=like.device) > self._drop_out_rate).to(like.dtype)
        return keep.repeat_interleave(like.shape[0] // groups).view(-1, 1, 1, 1)

    def _forward_2_frames_nhwc(self, feature_pyramid1, feature_pyramid2, groups=1):
        """forward_2_frames with channels-last conv stacks (see decoder_level_nhwc)."""
        ops = self._ops
        alpha = self._leaky_relu_alpha
        context = flow = flow_up = context_up = None
        flows = []
        for level in range(self._num_levels - 1, 0, -1):
            features1 = feature_pyramid1[level]                                         # channels-last
            f1 = to_nchw(features1)
            if feature_pyramid2 is None:     # stacked directions: the other half of the same batch
                f2 = to_nchw(features1, batch_shift=features1.shape[0] // 2)
            else:
                f2 = to_nchw(feature_pyramid2[level])
            warped2 = f2 if flow_up is None else warp_features(ops, f2, flow_up)
            f1n, w2n = normalize_features([f1, warped2], normalize=self._normalize_before_cost_volume,
                                          center=self._normalize_before_cost_volume, moments_across_channels=True,
                                          moments_across_images=True)
            # the cost volume's leaky ReLU (uflow_model.py:185-186) rides on its NCHW -> NHWC pack
            cost_volume = (ops.compute_cost_volume(f1n, w2n, max_displacement=4), alpha)
            if flow_up is None:
                parts = [cost_volume, features1]
            elif context_up is None:
                parts = [flow_up, cost_volume, features1]
            else:
                parts = [context_up, flow_up, cost_volume, features1]
            context, flow = decoder_level_nhwc(self._flow_layers[level], parts, alpha)

            keep = self._keep(flow, groups)
            if keep is not None:
                context = context * keep
                flow = flow * keep
            if flow_up is not None and self._accumulate_flow:
                flow = flow + flow_up
            flow_up = ops.upsample(flow, is_flow=True)
            context_up = context_up_nhwc(self._context_up_layers[level], context)
            flows.insert(0, flow)

        refinement = refine_nhwc(self._refine_model, context, flow, alpha)
        keep = self._keep(refinement, groups)
        if keep is not None:
            refinement = refinement * keep
        flows[0] = flow + refinement
        flows.insert(0, ops.upsample(flows[0], is_flow=True))
        flows.insert(0, ops.upsample(flows[0], is_flow=True))
        return flows

    def forward_2_frames(self, feature_pyramid1, feature_pyramid2, groups=1):
        if self._nhwc and feature_pyramid1[-1].is_cuda:
            return self._forward_2_frames_nhwc(feature_pyramid1, feature_pyramid2, groups)
        ops = self._ops
        context = flow = flow_up = context_up = None
        flows = []
        for level in range(self._num_levels - 1, 0, -1):
            features1, features2 = feature_pyramid1[level], feature_pyramid2[level]
            if flow_up is None:
                warped2 = features2
            else:
                warped2 = warp_features(ops, features2, flow_up)
            f1n, w2n = normalize_features([features1, warped2], normalize=self._normalize_before_cost_volume,
                                          center=self._normalize_before_cost_volume, moments_across_channels=True,
                                          moments_across_images=True)
            cost_volume = func.leaky_relu(ops.compute_cost_volume(f1n, w2n, max_displacement=4),
                                          negative_slope=self._leaky_relu_alpha)
            if flow_up is None:
                x_in = torch.cat([cost_volume, features1], dim=1)
            elif context_up is None:
                x_in = torch.cat([flow_up, cost_volume, features1], dim=1)
            else:
                x_in = torch.cat([context_up, flow_up, cost_volume, features1], dim=1)
            flow_layers = self._flow_layers[level]
            x_out = None
            dense = list(flow_layers)[:-1]
            for i, layer in enumerate(dense):
                x_out = conv_bias_leaky(layer[0], x_in, self._leaky_relu_alpha)   # layer = Sequential(Conv2d, LeakyReLU)
                if i + 1 < len(dense):   # the reference also concatenates after the last layer; that tensor is never read
                    x_in = torch.cat([x_in, x_out], dim=1)
            context = x_out
            flow = flow_layers[-1](context)

            keep = self._keep(flow, groups)
            if keep is not None:
                context = context * keep
                flow = flow * keep
            if flow_up is not None and self._accumulate_flow:
                flow = flow + flow_up
            flow_up = ops.upsample(flow, is_flow=True)
            context_up = self._context_up_layers[level](context)
            flows.insert(0, flow)

        refinement = torch.cat([context, flow], dim=1)
        refine = list(self._refine_model)          # conv, LeakyReLU, conv, LeakyReLU, ..., conv
        for conv in refine[:-1:2]:
            refinement = conv_bias_leaky(conv, refinement, self._leaky_relu_alpha)
        refinement = refine[-1](refinement)
        keep = self._keep(refinement, groups)
        if keep is not None:
            refinement = refinement * keep
        flows[0] = flow + refinement
        flows.insert(0, ops.upsample(flows[0], is_flow=True))
        flows.insert(0, ops.upsample(flows[0], is_flow=True))
        return flows

    def forward(self, x, with_bk=True):
        n_frames = x.size(1) // 3
        if n_frames != 2:
            raise NotImplementedError
        B = x.shape[0]
        res_dict = {}
        if with_bk and self._stack_directions:
            # one pyramid pass over [img1; img2], one decoder pass over [(1,2); (2,1)]
            if self._nhwc and x.is_cuda and x.dtype == torch.float32 and not x.requires_grad:
                # cat of the two images, x * 2 - 1 and the 8-channel NHWC pack in one pass
                feats = self._feature_pyramid_extractor(None, prepacked=(image_pair_nhwc(x), 3))
            else:
                feats = self._feature_pyramid_extractor(torch.cat([x[:, 0:3], x[:, 3:6]], dim=0), nhwc=self._nhwc)
            p1 = feats
            if self._nhwc and x.is_cuda:
                p2 = None       # the half-batch swap rides on the NHWC -> NCHW copy of the features (to_nchw)
            else:
                # level 0 is never read by the decoder (uflow_model.py:158): do not copy its features
                p2 = [None] + [torch.cat([f[B:], f[:B]], dim=0) for f in feats[1:]]
            flows = self.forward_2_frames(p1, p2, groups=2)
            res_dict['flows_fw'] = [f[:B] for f in flows]
            res_dict['flows_bw'] = [f[B:] for f in flows]
            return res_dict
        pyr = [self._feature_pyramid_extractor(x[:, 3 * i: 3 * i + 3], nhwc=self._nhwc) for i in range(2)]
        res_dict['flows_fw'] = self.forward_2_frames(pyr[0], pyr[1])
        if with_bk:
            res_dict['flows_bw'] = self.forward_2_frames(pyr[1], pyr[0])
        return res_dict

"""PWCLite — the ARFlow PWC-Lite network that calls the hot path (drop-in for models/pwclite.py:109-283).

Config 1 of BASELINE.json (two-view inference 384x640) and the C >= 64 regime of the correlation / warp kernels
live here: per level l = 0..4 (1/64 .. 1/4 resolution, C = 192, 128, 96, 64, 32) the reference up-samples the flow
(x2, align_corners=True, pwclite.py:178-179), warps the second feature map (`flow_warp`, :180), correlates
(`Correlation(pad_size=4, kernel_size=1, max_displacement=4, ...)`, :124-126, :183) and feeds shared estimator /
context networks.  The module tree and parameter names are the reference's (`feature_pyramid_extractor.convs`,
`flow_estimators.conv1..`, `context_networks.convs`, `conv_1x1`), so a reference state_dict loads by name and the
construction order (hence a seeded default initialisation) is the same.  Convolutions stay on cuDNN;
what changes is everything between them:
  * flow up-sampling, warp and cost volume go through the arflow_b200 kernels (no `mesh_grid` built on the host and
    copied per call, warp_utils.py:7-13, 85-86),
  * with `with_bk` both directions run as ONE pass over a 2B batch (the two `forward_2_frames` calls of
    pwclite.py:268-270 are independent and every op is per-sample).
CUDA only, like every arflow_b200 op; the CPU restatement used by tests and the CPU baseline is oracle/cpu_nets.py.
"""
import torch
import torch.nn as nn
import torch.nn.functional as func

from .correlation import Correlation
from .uflow_utils import interpolate_align_corners
from .warp_utils import flow_warp

_ALPHA = 0.1


def conv(in_planes, out_planes, kernel_size=3, stride=1, dilation=1, isReLU=True):
    """pwclite.py:10-23 — Conv2d ('same'-style padding for the dilation) [+ LeakyReLU(0.1)] as a Sequential, so the
    parameters are `<name>.0.weight` / `<name>.0.bias` as in the reference."""
    layers = [nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, dilation=dilation,
                        padding=((kernel_size - 1) * dilation) // 2, bias=True)]
    if isReLU:
        layers.append(nn.LeakyReLU(_ALPHA, inplace=True))
    return nn.Sequential(*layers)


class FeatureExtractor(nn.Module):
    """pwclite.py:26-45 — one (stride-2 conv, conv) pair per pyramid level; coarsest level first in the output."""

    def __init__(self, num_chs):
        super().__init__()
        self.num_chs = num_chs
        self.convs = nn.ModuleList(nn.Sequential(conv(c_in, c_out, stride=2), conv(c_out, c_out))
                                   for c_in, c_out in zip(num_chs[:-1], num_chs[1:]))

    def forward(self, x):
        pyramid = []
        for level in self.convs:
            x = level(x)
            pyramid.append(x)
        return pyramid[::-1]


class FlowEstimatorDense(nn.Module):
    """pwclite.py:48-66 — DenseNet-style estimator (every layer sees all earlier outputs)."""

    _WIDTHS = (128, 128, 96, 64, 32)

    def __init__(self, ch_in):
        super().__init__()
        c = ch_in
        for i, w in enumerate(self._WIDTHS, 1):
            setattr(self, "conv%d" % i, conv(c, w))
            c += w
        self.feat_dim = c
        self.conv_last = conv(c, 2, isReLU=False)

    def forward(self, x):
        for i in range(1, len(self._WIDTHS) + 1):
            x = torch.cat([getattr(self, "conv%d" % i)(x), x], dim=1)
        return x, self.conv_last(x)


class FlowEstimatorReduce(nn.Module):
    """pwclite.py:69-88 — the reduced estimator (each layer sees only the two previous outputs)."""

    def __init__(self, ch_in):
        super().__init__()
        self.conv1 = conv(ch_in, 128)
        self.conv2 = conv(128, 128)
        self.conv3 = conv(128 + 128, 96)
        self.conv4 = conv(128 + 96, 64)
        self.conv5 = conv(96 + 64, 32)
        self.feat_dim = 32
        self.predict_flow = conv(64 + 32, 2, isReLU=False)

    def forward(self, x):
        prev = self.conv1(x)
        cur = self.conv2(prev)
        for layer in (self.conv3, self.conv4, self.conv5):
            prev, cur = cur, layer(torch.cat([prev, cur], dim=1))
        return cur, self.predict_flow(torch.cat([prev, cur], dim=1))


class ContextNetwork(nn.Module):
    """pwclite.py:91-106 — seven dilated 3x3 convolutions."""

    def __init__(self, ch_in):
        super().__init__()
        spec = [(128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1)]
        layers, c = [], ch_in
        for width, dilation in spec:
            layers.append(conv(c, width, 3, 1, dilation))
            c = width
        layers.append(conv(c, 2, isReLU=False))
        self.convs = nn.Sequential(*layers)

    def forward(self, x):
        return self.convs(x)


class PWCLite(nn.Module):
    """pwclite.py:109-283.  cfg needs `upsample`, `n_frames`, `reduce_dense`."""

    def __init__(self, cfg, stack_directions=True):
        super().__init__()
        self.search_range = 4
        self.num_chs = [3, 16, 32, 64, 96, 128, 192]
        self.output_level = 4
        self.num_levels = 7
        self.leakyRELU = nn.LeakyReLU(_ALPHA, inplace=True)
        self._stack_directions = stack_directions

        self.feature_pyramid_extractor = FeatureExtractor(self.num_chs)
        self.upsample = cfg.upsample
        self.n_frames = cfg.n_frames
        self.reduce_dense = cfg.reduce_dense
        self.corr = Correlation(pad_size=self.search_range, kernel_size=1, max_displacement=self.search_range,
                                stride1=1, stride2=1, corr_multiply=1)
        self.dim_corr = (self.search_range * 2 + 1) ** 2
        self.num_ch_in = 32 + (self.dim_corr + 2) * (self.n_frames - 1)
        estimator = FlowEstimatorReduce if self.reduce_dense else FlowEstimatorDense
        self.flow_estimators = estimator(self.num_ch_in)
        self.context_networks = ContextNetwork((self.flow_estimators.feat_dim + 2) * (self.n_frames - 1))
        self.conv_1x1 = nn.ModuleList(conv(c, 32, kernel_size=1) for c in self.num_chs[:1:-1])

    def num_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def init_weights(self, kaiming=False):
        """pwclite.py:147-159.  The reference iterates `self.named_modules()` — (name, module) tuples — so its
        isinstance checks never match and PyTorch's default initialisation stays.  That observable behaviour is kept;
        `kaiming=True` applies what the code evidently intended."""
        if not kaiming:
            return
        for layer in self.modules():
            if isinstance(layer, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(layer.weight)
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, 0)

    # ------------------------------------------------------------------ decoders
    def _up2(self, flow):
        # F.interpolate(flow * 2, scale_factor=2, mode='bilinear', align_corners=True): the x2 rides on the resize
        return interpolate_align_corners(flow, 2, mul=2.0)

    def forward_2_frames(self, x1_pyramid, x2_pyramid):
        """pwclite.py:161-204."""
        flows = []
        b, _, h, w = x1_pyramid[0].shape
        flow = torch.zeros(b, 2, h, w, dtype=torch.float32, device=x1_pyramid[0].device)
        for level, (x1, x2) in enumerate(zip(x1_pyramid, x2_pyramid)):
            if level == 0:
                x2_warp = x2
            else:
                flow = self._up2(flow)
                x2_warp = flow_warp(x2, flow)
            cost = func.leaky_relu(self.corr(x1, x2_warp), _ALPHA)
            x_intm, flow_res = self.flow_estimators(torch.cat([cost, self.conv_1x1[level](x1), flow], dim=1))
            flow = flow + flow_res
            flow = flow + self.context_networks(torch.cat([x_intm, flow], dim=1))
            flows.append(flow)
            if level == self.output_level:
                break
        if self.upsample:
            # the reference appends the x4 up-sampled finest flow (the per-level variant is commented out, :199-203)
            flows.append(interpolate_align_corners(flow, 4, mul=4.0))
        return flows[::-1]

    def forward_3_frames(self, x0_pyramid, x1_pyramid, x2_pyramid):
        """pwclite.py:206-258 — the middle frame against its two neighbours, flow = [to frame 0 | to frame 2]."""
        flows = []
        b, _, h, w = x1_pyramid[0].shape
        flow = torch.zeros(b, 4, h, w, dtype=torch.float32, device=x1_pyramid[0].device)
        for level, (x0, x1, x2) in enumerate(zip(x0_pyramid, x1_pyramid, x2_pyramid)):
            if level == 0:
                x0_warp, x2_warp = x0, x2
            else:
                flow = self._up2(flow)
                x0_warp = flow_warp(x0, flow[:, :2].contiguous())
                x2_warp = flow_warp(x2, flow[:, 2:].contiguous())
            c10 = func.leaky_relu(self.corr(x1, x0_warp), _ALPHA)
            c12 = func.leaky_relu(self.corr(x1, x2_warp), _ALPHA)
            x1_1by1 = self.conv_1x1[level](x1)
            f10, f12 = flow[:, :2], flow[:, 2:]
            i10, r10 = self.flow_estimators(torch.cat([x1_1by1, c10, c12, f10, -f12], dim=1))
            i12, r12 = self.flow_estimators(torch.cat([x1_1by1, c12, c10, f12, -f10], dim=1))
            flow = flow + torch.cat([r10, r12], dim=1)
            f10, f12 = flow[:, :2], flow[:, 2:]
            r10 = self.context_networks(torch.cat([i10, i12, f10, -f12], dim=1))
            r12 = self.context_networks(torch.cat([i12, i10, f12, -f10], dim=1))
            flow = flow + torch.cat([r10, r12], dim=1)
            flows.append(flow)
            if level == self.output_level:
                break
        if self.upsample:
            flows = [interpolate_align_corners(f, 4, mul=4.0) for f in flows]
        return [f[:, :2] for f in flows[::-1]], [f[:, 2:] for f in flows[::-1]]

    def forward(self, x, with_bk=False):
        """pwclite.py:260-283."""
        if not x.is_cuda:
            raise RuntimeError("arflow_b200.PWCLite: input must be a CUDA tensor (no CPU fallback exists)")
        n_frames = x.size(1) // 3
        if x.size(1) != 3 * n_frames or n_frames not in (2, 3, 5):
            raise NotImplementedError
        imgs = [x[:, 3 * i: 3 * i + 3] for i in range(n_frames)]
        B = x.shape[0]
        res = {}
        if n_frames == 2 and with_bk and self._stack_directions:
            # one pyramid pass over [img1; img2], one decoder pass over [(1,2); (2,1)]
            feats = self.feature_pyramid_extractor(torch.cat(imgs, dim=0))
            p1 = feats + [torch.cat(imgs, dim=0)]
            p2 = [torch.cat([f[B:], f[:B]], dim=0) for f in p1]
            flows = self.forward_2_frames(p1, p2)
            res['flows_fw'] = [f[:B] for f in flows]
            res['flows_bw'] = [f[B:] for f in flows]
            return res
        pyr = [self.feature_pyramid_extractor(img) + [img] for img in imgs]
        if n_frames == 2:
            res['flows_fw'] = self.forward_2_frames(pyr[0], pyr[1])
            if with_bk:
                res['flows_bw'] = self.forward_2_frames(pyr[1], pyr[0])
        elif n_frames == 3:
            flows_10, flows_12 = self.forward_3_frames(pyr[0], pyr[1], pyr[2])
            res['flows_fw'], res['flows_bw'] = flows_12, flows_10
        else:
            flows_10, flows_12 = self.forward_3_frames(pyr[0], pyr[1], pyr[2])
            flows_21, flows_23 = self.forward_3_frames(pyr[1], pyr[2], pyr[3])
            res['flows_fw'] = [flows_12, flows_23]
            if with_bk:
                flows_32, flows_34 = self.forward_3_frames(pyr[2], pyr[3], pyr[4])
                res['flows_bw'] = [flows_21, flows_32]
        return res

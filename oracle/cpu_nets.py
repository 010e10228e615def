"""CPU restatements of the reference's networks (TEST INFRASTRUCTURE and CPU baseline only).

Nothing under arflow_b200/ imports this file, and this file imports nothing from arflow_b200/: the module trees are
declared again here, with the reference's parameter names, so that a state_dict of the product networks (or of the
reference's) loads by name and a seeded construction draws the same PyTorch default initialisation.  The hot-path
calls go to the oracle's plain-torch functions (oracle/arflow_oracle.py), the convolutions to torch's CPU kernels.

  PWCFlowCPU   models/uflow_model.py:96-470   (config 2 / 4: chairs_uflow, kitti_uflow)
  PWCLiteCPU   models/pwclite.py:109-283      (config 1: two-view inference; 3- and 5-frame variants)

Pinned by tests/golden/pwcflow_eval.npz and tests/golden/pwclite_eval.npz (outputs of the unmodified reference,
tests/golden/make_golden.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import arflow_oracle as orc


# ----------------------------------------------------------------------------- PWCFlow --------
def normalize_features(feature_list, normalize, center, moments_across_channels, moments_across_images):
    """uflow_model.py:8-50."""
    dim = [1, 2, 3] if moments_across_channels else [2, 3]
    means = [f.mean(dim=dim, keepdim=True) for f in feature_list]
    variances = [f.var(dim=dim, keepdim=True) for f in feature_list]      # unbiased, like torch.var there
    if moments_across_images:
        n = float(len(feature_list))
        means = [sum(means) / n] * len(means)
        variances = [sum(variances) / n] * len(variances)
    stds = [torch.sqrt(v + 1e-16) for v in variances]
    if center:
        feature_list = [f - m for f, m in zip(feature_list, means)]
    if normalize:
        feature_list = [f / s for f, s in zip(feature_list, stds)]
    return feature_list


class _PyramidCPU(nn.Module):
    """uflow_model.py:364-470 with its defaults: 5 levels x 3 convolutions of 32 channels, first of a level stride 2.
    (The reference zero-pads explicitly and convolves 'valid'; padding=1 is the same arithmetic.)"""

    def __init__(self):
        super().__init__()
        self._convs = nn.ModuleList()
        c = 3
        for _ in range(5):
            group = nn.ModuleList()
            for i in range(3):
                group.append(nn.Conv2d(c, 32, kernel_size=(3, 3), stride=2 if i == 0 else 1, padding=1))
                c = 32
            self._convs.append(group)

    def forward(self, x):
        x = x * 2. - 1.
        features = []
        for group in self._convs:
            for conv in group:
                x = F.leaky_relu(conv(x), 0.1)
            features.append(x)
        return features


class PWCFlowCPU(nn.Module):
    def __init__(self, level_dropout=0.1, feature_norm=True):
        super().__init__()
        self._drop_out_rate = level_dropout
        self._feature_norm = feature_norm
        # construction order of the reference (:111-123): refinement, flow layers, context up-sampling, pyramid
        layers, c_in = [], 34
        for c, d in [(128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1)]:
            layers += [nn.Conv2d(c_in, c, kernel_size=(3, 3), stride=1, padding='same', dilation=d), nn.LeakyReLU(0.1)]
            c_in = c
        layers.append(nn.Conv2d(c_in, 2, kernel_size=(3, 3), stride=1, padding='same'))
        self._refine_model = nn.ModuleList(layers)
        self._flow_layers = nn.ModuleList([None])
        for level in range(1, 5):
            block = nn.ModuleList()
            c_in = 81 + 32 + (0 if level == 4 else 2 + 32)
            for c in (128, 128, 96, 64, 32):
                block.append(nn.Sequential(nn.Conv2d(c_in, c, kernel_size=(3, 3), stride=1, padding='same'), nn.LeakyReLU(0.1)))
                c_in += c
            block.append(nn.Conv2d(32, 2, kernel_size=(3, 3), padding='same'))
            self._flow_layers.append(block)
        self._context_up_layers = nn.ModuleList(
            [nn.ConvTranspose2d(32, 32, kernel_size=(4, 4), stride=2, padding=1) for _ in range(5)])
        self._feature_pyramid_extractor = _PyramidCPU()

    def _dropout(self):
        if self.training and self._drop_out_rate > 0:
            return (torch.rand(1) > self._drop_out_rate).float().item()     # one host draw per level (:211-214)
        return None

    def forward_2_frames(self, pyr1, pyr2):
        """uflow_model.py:138-245."""
        context = flow = flow_up = context_up = None
        flows = []
        for level in range(4, 0, -1):
            f1, f2 = pyr1[level], pyr2[level]
            warped2 = f2 if flow_up is None else orc.warp(f2, orc.flow_to_warp(flow_up), kind="coords")
            f1n, w2n = normalize_features([f1, warped2], self._feature_norm, self._feature_norm, True, True)
            cost = F.leaky_relu(orc.cost_volume(f1n, w2n, 4), 0.1)
            if flow_up is None:
                x_in = torch.cat([cost, f1], dim=1)
            elif context_up is None:
                x_in = torch.cat([flow_up, cost, f1], dim=1)
            else:
                x_in = torch.cat([context_up, flow_up, cost, f1], dim=1)
            block = self._flow_layers[level]
            x_out = None
            for layer in list(block)[:-1]:
                x_out = layer(x_in)
                x_in = torch.cat([x_in, x_out], dim=1)
            context = x_out
            flow = block[-1](context)
            keep = self._dropout()
            if keep is not None:
                context, flow = context * keep, flow * keep
            if flow_up is not None:
                flow = flow + flow_up
            flow_up = orc.resize_bilinear(flow, 2.0, True)
            context_up = self._context_up_layers[level](context)
            flows.insert(0, flow)
        x = torch.cat([context, flow], dim=1)
        for layer in self._refine_model:
            x = layer(x)
        keep = self._dropout()
        if keep is not None:
            x = x * keep
        flows[0] = flow + x
        flows.insert(0, orc.resize_bilinear(flows[0], 2.0, True))
        flows.insert(0, orc.resize_bilinear(flows[0], 2.0, True))
        return flows

    def forward(self, x, with_bk=True):
        pyr = [self._feature_pyramid_extractor(x[:, 3 * i:3 * i + 3]) for i in range(2)]
        res = {'flows_fw': self.forward_2_frames(pyr[0], pyr[1])}
        if with_bk:
            res['flows_bw'] = self.forward_2_frames(pyr[1], pyr[0])
        return res


# ----------------------------------------------------------------------------- PWCLite --------
def _conv(c_in, c_out, k=3, stride=1, dilation=1, relu=True):
    """pwclite.py:10-23."""
    mods = [nn.Conv2d(c_in, c_out, kernel_size=k, stride=stride, dilation=dilation, padding=((k - 1) * dilation) // 2)]
    if relu:
        mods.append(nn.LeakyReLU(0.1))
    return nn.Sequential(*mods)


class _ExtractorCPU(nn.Module):
    def __init__(self, chs):
        super().__init__()
        self.convs = nn.ModuleList()
        for a, b in zip(chs[:-1], chs[1:]):
            self.convs.append(nn.Sequential(_conv(a, b, stride=2), _conv(b, b)))


class _ReduceCPU(nn.Module):
    def __init__(self, ch_in):
        super().__init__()
        self.conv1, self.conv2 = _conv(ch_in, 128), _conv(128, 128)
        self.conv3, self.conv4, self.conv5 = _conv(256, 96), _conv(224, 64), _conv(160, 32)
        self.predict_flow = _conv(96, 2, relu=False)
        self.feat_dim = 32

    def forward(self, x):
        x1 = self.conv1(x)
        x2 = self.conv2(x1)
        x3 = self.conv3(torch.cat([x1, x2], 1))
        x4 = self.conv4(torch.cat([x2, x3], 1))
        x5 = self.conv5(torch.cat([x3, x4], 1))
        return x5, self.predict_flow(torch.cat([x4, x5], 1))


class _DenseCPU(nn.Module):
    def __init__(self, ch_in):
        super().__init__()
        self.conv1, self.conv2 = _conv(ch_in, 128), _conv(ch_in + 128, 128)
        self.conv3, self.conv4, self.conv5 = _conv(ch_in + 256, 96), _conv(ch_in + 352, 64), _conv(ch_in + 416, 32)
        self.feat_dim = ch_in + 448
        self.conv_last = _conv(ch_in + 448, 2, relu=False)

    def forward(self, x):
        for layer in (self.conv1, self.conv2, self.conv3, self.conv4, self.conv5):
            x = torch.cat([layer(x), x], 1)
        return x, self.conv_last(x)


class _ContextCPU(nn.Module):
    def __init__(self, ch_in):
        super().__init__()
        self.convs = nn.Sequential(_conv(ch_in, 128, 3, 1, 1), _conv(128, 128, 3, 1, 2), _conv(128, 128, 3, 1, 4),
                                   _conv(128, 96, 3, 1, 8), _conv(96, 64, 3, 1, 16), _conv(64, 32, 3, 1, 1),
                                   _conv(32, 2, relu=False))


def _up_ac(flow, s):
    """F.interpolate(flow * s, scale_factor=s, mode='bilinear', align_corners=True) (pwclite.py:178-179, 203)."""
    return F.interpolate(flow * s, scale_factor=s, mode='bilinear', align_corners=True)


class PWCLiteCPU(nn.Module):
    def __init__(self, upsample=True, n_frames=2, reduce_dense=True):
        super().__init__()
        chs = [3, 16, 32, 64, 96, 128, 192]
        self.upsample, self.n_frames = upsample, n_frames
        self.feature_pyramid_extractor = _ExtractorCPU(chs)
        ch_in = 32 + (81 + 2) * (n_frames - 1)
        self.flow_estimators = _ReduceCPU(ch_in) if reduce_dense else _DenseCPU(ch_in)
        self.context_networks = _ContextCPU((self.flow_estimators.feat_dim + 2) * (n_frames - 1))
        self.conv_1x1 = nn.ModuleList([_conv(c, 32, k=1) for c in (192, 128, 96, 64, 32)])

    def _pyramid(self, img):
        feats, x = [], img
        for level in self.feature_pyramid_extractor.convs:
            x = level(x)
            feats.append(x)
        return feats[::-1] + [img]

    def _corr(self, a, b):
        return F.leaky_relu(orc.cost_volume(a, b, 4), 0.1)      # correlation_native.py:13-23 + pwclite.py:184

    def forward_2_frames(self, p1, p2):
        flows = []
        b, _, h, w = p1[0].shape
        flow = torch.zeros(b, 2, h, w, dtype=p1[0].dtype)
        for l, (x1, x2) in enumerate(zip(p1, p2)):
            if l == 0:
                x2w = x2
            else:
                flow = _up_ac(flow, 2)
                x2w = orc.warp(x2, flow, kind="flow")            # flow_warp, warp_utils.py:83-90
            x_intm, res = self.flow_estimators(torch.cat([self._corr(x1, x2w), self.conv_1x1[l](x1), flow], 1))
            flow = flow + res
            flow = flow + self.context_networks.convs(torch.cat([x_intm, flow], 1))
            flows.append(flow)
            if l == 4:
                break
        if self.upsample:
            flows.append(_up_ac(flow, 4))
        return flows[::-1]

    def forward_3_frames(self, p0, p1, p2):
        flows = []
        b, _, h, w = p1[0].shape
        flow = torch.zeros(b, 4, h, w, dtype=p1[0].dtype)
        for l, (x0, x1, x2) in enumerate(zip(p0, p1, p2)):
            if l == 0:
                x0w, x2w = x0, x2
            else:
                flow = _up_ac(flow, 2)
                x0w = orc.warp(x0, flow[:, :2], kind="flow")
                x2w = orc.warp(x2, flow[:, 2:], kind="flow")
            c10, c12 = self._corr(x1, x0w), self._corr(x1, x2w)
            f = self.conv_1x1[l](x1)
            i10, r10 = self.flow_estimators(torch.cat([f, c10, c12, flow[:, :2], -flow[:, 2:]], 1))
            i12, r12 = self.flow_estimators(torch.cat([f, c12, c10, flow[:, 2:], -flow[:, :2]], 1))
            flow = flow + torch.cat([r10, r12], 1)
            r10 = self.context_networks.convs(torch.cat([i10, i12, flow[:, :2], -flow[:, 2:]], 1))
            r12 = self.context_networks.convs(torch.cat([i12, i10, flow[:, 2:], -flow[:, :2]], 1))
            flow = flow + torch.cat([r10, r12], 1)
            flows.append(flow)
            if l == 4:
                break
        if self.upsample:
            flows = [_up_ac(f, 4) for f in flows]
        return [f[:, :2] for f in flows[::-1]], [f[:, 2:] for f in flows[::-1]]

    def forward(self, x, with_bk=False):
        n = x.size(1) // 3
        pyr = [self._pyramid(x[:, 3 * i:3 * i + 3]) for i in range(n)]
        res = {}
        if n == 2:
            res['flows_fw'] = self.forward_2_frames(pyr[0], pyr[1])
            if with_bk:
                res['flows_bw'] = self.forward_2_frames(pyr[1], pyr[0])
        elif n == 3:
            f10, f12 = self.forward_3_frames(pyr[0], pyr[1], pyr[2])
            res['flows_fw'], res['flows_bw'] = f12, f10
        elif n == 5:
            f10, f12 = self.forward_3_frames(pyr[0], pyr[1], pyr[2])
            f21, f23 = self.forward_3_frames(pyr[1], pyr[2], pyr[3])
            res['flows_fw'] = [f12, f23]
            if with_bk:
                f32, f34 = self.forward_3_frames(pyr[2], pyr[3], pyr[4])
                res['flows_bw'] = [f21, f32]
        else:
            raise NotImplementedError
        return res

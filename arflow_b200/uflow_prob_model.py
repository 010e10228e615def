"""PWCProbFlow — the probabilistic UFlow PWC network of the ELBO configurations (drop-in for
models/uflow_prob_model.py:149-500 of deu439/ARFlow; config 3 of BASELINE.json).

Every pyramid level predicts `out_channels = [L, M, N]` channels: L flow-mean channels (propagated and used for
warping, one cost volume per flow pair), M log-diagonal channels (propagated, biased by +-log 2 on every
x2 upsampling) and N extra channels produced by the output level only (the off-diagonal stencil taps of the
non-diagonal covariance, `uflow_elbo_loss.py:209-222`).  Module tree and parameter names are the reference's
(`_refine_model`, `_flow_layers`, `_context_up_layers`, `_feature_pyramid_extractor.{k}._convs`), so a
reference state_dict loads by name.  Convolutions stay on cuDNN; warp, cost volume and resizes run on the
arflow_b200 kernels; both flow directions run as one pass over a 2B batch and level dropout is drawn on the
device (see uflow_model.PWCFlow).  `mixture_weights=True` (MixtureWeightsNet) is not implemented.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as func

from .fused_conv import conv_bias_leaky, to_nchw
from .uflow_model import (PWCFeaturePyramid, _CudaOps, context_up_nhwc, decoder_level_nhwc, normalize_features,
                          refine_nhwc, warp_features)


class PWCProbFlow(nn.Module):
    """uflow_prob_model.py:149-500.  cfg needs: out_channels, inv_cov, n_pyramids, mixture_weights,
    feature_norm, level_dropout."""

    def __init__(self, cfg, ops=None, stack_directions=True, nhwc=True):
        super().__init__()
        if getattr(cfg, "mixture_weights", False):
            raise NotImplementedError("PWCProbFlow: mixture_weights=True (MixtureWeightsNet) is not implemented")
        self.cfg = cfg
        self._ops = ops if ops is not None else _CudaOps()
        self._stack_directions = stack_directions
        self._nhwc = nhwc      # CUDA only: channels-last conv stacks (fused_conv.py); results are unchanged
        self._leaky_relu_alpha = 0.1
        self._drop_out_rate = cfg.level_dropout
        self._num_context_up_channels = 32
        self._num_levels = 5
        self._normalize_before_cost_volume = cfg.feature_norm
        self._out_channels = list(cfg.out_channels)
        self._diag_bias = -math.log(2) if cfg.inv_cov else math.log(2)
        self._inv_cov = cfg.inv_cov

        self._refine_model = self._build_refinement_model()
        self._flow_layers = self._build_flow_layers()
        self._context_up_layers = nn.ModuleList(
            [nn.ConvTranspose2d(32, 32, kernel_size=(4, 4), stride=2, padding=1) for _ in range(self._num_levels)])
        self._feature_pyramid_extractor = nn.ModuleList([PWCFeaturePyramid() for _ in range(cfg.n_pyramids)])

    # ------------------------------------------------------------------ construction
    def _build_flow_layers(self):
        """uflow_prob_model.py:432-466 — unlike PWCFlow, every level (the coarsest too) takes flow, log-diagonal
        and context inputs; only the output level (1) predicts all L+M+N channels."""
        result = nn.ModuleList([None])
        block_layers = [128, 128, 96, 64, 32]
        L, M, _ = self._out_channels
        for i in range(1, self._num_levels):
            layers = nn.ModuleList()
            c_in = (L // 2) * 81 + 32 + L + M + self._num_context_up_channels
            for c in block_layers:
                layers.append(nn.Sequential(nn.Conv2d(c_in, c, kernel_size=(3, 3), stride=1, padding='same'),
                                            nn.LeakyReLU(negative_slope=self._leaky_relu_alpha)))
                c_in += c
            layers.append(nn.Conv2d(block_layers[-1], sum(self._out_channels) if i == 1 else L + M,
                                    kernel_size=(3, 3), padding='same'))
            result.append(layers)
        return result

    def _build_refinement_model(self):
        layers = []
        c_in = 32 + sum(self._out_channels)
        for c, d in [(128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1)]:
            layers.append(nn.Conv2d(c_in, c, kernel_size=(3, 3), stride=1, padding='same', dilation=d))
            layers.append(nn.LeakyReLU(negative_slope=self._leaky_relu_alpha))
            c_in = c
        layers.append(nn.Conv2d(c_in, sum(self._out_channels), kernel_size=(3, 3), stride=1, padding='same'))
        return nn.ModuleList(layers)

    def init_weights(self, kaiming=False):
        """uflow_prob_model.py:209-222 iterates `named_modules()` tuples, so — like PWCFlow.init_weights — it never
        re-initialises anything; that behaviour is kept.  `kaiming=True` applies what the code evidently intended."""
        if not kaiming:
            return
        for layer in self.modules():
            if isinstance(layer, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(layer.weight.data, mode='fan_in')
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, 0)

    # ------------------------------------------------------------------ forward
    def flows_cat(self, input_list):
        """uflow_prob_model.py:188-207 — regroup the pyramids' outputs as [means | log-diagonals | rest]."""
        L, M, _ = self._out_channels
        out_list = []
        for level in range(len(input_list[0])):
            parts = [torch.cat([f[level][:, 0:L] for f in input_list], dim=1),
                     torch.cat([f[level][:, L:L + M] for f in input_list], dim=1)]
            if input_list[0][level].size(1) > L + M:
                parts.append(torch.cat([f[level][:, L + M:sum(self._out_channels)] for f in input_list], dim=1))
            out_list.append(torch.cat(parts, dim=1))
        return out_list

    def upsample_out(self, out):
        """uflow_prob_model.py:224-253 — x2: flow values scale with the size, log-diagonals get the bias first."""
        L, M, N = self._out_channels
        up = self._ops.upsample
        parts = []
        if L > 0:
            parts.append(up(out[:, 0:L], is_flow=True))
        if M > 0:
            parts.append(up(out[:, L:L + M] + self._diag_bias, is_flow=False))
        if out.size(1) > L + M and N > 0:
            parts.append(up(out[:, L + M:L + M + N], is_flow=False))
        return torch.cat(parts, dim=1)

    def _keep(self, like, groups):
        if not (self.training and self._drop_out_rate > 0):
            return None
        keep = (torch.rand(groups, device=like.device) > self._drop_out_rate).to(like.dtype)
        return keep.repeat_interleave(like.shape[0] // groups).view(-1, 1, 1, 1)

    def forward_2_frames(self, feature_pyramid1, feature_pyramid2, groups=1):
        ops = self._ops
        L, M, N = self._out_channels
        context = context_up = out_up = out = None
        outs = []
        for level in range(self._num_levels - 1, 0, -1):
            features1 = feature_pyramid1[level]
            features2 = feature_pyramid2[level] if feature_pyramid2 is not None else None
            if out_up is None:   # coarsest level: zero flow, constant log-diagonal, zero context (:265-277)
                b, _, h, w = features1.shape
                out_up = torch.cat([features1.new_zeros(b, L, h, w),
                                    features1.new_full((b, M, h, w), -(self._num_levels - 3) * self._diag_bias)], dim=1)
                context_up = features1.new_zeros(b, self._num_context_up_channels, h, w)

            nhwc = self._nhwc and features1.is_cuda
            f1 = to_nchw(features1) if nhwc else features1      # NCHW copies for the hot-path kernels
            if features2 is None:      # stacked directions (channels-last only): the other half of the same batch
                f2 = to_nchw(features1, batch_shift=features1.shape[0] // 2)
            else:
                f2 = to_nchw(features2) if nhwc else features2
            cost_volumes = []
            for k in range(L // 2):
                warped2 = warp_features(ops, f2, out_up[:, 2 * k:2 * k + 2])
                f1n, w2n = normalize_features([f1, warped2], normalize=self._normalize_before_cost_volume,
                                              center=self._normalize_before_cost_volume, moments_across_channels=True,
                                              moments_across_images=True)
                cv = ops.compute_cost_volume(f1n, w2n, max_displacement=4)
                # channels-last: the leaky ReLU rides on the NCHW -> NHWC pack of the concat (fused_conv.nhwc_concat)
                cost_volumes.append((cv, self._leaky_relu_alpha) if nhwc else
                                    func.leaky_relu(cv, negative_slope=self._leaky_relu_alpha))
            parts = [context_up, out_up] + cost_volumes + [features1]
            if nhwc:
                context, out = decoder_level_nhwc(self._flow_layers[level], parts, self._leaky_relu_alpha)
            else:
                x_in = torch.cat(parts, dim=1)
                dense = list(self._flow_layers[level])[:-1]
                x_out = None
                for i, layer in enumerate(dense):
                    x_out = conv_bias_leaky(layer[0], x_in, self._leaky_relu_alpha)
                    if i + 1 < len(dense):
                        x_in = torch.cat([x_in, x_out], dim=1)
                context = x_out
                out = self._flow_layers[level][-1](context)

            keep = self._keep(out, groups)
            if keep is not None:
                context = context * keep
                out = out * keep
            if out.shape[1] > L + M:   # output level: the propagated tensor has no "rest" channels yet (:337-342)
                out_up = torch.cat([out_up, out_up.new_zeros(out_up.shape[0], L + M + N - out_up.shape[1],
                                                             *out_up.shape[2:])], dim=1)
            out = out + out_up
            out_up = self.upsample_out(out)
            up = self._context_up_layers[level]
            context_up = context_up_nhwc(up, context) if nhwc else up(context)
            outs.insert(0, out)

        if out.shape[1] < L + M + N:
            out = torch.cat([out, out.new_zeros(out.shape[0], L + M + N - out.shape[1], *out.shape[2:])], dim=1)
        if self._nhwc and out.is_cuda:
            refinement = refine_nhwc(self._refine_model, context, out, self._leaky_relu_alpha)
        else:
            refinement = torch.cat([context, out], dim=1)
            refine = list(self._refine_model)          # conv, LeakyReLU, ..., conv
            for conv in refine[:-1:2]:
                refinement = conv_bias_leaky(conv, refinement, self._leaky_relu_alpha)
            refinement = refine[-1](refinement)
        keep = self._keep(refinement, groups)
        if keep is not None:
            refinement = refinement * keep
        refined = out + refinement
        flow, log_diag, rest = torch.split(refined, [L, M, N], dim=1)
        log_diag = torch.clamp(log_diag, min=-5.0) if self._inv_cov else torch.clamp(log_diag, max=10.0, min=-10.0)
        outs[0] = torch.cat([flow, log_diag, rest], dim=1)
        out_1 = self.upsample_out(outs[0])
        out_0 = self.upsample_out(out_1)
        outs.insert(0, out_1)
        outs.insert(0, out_0)
        return outs

    def forward(self, img1, img2, with_bk=True):
        B = img1.shape[0]
        flows_fw, flows_bw = [], []
        for extractor in self._feature_pyramid_extractor:
            if with_bk and self._stack_directions:
                feats = extractor(torch.cat([img1, img2], dim=0), nhwc=self._nhwc)
                if self._nhwc and img1.is_cuda:
                    swapped = None     # the half-batch swap rides on the NHWC -> NCHW copy of the features
                else:
                    swapped = [None] + [torch.cat([f[B:], f[:B]], dim=0) for f in feats[1:]]   # level 0 is never read
                outs = self.forward_2_frames(feats, swapped, groups=2)
                flows_fw.append([o[:B] for o in outs])
                flows_bw.append([o[B:] for o in outs])
            else:
                feat1, feat2 = extractor(img1, nhwc=self._nhwc), extractor(img2, nhwc=self._nhwc)
                flows_fw.append(self.forward_2_frames(feat1, feat2))
                if with_bk:
                    flows_bw.append(self.forward_2_frames(feat2, feat1))
        res_dict = {'flows_fw': self.flows_cat(flows_fw)}
        if with_bk:
            res_dict['flows_bw'] = self.flows_cat(flows_bw)
        return res_dict

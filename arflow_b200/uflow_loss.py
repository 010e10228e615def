"""Drop-in for losses/uflow_loss.py of deu439/ARFlow: UFlowLoss(cfg)(output, target)."""
import torch.nn as nn

from .loss_blocks import _SmoothFunction
from .uflow_utils import (census_loss, clamp01, compute_range_map, downsample, mask_invalid_flow, resample_flow,
                          upsample)


class UFlowLoss(nn.modules.Module):
    """uflow_loss.py:8-109.  cfg needs: w_census, w_smooth, edge_constant, with_bk, smooth_order."""

    def __init__(self, cfg):
        super(UFlowLoss, self).__init__()
        self.cfg = cfg

    def _direction(self, im_a, im_b, flow_ab_0, flow_ba_2):
        """photometric term of one direction: warp im_b towards im_a, masks, fused census (uflow_loss.py:28-54)."""
        # resample(im_b, flow_to_warp(flow)) and mask_invalid(flow_to_warp(flow)) with the grid added in-kernel
        recons = resample_flow(im_b.detach(), flow_ab_0)
        valid = mask_invalid_flow(flow_ab_0)
        occu = upsample(clamp01(compute_range_map(flow_ba_2.detach())), is_flow=False, scale_factor=4.0)
        mask = (occu * valid).detach()
        return census_loss(im_a, recons, mask), mask

    def _smooth(self, im_0, flow_2):
        """edge-aware smoothness of one direction at level 2 (uflow_loss.py:58-102)."""
        order = getattr(self.cfg, 'smooth_order', 1) if not isinstance(self.cfg, dict) else self.cfg.get('smooth_order', 1)
        im_2 = downsample(im_0.detach(), is_flow=False, scale_factor=4.0)
        if order == 1:
            return _SmoothFunction.apply(flow_2, im_2, 1, 1, 0, 0, self.cfg.edge_constant, 0.001 ** 2,
                                         self.cfg.w_smooth / 2.)
        if order == 2:
            return _SmoothFunction.apply(flow_2, im_2, 2, 2, 0, 0, self.cfg.edge_constant, 0.001 ** 2,
                                         self.cfg.w_smooth / 2.)
        raise NotImplementedError("smooth_order must be 1 or 2")

    def forward(self, output, target):
        """
        :param output: multi-scale forward/backward flows, n * [B x 4 x h x w]
        :param target: image pairs B x 6 x H x W
        :return: total_loss, loss_warp, loss_smooth, mean |flow| at level 0, forward mask
        """
        # every slice below is read by two or three kernels that need dense NCHW: one copy each here instead of one per
        # consumer (six image and four full-resolution flow copies per step in the launch list otherwise)
        flow12_0 = output[0][:, 0:2].contiguous()
        flow21_0 = output[0][:, 2:4].contiguous()
        flow12_2 = output[2][:, 0:2].contiguous()
        flow21_2 = output[2][:, 2:4].contiguous()
        im1_0 = target[:, :3].contiguous()
        im2_0 = target[:, 3:].contiguous()

        l1, mask1 = self._direction(im1_0, im2_0, flow12_0, flow21_2)
        loss_warp = self.cfg.w_census * l1
        if self.cfg.with_bk:
            l2, _ = self._direction(im2_0, im1_0, flow21_0, flow12_2)
            loss_warp = loss_warp + self.cfg.w_census * l2

        loss_smooth = self._smooth(im1_0, flow12_2)
        if self.cfg.with_bk:
            loss_smooth = loss_smooth + self._smooth(im2_0, flow21_2)

        total_loss = loss_warp + loss_smooth
        return total_loss, loss_warp, loss_smooth, output[0].abs().mean(), mask1

"""Drop-in for losses/uflow_loss.py of deu439/ARFlow: UFlowLoss(cfg)(output, target)."""
import torch
import torch.nn as nn

from .loss_blocks import _SmoothFunction
from .uflow_utils import (census_loss, census_loss_groups, clamp01, compute_range_map, downsample, mask_invalid_flow, resample_flow,
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
        B = target.shape[0]
        if not self.cfg.with_bk:
            # every slice below is read by two or three kernels that need dense NCHW: one copy each here instead of one
            # per consumer
            flow12_0 = output[0][:, 0:2].contiguous()
            flow21_2 = output[2][:, 2:4].contiguous()
            flow12_2 = output[2][:, 0:2].contiguous()
            im1_0 = target[:, :3].contiguous()
            im2_0 = target[:, 3:].contiguous()
            l1, mask1 = self._direction(im1_0, im2_0, flow12_0, flow21_2)
            loss_warp = self.cfg.w_census * l1
            loss_smooth = self._smooth(im1_0, flow12_2)
            total_loss = loss_warp + loss_smooth
            return total_loss, loss_warp, loss_smooth, output[0].abs().mean(), mask1

        # Both directions are the same computation on swapped roles (uflow_loss.py:28-54 twice), and every op in it is
        # per-sample: they run STACKED on the batch - one warp, one validity mask, one range map, one up-sampling, one
        # down-sampling, one smoothness launch for 2B samples instead of two for B (half the launches of the loss, and
        # twice the parallelism for kernels that one batch of 8 images does not fill the machine with).  Only the census
        # term keeps one kernel call per direction: its normaliser sum(mask) is per direction (uflow_utils.py:293); the
        # calls read and write batch slices of the stacked tensors (census_loss_groups: views, no copies).
        flow_0 = torch.cat([output[0][:, 0:2], output[0][:, 2:4]], dim=0)          # [flow12; flow21] at level 0
        flow_2 = torch.cat([output[2][:, 0:2], output[2][:, 2:4]], dim=0)          # [flow12; flow21] at level 2
        flow_2_other = torch.cat([output[2][:, 2:4], output[2][:, 0:2]], dim=0).detach()   # [flow21; flow12]
        im_a = torch.cat([target[:, :3], target[:, 3:]], dim=0)                    # [im1; im2]: the image each flow starts from
        im_b = torch.cat([target[:, 3:], target[:, :3]], dim=0)                    # [im2; im1]: the image it is warped from

        recons = resample_flow(im_b.detach(), flow_0)
        valid = mask_invalid_flow(flow_0)
        occu = upsample(clamp01(compute_range_map(flow_2_other)), is_flow=False, scale_factor=4.0)
        mask = (occu * valid).detach()
        l12 = census_loss_groups(im_a, recons, mask, 2)
        loss_warp = self.cfg.w_census * l12[0] + self.cfg.w_census * l12[1]

        # mean over 2B samples = (mean_1 + mean_2) / 2, so the per-direction factor w_smooth / 2 becomes w_smooth
        order = getattr(self.cfg, 'smooth_order', 1) if not isinstance(self.cfg, dict) else self.cfg.get('smooth_order', 1)
        if order not in (1, 2):
            raise NotImplementedError("smooth_order must be 1 or 2")
        im_2 = downsample(im_a.detach(), is_flow=False, scale_factor=4.0)
        loss_smooth = _SmoothFunction.apply(flow_2, im_2, order, order, 0, 0, self.cfg.edge_constant, 0.001 ** 2,
                                            self.cfg.w_smooth)

        total_loss = loss_warp + loss_smooth
        return total_loss, loss_warp, loss_smooth, output[0].abs().mean(), mask[:B]

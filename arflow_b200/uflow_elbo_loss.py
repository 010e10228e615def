"""Drop-in for losses/uflow_elbo_loss.py of deu439/ARFlow: the variational (ELBO) UFlow loss.

`UFlowElboLoss(cfg)(res_dict, im1_0, im2_0)` returns the reference's 8-tuple
(total, loss_warp, loss_smooth, loss_entropy, loss_oof, flow12_2, occu_mask12, valid_mask12).
Supported variational families: `approx` = 'diag' (with or without `inv_cov`) and 'sparse' (the
non-diagonal stencil-triangular covariance, config chairs_uflow_elbo_nondiag.json) — the ones the
reference can actually run end to end besides 'mixture'/'lowrank', which are not implemented here.
All heavy lifting goes through the arflow_b200 kernels: stencil mat-vec for the reparameterisation
(uflow_elbo_loss.py:142-147), x4 upsampling, warp, masks, range map, fused census / SSIM, x1/4 downsampling.
"""
import torch
import torch.nn as nn

from .penalty_functions import get_penalty
from .triag_solve import matrix_vector_product_general
from .uflow_utils import (census_loss_no_penalty, compute_range_map, downsample, flow_to_warp, image_grads,
                          mask_invalid, resample, ssim_loss, upsample)


def data_loss_no_penalty(im1_0, im2_0, flow12_2, flow21_2, occ_type, data_loss, mean12_2=None, mean21_2=None):
    """uflow_elbo_loss.py:18-78 — per-pixel data terms and weight maps before the penalty functions."""
    flow12_0 = upsample(flow12_2, is_flow=True, scale_factor=4.0)
    warp12_0 = flow_to_warp(flow12_0)
    im1_recons = resample(im2_0.detach(), warp12_0)

    occu_mask_2 = None
    if occ_type == 'mean':
        mean_warp12_0 = flow_to_warp(upsample(mean12_2, is_flow=True, scale_factor=4.0))
        valid_mask_0 = mask_invalid(mean_warp12_0)
        occu_mask_2 = torch.clamp(compute_range_map(mean21_2), min=0., max=1.)
        mask_0 = (upsample(occu_mask_2, is_flow=False, scale_factor=4.0) * valid_mask_0).detach()
    elif occ_type == 'sample':
        valid_mask_0 = mask_invalid(warp12_0)
        occu_mask_2 = torch.clamp(compute_range_map(flow21_2), min=0., max=1.)
        mask_0 = (upsample(occu_mask_2, is_flow=False, scale_factor=4.0) * valid_mask_0).detach()
    elif occ_type == 'none':
        valid_mask_0 = mask_invalid(warp12_0)
        mask_0 = valid_mask_0.detach()
    else:
        raise NotImplementedError('Occlusion type {} not implemented!'.format(occ_type))

    pixel_loss, pixel_weight = [], []
    for kind in data_loss:
        if kind == "census":
            l, w = census_loss_no_penalty(im1_0, im1_recons, mask_0)
        elif kind == "ssim":
            l, w = ssim_loss(im1_0, im1_recons, mask_0)
        else:
            raise NotImplementedError('Data loss {} not implemented!'.format(kind))
        pixel_loss.append(l)
        pixel_weight.append(w)
    return pixel_loss, pixel_weight, occu_mask_2, valid_mask_0


def _edge_weights(im_0, edge_constant, edge_asymp, stride=1):
    im_2 = downsample(im_0, is_flow=False, scale_factor=4.0)
    gx, gy = image_grads(im_2.detach(), stride=stride)
    wx = edge_asymp + (1.0 - edge_asymp) * torch.exp(-torch.mean(torch.abs(edge_constant * gx), 1, keepdim=True))
    wy = edge_asymp + (1.0 - edge_asymp) * torch.exp(-torch.mean(torch.abs(edge_constant * gy), 1, keepdim=True))
    return wx, wy


def smooth_loss_no_penalty(im1_0, flow12_2, edge_constant, edge_asymp):
    """uflow_elbo_loss.py:81-96 — flow first differences and halved edge-aware weights."""
    wx, wy = _edge_weights(im1_0, edge_constant, edge_asymp)
    flow12_x, flow12_y = image_grads(flow12_2)
    return flow12_x, wx / 2., flow12_y, wy / 2.


class UFlowElboLoss(nn.modules.Module):
    def __init__(self, cfg):
        super(UFlowElboLoss, self).__init__()
        self.cfg = cfg

    # -- sampling -------------------------------------------------------------------------
    def _normal(self, size, like):
        """Standard normal noise on the data's device (the reference moves a Normal(0,1) to the GPU, :112-116)."""
        return torch.randn(size, device=like.device, dtype=like.dtype)

    def reparam_diag(self, mean, log_diag, nsamples=1):
        """:118-128"""
        mean = mean.repeat(nsamples, 1, 1, 1)
        log_diag = log_diag.repeat(nsamples, 1, 1, 1)
        return mean + torch.exp(log_diag) * self._normal(mean.size(), mean)

    def reparam_diag_inv(self, mean, log_diag, nsamples=1):
        """:130-140"""
        mean = mean.repeat(nsamples, 1, 1, 1)
        log_diag = log_diag.repeat(nsamples, 1, 1, 1)
        return mean + torch.exp(-log_diag) * self._normal(mean.size(), mean)

    def reparam_triag(self, mean, std, nsamples=1):
        """:142-147 — z = mean + L eps with the stencil-triangular factor L (csrc/stencil.cu)."""
        mean = mean.repeat(nsamples, 1, 1, 1)
        std = std.repeat(nsamples, 1, 1, 1)
        eps = self._normal(mean.size(), mean)
        return mean + matrix_vector_product_general(std, eps, k=self.cfg.cov_supp)

    # -- forward ----------------------------------------------------------------------------
    def forward(self, res_dict, im1_0, im2_0):
        cfg = self.cfg
        fw, bw = res_dict['flows_fw'][2], res_dict['flows_bw'][2]
        if cfg.approx == 'diag':
            mean12_2, log_diag12_2 = fw[:, 0:2], fw[:, 2:4]
            mean21_2, log_diag21_2 = bw[:, 0:2], bw[:, 2:4]
            diag12_2, diag21_2 = torch.exp(log_diag12_2), torch.exp(log_diag21_2)
        elif cfg.approx == 'sparse':
            n_off = (cfg.cov_supp + 1) ** 2 - 1
            mean12_2, log_diag12_2, offdiag12_2 = fw[:, 0:2], fw[:, 2:4], fw[:, 4:4 + n_off * 2]
            mean21_2, log_diag21_2, offdiag21_2 = bw[:, 0:2], bw[:, 2:4], bw[:, 4:4 + n_off * 2]
            diag12_2, diag21_2 = torch.exp(log_diag12_2), torch.exp(log_diag21_2)
            full12_2 = torch.cat((diag12_2, offdiag12_2), dim=1)
            full21_2 = torch.cat((diag21_2, offdiag21_2), dim=1)
        else:
            raise NotImplementedError("arflow_b200 UFlowElboLoss: approx='%s' is not implemented" % cfg.approx)
        if cfg.natural_grad:
            raise NotImplementedError("Natural gradient is not implemented!")

        loss_offdiag = 0
        if cfg.approx == 'sparse':
            loss_offdiag = torch.mean(torch.square(offdiag12_2))
            if cfg.with_bk:
                loss_offdiag = loss_offdiag + torch.mean(torch.square(offdiag21_2))

        # reparameterisation (forward direction first, then backward: same noise order as the reference)
        ns = cfg.n_samples
        if cfg.approx == 'diag' and not cfg.inv_cov:
            flow12_2 = self.reparam_diag(mean12_2, log_diag12_2, nsamples=ns)
            flow21_2 = self.reparam_diag(mean21_2, log_diag21_2, nsamples=ns)
        elif cfg.approx == 'diag':
            flow12_2 = self.reparam_diag_inv(mean12_2, log_diag12_2, nsamples=ns)
            flow21_2 = self.reparam_diag_inv(mean21_2, log_diag21_2, nsamples=ns)
        elif not cfg.inv_cov:
            flow12_2 = self.reparam_triag(mean12_2, full12_2, nsamples=ns)
            flow21_2 = self.reparam_triag(mean21_2, full21_2, nsamples=ns)
        else:
            raise NotImplementedError("Sparse precision matrix representation is not implemented!")

        im1_0 = im1_0.repeat(ns, 1, 1, 1)
        im2_0 = im2_0.repeat(ns, 1, 1, 1)
        mean12_2_rep = mean12_2.repeat(ns, 1, 1, 1)
        mean21_2_rep = mean21_2.repeat(ns, 1, 1, 1)

        # entropy
        if cfg.approx == 'diag' and not cfg.inv_cov and cfg.approx_entropy:
            tmp12 = (flow12_2 - mean12_2_rep.detach()) / diag12_2.detach().repeat(ns, 1, 1, 1)
            loss_entropy = cfg.w_entropy * torch.sum(tmp12 * tmp12 / 2, dim=1).mean()
            if cfg.with_bk:
                tmp21 = (flow21_2 - mean21_2_rep.detach()) / diag21_2.detach().repeat(ns, 1, 1, 1)
                loss_entropy = loss_entropy + cfg.w_entropy * torch.sum(tmp21 * tmp21 / 2, dim=1).mean()
        else:
            sign = -1.0 if cfg.inv_cov else 1.0
            loss_entropy = sign * cfg.w_entropy * torch.sum(log_diag12_2, dim=1).mean()
            if cfg.with_bk:
                loss_entropy = loss_entropy + sign * cfg.w_entropy * torch.sum(log_diag21_2, dim=1).mean()

        # data term at level 0
        penalties = [get_penalty(t) for t in cfg.data_penalty]
        loss_warp = 0
        pl12, pw12, occu_mask12, valid_mask12 = data_loss_no_penalty(
            im1_0, im2_0, flow12_2, flow21_2, cfg.occ_type, cfg.data_loss, mean12_2_rep, mean21_2_rep)
        for l, w, weight, pen in zip(pl12, pw12, cfg.data_weight, penalties):
            loss_warp = loss_warp + torch.sum(w * weight * pen(l))
        occu_mask21 = None
        if cfg.with_bk:
            pl21, pw21, occu_mask21, _ = data_loss_no_penalty(
                im2_0, im1_0, flow21_2, flow12_2, cfg.occ_type, cfg.data_loss, mean21_2_rep, mean12_2_rep)
            for l, w, weight, pen in zip(pl21, pw21, cfg.data_weight, penalties):
                loss_warp = loss_warp + torch.sum(w * weight * pen(l))

        # smoothness at level 2
        pen_s = get_penalty(cfg.penalty_smooth)
        iso = getattr(cfg, 'isotropic_smooth', False)

        def weighted(wx, wy, ex, ey):
            if iso:
                ex, ey = torch.mean(ex, dim=1), torch.mean(ey, dim=1)
            return torch.mean(wx * cfg.w_smooth * pen_s(ex)) + torch.mean(wy * cfg.w_smooth * pen_s(ey))

        def closed_form(im_0, mean, diag):
            if cfg.order_smooth == 1:
                _, wx, _, wy = smooth_loss_no_penalty(im_0, mean, cfg.edge_constant, cfg.edge_asymp)
                ex = (mean[:, :, :, 1:] - mean[:, :, :, :-1]) ** 2 + diag[:, :, :, 1:] ** 2 + diag[:, :, :, :-1] ** 2
                ey = (mean[:, :, 1:] - mean[:, :, :-1]) ** 2 + diag[:, :, 1:] ** 2 + diag[:, :, :-1] ** 2
            elif cfg.order_smooth == 2:
                wx, wy = _edge_weights(im_0, cfg.edge_constant, cfg.edge_asymp, stride=2)
                ex = ((mean[:, :, :, :-2] - 2 * mean[:, :, :, 1:-1] + mean[:, :, :, 2:]) ** 2
                      + diag[:, :, :, 0:-2] ** 2 + 4 * diag[:, :, :, 1:-1] ** 2 + diag[:, :, :, 2:] ** 2)
                ey = ((mean[:, :, :-2] - 2 * mean[:, :, 1:-1] + mean[:, :, 2:]) ** 2
                      + diag[:, :, 0:-2] ** 2 + 4 * diag[:, :, 1:-1] ** 2 + diag[:, :, 2:] ** 2)
            else:
                raise NotImplementedError()
            return weighted(wx, wy, ex, ey)

        def sampled(im_0, flow):
            fx, wx, fy, wy = smooth_loss_no_penalty(im_0, flow, cfg.edge_constant, cfg.edge_asymp)
            return weighted(wx, wy, fx ** 2, fy ** 2)

        if cfg.closed_form_smooth:
            if cfg.approx != 'diag':
                raise NotImplementedError()
            # the reference passes the sample-repeated images with the un-repeated means; the weights broadcast
            # only for n_samples == 1 there, so the first B images are used
            B = mean12_2.shape[0]
            loss_smooth = closed_form(im1_0[:B], mean12_2, diag12_2)
            if cfg.with_bk:
                loss_smooth = loss_smooth + closed_form(im2_0[:B], mean21_2, diag21_2)
        else:
            loss_smooth = sampled(im1_0, flow12_2)
            if cfg.with_bk:
                loss_smooth = loss_smooth + sampled(im2_0, flow21_2)

        # out-of-frame and occlusion penalties
        loss_oof = 0
        if cfg.w_oof > 0.0:
            def oof(flow):
                wrp = flow_to_warp(flow)
                mh, mw = float(wrp.shape[2] - 1), float(wrp.shape[3] - 1)
                u = torch.clamp(wrp[:, 0], max=0) ** 2 + torch.clamp(wrp[:, 0] - mw, min=0) ** 2
                v = torch.clamp(wrp[:, 1], max=0) ** 2 + torch.clamp(wrp[:, 1] - mh, min=0) ** 2
                return cfg.w_oof * (u + v).mean()
            loss_oof = oof(flow12_2)
            if cfg.with_bk:
                loss_oof = loss_oof + oof(flow21_2)

        loss_occ = 0
        if cfg.w_occ > 0.0:
            loss_occ = cfg.w_occ * (1 / (100.0 * occu_mask12 + 1) * torch.square(flow12_2)).mean()
            if cfg.with_bk:
                loss_occ = loss_occ + cfg.w_occ * (1 / (100.0 * occu_mask21 + 1) * torch.square(flow21_2)).mean()

        total_loss = loss_warp + loss_smooth - loss_entropy + loss_oof + loss_occ
        if cfg.approx == 'sparse':
            total_loss = total_loss + cfg.offdiag_reg * loss_offdiag
        return total_loss, loss_warp, loss_smooth, loss_entropy, loss_oof, flow12_2, occu_mask12, valid_mask12

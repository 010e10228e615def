"""Times one training step of BASELINE.json's configs 3 and 4 on one B200 (evidence that the other configurations run
at their full sizes; the bench line itself stays config 2, see bench.py).

    python tools/step_configs.py [elbo] [kitti]

config 3: sintel_uflow_elbo non-diagonal covariance: PWCProbFlow [2,2,30] + UFlowElboLoss(approx='sparse', cov_supp=3,
          n_samples=4), 448x1024, batch 8 (loss block of configs/chairs_uflow_elbo_nondiag.json:23-46), eager steps.
config 4: kitti_uflow: PWCFlow + UFlowLoss(smooth_order=2), 320x1024, per-GPU batch 16 / 8 / 4, CUDA-graph replay.
"""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n, out


def elbo():
    from arflow_b200.uflow_elbo_loss import UFlowElboLoss
    from arflow_b200.uflow_prob_model import PWCProbFlow
    torch.manual_seed(0)
    net = PWCProbFlow(types.SimpleNamespace(out_channels=[2, 2, 30], inv_cov=False, n_pyramids=1, mixture_weights=False,
                                            feature_norm=True, level_dropout=0.1)).cuda().train()
    lcfg = dict(edge_constant=150, edge_asymp=0.01, w_smooth=4.0, penalty_smooth="charbonnier", closed_form_smooth=False,
                data_loss=["census"], data_weight=[1.0], data_penalty=["abs_robust_loss"], w_entropy=0.1, w_oof=0.0,
                w_occ=0.0, with_bk=True, approx="sparse", n_components=1, cov_supp=3, inv_cov=False,
                approx_entropy=False, occ_type="sample", n_samples=4, offdiag_reg=0.0, natural_grad=False,
                isotropic_smooth=False)
    loss_fn = UFlowElboLoss(types.SimpleNamespace(**lcfg))
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    B, H, W = 8, 448, 1024
    gen = torch.Generator().manual_seed(1)
    im1, im2 = torch.rand(B, 3, H, W, generator=gen).cuda(), torch.rand(B, 3, H, W, generator=gen).cuda()

    def step():
        opt.zero_grad(set_to_none=True)
        out = loss_fn(net(im1, im2, with_bk=True), im1, im2)
        out[0].backward()
        opt.step()
        return out[0].detach()
    ms, loss = timed(step)
    print(json.dumps({"config": "3: sintel_uflow_elbo nondiag, PWCProbFlow[2,2,30] + UFlowElboLoss sparse k=3 n_samples=4",
                      "shape": [B, H, W], "mode": "eager", "ms_per_step": ms, "pairs_per_s": B / ms * 1e3,
                      "loss": float(loss), "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


def kitti():
    from arflow_b200.train_step import UFlowTrainStep
    from arflow_b200.uflow_loss import UFlowLoss
    from arflow_b200.uflow_model import PWCFlow
    H, W = 320, 1024
    for B in (16, 8, 4):
        torch.manual_seed(0)
        net = PWCFlow(types.SimpleNamespace(level_dropout=0.1, feature_norm=True)).cuda().train()
        loss_fn = UFlowLoss(types.SimpleNamespace(edge_constant=150, w_smooth=4.0, w_census=1.0, with_bk=True, smooth_order=2))
        step = UFlowTrainStep(net, loss_fn, lr=1e-4, use_graph=True)
        x = torch.rand(B, 6, H, W, generator=torch.Generator().manual_seed(2)).cuda()
        ms, out = timed(lambda: step(x), n=10)
        print(json.dumps({"config": "4: kitti_uflow, PWCFlow + UFlowLoss(smooth_order=2)", "shape": [B, H, W],
                          "mode": "cuda graph", "ms_per_step": ms, "pairs_per_s": B / ms * 1e3,
                          "loss": float(out[0])}), flush=True)
        del step, net
        torch.cuda.empty_cache()


if __name__ == "__main__":
    torch.backends.cudnn.benchmark = True
    what = sys.argv[1:] or ["elbo", "kitti"]
    if "elbo" in what:
        elbo()
    if "kitti" in what:
        kitti()

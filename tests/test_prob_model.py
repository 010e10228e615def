"""The PWCProbFlow caller (config 3's network): same module tree / state-dict keys / outputs as the reference
network (models/uflow_prob_model.py), goldens from tests/golden/make_golden.py."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, assert_close

CFGS = {
    "nondiag": dict(out_channels=[2, 2, 30], inv_cov=False, n_pyramids=1, mixture_weights=False, feature_norm=True,
                    level_dropout=0.1),
    "diag2pyr": dict(out_channels=[2, 2, 0], inv_cov=True, n_pyramids=2, mixture_weights=False, feature_norm=True,
                     level_dropout=0.1),
}


def _golden(tag):
    with np.load(os.path.join(GOLDEN, "pwcprobflow_%s.npz" % tag), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def _build(tag, seed, ops=None, device="cpu", stack=True, nhwc=True):
    from arflow_b200.uflow_prob_model import PWCProbFlow
    torch.manual_seed(seed)
    net = PWCProbFlow(types.SimpleNamespace(**CFGS[tag]), ops=ops, stack_directions=stack, nhwc=nhwc)
    net.init_weights()
    return net.to(device).eval()


def _inputs(seed):
    gen = torch.Generator().manual_seed(seed + 1000)
    return torch.rand(1, 3, 192, 256, generator=gen), torch.rand(1, 3, 192, 256, generator=gen)


@pytest.mark.parametrize("tag", ["nondiag", "diag2pyr"])
@pytest.mark.parametrize("stack", [False, True])
def test_prob_model_matches_reference_on_cpu_ops(oracle, tag, stack):
    g = _golden(tag)
    seed = int(g["in0"])
    net = _build(tag, seed, ops=oracle.OracleOps(), stack=stack)
    assert [str(k) for k in g["keys"]] == list(net.state_dict().keys())
    assert int(g["n_params"]) == sum(p.numel() for p in net.parameters())
    im1, im2 = _inputs(seed)
    with torch.no_grad():
        r = net(im1, im2, with_bk=True)
    assert len(r["flows_fw"]) == 6
    assert_close(r["flows_fw"][2], torch.from_numpy(g["fw2"]), 1e-4, "flows_fw[2]")
    assert_close(r["flows_bw"][2], torch.from_numpy(g["bw2"]), 1e-4, "flows_bw[2]")
    assert_close(r["flows_fw"][4], torch.from_numpy(g["fw4"]), 1e-4, "flows_fw[4]")
    assert_close(r["flows_fw"][0].abs().mean(dim=(0, 2, 3)), torch.from_numpy(g["fw0_absmean"]), 1e-4, "flows_fw[0]")


def test_mixture_weights_is_refused():
    from arflow_b200.uflow_prob_model import PWCProbFlow
    cfg = dict(CFGS["nondiag"], mixture_weights=True)
    with pytest.raises(NotImplementedError):
        PWCProbFlow(types.SimpleNamespace(**cfg), ops=object())


@pytest.mark.gpu
@pytest.mark.parametrize("nhwc", [False, True])
@pytest.mark.parametrize("tag", ["nondiag", "diag2pyr"])
def test_prob_model_matches_reference_on_b200(tag, nhwc):
    g = _golden(tag)
    seed = int(g["in0"])
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False      # compare against an fp32 CPU run of the reference
    try:
        net = _build(tag, seed, device="cuda", nhwc=nhwc)
        im1, im2 = _inputs(seed)
        with torch.no_grad():
            r = net(im1.cuda(), im2.cuda(), with_bk=True)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert_close(r["flows_fw"][2], torch.from_numpy(g["fw2"]), 1e-3, "flows_fw[2]")
    assert_close(r["flows_bw"][2], torch.from_numpy(g["bw2"]), 1e-3, "flows_bw[2]")


@pytest.mark.gpu
def test_elbo_nondiag_train_step_runs_on_b200():
    """Config 3 end to end at reduced size: PWCProbFlow [2,2,30] -> UFlowElboLoss (sparse / stencil mat-vec) ->
    backward -> Adam; losses finite, every parameter that feeds the loss receives a finite gradient, loss moves."""
    from arflow_b200.uflow_elbo_loss import UFlowElboLoss
    from arflow_b200.uflow_prob_model import PWCProbFlow
    torch.manual_seed(3)
    net = PWCProbFlow(types.SimpleNamespace(**dict(CFGS["nondiag"], level_dropout=0.0))).cuda().train()
    import json
    with np.load(os.path.join(GOLDEN, "elbo_sparse.npz")) as z:
        lcfg = json.loads(str(z["cfg"]))     # the loss block of configs/chairs_uflow_elbo_nondiag.json (cov_supp 3)
    loss_fn = UFlowElboLoss(types.SimpleNamespace(**lcfg))
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    gen = torch.Generator().manual_seed(9)
    im1, im2 = torch.rand(2, 3, 192, 256, generator=gen).cuda(), torch.rand(2, 3, 192, 256, generator=gen).cuda()
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        res = net(im1, im2, with_bk=True)
        out = loss_fn(res, im1, im2)
        out[0].backward()
        grads = [p.grad for p in net.parameters() if p.grad is not None]
        assert len(grads) > 60 and all(torch.isfinite(gp).all() for gp in grads)
        opt.step()
        losses.append(float(out[0]))
    assert all(np.isfinite(losses)) and losses[2] != losses[0]

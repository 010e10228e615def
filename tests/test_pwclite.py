"""The PWC-Lite caller (config 1): same module tree / state-dict keys / outputs as the reference network
(models/pwclite.py:109-283).  CPU: the oracle's restatement against the reference's golden outputs; GPU: the product."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, assert_close

CASES = {"pwclite_eval": dict(upsample=True, n_frames=2, reduce_dense=True),
         "pwclite3_eval": dict(upsample=True, n_frames=3, reduce_dense=False)}


def _golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def _input(g):
    shape = tuple(int(v) for v in g["shape"])
    return torch.rand(*shape, generator=torch.Generator().manual_seed(int(g["in0"]) + 1000))


def _outputs(g):
    return sorted(k for k in g if k[:2] in ("fw", "bw") and k[2:].isdigit())


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_twin_matches_reference(name):
    from oracle.cpu_nets import PWCLiteCPU
    g = _golden(name)
    torch.manual_seed(int(g["in0"]))
    net = PWCLiteCPU(**CASES[name]).eval()
    assert [str(k) for k in g["keys"]] == list(net.state_dict().keys())
    assert int(g["n_params"]) == sum(p.numel() for p in net.parameters())
    with torch.no_grad():
        r = net(_input(g), with_bk=True)
    for k in _outputs(g):
        assert_close(r["flows_" + k[:2]][int(k[2:])], torch.from_numpy(g[k]), 1e-4, k)
    assert abs(float(r["flows_fw"][0].abs().mean()) - float(g["fw0_absmean"])) < 1e-4 * float(g["fw0_absmean"])


@pytest.mark.parametrize("name", list(CASES))
def test_product_state_dict_matches_reference(name):
    from arflow_b200.pwclite import PWCLite
    g = _golden(name)
    net = PWCLite(types.SimpleNamespace(**CASES[name]))
    assert [str(k) for k in g["keys"]] == list(net.state_dict().keys())
    assert int(g["n_params"]) == net.num_parameters()
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3 * CASES[name]["n_frames"], 64, 64))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("stack", [False, True])
def test_product_forward_matches_reference_on_b200(name, stack):
    from arflow_b200.pwclite import PWCLite
    g = _golden(name)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False      # compare against an fp32 CPU run of the reference
    try:
        torch.manual_seed(int(g["in0"]))
        net = PWCLite(types.SimpleNamespace(**CASES[name]), stack_directions=stack).cuda().eval()
        with torch.no_grad():
            r = net(_input(g).cuda(), with_bk=True)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for k in _outputs(g):
        assert_close(r["flows_" + k[:2]][int(k[2:])], torch.from_numpy(g[k]), 1e-3, k)


@pytest.mark.gpu
def test_product_backward_runs_and_matches_oracle_twin():
    """Gradients through warp / cost volume / align_corners up-sampling of the PWC-Lite decoder (training path)."""
    from arflow_b200.pwclite import PWCLite
    from oracle.cpu_nets import PWCLiteCPU
    cfg = CASES["pwclite_eval"]
    torch.manual_seed(5)
    net = PWCLite(types.SimpleNamespace(**cfg)).cuda().train()
    twin = PWCLiteCPU(**cfg).train()
    twin.load_state_dict(net.state_dict())
    x = torch.rand(1, 6, 128, 128, generator=torch.Generator().manual_seed(6))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        r = net(x.cuda(), with_bk=False)
        sum(f.abs().mean() for f in r["flows_fw"]).backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    t = twin(x, with_bk=False)
    sum(f.abs().mean() for f in t["flows_fw"]).backward()
    for (n, p), (_, q) in zip(net.named_parameters(), twin.named_parameters()):
        if q.grad is not None:
            assert_close(p.grad, q.grad, 2e-3, n)

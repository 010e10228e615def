"""world_size-2 gloo test of the data-parallel host logic of arflow_b200.train_step.UFlowTrainStep (flat gradient
buffer, bucket hooks firing during backward, averaged all-reduce, discovery of gradient-less parameters) on the
CPU with the oracle's ops: both ranks must hold identical gradients equal to the single-process gradient of the
mean of the two per-rank losses, on the discovery step and on the overlapped steps after it."""
import os
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _model(orc):
    from arflow_b200.uflow_model import PWCFlow
    torch.manual_seed(11)
    net = PWCFlow(types.SimpleNamespace(level_dropout=0.0, feature_norm=True), ops=orc.OracleOps(),
                  stack_directions=True)
    return net.train()


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import oracle.arflow_oracle as orc
    from arflow_b200.train_step import UFlowTrainStep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    net = _model(orc)
    step = UFlowTrainStep(net, lambda flows, pair: orc.uflow_loss(flows, pair), lr=0.0, use_graph=False,
                          world_size=world, n_buckets=3)
    step.optimizer.step = lambda: None            # keep the weights fixed: gradients of step 1 and 2 must agree
    pair = torch.rand(1, 6, 160, 192, generator=torch.Generator().manual_seed(100 + rank))
    step(pair)
    g1, log1 = step.flat_grad.clone(), list(step.reduced_log)
    step(pair)
    g2, log2 = step.flat_grad.clone(), list(step.reduced_log)
    torch.save({"g1": g1, "g2": g2, "log1": log1, "log2": log2, "pair": pair, "spans": step._spans,
                "counts": step._counts, "nb": len(step._buckets)}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_gradient_average(tmp_path, oracle):
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert torch.equal(r0["g1"], r1["g1"]) and torch.equal(r0["g2"], r1["g2"]), "ranks disagree after the all-reduce"
    assert sorted(r0["log1"]) == list(range(r0["nb"])) and sorted(r0["log2"]) == list(range(r0["nb"]))
    # overlapped step: buckets are issued from inside backward in the order their gradients complete
    assert r0["log2"][0] == 0 and r0["log2"][-1] == r0["nb"] - 1   # refinement/decoder bucket first, feature pyramid last
    assert sum(r0["counts"]) < len(r0["spans"]), "some parameters receive no gradient and must be discovered"
    assert (r0["g1"] - r0["g2"]).abs().max() <= 1e-6 * r0["g1"].abs().max()
    # single-process reference: mean of the two per-rank losses
    net = _model(oracle)
    total = 0
    for r in (r0, r1):
        res = net(r["pair"], with_bk=True)
        flows = [torch.cat([a, b], 1) for a, b in zip(res['flows_fw'], res['flows_bw'])]
        total = total + oracle.uflow_loss(flows, r["pair"])[0] / 2
    total.backward()
    params = [p for p in net.parameters() if p.requires_grad]
    scale = max(float(p.grad.abs().max()) for p in params if p.grad is not None)
    worst = 0.0
    for p, (s0, e0) in zip(params, r0["spans"]):
        ref = p.grad if p.grad is not None else torch.zeros_like(p)
        worst = max(worst, float((r0["g2"][s0:e0].view_as(p) - ref).abs().max()))
    assert worst <= 1e-4 * scale, (worst, scale)


def _norm_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from arflow_b200.uflow_utils import globalise_census_sums
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    num, den = (3.0, 10.0) if rank == 0 else (5.0, 30.0)       # per-rank masked sums
    sums = torch.tensor([num, den, num / (den + 1e-6)])
    torch.save(globalise_census_sums(sums, dist.group.WORLD), os.path.join(out_dir, "norm%d.pt" % rank))
    dist.destroy_process_group()


def test_global_census_normaliser_two_ranks(tmp_path):
    """with the batch-global denominator the AVERAGE of the per-rank losses (what the averaged
    gradient all-reduce optimises) equals the single-process loss sum(num) / (sum(den) + 1e-6)."""
    port = 29800 + os.getpid() % 200
    mp.spawn(_norm_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    s0 = torch.load(os.path.join(tmp_path, "norm0.pt"))
    s1 = torch.load(os.path.join(tmp_path, "norm1.pt"))
    want = (3.0 + 5.0) / (10.0 + 30.0 + 1e-6)
    assert abs(float(s0[2] + s1[2]) / 2 - want) < 1e-6
    # the backward kernel divides by (sums[1] + 1e-6): that must be world / (global den + 1e-6)
    assert abs(1.0 / (float(s0[1]) + 1e-6) - 2.0 / (40.0 + 1e-6)) < 1e-7
    assert float(s0[0]) == 3.0 and float(s1[0]) == 5.0

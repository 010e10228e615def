"""Peer-memory all-reduce (csrc/comm.cu, arflow_b200/comm.py) and the overlapped, fully captured data-parallel train
step.  Needs two GPUs: one process per GPU over NCCL for the handle exchange (`gpurun --gpus 2`); skipped on a
single-GPU box, where only the one-rank degenerate case runs."""
import os
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _init(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))


def _allreduce_worker(rank, world, port, out_dir):
    _init(rank, world, port)
    from arflow_b200.comm import PeerAllReduce
    dev = torch.device("cuda", rank)
    n = 1_000_003
    comm = PeerAllReduce(n, dev, ctas=16)
    res = {}
    g = torch.Generator().manual_seed(7 + rank)
    x = torch.randn(comm.numel, generator=g)
    # whole buffer, mean
    comm.buffer.copy_(x)
    torch.cuda.synchronize()
    dist.barrier()
    comm.all_reduce_(0, comm.numel, average=True)
    comm.check()
    res["mean"] = comm.buffer.cpu().clone()
    # a ragged sub-range, sum; the rest must stay untouched
    comm.buffer.copy_(x)
    torch.cuda.synchronize()
    dist.barrier()
    comm.all_reduce_(1000, 77780, average=False)
    comm.check()
    res["part"] = comm.buffer.cpu().clone()
    # many back-to-back calls inside one CUDA graph, replayed (epochs live on the device)
    comm.buffer.copy_(x)
    torch.cuda.synchronize()
    dist.barrier()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        comm.all_reduce_(0, 4096, average=False)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    comm.buffer.copy_(x)
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        comm.all_reduce_(0, 4096, average=False)
        comm.all_reduce_(4096, 8192, average=False)
    for _ in range(3):
        graph.replay()
    comm.check()
    res["graph"] = comm.buffer[:8192].cpu().clone()
    res["x"] = x
    torch.save(res, os.path.join(out_dir, "rank%d.pt" % rank))
    comm.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_allreduce_multi_gpu(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29700 + os.getpid() % 200 + world
    mp.spawn(_allreduce_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rs = [torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world)]
    total = rs[0]["x"].clone()
    for r in rs[1:]:                      # the kernel sums in rank order
        total = total + r["x"]
    for r in rs:
        assert torch.equal(r["mean"], total * (1.0 / world))
        assert torch.equal(r["part"][1000:77780], total[1000:77780])
        assert torch.equal(r["part"][:1000], r["x"][:1000]) and torch.equal(r["part"][77780:], r["x"][77780:])
        # three replays of sum: every replay sums `world` (already equal) buffers again -> world^2 * total
        assert torch.equal(r["graph"], rs[0]["graph"])
    ref = total[:8192]
    for _ in range(2):
        acc = ref.clone()
        for _ in range(world - 1):
            acc = acc + ref
        ref = acc
    assert torch.equal(rs[0]["graph"], ref)


def _step_worker(rank, world, port, out_dir):
    _init(rank, world, port)
    from arflow_b200 import uflow_utils
    from arflow_b200.train_step import UFlowTrainStep
    from arflow_b200.uflow_loss import UFlowLoss
    from arflow_b200.uflow_model import PWCFlow
    dev = torch.device("cuda", rank)
    out = {}
    # (collective, captured?, batch-global census normaliser?, learning rate)
    for key, mode, graph, gnorm, lr in (("nccl_eager", "nccl", False, False, 0.0), ("peer_eager", "peer", False, False, 0.0),
                                        ("peer_graph", "peer", True, False, 0.0), ("peer_eager_gn", "peer", False, True, 0.0),
                                        ("peer_graph_gn", "peer", True, True, 0.0), ("peer_graph_train", "peer", True, True, 1e-4)):
        torch.manual_seed(3)
        model = PWCFlow(types.SimpleNamespace(level_dropout=0.0, feature_norm=True)).to(dev).train()
        loss_fn = UFlowLoss(types.SimpleNamespace(edge_constant=150, w_smooth=4.0, w_census=1.0, with_bk=True, smooth_order=1))
        step = UFlowTrainStep(model, loss_fn, lr=lr, use_graph=graph, world_size=world, allreduce=mode,
                              global_census_norm=gnorm)
        gen = torch.Generator().manual_seed(50 + rank)
        losses = []
        for _ in range(3):
            pair = torch.rand(2, 6, 320, 384, generator=gen).to(dev)
            losses.append(step(pair).cpu())
        torch.cuda.synchronize()
        out[key] = {"losses": torch.stack(losses), "flat": step.flat_grad.cpu().clone(), "log": list(step.reduced_log),
                    "w": torch.cat([p.detach().flatten().cpu() for p in model.parameters()])}
        uflow_utils.set_census_normaliser_group(None)
    torch.save(out, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_train_step_peer_overlap_matches_nccl_two_gpus(tmp_path):
    """Fixed weights (lr = 0): the reduced gradient of the last step must not depend on the collective (NCCL / peer
    kernel) nor on capture + overlap; the batch-global census normaliser works inside the captured graph; and with a
    learning rate the replicas stay bit-identical."""
    port = 29750 + os.getpid() % 200
    mp.spawn(_step_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for key in r0:
        assert torch.equal(r0[key]["flat"], r1[key]["flat"]), key + ": ranks hold different reduced gradients"
        assert torch.equal(r0[key]["w"], r1[key]["w"]), key + ": replicas diverged"

    def close(x, y, tol, what):
        e = (x - y).abs().max() / y.abs().max()
        assert e <= tol, "%s: %.3e" % (what, float(e))
    close(r0["peer_eager"]["flat"], r0["nccl_eager"]["flat"], 1e-3, "peer kernel vs NCCL")               # separate runs: cuDNN picks its TF32 algorithms per run
    close(r0["peer_graph"]["flat"], r0["peer_eager"]["flat"], 1e-3, "captured + overlapped vs eager")   # cuDNN picks TF32 algorithms per run
    close(r0["peer_graph_gn"]["flat"], r0["peer_eager_gn"]["flat"], 1e-3, "global census normaliser, captured vs eager")
    assert sorted(r0["peer_eager"]["log"]) == [0, 1, 2, 3] and r0["peer_eager"]["log"][0] == 0 and r0["peer_eager"]["log"][-1] == 3
    # the global normaliser changes the gradient (ranks see different mask sums), but not by much
    e = (r0["peer_eager_gn"]["flat"] - r0["peer_eager"]["flat"]).abs().max() / r0["peer_eager"]["flat"].abs().max()
    assert 0 < e < 0.2
    assert (r0["peer_graph_train"]["w"] - r0["peer_graph"]["w"]).abs().max() > 0       # it did train


def test_peer_allreduce_single_rank_is_identity(tmp_path):
    """world size 1: the kernel degenerates to scale * x (no peers) — runs on the single-GPU box."""
    port = 29790 + os.getpid() % 200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if not dist.is_initialized():
        dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from arflow_b200.comm import PeerAllReduce
        comm = PeerAllReduce(5000, torch.device("cuda", 0), ctas=4)
        x = torch.randn(comm.numel, device="cuda")
        comm.buffer.copy_(x)
        comm.all_reduce_(0, comm.numel, average=True)
        comm.all_reduce_(8, 1024, average=False)
        comm.check()
        assert torch.equal(comm.buffer, x)
        comm.close()
    finally:
        dist.destroy_process_group()

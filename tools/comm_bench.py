"""Time the peer-memory all-reduce (and NCCL beside it) on the gradient payload of the train step.
    python -m torch.distributed.run --nproc-per-node N tools/comm_bench.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from arflow_b200.comm import PeerAllReduce
    n = 5_734_636          # PWCFlow parameters (22.9 MB)
    comm = PeerAllReduce(n, dev)
    x = torch.randn(comm.numel, device=dev)
    y = torch.randn(comm.numel, device=dev)

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) * 1e3

    rows = []
    for numel in (comm.numel, comm.numel // 10 // 4 * 4, 4):
        for ctas in (8, 16, 32, 64):
            rows.append(("peer", numel, ctas, timeit(lambda: comm.all_reduce_(0, numel, average=True, ctas=ctas))))
        rows.append(("nccl", numel, 0, timeit(lambda: dist.all_reduce(y[:numel], op=dist.ReduceOp.AVG))))
    comm.check()
    if rank == 0:
        for kind, numel, ctas, us in rows:
            bus = 2 * (world - 1) / world * numel * 4 / (us * 1e-6) / 1e9
            print("%s  world %d  %9d floats  ctas %2d  %8.1f us  bus %7.1f GB/s" % (kind, world, numel, ctas, us, bus), flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

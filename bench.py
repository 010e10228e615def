#!/usr/bin/env python
"""bench.py — train image-pairs/s of the chairs_uflow UFlow step (BASELINE.json `metric`, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo, one rank per GPU (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference path on the host CPUs

A step = PWCFlow forward (both directions) + UFlowLoss + backward + Adam on a synthetic batch of 8 image
pairs of 384x512 per GPU (weak scaling: per-GPU work fixed).  Rank 0 prints ONE JSON line.
  value     pairs/s with the inputs already in HBM (CUDA-graph replay, CUDA events, max over ranks)
  e2e       pairs/s through the public API with HOST inputs: pinned-host -> device copy of every batch
            and a device -> host read of the loss inside the timed region
  roofline  the dominant arflow_b200 kernel of the step, timed in situ with CUDA events around its C-ABI
            call on the launching stream during a few extra eager (non-graph) steps
  cpu_baseline  the reference path (oracle port: torch-CPU restatement, see oracle/) on the host cores,
            bounded sample, rank 0 / N=1 only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, PER_GPU_BATCH = 384, 512, 8
WORKLOAD = "chairs_uflow PWCFlow+UFlowLoss train step, 384x512, batch 8 per GPU, synthetic pairs, random init"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch (default: the named config)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="pairs per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nchw", action="store_true", help="experiment: NCHW conv stacks (the default is channels-last)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true",
                    help="disable cuDNN autotuning of the (out-of-scope) convolutions; about 11 % faster with it")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`) and exit")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers -------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def alg_bytes(name, a):
    """Algorithmic bytes of one C-ABI call (SURVEY §8d / DESIGN.md byte counts); a = raw argument tuple."""
    if name == "arf_corr_fwd":
        B, C, Hh, Ww = a[3:7]
        return B * Hh * Ww * (8 * C + 324)
    if name == "arf_corr_bwd":
        B, C, Hh, Ww = a[5:9]
        return B * Hh * Ww * (16 * C + 324)
    if name == "arf_warp_fwd":
        B, C, _, _, Ho, Wo = a[3:9]
        return B * Ho * Wo * (8 * C + 8)
    if name == "arf_warp_bwd":
        B, C, _, _, Ho, Wo = a[5:11]
        return B * Ho * Wo * ((12 if a[3] else 8) * C + 16)
    if name == "arf_census_fwd":
        B, Hh, Ww = a[6:9]
        return B * Hh * Ww * 32
    if name == "arf_census_bwd":
        B, Hh, Ww = a[9:12]
        return B * Hh * Ww * (24 + 8 + 12 * ((a[7] is not None) + (a[8] is not None)))
    if name in ("arf_resize_bilinear_fwd", "arf_resize_bilinear_bwd"):
        n, Hi, Wi, Ho, Wo = a[2:7]
        return n * (Hi * Wi + Ho * Wo) * 4
    if name in ("arf_smooth_fwd", "arf_smooth_bwd"):
        B, Ci, Hh, Ww = a[4:8]
        return B * Hh * Ww * (4 * Ci + 8 + (8 if name.endswith("bwd") else 0))
    if name == "arf_range_map":
        B, Hh, Ww = a[2:5]
        return B * Hh * Ww * 16
    if name == "arf_inside_mask":
        B, Hh, Ww = a[2:5]
        return B * Hh * Ww * 12
    if name == "arf_count_to_mask":
        return a[2] * 8
    if name in ("arf_nhwc_pack", "arf_nhwc_unpack"):
        return a[2] * a[3] * a[4] * 8                      # one part: read + write
    if name == "arf_bias_leaky_fwd":
        return a[2] * a[3] * a[4] * 8                      # in place: read + write
    if name == "arf_bias_leaky_bwd":
        return a[5] * a[6] * a[7] * 12                     # gy, y in; g out
    if name == "arf_bias_leaky_nhwc_fwd":
        return a[2] * a[3] * 8
    if name == "arf_bias_leaky_nhwc_bwd":
        return a[5] * a[6] * 12
    if name == "arf_bias_leaky_nhwc_bwd_ld":
        return a[7] * a[8] * 12
    if name == "arf_bias_leaky_nhwc_fwd_ld":
        return a[4] * a[5] * 8
    if name == "arf_nhwc_unpack_add":
        return a[2] * a[3] * a[4] * 12                     # part read + packed slice read + part write
    if name == "arf_featnorm_fwd":
        return a[6] * a[7] * 24                            # two maps: read twice (moments, apply), written once
    if name == "arf_featnorm_bwd":
        return a[9] * a[10] * 40
    if name == "arf_pad_weight":
        return a[6] * a[7] * a[4] * a[5] * 8
    if name == "arf_conv3x3_small_fwd":
        N, Hh, Ww, Ci, Co = a[4:9]
        return N * Hh * Ww * (Ci + Co) * 4                 # input read once, output written
    if name == "arf_conv3x3_small_bwd":
        N, Hh, Ww, Ci, Co = a[6:11]
        return N * Hh * Ww * (2 * Ci + Co) * 4             # input and output gradient read, input gradient written
    return 0


FP32_FMA_TFLOPS = 71.3   # tools/fma_peak.cu on this pool's B200 (1965 MHz), FFMA with 2 register operands
MUFU_GOPS = 4544.0       # same tool, MUFU.RSQ


def alg_work(name, a):
    """(shape label, fp32 flops, MUFU ops) of one C-ABI call - the compute-side roofline numerators."""
    if name == "arf_corr_fwd":
        B, C, Hh, Ww = a[3:7]
        return "B%d C%d %dx%d" % (B, C, Hh, Ww), B * Hh * Ww * C * 162, 0
    if name == "arf_corr_bwd":
        B, C, Hh, Ww = a[5:9]
        return "B%d C%d %dx%d" % (B, C, Hh, Ww), B * Hh * Ww * C * 324, 0
    if name == "arf_warp_fwd":
        B, C, _, _, Ho, Wo = a[3:9]
        return "B%d C%d %dx%d" % (B, C, Ho, Wo), B * Ho * Wo * C * 8, 0
    if name == "arf_warp_bwd":
        B, C, _, _, Ho, Wo = a[5:11]
        return "B%d C%d %dx%d%s" % (B, C, Ho, Wo, "" if a[3] else " flow-grad only"), B * Ho * Wo * C * 16, 0
    if name == "arf_census_fwd":
        B, Hh, Ww, patch = a[6:10]
        return "B%d %dx%d p%d" % (B, Hh, Ww, patch), B * Hh * Ww * 13 * (patch * patch - 1), B * Hh * Ww * 3 * (patch * patch - 1)
    if name == "arf_census_bwd":
        B, Hh, Ww, patch = a[9:13]
        return "B%d %dx%d p%d" % (B, Hh, Ww, patch), B * Hh * Ww * 22 * (patch * patch - 1), B * Hh * Ww * 3 * (patch * patch - 1)
    if name in ("arf_resize_bilinear_fwd", "arf_resize_bilinear_bwd"):
        n, Hi, Wi, Ho, Wo = a[2:7]
        return "N%d %dx%d->%dx%d" % (n, Hi, Wi, Ho, Wo), 0, 0
    if name in ("arf_nhwc_pack", "arf_nhwc_unpack"):
        return "N%d HW%d C%d of %d %s" % (a[2], a[3], a[4], a[5], "nhwc" if a[7] else "nchw"), 0, 0
    if name in ("arf_bias_leaky_nhwc_fwd", "arf_bias_leaky_nhwc_bwd", "arf_bias_leaky_nhwc_bwd_ld",
                "arf_bias_leaky_nhwc_fwd_ld"):
        i = {"arf_bias_leaky_nhwc_fwd": 2, "arf_bias_leaky_nhwc_bwd": 5, "arf_bias_leaky_nhwc_bwd_ld": 7,
             "arf_bias_leaky_nhwc_fwd_ld": 4}[name]
        rows, C = a[i], a[i + 1]
        return "rows%d C%d" % (rows, C), 0, 0
    if name == "arf_nhwc_unpack_add":
        return "N%d HW%d C%d of %d" % (a[2], a[3], a[4], a[5]), 0, 0
    if name in ("arf_featnorm_fwd", "arf_featnorm_bwd"):
        B, n = (a[6], a[7]) if name.endswith("fwd") else (a[9], a[10])
        return "B%d n%d" % (B, n), 0, 0
    if name in ("arf_bias_leaky_fwd", "arf_bias_leaky_bwd"):
        B, C, HW = a[2:5] if name.endswith("fwd") else a[5:8]
        return "B%d C%d HW%d" % (B, C, HW), 0, 0
    if name in ("arf_conv3x3_small_fwd", "arf_conv3x3_small_bwd"):
        N, Hh, Ww, Ci, Co = a[4:9] if name.endswith("fwd") else a[6:11]
        return "N%d %dx%d %d->%d" % (N, Hh, Ww, Ci, Co), (2 if name.endswith("fwd") else 4) * N * Hh * Ww * 9 * Ci * Co, 0
    return "other", 0, 0


# ----------------------------------------------------------------------------- reference arm --
def run_cpu_reference(steps, warmup, batch):
    """The reference's path on the host CPU cores (oracle port; the Python reference cannot travel to the
    GPU box).  Each step = one full train step on `batch` synthetic pairs of 384x512."""
    import torch
    import oracle.arflow_oracle as orc
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its children)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:
        torch.set_num_threads(avail)
    cores = torch.get_num_threads()
    step = orc.CpuTrainStep(seed=0)
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(batch, 6, H, W, generator=gen)
    for _ in range(warmup):
        step(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": batch / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d train step(s) of %d pairs 384x512 after %d warm-up (oracle port of PWCFlow+UFlowLoss+Adam, "
                      "torch %s CPU, %d threads)" % (steps, batch, warmup, torch.__version__, cores)}, dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    cb, dt = run_cpu_reference(steps, warmup, args.cpu_batch)
    line = {"impl": "reference", "metric": "train pairs/s (chairs_uflow)", "value": cb["value"], "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_batch": args.cpu_batch},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm -------
def main_b200(args):
    import torch
    import torch.distributed as dist
    from arflow_b200 import _lib
    from arflow_b200.train_step import UFlowTrainStep
    from arflow_b200.uflow_loss import UFlowLoss
    from arflow_b200.uflow_model import PWCFlow

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device — the arflow_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark   # static shapes: let cuDNN pick its kernels
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B = args.batch
    torch.manual_seed(0)  # same weights on every rank (the reference broadcasts through DataParallel)
    model = PWCFlow(types.SimpleNamespace(level_dropout=0.1, feature_norm=True), nhwc=not args.nchw).to(dev)
    model.init_weights()
    model.train()
    loss_fn = UFlowLoss(types.SimpleNamespace(edge_constant=150, w_smooth=4.0, w_census=1.0, with_bk=True,
                                               smooth_order=1))
    step = UFlowTrainStep(model, loss_fn, lr=1e-4, use_graph=not args.no_graph, world_size=world)

    gen = torch.Generator().manual_seed(1000 + rank)
    n_host = 4
    host = [torch.rand(B, 6, H, W, generator=gen).pin_memory() for _ in range(n_host)]
    devb = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.profile_step:
        eager = UFlowTrainStep(model, loss_fn, lr=1e-4, use_graph=False, world_size=1)
        for i in range(3):
            eager(devb[i % n_host])
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        eager(devb[3])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled": "one eager train step", "launches": None}), flush=True)
        return

    # ---- warm-up (also captures the graph) ----
    for i in range(max(args.warmup, 3)):
        out = step(devb[i % n_host])
    barrier()
    launches_per_step = step.launches_per_step   # kernels of this library inside one graph replay

    # ---- timed region: inputs resident in HBM ----
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(devb[i % n_host])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    eager_launches = _lib.launch_count() - l0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    last = [float(v) for v in out.tolist()]

    # ---- end to end: pinned host -> device every step, loss read back every step ----
    # Like a data loader would, the copy of batch i+1 is issued on a copy stream while step i computes; every batch
    # still crosses PCIe inside the timed region and every step's loss is read back to the host before the next one.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(devb[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])       # the step that last read this buffer has finished
            bufs[i % 2].copy_(host[i % n_host], non_blocking=True)
            ready[i % 2].record(copy_stream)

    for ev in consumed:
        ev.record()
    barrier()
    t0 = time.perf_counter()
    issue_copy(0)
    for i in range(args.steps):
        torch.cuda.current_stream().wait_event(ready[i % 2])
        o = step(bufs[i % 2])
        consumed[i % 2].record()
        if i + 1 < args.steps:
            issue_copy(i + 1)
        _ = o[0].item()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())

    # ---- in-situ kernel timings (eager, events around each C-ABI call), rank 0 ----
    kernels, roof = {}, None
    launches = launches_per_step * args.steps if launches_per_step is not None else eager_launches
    if rank == 0:
        eager = UFlowTrainStep(model, loss_fn, lr=1e-4, use_graph=False, world_size=1) if world == 1 else None
        if eager is not None:
            eager(devb[0])
            _lib.profile_start()
            n_prof = 3
            for i in range(n_prof):
                eager(devb[i % n_host])
            rec = _lib.profile_stop()
            per_shape = {}
            for name, a, dt_ms in rec:
                if name.endswith(("_out_dims", "_num_partials")):
                    continue                       # host-only helpers, no launch
                k = kernels.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0})
                k["calls"] += 1
                k["ms"] += dt_ms
                k["bytes"] += alg_bytes(name, a)
                label, flops, mufu = alg_work(name, a)
                q = per_shape.setdefault((name, label), {"calls": 0, "ms": 0.0, "bytes": 0, "flops": 0, "mufu": 0})
                q["calls"] += 1
                q["ms"] += dt_ms
                q["bytes"] += alg_bytes(name, a)
                q["flops"] += flops
                q["mufu"] += mufu
            hbm, how = measured_peaks()
            # dominant = the (entry point, problem shape) with the most device time per step
            # (launch-latency-sized helpers such as the per-layer weight re-layout are not candidates: their event-bracketed
            # time is mostly launch gap)
            (name, label), k = max(((kk, vv) for kk, vv in per_shape.items()
                                    if vv["bytes"] > 0 and kk[1] != "other" and vv["ms"] / vv["calls"] > 0.015),
                                   key=lambda kv: kv[1]["ms"])
            ach = k["bytes"] / (k["ms"] * 1e-3) / 1e9
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get("%s [%s]" % (name, label), {}).get("dram_bytes")
            roof = {"kernel": "%s [%s]" % (name, label), "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                    "frac": ach / hbm, "traffic": traffic, "peak_source": how,
                    "avg_launch_us": 1e3 * k["ms"] / k["calls"], "bytes_per_launch": k["bytes"] / k["calls"],
                    "step_share": k["ms"] / n_prof / (ms_total / args.steps),
                    "fp32_tflops": k["flops"] / (k["ms"] * 1e-3) / 1e12,
                    "fp32_frac": k["flops"] / (k["ms"] * 1e-3) / 1e12 / FP32_FMA_TFLOPS,
                    "mufu_frac": k["mufu"] / (k["ms"] * 1e-3) / 1e9 / MUFU_GOPS,
                    "note": "algorithmic bytes / CUDA-event time of the C-ABI call, in situ (L2-warm, inside an eager step; "
                            "a C-ABI call may launch two kernels, e.g. the two correlation gradients). fp32_frac / "
                            "mufu_frac: the same call against the measured FP32-FMA (%.1f TFLOP/s) and MUFU (%.0f Gop/s) "
                            "peaks - the binding roof for the correlation gradients and the census kernels. "
                            "L2-cold numbers at the sweep sizes: profiles/" % (FP32_FMA_TFLOPS, MUFU_GOPS)}
            for name, k in kernels.items():
                k["us_per_step"] = 1e3 * k["ms"] / n_prof
                k["alg_GBps"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
                k["calls_per_step"] = k["calls"] / n_prof
                del k["ms"], k["bytes"], k["calls"]
            shapes = []
            for (name, label), q in sorted(per_shape.items(), key=lambda kv: -kv[1]["ms"])[:12]:
                sec = q["ms"] * 1e-3
                shapes.append({"call": "%s [%s]" % (name, label), "us_per_launch": 1e3 * q["ms"] / q["calls"],
                               "launches_per_step": q["calls"] / n_prof,
                               "hbm_frac": q["bytes"] / sec / 1e9 / hbm,
                               "fp32_frac": q["flops"] / sec / 1e12 / FP32_FMA_TFLOPS,
                               "mufu_frac": q["mufu"] / sec / 1e9 / MUFU_GOPS})
            kernels["_by_shape"] = shapes

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = run_cpu_reference(1, 1, args.cpu_batch)

    if rank == 0:
        gb = B * world
        line = {"metric": "train pairs/s (chairs_uflow)", "value": gb * args.steps / (ms_total * 1e-3),
                "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": gb, "per_gpu_batch": B, "height": H, "width": W,
                           "parallelism": "dp%d" % world, "cuda_graph": not args.no_graph, "conv_layout": "nchw" if args.nchw else "nhwc (channels-last conv stacks, arflow_b200/fused_conv.py)",
                           "conv_math": "cuDNN fp32 tensors, torch default allow_tf32=%s, cudnn.benchmark=%s"
                                        % (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark),
                           "l2": "per-step working set (activations) is several GB >> 126 MB L2; 4 input batches rotate",
                           "loss_last_step": last},
                "clocks": clocks,
                "e2e": {"value": gb * args.steps / e2e_s, "unit": "pairs/s",
                        "h2d_bytes_per_step": B * 6 * H * W * 4, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)

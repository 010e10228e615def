#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its five configurations; the default is the headline (configs[1]).

    python bench.py --gpus N --steps K --warmup W                   # config 2: chairs_uflow train step, weak scaling
    python bench.py --config {1,3,4,5} [--gpus N] ...               # the other configurations (profiles/r2_config*.json)
    python bench.py --impl reference [--config C] --gpus N ...      # the reference path on the host CPUs

  config 2  chairs_uflow: PWCFlow + UFlowLoss + Adam, 384x512, batch 8 per GPU (weak scaling)          [headline]
  config 4  kitti_uflow: same step, 320x1024, smooth_order 2, GLOBAL batch 32 sharded over N GPUs (strong scaling)
  config 3  sintel_uflow_elbo non-diagonal: PWCProbFlow[2,2,30] + UFlowElboLoss(sparse, k=3, 4 samples), 448x1024, batch 8
  config 1  PWC-Lite two-view inference 384x640, batch 1 (replicas only for N > 1)
  config 5  correlation + warp kernel sweep (one GPU)

A step = one pass of the hot path's caller over one synthetic batch.  Rank 0 prints ONE JSON line:
  value          pairs/s with the inputs already in HBM (CUDA-graph replay where the step is capturable, CUDA events,
                 max over ranks)
  e2e            pairs/s through the public API with HOST inputs: pinned-host -> device copy of every batch and a
                 device -> host read of the step's result inside the timed region
  roofline       the dominant HOT-PATH kernel of the step (SURVEY §8a entry points only: correlation, warp, census,
                 smoothness, masks, stencil, normalize), timed in situ with CUDA events around its C-ABI call on the
                 launching stream during extra eager steps; against the roof that binds it (measured peaks)
  roofline_hotpath  correlation / warp / census forward and backward at the step's finest-level shapes, timed alone and
                 L2-cold (CUDA graph over rotating buffer sets), each with its HBM, FP32-FMA and MUFU fractions
  cpu_baseline   the reference path (oracle/: torch-CPU restatement, no arflow_b200 code) on the host cores, bounded sample
Peaks: HBM from MEASURED_PEAKS.json (driver-written); FP32 FMA, MUFU, shared-memory and mma.sync peaks measured live in
this process by tools/libarf_peaks.so (tools/peaks.cu).
"""
import argparse
import ctypes
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

CONFIGS = {
    1: dict(H=384, W=640, batch=1, metric="inference pairs/s (PWC-Lite two-view)", scaling="weak",
            workload="PWC-Lite two-view inference 384x640, batch 1, synthetic pair, random init (BASELINE configs[0])"),
    2: dict(H=384, W=512, batch=8, metric="train pairs/s (chairs_uflow)", scaling="weak", smooth_order=1,
            workload="chairs_uflow PWCFlow+UFlowLoss train step, 384x512, batch 8 per GPU, synthetic pairs, random init"),
    3: dict(H=448, W=1024, batch=8, metric="train pairs/s (sintel_uflow_elbo non-diagonal)", scaling="weak",
            workload="sintel_uflow_elbo non-diagonal covariance (PWCProbFlow[2,2,30] + UFlowElboLoss sparse k=3, 4 samples) "
                     "train step, 448x1024, batch 8 per GPU, synthetic pairs, random init (BASELINE configs[2])"),
    4: dict(H=320, W=1024, batch=32, metric="train pairs/s (kitti_uflow, global batch 32)", scaling="strong", smooth_order=2,
            workload="kitti_uflow PWCFlow+UFlowLoss(smooth_order=2) train step, 320x1024, GLOBAL batch 32 sharded over the "
                     "GPUs, synthetic pairs, random init (BASELINE configs[3])"),
    5: dict(H=0, W=0, batch=0, metric="correlation+warp sweep, geometric-mean fraction of the binding roof", scaling="weak",
            workload="correlation + warp kernel sweep C in {32,64,96,128,192,196}, md=4, pyramid levels 1/4..1/64 of "
                     "384x512 and 448x1024, batch 1..64 (BASELINE configs[4])"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the named config)")
    ap.add_argument("--cpu-batch", type=int, default=None, help="pairs per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hotpath", action="store_true", help="skip the isolated hot-path kernel timings")
    ap.add_argument("--nchw", action="store_true", help="experiment: NCHW conv stacks (the default is channels-last)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true",
                    help="disable cuDNN autotuning of the (out-of-scope) convolutions; about 11 %% faster with it")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nccl", "peer", "fused"],
                    help="gradient all-reduce of the multi-GPU step: the library's own NVLink kernel captured in the step "
                         "graph (fused), or NCCL between two graphs")
    ap.add_argument("--comm-ctas", type=int, default=None, help="CTAs of an overlapped bucket all-reduce (peer mode; tuning)")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`) and exit")
    return ap.parse_args()


# ----------------------------------------------------------------------------- peaks ---------
def measured_peaks(live=True):
    """HBM from the driver's file; FMA / MUFU / LDS / mma.sync measured live on this GPU (tools/peaks.cu)."""
    pk = {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)", "fp32_tflops": 71.3, "mufu_gops": 4544.0,
          "compute_source": "fallback constants (round-1 measurement with tools/fma_peak.cu)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        pk["hbm_gbs"] = json.load(open(p)).get("hbm_gbs", 6650.0)
        pk["hbm_source"] = "measured (MEASURED_PEAKS.json, burst copy)"
    so = os.path.join(ROOT, "tools", "libarf_peaks.so")
    if live and os.path.exists(so):
        try:
            lib = ctypes.CDLL(so)
            out = (ctypes.c_double * 8)()
            if lib.arf_peaks_measure(out) == 0:
                pk.update(fp32_tflops=max(out[0], out[1]), fp32_ffma2_tflops=out[1], mufu_gops=out[2], lds_gbs=out[3],
                          mma_sync_tf32_tflops=out[4], sms=int(out[6]),
                          compute_source="measured live in this process (tools/peaks.cu: FFMA best of two loops, MUFU.RSQ)")
        except OSError:
            pass
    return pk


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.  Started before the warm-up (nvidia-smi needs ~0.5 s to
    print its first row); `mark()` opens the window at the start of the timed region and `stop()` closes it, so only rows
    taken under the timed load are reported.  A region shorter than three samples is widened by the rows just before it
    (the warm-up steps, same load)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.first = 0
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

    def mark(self):
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = self.rows[self.first:]
        if len(rows) < 3:
            rows = self.rows[max(0, len(self.rows) - 3):]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- algorithmic work -----
# SURVEY §8(a) entry points: the only candidates for `roofline` (conv glue such as arf_bias_leaky_* / arf_nhwc_* is
# timed and listed under `kernels`, but it is not the hot path the metric names).
HOTPATH = ("arf_corr_fwd", "arf_corr_bwd", "arf_warp_fwd", "arf_warp_bwd", "arf_census_fwd", "arf_census_bwd",
           "arf_census_fwd_groups", "arf_census_bwd_groups",
           "arf_smooth_fwd", "arf_smooth_bwd", "arf_range_map", "arf_range_map_bwd", "arf_inside_mask",
           "arf_count_to_mask", "arf_occ_bidir", "arf_stencil_mv_fwd", "arf_stencil_mv_bwd", "arf_trisolve",
           "arf_ssim_fwd", "arf_ssim_bwd", "arf_featnorm_fwd", "arf_featnorm_bwd", "arf_resize_bilinear_fwd",
           "arf_resize_bilinear_bwd", "arf_corr_level_fwd")
CENSUS_MUFU_FWD = 1.5   # MUFU per pixel and offset: each unordered pixel pair is evaluated once (3 MUFU per pair)
CENSUS_MUFU_BWD = 1.5   # same for the backward: one evaluation per unordered pair, +X to one end and -X to the other


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
    if name in ("arf_census_fwd", "arf_census_fwd_groups"):
        B, Hh, Ww = a[6:9]
        return B * Hh * Ww * 32
    if name in ("arf_census_bwd", "arf_census_bwd_groups"):
        B, Hh, Ww = a[9:12]
        return B * Hh * Ww * (24 + 8 + 12 * ((a[7] is not None) + (a[8] is not None)))
    if name in ("arf_resize_bilinear_fwd", "arf_resize_bilinear_bwd"):
        n, Hi, Wi, Ho, Wo = a[2:7]
        return n * (Hi * Wi + Ho * Wo) * 4
    if name in ("arf_smooth_fwd", "arf_smooth_bwd"):
        B, Ci, Hh, Ww = a[4:8]
        return B * Hh * Ww * (4 * Ci + 8 + (8 if name.endswith("bwd") else 0))
    if name in ("arf_range_map", "arf_range_map_bwd"):
        B, Hh, Ww = a[2:5] if name == "arf_range_map" else a[3:6]
        return B * Hh * Ww * 16
    if name == "arf_inside_mask":
        B, Hh, Ww = a[2:5]
        return B * Hh * Ww * 12
    if name == "arf_count_to_mask":
        return a[2] * 8
    if name == "arf_stencil_mv_fwd":
        N, Hh, Ww, k = a[3:7]
        return N * Hh * Ww * (8 * (k + 1) ** 2 + 16)
    if name == "arf_stencil_mv_bwd":
        N, Hh, Ww, k = a[5:9]
        return N * Hh * Ww * (16 * (k + 1) ** 2 + 24)
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
    if name in ("arf_census_fwd", "arf_census_fwd_groups"):
        B, Hh, Ww = a[6:9]
        patch = a[9] if name == "arf_census_fwd" else a[10]
        n = patch * patch - 1
        return "B%d %dx%d p%d" % (B, Hh, Ww, patch), B * Hh * Ww * 7 * n, int(B * Hh * Ww * CENSUS_MUFU_FWD * n)
    if name in ("arf_census_bwd", "arf_census_bwd_groups"):
        B, Hh, Ww = a[9:12]
        patch = a[12] if name == "arf_census_bwd" else a[13]
        n = patch * patch - 1
        return "B%d %dx%d p%d" % (B, Hh, Ww, patch), B * Hh * Ww * 22 * n, int(B * Hh * Ww * CENSUS_MUFU_BWD * n)
    if name in ("arf_resize_bilinear_fwd", "arf_resize_bilinear_bwd"):
        n, Hi, Wi, Ho, Wo = a[2:7]
        return "N%d %dx%d->%dx%d" % (n, Hi, Wi, Ho, Wo), 0, 0
    if name in ("arf_smooth_fwd", "arf_smooth_bwd"):
        B, Ci, Hh, Ww = a[4:8]
        return "B%d C%d %dx%d" % (B, Ci, Hh, Ww), 0, 0
    if name in ("arf_stencil_mv_fwd", "arf_stencil_mv_bwd"):
        N, Hh, Ww, k = a[3:7] if name.endswith("fwd") else a[5:9]
        return "N%d %dx%d k%d" % (N, Hh, Ww, k), (4 if name.endswith("fwd") else 8) * N * Hh * Ww * (k + 1) ** 2, 0
    if name in ("arf_range_map", "arf_range_map_bwd", "arf_inside_mask"):
        B, Hh, Ww = a[3:6] if name == "arf_range_map_bwd" else a[2:5]
        return "B%d %dx%d" % (B, Hh, Ww), 0, 0
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


def roof_entry(label, nbytes, flops, mufu, sec, pk, traffic=None, extra=None):
    """One roofline record: achieved rate and fraction against the roof that binds this kernel at this shape."""
    t_h = nbytes / (pk["hbm_gbs"] * 1e9)
    t_f = flops / (pk["fp32_tflops"] * 1e12)
    t_m = mufu / (pk["mufu_gops"] * 1e9)
    bound = max((t_h, "hbm"), (t_f, "fp32_fma"), (t_m, "mufu"))[1]
    if bound == "hbm":
        ach, peak, unit = nbytes / sec / 1e9, pk["hbm_gbs"], "GB/s"
    elif bound == "fp32_fma":
        ach, peak, unit = flops / sec / 1e12, pk["fp32_tflops"], "TFLOP/s"
    else:
        ach, peak, unit = mufu / sec / 1e9, pk["mufu_gops"], "Gop/s"
    r = {"kernel": label, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
         "traffic": traffic, "us": sec * 1e6, "hbm_frac": nbytes / sec / 1e9 / pk["hbm_gbs"],
         "fp32_frac": flops / sec / 1e12 / pk["fp32_tflops"], "mufu_frac": mufu / sec / 1e9 / pk["mufu_gops"],
         "bytes_per_launch": nbytes}
    if extra:
        r.update(extra)
    return r


def load_traffic():
    for n in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", n)
        if os.path.exists(p):
            return json.load(open(p)), n
    return {}, None


def hotpath_rooflines(B2, C, h, w, Bimg, H, W, pk):
    """corr / warp / census forward and backward at the step's finest-level shapes, each alone and L2-cold."""
    import torch
    from arflow_b200 import _lib
    from tools.microbench import time_graph
    lib = _lib.load()
    cs = lambda: torch.cuda.current_stream().cuda_stream
    traffic, tsrc = load_traffic()
    out = []

    def add(name, label, nbytes, flops, mufu, make):
        med, _ = time_graph(make, nbytes)
        t = traffic.get("%s [%s]" % (name, label), {}).get("dram_bytes")
        out.append(roof_entry("%s [%s]" % (name, label), nbytes, flops, mufu, med, pk, t,
                              {"timing": "alone, L2-cold (rotating buffers inside one CUDA graph)"}))

    px = B2 * h * w
    lab = "B%d C%d %dx%d" % (B2, C, h, w)

    def mk_cf():
        f1, f2 = torch.randn(B2, C, h, w, device="cuda"), torch.randn(B2, C, h, w, device="cuda")
        o = torch.empty(B2, 81, h, w, device="cuda")
        return lambda: lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), o.data_ptr(), B2, C, h, w, 4, 1, 4, 1, 1, cs())

    def mk_cb():
        f1, f2 = torch.randn(B2, C, h, w, device="cuda"), torch.randn(B2, C, h, w, device="cuda")
        go = torch.randn(B2, 81, h, w, device="cuda")
        g1, g2 = torch.empty_like(f1), torch.empty_like(f2)
        return lambda: lib.arf_corr_bwd(f1.data_ptr(), f2.data_ptr(), go.data_ptr(), g1.data_ptr(), g2.data_ptr(),
                                        B2, C, h, w, 4, 1, 4, 1, 1, cs())
    add("arf_corr_fwd", lab, px * (8 * C + 324), px * C * 162, 0, mk_cf)
    add("arf_corr_bwd", lab, px * (16 * C + 324), px * C * 324, 0, mk_cb)
    wa = (B2, C, h, w, h, w, float(w - 1), float(h - 1), 0, 0, 0, 1)

    def mk_w(kind):
        def make():
            x = torch.randn(B2, C, h, w, device="cuda")
            fl = torch.nn.functional.interpolate(torch.randn(B2, 2, max(h // 8, 1), max(w // 8, 1), device="cuda") * 2,
                                                 size=(h, w), mode="bilinear").contiguous()
            y, gy = torch.empty_like(x), torch.randn_like(x)
            gx, gf = torch.empty_like(x), torch.empty_like(fl)
            if kind == "fwd":
                return lambda: lib.arf_warp_fwd(x.data_ptr(), fl.data_ptr(), y.data_ptr(), *wa, cs())
            if kind == "bwd":
                return lambda: lib.arf_warp_bwd(x.data_ptr(), fl.data_ptr(), gy.data_ptr(), gx.data_ptr(), gf.data_ptr(), *wa, cs())
            return lambda: lib.arf_warp_bwd(x.data_ptr(), fl.data_ptr(), gy.data_ptr(), None, gf.data_ptr(), *wa, cs())
        return make
    add("arf_warp_fwd", lab + " smooth flow", px * (8 * C + 8), px * C * 8, 0, mk_w("fwd"))
    add("arf_warp_bwd", lab + " smooth flow", px * (12 * C + 16), px * C * 16, 0, mk_w("bwd"))
    add("arf_warp_bwd", lab + " smooth flow, flow-grad only", px * (8 * C + 16), px * C * 16, 0, mk_w("bwdF"))
    pxi = Bimg * H * W
    npart = lib.arf_census_num_partials(Bimg, H, W)

    def mk_c(kind):
        def make():
            a, b = torch.rand(Bimg, 3, H, W, device="cuda"), torch.rand(Bimg, 3, H, W, device="cuda")
            m = torch.rand(Bimg, 1, H, W, device="cuda")
            ham = torch.empty(Bimg, 1, H, W, device="cuda")
            part = torch.empty(2 * npart, device="cuda")
            ng = 2 if Bimg % 2 == 0 else 1
            sums, gl = torch.ones(3 * ng, device="cuda"), torch.ones(ng, device="cuda")
            gb = torch.empty_like(b)
            if kind == "fwd":
                return lambda: lib.arf_census_fwd_groups(a.data_ptr(), b.data_ptr(), m.data_ptr(), ham.data_ptr(), part.data_ptr(),
                                                         sums.data_ptr(), Bimg, H, W, ng, 7, 1.0, 0.01, 0.4, cs())
            lib.arf_census_fwd_groups(a.data_ptr(), b.data_ptr(), m.data_ptr(), ham.data_ptr(), part.data_ptr(), sums.data_ptr(),
                                      Bimg, H, W, ng, 7, 1.0, 0.01, 0.4, cs())
            return lambda: lib.arf_census_bwd_groups(a.data_ptr(), b.data_ptr(), None, ham.data_ptr(), m.data_ptr(), sums.data_ptr(),
                                                     gl.data_ptr(), None, gb.data_ptr(), Bimg, H, W, ng, 7, 1.0, 0.01, 0.4, cs())
        return make
    labc = "B%d %dx%d p7" % (Bimg, H, W)
    add("arf_census_fwd", labc, pxi * 32, pxi * 7 * 48, int(pxi * CENSUS_MUFU_FWD * 48), mk_c("fwd"))
    add("arf_census_bwd", labc, pxi * 44, pxi * 22 * 48, int(pxi * CENSUS_MUFU_BWD * 48), mk_c("bwd"))
    return out, tsrc


# ----------------------------------------------------------------------------- CPU arm --------
def host_threads():
    import torch
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:    # torchrun exports OMP_NUM_THREADS=1 for its children
        torch.set_num_threads(avail)
    return torch.get_num_threads()


def run_cpu_reference(cfg_id, steps, warmup, batch, budget_s=150.0):
    """The reference's path for this configuration on the host CPU cores (oracle/: a torch-CPU restatement; the Python
    reference itself cannot travel to the GPU box).  Each step = the configuration's step on `batch` synthetic pairs;
    the step count is cut so that the run stays inside `budget_s`."""
    import torch
    import oracle.arflow_oracle as orc
    cfg = CONFIGS[cfg_id]
    cores = host_threads()
    gen = torch.Generator().manual_seed(0)
    H, W = cfg["H"], cfg["W"]
    if cfg_id in (2, 4):
        step = orc.CpuTrainStep(smooth_order=cfg["smooth_order"], seed=0)
        what = "train step(s) (oracle restatement of PWCFlow+UFlowLoss+Adam)"
    elif cfg_id == 1:
        step = orc.CpuPwcLiteInference(seed=0)
        what = "two-view inference(s) (oracle restatement of PWCLite via correlation_native + flow_warp)"
    else:
        return None, None
    x = torch.rand(batch, 6, H, W, generator=gen)
    t0 = time.perf_counter()
    for _ in range(max(warmup, 1)):
        step(x)
    per = (time.perf_counter() - t0) / max(warmup, 1)
    done = max(1, min(steps, int(budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(done):
        step(x)
    dt = (time.perf_counter() - t0) / done
    return {"value": batch / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d %s of %d pair(s) %dx%d after %d warm-up, torch %s CPU, %d threads"
                      % (done, what, batch, H, W, max(warmup, 1), torch.__version__, cores)}, (dt, done)


def cpu_sweep(budget_s=30.0):
    """correlation_native + flow_warp (oracle restatements) fwd and fwd+bwd on config-5 shapes, host cores."""
    import torch
    import oracle.arflow_oracle as orc
    cores = host_threads()
    rows = []
    t_all = time.perf_counter()
    for (B, C, h, w) in [(1, 32, 96, 128), (8, 32, 96, 128), (1, 64, 48, 64), (1, 128, 12, 16), (1, 192, 6, 8)]:
        if time.perf_counter() - t_all > budget_s:
            break
        g = torch.Generator().manual_seed(0)
        f1 = torch.randn(B, C, h, w, generator=g, requires_grad=True)
        f2 = torch.randn(B, C, h, w, generator=g, requires_grad=True)
        fl = (torch.randn(B, 2, h, w, generator=g) * 2).requires_grad_(True)

        def best(fn, n=3):
            fn()
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return min(ts)
        t_cf = best(lambda: orc.cost_volume(f1.detach(), f2.detach(), 4))
        t_cfb = best(lambda: orc.cost_volume(f1, f2, 4).square().sum().backward())
        t_wf = best(lambda: orc.warp(f2.detach(), fl.detach(), kind="flow"))
        t_wfb = best(lambda: orc.warp(f2, fl, kind="flow").square().sum().backward())
        rows.append({"shape": [B, C, h, w], "corr_fwd_ms": t_cf * 1e3, "corr_fwd_bwd_ms": t_cfb * 1e3,
                     "warp_fwd_ms": t_wf * 1e3, "warp_fwd_bwd_ms": t_wfb * 1e3})
    return {"cores": cores, "kind": "port", "rows": rows,
            "sample": "oracle restatement of correlation_native.Correlation and flow_warp, best of 3 after 1 warm-up"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    base = {"impl": "reference", "metric": cfg["metric"], "unit": "pairs/s", "n_gpus": args.gpus, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gpu_launches": 0}
    if args.config == 5:
        sw = cpu_sweep()
        t = sum(r["corr_fwd_bwd_ms"] + r["warp_fwd_bwd_ms"] for r in sw["rows"])
        line = dict(base, value=len(sw["rows"]) / (t * 1e-3) if t else 0.0, unit="shapes/s", steps=1, warmup=1, ms_per_step=t,
                    config={"workload": cfg["workload"]}, cpu_baseline=sw,
                    e2e={"value": len(sw["rows"]) / (t * 1e-3) if t else 0.0, "unit": "shapes/s", "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    if args.config == 3:
        print(json.dumps(dict(base, unavailable="no CPU restatement of PWCProbFlow + UFlowElboLoss in oracle/ (SURVEY App. B: the "
                                                 "reference itself needs 19 s/step at batch 2 and > 60 GB at batch 8)",
                              config={"workload": cfg["workload"]})), flush=True)
        return
    # same configuration as the B200 arm (its per-GPU batch); --cpu-batch bounds the sample on a slow host
    batch = args.cpu_batch if args.cpu_batch is not None else (1 if args.config == 1 else 8)
    cb, (dt, done) = run_cpu_reference(args.config, max(1, args.steps), max(1, min(args.warmup, 2)), batch)
    line = dict(base, value=cb["value"], steps=done, warmup=max(1, min(args.warmup, 2)), ms_per_step=dt * 1e3,
                config={"workload": cfg["workload"], "sample_batch": batch,
                        "note": "each step processes sample_batch pairs of the configuration's shape; the number of timed "
                                "steps is cut to a 150 s budget"},
                cpu_baseline=cb,
                e2e={"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm -------
def build_step(args, dev, world):
    """Returns (step callable batch->tensor, eager factory or None, per-GPU batch, input channels, model)."""
    import torch
    cfg = CONFIGS[args.config]
    cid = args.config
    if cid in (2, 4):
        from arflow_b200.train_step import UFlowTrainStep
        from arflow_b200.uflow_loss import UFlowLoss
        from arflow_b200.uflow_model import PWCFlow
        B = args.batch if args.batch else (cfg["batch"] if cid == 2 else max(1, cfg["batch"] // world))
        torch.manual_seed(0)  # same weights on every rank (the reference broadcasts through DataParallel)
        model = PWCFlow(types.SimpleNamespace(level_dropout=0.1, feature_norm=True), nhwc=not args.nchw).to(dev)
        model.init_weights()
        model.train()
        loss_fn = UFlowLoss(types.SimpleNamespace(edge_constant=150, w_smooth=4.0, w_census=1.0, with_bk=True,
                                                   smooth_order=cfg["smooth_order"]))
        # N > 1: the census loss keeps the reference's batch-global normaliser (uflow_utils.py:293) through a 16-byte
        # peer all-reduce inside the captured step, so every N optimises the same loss as the single-process reference
        extra = {"comm_ctas": args.comm_ctas} if args.comm_ctas else {}
        step = UFlowTrainStep(model, loss_fn, lr=1e-4, use_graph=not args.no_graph, world_size=world,
                              allreduce=args.allreduce, global_census_norm=(world > 1 and args.allreduce != "nccl"), **extra)
        import copy
        fresh = copy.deepcopy(model)      # in-situ kernel timings run on the initial weights, see main_b200

        def eager():
            return UFlowTrainStep(fresh, loss_fn, lr=1e-4, use_graph=False, world_size=1)
        return step, eager, B, 6, model
    if cid == 3:
        from arflow_b200.uflow_elbo_loss import UFlowElboLoss
        from arflow_b200.uflow_prob_model import PWCProbFlow
        B = args.batch if args.batch else cfg["batch"]
        torch.manual_seed(0)
        net = PWCProbFlow(types.SimpleNamespace(out_channels=[2, 2, 30], inv_cov=False, n_pyramids=1, mixture_weights=False,
                                                feature_norm=True, level_dropout=0.1)).to(dev).train()
        lcfg = dict(edge_constant=150, edge_asymp=0.01, w_smooth=4.0, penalty_smooth="charbonnier", closed_form_smooth=False,
                    data_loss=["census"], data_weight=[1.0], data_penalty=["abs_robust_loss"], w_entropy=0.1, w_oof=0.0,
                    w_occ=0.0, with_bk=True, approx="sparse", n_components=1, cov_supp=3, inv_cov=False,
                    approx_entropy=False, occ_type="sample", n_samples=4, offdiag_reg=0.0, natural_grad=False,
                    isotropic_smooth=False)
        loss_fn = UFlowElboLoss(types.SimpleNamespace(**lcfg))
        params = [p for p in net.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-4, fused=True, capturable=True)
        import torch.distributed as dist

        def step(x):
            for p in params:
                p.grad = None
            im1, im2 = x[:, :3].contiguous(), x[:, 3:].contiguous()
            out = loss_fn(net(im1, im2, with_bk=True), im1, im2)
            out[0].backward()
            if world > 1:
                for p in params:
                    if p.grad is not None:
                        dist.all_reduce(p.grad, op=dist.ReduceOp.AVG)
            opt.step()
            return torch.stack([out[0].detach()])
        return step, (lambda: step), B, 6, net
    if cid == 1:
        from arflow_b200.pwclite import PWCLite
        B = args.batch if args.batch else cfg["batch"]
        torch.manual_seed(0)
        net = PWCLite(types.SimpleNamespace(upsample=True, n_frames=2, reduce_dense=True)).to(dev).eval()
        state = {"graph": None, "in": None, "out": None}

        def eager_step(x):
            with torch.no_grad():
                flow = net(x, with_bk=False)['flows_fw'][0]
            return flow

        def step(x):
            if args.no_graph:
                return eager_step(x).abs().mean().reshape(1)
            if state["graph"] is None:
                state["in"] = x.clone()
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(3):
                        eager_step(state["in"])
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                state["graph"] = torch.cuda.CUDAGraph()
                from arflow_b200 import _lib
                n0 = _lib.launch_count()
                with torch.cuda.graph(state["graph"]):
                    state["out"] = eager_step(state["in"]).abs().mean().reshape(1)
                step.launches_per_step = _lib.launch_count() - n0
            state["in"].copy_(x, non_blocking=True)
            state["graph"].replay()
            return state["out"]
        step.launches_per_step = None
        return step, (lambda: (lambda x: eager_step(x))), B, 6, net
    raise ValueError(cid)


def main_sweep(args):
    """Config 5: the kernel sweep (one GPU).  value = geometric mean over the sweep of each kernel's fraction of the roof
    that binds it at that shape."""
    import math
    import torch
    from arflow_b200 import _lib
    from tools.microbench import time_graph
    lib = _lib.load()
    torch.cuda.set_device(0)
    pk = measured_peaks()
    cs = lambda: torch.cuda.current_stream().cuda_stream
    shapes = []
    for (H, W) in ((384, 512), (448, 1024)):
        for lvl, C in ((4, 32), (8, 64), (16, 96), (32, 128), (64, 192)):
            h, w = H // lvl, W // lvl
            if h <= 4:
                continue
            for B in (1, 8, 64):
                if B * C * h * w * 4 * 3 + B * 81 * h * w * 4 < 6e9:
                    shapes.append((B, C, h, w))
    shapes += [(16, 196, 12, 16), (2, 196, 6, 8)]
    rows, l0 = [], _lib.launch_count()
    sampler = ClockSampler(0)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for (B, C, h, w) in shapes:
        px = B * h * w

        def mk(kind):
            def make():
                f1, f2 = torch.randn(B, C, h, w, device="cuda"), torch.randn(B, C, h, w, device="cuda")
                if kind == "cf":
                    o = torch.empty(B, 81, h, w, device="cuda")
                    return lambda: lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), o.data_ptr(), B, C, h, w, 4, 1, 4, 1, 1, cs())
                if kind == "cb":
                    go = torch.randn(B, 81, h, w, device="cuda")
                    g1, g2 = torch.empty_like(f1), torch.empty_like(f2)
                    return lambda: lib.arf_corr_bwd(f1.data_ptr(), f2.data_ptr(), go.data_ptr(), g1.data_ptr(), g2.data_ptr(),
                                                    B, C, h, w, 4, 1, 4, 1, 1, cs())
                fl = torch.randn(B, 2, h, w, device="cuda") * 2
                wa = (B, C, h, w, h, w, float(w - 1), float(h - 1), 0, 0, 0, 1)
                if kind == "wf":
                    return lambda: lib.arf_warp_fwd(f1.data_ptr(), fl.data_ptr(), f2.data_ptr(), *wa, cs())
                gx, gf = torch.empty_like(f1), torch.empty_like(fl)
                return lambda: lib.arf_warp_bwd(f1.data_ptr(), fl.data_ptr(), f2.data_ptr(), gx.data_ptr(), gf.data_ptr(), *wa, cs())
            return make
        for kind, name, nb, fl in (("cf", "corr_fwd", px * (8 * C + 324), px * C * 162), ("cb", "corr_bwd", px * (16 * C + 324), px * C * 324),
                                   ("wf", "warp_fwd", px * (8 * C + 8), px * C * 8), ("wb", "warp_bwd", px * (12 * C + 16), px * C * 16)):
            med, _ = time_graph(mk(kind), nb, reps=2)
            r = roof_entry("%s %dx%dx%dx%d" % (name, B, C, h, w), nb, fl, 0, med, pk)
            rows.append({k: r[k] for k in ("kernel", "us", "bound", "frac", "hbm_frac", "fp32_frac")})
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = _lib.launch_count() - l0
    gm = math.exp(sum(math.log(max(r["frac"], 1e-9)) for r in rows) / len(rows))
    big = [r for r in rows if r["us"] >= 20.0]
    gm_big = math.exp(sum(math.log(max(r["frac"], 1e-9)) for r in big) / max(len(big), 1))
    cpu = None if args.no_cpu_baseline else cpu_sweep()
    line = {"metric": CONFIGS[5]["metric"], "value": gm, "unit": "fraction of roof", "n_gpus": 1, "steps": len(rows),
            "warmup": 3, "ms_per_step": wall * 1e3 / len(rows), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[5]["workload"], "l2": "L2-cold: rotating buffer sets > 2.5x L2 inside one CUDA graph",
                       "geomean_frac_launches_over_20us": gm_big, "n_shapes": len(shapes)},
            "clocks": clocks, "e2e": None, "gpu_launches": int(launches), "peaks": pk,
            "roofline": max(rows, key=lambda r: r["us"]), "sweep": rows, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main_b200(args):
    import torch
    import torch.distributed as dist
    from arflow_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device — the arflow_b200 path has no CPU fallback")
    if args.config == 5:
        if rank == 0:
            main_sweep(args)
        return
    cfg = CONFIGS[args.config]
    H, W = cfg["H"], cfg["W"]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark   # static shapes: let cuDNN pick its kernels
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    step, eager_factory, B, cin, model = build_step(args, dev, world)
    gen = torch.Generator().manual_seed(1000 + rank)
    n_host = 4
    host = [torch.rand(B, cin, H, W, generator=gen).pin_memory() for _ in range(n_host)]
    devb = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.profile_step:
        eager = eager_factory()
        for i in range(3):
            eager(devb[i % n_host])
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        eager(devb[3])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled": "one eager step of config %d" % args.config, "launches": None}), flush=True)
        return

    # ---- warm-up (also captures the graph) ----
    sampler = ClockSampler(local) if rank == 0 else None
    W_ = max(args.warmup, 3)
    for i in range(W_):
        out = step(devb[i % n_host])
    barrier()
    launches_per_step = getattr(step, "launches_per_step", None)   # kernels of this library inside one graph replay

    # ---- timed region: inputs resident in HBM ----
    l0 = _lib.launch_count()
    barrier()
    if sampler:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(devb[i % n_host])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    eager_launches = _lib.launch_count() - l0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    last = [float(v) for v in out.flatten().tolist()]

    # ---- end to end: pinned host -> device every step, result read back every step ----
    # Like a data loader would, the copy of batch i+1 is issued on a copy stream while step i computes; every batch
    # still crosses PCIe inside the timed region and every step's result is read back to the host before the next one.
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
        _ = o.flatten()[0].item()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())
    clocks = sampler.stop() if sampler else None     # window = timed region + end-to-end region (the same steps)

    # ---- in-situ kernel timings (eager, events around each C-ABI call), rank 0, single GPU ----
    kernels, roof, hot, pk = {}, None, None, None
    launches = launches_per_step * args.steps if launches_per_step is not None else eager_launches
    if rank == 0 and world == 1:
        pk = measured_peaks()
        eager = eager_factory()
        eager(devb[0])
        _lib.profile_start()
        n_prof = 3
        for i in range(n_prof):
            eager(devb[i % n_host])
        rec = _lib.profile_stop()
        per_shape = {}
        for name, a, dt_ms in rec:
            if name.endswith(("_out_dims", "_num_partials", "_workspace")):
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
        traffic, tsrc = load_traffic()
        # dominant = the HOT-PATH (entry point, problem shape) with the most device time per step
        cand = [(kk, vv) for kk, vv in per_shape.items() if kk[0] in HOTPATH and vv["bytes"] > 0]
        if cand:
            (name, label), k = max(cand, key=lambda kv: kv[1]["ms"])
            sec = k["ms"] * 1e-3 / k["calls"]
            roof = roof_entry("%s [%s]" % (name, label), k["bytes"] / k["calls"], k["flops"] / k["calls"],
                              k["mufu"] / k["calls"], sec, pk,
                              traffic.get("%s [%s]" % (name, label), {}).get("dram_bytes"))
            roof["peak_source"] = pk["hbm_source"] if roof["bound"] == "hbm" else pk["compute_source"]
            roof["avg_launch_us"] = roof.pop("us")
            roof["launches_per_step"] = k["calls"] / n_prof
            roof["step_share"] = k["ms"] / n_prof / (ms_total / args.steps)
            roof["traffic_source"] = tsrc
            roof["note"] = ("dominant SURVEY-8(a) entry point of the step by device time; algorithmic bytes (flops) / CUDA-event "
                            "time of the C-ABI call, in situ (L2-warm, inside eager steps of a copy of the freshly initialised "
                            "network; one C-ABI call may launch more than one kernel). bound = whichever of HBM bytes, FP32 flops, MUFU ops takes longest at the measured "
                            "peaks; all three fractions are given.")
        for name, k in kernels.items():
            k["us_per_step"] = 1e3 * k["ms"] / n_prof
            k["alg_GBps"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
            k["calls_per_step"] = k["calls"] / n_prof
            k["hot_path"] = name in HOTPATH
            del k["ms"], k["bytes"], k["calls"]
        shapes = []
        for (name, label), q in sorted(per_shape.items(), key=lambda kv: -kv[1]["ms"])[:14]:
            sec = q["ms"] * 1e-3
            shapes.append({"call": "%s [%s]" % (name, label), "hot_path": name in HOTPATH,
                           "us_per_launch": 1e3 * q["ms"] / q["calls"], "launches_per_step": q["calls"] / n_prof,
                           "hbm_frac": q["bytes"] / sec / 1e9 / pk["hbm_gbs"],
                           "fp32_frac": q["flops"] / sec / 1e12 / pk["fp32_tflops"],
                           "mufu_frac": q["mufu"] / sec / 1e9 / pk["mufu_gops"]})
        kernels["_by_shape"] = shapes
        # every HOT-PATH (entry point, shape) of the step, in situ: the table the metric's "% of roofline" refers to
        hot_shapes = []
        for (name, label), q in sorted(per_shape.items(), key=lambda kv: -kv[1]["ms"]):
            if name not in HOTPATH or q["bytes"] <= 0:
                continue
            r = roof_entry("%s [%s]" % (name, label), q["bytes"] / q["calls"], q["flops"] / q["calls"], q["mufu"] / q["calls"],
                           q["ms"] * 1e-3 / q["calls"], pk)
            hot_shapes.append({"call": r["kernel"], "us_per_launch": r["us"], "launches_per_step": q["calls"] / n_prof,
                               "bound": r["bound"], "frac": r["frac"]})
        kernels["_hot_by_shape"] = hot_shapes
        hot_us = sum(v["us_per_step"] for n, v in kernels.items() if not n.startswith("_") and v["hot_path"])
        kernels["_hot_path_us_per_step"] = hot_us
        if not args.no_hotpath:
            del eager
            torch.cuda.empty_cache()
            # finest-level shapes of this configuration: both directions stacked on the batch for the feature-level
            # kernels (one direction in config 1), census per direction (n_samples * B in the ELBO loss)
            b_feat = B if args.config == 1 else 2 * B
            # census: UFlowLoss stacks its two directions (2B images, two normaliser groups); the ELBO loss evaluates
            # n_samples * B images per direction
            b_img = 4 * B if args.config == 3 else (2 * B if args.config in (2, 4) else B)
            hot, _ = hotpath_rooflines(b_feat, 32, H // 4, W // 4, b_img, H, W, pk)

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config in (1, 2, 4):
        cb = args.cpu_batch if args.cpu_batch is not None else (1 if args.config == 1 else 2)
        cpu, _ = run_cpu_reference(args.config, 1, 1, cb, budget_s=40.0)

    if rank == 0:
        gb = B * world
        line = {"metric": cfg["metric"], "value": gb * args.steps / (ms_total * 1e-3),
                "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": W_,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "global_batch": gb, "per_gpu_batch": B, "height": H, "width": W,
                           "parallelism": ("dp%d" % world) if args.config != 1 else ("replicas x%d" % world),
                           "cuda_graph": not args.no_graph and args.config != 3,
                           "allreduce": getattr(step, "allreduce_mode", None),
                           "conv_layout": "nchw" if args.nchw else "nhwc (channels-last conv stacks, arflow_b200/fused_conv.py)",
                           "conv_math": "cuDNN fp32 tensors, torch default allow_tf32=%s, cudnn.benchmark=%s"
                                        % (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark),
                           "l2": "per-step working set (activations) is several GB >> 126 MB L2; 4 input batches rotate",
                           "result_last_step": last},
                "clocks": clocks,
                "e2e": {"value": gb * args.steps / e2e_s, "unit": "pairs/s",
                        "h2d_bytes_per_step": B * cin * H * W * 4, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches),
                "roofline": roof, "roofline_hotpath": hot, "peaks": pk, "cpu_baseline": cpu, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)

"""Kernel micro-benchmarks (config 5 of BASELINE.json): CUDA-event timings of the C-ABI calls,
algorithmic GB/s (SURVEY §8d byte counts) and fraction of the measured HBM peak.

    python tools/microbench.py [corr|warp|all] [--csv gpurun_out/micro.csv]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


L2_BYTES = 126 * 1024 * 1024


def time_graph(make_call, set_bytes, reps=3):
    """Device time per launch, L2-cold: K independent buffer sets (K * set_bytes > 2.5 * L2) are
    visited round-robin inside one CUDA graph, so no launch ever finds its inputs in L2 and no host
    launch latency sits between the events."""
    k = max(2, int(2.5 * L2_BYTES / max(set_bytes, 1)) + 1)
    k = min(k, 64)
    calls = [make_call() for _ in range(k)]
    for c in calls:
        c()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for c in calls:
                c()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3 / (reps * k))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def time_eager(fn, reps=10):
    """Comparator timing for multi-launch ATen paths (not graph-capturable as written): CUDA events around `reps` eager
    calls after two warm-up calls, L2 flushed before each by writing a 256 MB buffer outside the events."""
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def load_ref_cuda():
    """The reference's own CUDA correlation package, compiled unmodified for sm_100a (oracle/build_ref.py)."""
    import importlib.util
    so = os.path.join(ROOT, "oracle", "_ref", "correlation_cuda.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("correlation_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def native_cost_volume(x1, x2, md=4):
    """models/correlation_native.py:13-23 as written (81 shifted products + channel mean + cat) - the ATen comparator."""
    pad = torch.nn.functional.pad(x2, [md] * 4)
    H, W = x1.shape[2:]
    cv = []
    for i in range(2 * md + 1):
        for j in range(2 * md + 1):
            cv.append(torch.mean(x1 * pad[:, :, i:i + H, j:j + W], 1, keepdim=True))
    return torch.cat(cv, 1)


def grid_sample_warp(x, flow):
    """utils/warp_utils.py:83-90 with the base grid already on the device (the reference builds it on the host per call)."""
    B, _, H, W = flow.shape
    jj = torch.arange(W, device=x.device, dtype=x.dtype).view(1, 1, 1, W).expand(B, 1, H, W)
    ii = torch.arange(H, device=x.device, dtype=x.dtype).view(1, 1, H, 1).expand(B, 1, H, W)
    v = torch.cat([jj, ii], 1) + flow
    g = torch.stack([2.0 * v[:, 0] / (W - 1) - 1.0, 2.0 * v[:, 1] / (H - 1) - 1.0], -1)
    return torch.nn.functional.grid_sample(x, g, mode="bilinear", padding_mode="zeros", align_corners=True)


def comparators(shapes, rows_out):
    """BASELINE.md §2 comparators on the same GPU: (i) the reference's Python/ATen path (correlation_native, grid_sample),
    (ii) the reference's own correlation_cuda rebuilt for sm_100a; next to the arf_* calls."""
    from arflow_b200 import _lib
    lib = _lib.load()
    ref = load_ref_cuda()
    cs = lambda: torch.cuda.current_stream().cuda_stream
    for (B, C, H, W) in shapes:
        f1 = torch.randn(B, C, H, W, device="cuda")
        f2 = torch.randn(B, C, H, W, device="cuda")
        go = torch.randn(B, 81, H, W, device="cuda")
        out = torch.empty(B, 81, H, W, device="cuda")
        g1, g2 = torch.empty_like(f1), torch.empty_like(f2)
        fl = torch.randn(B, 2, H, W, device="cuda") * 2
        wa = (B, C, H, W, H, W, float(W - 1), float(H - 1), 0, 0, 0, 1)
        r = {"shape": "%dx%dx%dx%d" % (B, C, H, W)}
        r["arf_corr_fwd_us"] = time_eager(lambda: lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), out.data_ptr(), B, C, H, W, 4, 1, 4, 1, 1, cs()))[0] * 1e6
        r["arf_corr_bwd_us"] = time_eager(lambda: lib.arf_corr_bwd(f1.data_ptr(), f2.data_ptr(), go.data_ptr(), g1.data_ptr(), g2.data_ptr(), B, C, H, W, 4, 1, 4, 1, 1, cs()))[0] * 1e6
        if ref is not None:
            def rf():
                rb1, rb2, o = f1.new_empty(0), f1.new_empty(0), f1.new_empty(0)
                ref.forward(f1, f2, rb1, rb2, o, 4, 1, 4, 1, 1, 1)

            def rb():
                rb1, rb2, a, b = f1.new_empty(0), f1.new_empty(0), f1.new_empty(0), f1.new_empty(0)
                ref.backward(f1, f2, rb1, rb2, go, a, b, 4, 1, 4, 1, 1, 1)
            r["ref_cuda_corr_fwd_us"] = time_eager(rf)[0] * 1e6
            r["ref_cuda_corr_bwd_us"] = time_eager(rb)[0] * 1e6
        if B * C * H * W * 81 * 4 < 12e9:
            r["aten_native_corr_fwd_us"] = time_eager(lambda: native_cost_volume(f1, f2), reps=5)[0] * 1e6
            a1, a2 = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)

            def nb():
                o = native_cost_volume(a1, a2)
                torch.autograd.grad(o, [a1, a2], go)
            r["aten_native_corr_fwd_bwd_us"] = time_eager(nb, reps=5)[0] * 1e6
        y = torch.empty_like(f1)
        gx, gf = torch.empty_like(f1), torch.empty_like(fl)
        r["arf_warp_fwd_us"] = time_eager(lambda: lib.arf_warp_fwd(f1.data_ptr(), fl.data_ptr(), y.data_ptr(), *wa, cs()))[0] * 1e6
        r["arf_warp_bwd_us"] = time_eager(lambda: lib.arf_warp_bwd(f1.data_ptr(), fl.data_ptr(), f2.data_ptr(), gx.data_ptr(), gf.data_ptr(), *wa, cs()))[0] * 1e6
        r["aten_grid_sample_fwd_us"] = time_eager(lambda: grid_sample_warp(f1, fl))[0] * 1e6
        b1, bf = f1.clone().requires_grad_(True), fl.clone().requires_grad_(True)

        def gb():
            o = grid_sample_warp(b1, bf)
            torch.autograd.grad(o, [b1, bf], f2)
        r["aten_grid_sample_fwd_bwd_us"] = time_eager(gb)[0] * 1e6
        rows_out.append(r)
        print(json.dumps(r), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--csv", default=None)
    ap.add_argument("--shapes", default=None, help="comma list of BxCxHxW")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (debug hook 1)")
    ap.add_argument("--probe", type=int, default=0, help="corr fwd probe mode (debug hook 2)")
    ap.add_argument("--flow", default="both", choices=["iid", "smooth", "wild", "both"],
                    help="warp flow field: iid N(0,2^2) px per pixel (config 5, worst case for gathers) or a smooth field "
                         "(1/8-resolution N(0,2^2) noise, bilinearly upsampled: what a flow network produces)")
    ap.add_argument("--warp-variant", type=int, default=0, help="debug hook 3: warp backward tuning variants")
    ap.add_argument("--census-variant", type=int, default=0, help="debug hook 5: 1 = per-pixel census kernels, 8..64 = strip height of the pair-symmetric ones")
    args = ap.parse_args()
    from arflow_b200 import _lib
    from arflow_b200.correlation import corr_out_dims
    lib = _lib.load()
    lib.arf_debug_set(1, args.variant)
    lib.arf_debug_set(2, args.probe)
    lib.arf_debug_set(5, args.census_variant)
    lib.arf_debug_set(3, args.warp_variant)
    hbm, src = peaks()
    rows = []

    def report(kind, shape, bytes_alg, flops, med, best):
        gbs = bytes_alg / med / 1e9
        rows.append((kind, "x".join(map(str, shape)), med * 1e6, best * 1e6, gbs, gbs / hbm, flops / med / 1e12))
        print("%-10s %-18s med %8.1f us  best %8.1f us  %8.1f GB/s  %5.1f%% HBM(%s)  %6.2f TFLOP/s"
              % (kind, rows[-1][1], med * 1e6, best * 1e6, gbs, 100 * gbs / hbm, src, flops / med / 1e12), flush=True)

    shapes = [(8, 32, 96, 128), (64, 32, 96, 128), (8, 32, 48, 64), (8, 32, 24, 32), (8, 32, 12, 16),
              (32, 32, 112, 256), (16, 64, 48, 64), (16, 96, 24, 32), (16, 128, 12, 16), (16, 196, 6, 8),
              (1, 32, 96, 160), (1, 192, 6, 10)]
    if args.shapes:
        shapes = [tuple(int(v) for v in sh.split("x")) for sh in args.shapes.split(",")]
    cs = lambda: torch.cuda.current_stream().cuda_stream
    if args.what in ("corr", "corr_fwd", "corr_bwd", "all"):
        for (B, C, H, W) in shapes:
            px = B * H * W

            def mk_fwd():
                f1 = torch.randn(B, C, H, W, device="cuda")
                f2 = torch.randn(B, C, H, W, device="cuda")
                out = torch.empty(B, 81, H, W, device="cuda")
                return lambda: lib.arf_corr_fwd(f1.data_ptr(), f2.data_ptr(), out.data_ptr(), B, C, H, W, 4, 1, 4, 1, 1, cs())

            def mk_bwd():
                f1 = torch.randn(B, C, H, W, device="cuda")
                f2 = torch.randn(B, C, H, W, device="cuda")
                go = torch.randn(B, 81, H, W, device="cuda")
                g1, g2 = torch.empty_like(f1), torch.empty_like(f2)
                return lambda: lib.arf_corr_bwd(f1.data_ptr(), f2.data_ptr(), go.data_ptr(), g1.data_ptr(),
                                                g2.data_ptr(), B, C, H, W, 4, 1, 4, 1, 1, cs())
            if args.what != "corr_bwd":
                report("corr_fwd", (B, C, H, W), px * (8 * C + 324), px * C * 162,
                       *time_graph(mk_fwd, px * (8 * C + 324)))
            if args.what != "corr_fwd":
                report("corr_bwd", (B, C, H, W), px * (16 * C + 324), px * C * 324,
                       *time_graph(mk_bwd, px * (16 * C + 324)))
    if args.what in ("warp", "all"):
        wshapes = shapes if args.shapes else shapes[:8] + [(8, 3, 384, 512), (32, 3, 448, 1024)]
        for (B, C, H, W) in wshapes:
            px = B * H * W
            a = (B, C, H, W, H, W, float(W - 1), float(H - 1), 0, 0, 0, 1)

            for flow_kind in (["iid", "smooth"] if args.flow == "both" else [args.flow]):
                def mk(kind):
                    def make():
                        x = torch.randn(B, C, H, W, device="cuda")
                        if flow_kind == "wild":
                            # what a random-init network produces on noise images (the bench's in-situ fields): 1/4-resolution
                            # noise of tens of pixels, bilinearly upsampled - neighbouring pixels sample far-apart places
                            fl = torch.nn.functional.interpolate(torch.randn(B, 2, max(H // 4, 1), max(W // 4, 1), device="cuda") * 40,
                                                                 size=(H, W), mode="bilinear").contiguous()
                        elif flow_kind == "iid" or H < 16 or W < 16:
                            fl = torch.randn(B, 2, H, W, device="cuda") * 2
                        else:
                            fl = torch.nn.functional.interpolate(torch.randn(B, 2, H // 8, W // 8, device="cuda") * 2,
                                                                 size=(H, W), mode="bilinear").contiguous()
                        y = torch.empty_like(x)
                        gy = torch.randn_like(x)
                        gx, gf = torch.empty_like(x), torch.empty_like(fl)
                        if kind == "fwd":
                            return lambda: lib.arf_warp_fwd(x.data_ptr(), fl.data_ptr(), y.data_ptr(), *a, cs())
                        if kind == "bwd":
                            return lambda: lib.arf_warp_bwd(x.data_ptr(), fl.data_ptr(), gy.data_ptr(), gx.data_ptr(),
                                                            gf.data_ptr(), *a, cs())
                        return lambda: lib.arf_warp_bwd(x.data_ptr(), fl.data_ptr(), gy.data_ptr(), None, gf.data_ptr(), *a, cs())
                    return make
                tag = {"iid": "", "smooth": "_smooth", "wild": "_wild"}[flow_kind]
                report("warp_fwd" + tag, (B, C, H, W), px * (8 * C + 8), px * C * 8, *time_graph(mk("fwd"), px * (8 * C + 8)))
                report("warp_bwd" + tag, (B, C, H, W), px * (12 * C + 16), px * C * 16, *time_graph(mk("bwd"), px * (12 * C + 16)))
                report("warp_bwdF" + tag, (B, C, H, W), px * (8 * C + 16), px * C * 16, *time_graph(mk("bwdF"), px * (8 * C + 16)))
    if args.what in ("census", "all"):
        cshapes = shapes if args.shapes else [(8, 3, 384, 512), (32, 3, 448, 1024)]
        for (B, C, H, W) in cshapes:
            px = B * H * W
            npart = lib.arf_census_num_partials(B, H, W)

            def mkc(kind):
                def make():
                    a = torch.rand(B, 3, H, W, device="cuda")
                    b = torch.rand(B, 3, H, W, device="cuda")
                    m = torch.rand(B, 1, H, W, device="cuda")
                    ham = torch.empty(B, 1, H, W, device="cuda")
                    part = torch.empty(2 * npart, device="cuda")
                    sums = torch.ones(3, device="cuda")
                    gl = torch.ones(1, device="cuda")
                    gb = torch.empty_like(b)
                    if kind == "fwd":
                        return lambda: lib.arf_census_fwd(a.data_ptr(), b.data_ptr(), m.data_ptr(), ham.data_ptr(),
                                                          part.data_ptr(), sums.data_ptr(), B, H, W, 7, 1.0, 0.01, 0.4, cs())
                    lib.arf_census_fwd(a.data_ptr(), b.data_ptr(), m.data_ptr(), ham.data_ptr(), part.data_ptr(),
                                       sums.data_ptr(), B, H, W, 7, 1.0, 0.01, 0.4, cs())
                    return lambda: lib.arf_census_bwd(a.data_ptr(), b.data_ptr(), None, ham.data_ptr(), m.data_ptr(),
                                                      sums.data_ptr(), gl.data_ptr(), None, gb.data_ptr(), B, H, W, 7,
                                                      1.0, 0.01, 0.4, cs())
                return make
            report("census_fwd", (B, 3, H, W), px * 32, px * 640, *time_graph(mkc("fwd"), px * 32))
            report("census_bwd", (B, 3, H, W), px * 44, px * 1280, *time_graph(mkc("bwd"), px * 44))
    if args.what in ("loss", "all"):
        # the remaining SURVEY §8(a) rows at the chairs_uflow loss shapes: SSIM (P2/P4), NHWC resampler (W3), masks (M1-M3),
        # smoothness (S1), resize (U1), inverse_diagonal (T4)
        B, H, W = (shapes[0][0], shapes[0][2], shapes[0][3]) if args.shapes else (8, 384, 512)     # full-resolution batch
        px = B * H * W

        def mk_ssim(kind):
            def make():
                x, y = torch.rand(B * 3, H, W, device="cuda"), torch.rand(B * 3, H, W, device="cuda")
                o1, o2 = torch.empty_like(x), torch.empty_like(x)
                g1, g2 = torch.randn_like(x), torch.randn_like(x)
                coef = torch.empty(5 * B * 3 * H * W, device="cuda")
                gx, gy = torch.empty_like(x), torch.empty_like(x)
                if kind == "fwd":
                    return lambda: lib.arf_ssim_fwd(x.data_ptr(), y.data_ptr(), o1.data_ptr(), o2.data_ptr(), B * 3, H, W, 3, 0, 0, cs())
                return lambda: lib.arf_ssim_bwd(x.data_ptr(), y.data_ptr(), g1.data_ptr(), g2.data_ptr(), coef.data_ptr(),
                                                gx.data_ptr(), gy.data_ptr(), B * 3, H, W, 3, 0, 0, cs())
            return make
        report("ssim_fwd", (B, 3, H, W), px * 3 * 16, px * 3 * 90, *time_graph(mk_ssim("fwd"), px * 3 * 16))
        report("ssim_bwd", (B, 3, H, W), px * 3 * 24, px * 3 * 200, *time_graph(mk_ssim("bwd"), px * 3 * 24))

        def mk_res(kind):
            def make():
                data = torch.randn(B, H, W, 3, device="cuda")
                wxy = torch.rand(B, H, W, 2, device="cuda") * torch.tensor([W - 1.0, H - 1.0], device="cuda")
                out = torch.empty(B, H * W, 3, device="cuda")
                go = torch.randn_like(out)
                gd, gw = torch.empty_like(data), torch.empty_like(wxy)
                wp = wxy.data_ptr()
                # (the lambdas keep `wxy` alive: only its address is passed on)
                if kind == "fwd":
                    return lambda keep=wxy: lib.arf_resampler_fwd(data.data_ptr(), wp, wp + 4, 2, out.data_ptr(), B, H, W, 3, H * W, cs())
                return lambda keep=wxy: lib.arf_resampler_bwd(data.data_ptr(), wp, wp + 4, 2, go.data_ptr(), gd.data_ptr(),
                                                              gw.data_ptr(), gw.data_ptr() + 4, 2, B, H, W, 3, H * W, cs())
            return make
        report("resampler_fwd", (B, H, W, 3), px * (12 + 8 + 12), px * 3 * 8, *time_graph(mk_res("fwd"), px * 32))
        report("resampler_bwd", (B, H, W, 3), px * (12 + 8 + 12 + 12 + 8), px * 3 * 16, *time_graph(mk_res("bwd"), px * 52))

        hq, wq = H // 4, W // 4
        pq = B * hq * wq

        def mk_mask(kind):
            def make():
                fl = torch.randn(B, 2, hq, wq, device="cuda") * 3
                fl2 = torch.randn(B, 2, hq, wq, device="cuda") * 3
                m = torch.empty(B, 1, hq, wq, device="cuda")
                gc = torch.randn(B, 1, hq, wq, device="cuda")
                gf = torch.empty_like(fl)
                if kind == "inside":
                    return lambda: lib.arf_inside_mask(fl.data_ptr(), m.data_ptr(), B, hq, wq, 0, 0, cs())
                if kind == "range":
                    return lambda: lib.arf_range_map(fl.data_ptr(), m.data_ptr(), B, hq, wq, 0, cs())
                if kind == "range_bwd":
                    return lambda: lib.arf_range_map_bwd(fl.data_ptr(), gc.data_ptr(), gf.data_ptr(), B, hq, wq, 0, cs())
                if kind == "count":
                    return lambda: lib.arf_count_to_mask(gc.data_ptr(), m.data_ptr(), B * hq * wq, 0, 0.0, cs())
                return lambda: lib.arf_occ_bidir(fl.data_ptr(), fl2.data_ptr(), m.data_ptr(), B, hq, wq, 0.01, 0.5, cs())
            return make
        for kind, nb in (("inside", 12), ("range", 16), ("range_bwd", 20), ("count", 8), ("occ_bidir", 20)):
            report("mask_" + kind, (B, 2, hq, wq), pq * nb, 0, *time_graph(mk_mask(kind), pq * nb))

        def mk_smooth(kind, order):
            def make():
                img = torch.rand(B, 3, hq, wq, device="cuda")
                fl = torch.randn(B, 2, hq, wq, device="cuda")
                npart = lib.arf_smooth_num_partials(B, hq, wq)
                part, out = torch.empty(2 * npart, device="cuda"), torch.empty(1, device="cuda")
                gl, gfl = torch.ones(1, device="cuda"), torch.empty_like(fl)
                a = (B, 3, hq, wq, order, 1, 0, 0, 150.0, 1e-6, 1.0)
                if kind == "fwd":
                    return lambda: lib.arf_smooth_fwd(img.data_ptr(), fl.data_ptr(), out.data_ptr(), part.data_ptr(), *a, cs())
                return lambda: lib.arf_smooth_bwd(img.data_ptr(), fl.data_ptr(), gl.data_ptr(), gfl.data_ptr(), *a, cs())
            return make
        for order in (1, 2):
            report("smooth%d_fwd" % order, (B, 3, hq, wq), pq * 20, 0, *time_graph(mk_smooth("fwd", order), pq * 20))
            report("smooth%d_bwd" % order, (B, 3, hq, wq), pq * 28, 0, *time_graph(mk_smooth("bwd", order), pq * 28))

        for (N, Hi, Wi, Ho, Wo, tag) in ((32, 192, 256, 384, 512, "up2"), (16, 96, 128, 384, 512, "up4"), (24, 384, 512, 96, 128, "down4")):
            def mk_rs(kind):
                def make():
                    a = torch.randn(N, Hi, Wi, device="cuda")
                    b = torch.randn(N, Ho, Wo, device="cuda")
                    r = (N, Hi, Wi, Ho, Wo, Hi / Ho, Wi / Wo, 1.0, 0)
                    if kind == "fwd":
                        return lambda: lib.arf_resize_bilinear_fwd(a.data_ptr(), b.data_ptr(), *r, cs())
                    return lambda: lib.arf_resize_bilinear_bwd(b.data_ptr(), a.data_ptr(), *r, cs())
                return make
            nb = N * (Hi * Wi + Ho * Wo) * 4
            report("resize_%s_fwd" % tag, (N, Hi, Wi), nb, 0, *time_graph(mk_rs("fwd"), nb))
            report("resize_%s_bwd" % tag, (N, Hi, Wi), nb, 0, *time_graph(mk_rs("bwd"), nb))

        def mk_inv():
            S, M, Nn = 16, 24, 32
            A = torch.rand(S, M, Nn, device="cuda") + 1.5
            Bm, Cm = torch.randn(S, M, Nn - 1, device="cuda") * 0.3, torch.randn(S, M - 1, Nn, device="cuda") * 0.3
            Hh = torch.empty(S, M, Nn, device="cuda")
            return lambda: lib.arf_inv_diag(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), Hh.data_ptr(), S, M, Nn, cs())
        report("inv_diag", (16, 24, 32), 16 * 24 * 32 * 16, 0, *time_graph(mk_inv, 16 * 24 * 32 * 16))
    if args.what in ("stencil", "all"):
        for (N, k, H, W) in [(32, 3, 112, 256), (8, 3, 96, 128)]:
            px = N * H * W
            taps = (k + 1) ** 2

            def mks(kind):
                def make():
                    A = torch.randn(N, 2 * taps, H, W, device="cuda")
                    X = torch.randn(N, 2, H, W, device="cuda")
                    Y = torch.empty_like(X)
                    dA, dX = torch.empty_like(A), torch.empty_like(X)
                    if kind == "fwd":
                        return lambda: lib.arf_stencil_mv_fwd(A.data_ptr(), X.data_ptr(), Y.data_ptr(), N, H, W, k, 0, cs())
                    return lambda: lib.arf_stencil_mv_bwd(A.data_ptr(), X.data_ptr(), Y.data_ptr(), dA.data_ptr(),
                                                          dX.data_ptr(), N, H, W, k, 0, cs())
                return make
            report("stencil_fwd", (N, 2 * taps, H, W), px * (8 * taps + 16), px * 4 * taps, *time_graph(mks("fwd"), px * (8 * taps + 16)))
            report("stencil_bwd", (N, 2 * taps, H, W), px * (16 * taps + 24), px * 8 * taps, *time_graph(mks("bwd"), px * (16 * taps + 24)))

            def mkt():
                A = 1 + torch.rand(N, 2, H, W, device="cuda")
                Bc = 0.3 * torch.randn(N, 2, H, W - 1, device="cuda")
                Cc = 0.3 * torch.randn(N, 2, H - 1, W, device="cuda")
                Dc = 0.2 * torch.randn(N, 2, H - 1, W - 1, device="cuda")
                X = torch.randn(N, 2, H, W, device="cuda")
                Y = torch.empty_like(X)
                return lambda: lib.arf_trisolve(A.data_ptr(), Bc.data_ptr(), Cc.data_ptr(), Dc.data_ptr(), X.data_ptr(),
                                                Y.data_ptr(), N * 2, H, W, 0, cs())
            report("trisolve", (N, 2, H, W), N * 2 * H * W * 24, N * 2 * H * W * 8, *time_graph(mkt, N * 2 * H * W * 24))
            lib.arf_debug_set(4, 1)
            report("trisolve_wavefront", (N, 2, H, W), N * 2 * H * W * 24, N * 2 * H * W * 8, *time_graph(mkt, N * 2 * H * W * 24))
            lib.arf_debug_set(4, 0)
    if args.what in ("glue", "all"):
        # the streaming kernels around the cuDNN convolutions (fused_conv.py): L1-level shapes of the chairs_uflow step
        nrows, C = 16 * 96 * 128, 128

        def mk_epi(kind):
            def make():
                y = torch.randn(nrows, C, device="cuda")
                gy = torch.randn(nrows, C, device="cuda")
                g = torch.empty_like(y)
                b = torch.randn(C, device="cuda")
                db = torch.empty(C, device="cuda")
                part = torch.empty(lib.arf_bias_leaky_nhwc_num_partials(nrows, C), device="cuda")
                if kind == "fwd":
                    return lambda: lib.arf_bias_leaky_nhwc_fwd(y.data_ptr(), b.data_ptr(), nrows, C, 0.1, cs())
                return lambda: lib.arf_bias_leaky_nhwc_bwd(gy.data_ptr(), y.data_ptr(), g.data_ptr(), part.data_ptr(),
                                                           db.data_ptr(), nrows, C, 0.1, cs())
            return make
        report("epilogue_fwd", (nrows, C), nrows * C * 8, nrows * C * 2, *time_graph(mk_epi("fwd"), nrows * C * 8))
        report("epilogue_bwd", (nrows, C), nrows * C * 12, nrows * C * 2, *time_graph(mk_epi("bwd"), nrows * C * 12))
        for (Cs, Cd, nhwc) in [(408, 504, 1), (81, 152, 0), (32, 152, 1)]:
            N, HW = 16, 96 * 128

            def mk_pack(kind):
                def make():
                    src = torch.randn(N * HW * Cs, device="cuda")
                    dst = torch.empty(N * HW * Cd, device="cuda")
                    if kind == "pack":
                        return lambda: lib.arf_nhwc_pack(dst.data_ptr(), src.data_ptr(), N, HW, Cs, Cd, 0 if nhwc else 34, nhwc, cs())
                    return lambda: lib.arf_nhwc_unpack(src.data_ptr(), dst.data_ptr(), N, HW, Cs, Cd, 0 if nhwc else 34, nhwc, cs())
                return make
            tag = "nhwc" if nhwc else "nchw"
            report("pack_" + tag, (N, Cs, Cd, HW), N * HW * Cs * 8, 0, *time_graph(mk_pack("pack"), N * HW * Cs * 8))
            report("unpack_" + tag, (N, Cs, Cd, HW), N * HW * Cs * 8, 0, *time_graph(mk_pack("unpack"), N * HW * Cs * 8))
        n = 32 * 96 * 128

        def mk_norm(kind):
            def make():
                B = 16
                f1, f2 = torch.randn(B, n, device="cuda"), torch.randn(B, n, device="cuda")
                g1, g2 = torch.randn(B, n, device="cuda"), torch.randn(B, n, device="cuda")
                y1, y2 = torch.empty_like(f1), torch.empty_like(f2)
                stats, coef = torch.empty(B * 4, device="cuda"), torch.empty(B * 2, device="cuda")
                ws = torch.empty(lib.arf_featnorm_workspace(B, n) // 8, dtype=torch.float64, device="cuda")
                lib.arf_featnorm_fwd(f1.data_ptr(), f2.data_ptr(), y1.data_ptr(), y2.data_ptr(), stats.data_ptr(), ws.data_ptr(), B, n, cs())
                if kind == "fwd":
                    return lambda: lib.arf_featnorm_fwd(f1.data_ptr(), f2.data_ptr(), y1.data_ptr(), y2.data_ptr(), stats.data_ptr(),
                                                        ws.data_ptr(), B, n, cs())
                return lambda: lib.arf_featnorm_bwd(f1.data_ptr(), f2.data_ptr(), g1.data_ptr(), g2.data_ptr(), stats.data_ptr(),
                                                    y1.data_ptr(), y2.data_ptr(), coef.data_ptr(), ws.data_ptr(), B, n, cs())
            return make
        # fwd: read both maps twice (moments, apply) + write both; bwd: read f, g twice + write both
        report("featnorm_fwd", (16, 32, 96, 128), 16 * n * 4 * 6, 16 * n * 8, *time_graph(mk_norm("fwd"), 16 * n * 24))
        report("featnorm_bwd", (16, 32, 96, 128), 16 * n * 4 * 10, 16 * n * 12, *time_graph(mk_norm("bwd"), 16 * n * 40))
    if args.what in ("smallconv", "all"):
        # the own fp32 kernels for the thin convolutions (csrc/smallconv.cu, layout.cu): chairs_uflow level-1 / level-0 shapes
        N, H, W, Ci = 16, 96, 128, 32
        px = N * H * W

        def mk_head(kind):
            def make():
                x = torch.randn(N, H, W, Ci, device="cuda")
                w = torch.randn(2, 3, 3, Ci, device="cuda")
                b = torch.randn(2, device="cuda")
                y = torch.empty(N, 2, H, W, device="cuda")
                gy = torch.randn(N, 2, H, W, device="cuda")
                gx = torch.empty_like(x)
                out = torch.empty(2 * 9 * Ci + 2, device="cuda")
                part = torch.empty(lib.arf_conv3x3_small_bwd_workspace(N, H, W, Ci, 2), device="cuda")
                if kind == "fwd":
                    return lambda: lib.arf_conv3x3_small_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, H, W, Ci, 2, cs())
                return lambda: lib.arf_conv3x3_small_bwd(x.data_ptr(), gy.data_ptr(), w.data_ptr(), gx.data_ptr(), out.data_ptr(),
                                                         part.data_ptr(), N, H, W, Ci, 2, cs())
            return make
        report("flowhead_fwd", (N, Ci, H, W), px * (Ci + 2) * 4, px * 36 * Ci, *time_graph(mk_head("fwd"), px * (Ci + 2) * 4))
        report("flowhead_bwd", (N, Ci, H, W), px * (2 * Ci + 2) * 4, px * 72 * Ci, *time_graph(mk_head("bwd"), px * (2 * Ci + 2) * 4))
        Hi, Wi = 384, 512
        Ho, Wo = Hi // 2, Wi // 2

        def mk_first():
            x = torch.randn(N, Hi, Wi, 8, device="cuda")
            g = torch.randn(N, Ho, Wo, 32, device="cuda")
            out = torch.empty(27 * 32, device="cuda")
            part = torch.empty(lib.arf_conv3x3s2_first_wgrad_workspace(N, Hi, Wi), device="cuda")
            return lambda: lib.arf_conv3x3s2_first_wgrad(x.data_ptr(), g.data_ptr(), out.data_ptr(), part.data_ptr(), N, Hi, Wi, 3, 32, cs())
        by = N * (Hi * Wi * 8 + Ho * Wo * 32) * 4
        report("first_wgrad", (N, 8, Hi, Wi), by, N * Ho * Wo * 32 * 54, *time_graph(mk_first, by))

        def mk_pair():
            src = torch.rand(N // 2, 6, Hi, Wi, device="cuda")
            dst = torch.empty(N, Hi, Wi, 8, device="cuda")
            return lambda: lib.arf_image_pair_pack(dst.data_ptr(), src.data_ptr(), N // 2, Hi * Wi, 3, 8, 2.0, -1.0, cs())
        by = N * Hi * Wi * (3 + 8) * 4
        report("pair_pack", (N // 2, 6, Hi, Wi), by, 0, *time_graph(mk_pair, by))
    if args.what == "compare":
        crow = []
        comparators(shapes if args.shapes else [(8, 32, 96, 128), (16, 32, 96, 128), (64, 32, 96, 128), (16, 64, 48, 64),
                                                (16, 96, 24, 32), (16, 128, 12, 16), (1, 192, 6, 10), (1, 32, 96, 160)], crow)
        if args.csv:
            os.makedirs(os.path.dirname(args.csv), exist_ok=True)
            keys = sorted({k for r in crow for k in r if k != "shape"})
            with open(args.csv, "w") as f:
                f.write("shape," + ",".join(keys) + ",vs_ref_cuda_fwd,vs_ref_cuda_bwd,vs_aten_native_fwd,vs_grid_sample_fwd\n")
                for r in crow:
                    q = lambda a, b: ("%.1f" % (r[a] / r[b])) if a in r and b in r else ""
                    f.write(r["shape"] + "," + ",".join("%.1f" % r[k] if k in r else "" for k in keys) + "," +
                            ",".join([q("ref_cuda_corr_fwd_us", "arf_corr_fwd_us"), q("ref_cuda_corr_bwd_us", "arf_corr_bwd_us"),
                                      q("aten_native_corr_fwd_us", "arf_corr_fwd_us"), q("aten_grid_sample_fwd_us", "arf_warp_fwd_us")]) + "\n")
        return
    if args.csv:
        os.makedirs(os.path.dirname(args.csv), exist_ok=True)
        with open(args.csv, "w") as f:
            f.write("kernel,shape,median_us,best_us,alg_GBps,frac_hbm,TFLOPs\n")
            for r in rows:
                f.write("%s,%s,%.2f,%.2f,%.1f,%.4f,%.3f\n" % r)


if __name__ == "__main__":
    main()

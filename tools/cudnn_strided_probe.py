"""Probe: does cuDNN keep its fast sm_100 kernels when a convolution operand is a COLUMN SLICE of a wider channels-last
buffer (pixel stride ld > C)?  If it does, the dense blocks of the PWC decoders need no prefix re-pack (arf_nhwc_pack,
arf_nhwc_unpack_add: ~1.4 ms of the 14.5 ms chairs_uflow step) — torch's own cuDNN binding always packs first.

Uses the cuDNN frontend's Python graph API (the same library torch calls); builds every plan the heuristics offer for
packed and for strided descriptors of the level-1 dense-block layer and times each.  Run under gpurun:
    python tools/cudnn_strided_probe.py
"""
import sys
import time

import cudnn
import torch

N, H, W = 16, 96, 128
F = cudnn.data_type.FLOAT


def nhwc_strides(C_ld):
    return [H * W * C_ld, 1, W * C_ld, C_ld]


def time_plans(graph, pack, tag, handle, iters=5):
    graph.validate()
    graph.build_operation_graph()
    graph.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
    try:
        graph.check_support()
    except Exception as e:  # noqa: BLE001
        print("%-34s unsupported: %s" % (tag, str(e)[:120]))
        return None
    n = graph.get_execution_plan_count()
    ptrs = {k: v.data_ptr() for k, v in pack.items()}   # raw pointers: the descriptors carry the strides
    best = None
    for i in range(n):
        try:
            graph.build_plan_at_index(i)
            ws = torch.empty(max(graph.get_workspace_size_plan_at_index(i), 1), device="cuda", dtype=torch.uint8)
            for _ in range(2):
                graph.execute_plan_at_index(ptrs, ws, i, handle=handle)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                graph.execute_plan_at_index(ptrs, ws, i, handle=handle)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            name = graph.get_plan_name_at_index(i)
            if best is None or us < best[0]:
                best = (us, name, i)
        except Exception as e:  # noqa: BLE001
            continue
    if best is None:
        print("%-34s no plan ran" % tag)
    else:
        print("%-34s best %8.1f us  (%d plans)  %s" % (tag, best[0], n, best[1][:70]))
    return best


def new_graph(handle):
    return cudnn.pygraph(io_data_type=F, intermediate_data_type=F, compute_data_type=F, handle=handle)


def main():
    torch.manual_seed(0)
    handle = cudnn.create_handle()
    stream = torch.cuda.current_stream().cuda_stream
    cudnn.set_stream(handle=handle, stream=stream)
    C, K, LD = 408, 128, 568          # layer 3 of the level-1 block: 403 real input channels (8-aligned: 408)
    big = torch.randn(N, H, W, LD, device="cuda")                 # the whole block as one NHWC matrix
    xs = big[..., LD - C:]                                          # column slice: pixel stride LD
    xp = xs.contiguous()                                            # packed copy: pixel stride C
    w = torch.randn(K, 3, 3, C, device="cuda") * 0.05               # KRSC
    yp = torch.empty(N, H, W, K, device="cuda")
    ys = big[..., 32:32 + K]                                        # output as a column slice of the same buffer
    gy = torch.randn(N, H, W, K, device="cuda")
    dxp = torch.empty(N, H, W, C, device="cuda")
    dbig = torch.zeros(N, H, W, LD, device="cuda")
    dxs = dbig[..., LD - C:]
    dw = torch.empty(K, 3, 3, C, device="cuda")

    def conv_graph(kind, x_ld, y_ld):
        g = new_graph(handle)
        X = g.tensor(name="X", dim=[N, C, H, W], stride=nhwc_strides(x_ld), data_type=F)
        Wt = g.tensor(name="W", dim=[K, C, 3, 3], stride=[9 * C, 1, 3 * C, C], data_type=F)
        Y = g.tensor(name="Y", dim=[N, K, H, W], stride=nhwc_strides(y_ld), data_type=F)
        if kind == "fprop":
            out = g.conv_fprop(image=X, weight=Wt, padding=[1, 1], stride=[1, 1], dilation=[1, 1], compute_data_type=F)
            out.set_output(True).set_dim([N, K, H, W]).set_stride(nhwc_strides(y_ld)).set_data_type(F)
            return g, (X, Wt, out)
        if kind == "dgrad":
            out = g.conv_dgrad(loss=Y, filter=Wt, padding=[1, 1], stride=[1, 1], dilation=[1, 1], compute_data_type=F)
            out.set_output(True).set_dim([N, C, H, W]).set_stride(nhwc_strides(x_ld)).set_data_type(F)
            return g, (Y, Wt, out)
        out = g.conv_wgrad(image=X, loss=Y, padding=[1, 1], stride=[1, 1], dilation=[1, 1], compute_data_type=F)
        out.set_output(True).set_dim([K, C, 3, 3]).set_stride([9 * C, 1, 3 * C, C]).set_data_type(F)
        return g, (X, Y, out)

    print("cuDNN backend", cudnn.backend_version(), "frontend", cudnn.__version__)
    # ---- fprop
    g, (X, Wt, Y) = conv_graph("fprop", C, K)
    time_plans(g, {X: xp, Wt: w, Y: yp}, "fprop packed in / packed out", handle)
    g, (X, Wt, Y) = conv_graph("fprop", LD, K)
    time_plans(g, {X: xs, Wt: w, Y: yp}, "fprop STRIDED in / packed out", handle)
    g, (X, Wt, Y) = conv_graph("fprop", LD, LD)
    time_plans(g, {X: xs, Wt: w, Y: ys}, "fprop STRIDED in / STRIDED out", handle)
    # ---- dgrad
    g, (Yt, Wt, DX) = conv_graph("dgrad", C, K)
    time_plans(g, {Yt: gy, Wt: w, DX: dxp}, "dgrad packed dx", handle)
    g, (Yt, Wt, DX) = conv_graph("dgrad", LD, K)
    time_plans(g, {Yt: gy, Wt: w, DX: dxs}, "dgrad STRIDED dx", handle)
    # ---- wgrad
    g, (X, Yt, DW) = conv_graph("wgrad", C, K)
    time_plans(g, {X: xp, Yt: gy, DW: dw}, "wgrad packed x", handle)
    g, (X, Yt, DW) = conv_graph("wgrad", LD, K)
    time_plans(g, {X: xs, Yt: gy, DW: dw}, "wgrad STRIDED x", handle)

    # ---- torch's own binding on the packed operands, for reference
    torch.backends.cudnn.benchmark = True
    xt = xp.permute(0, 3, 1, 2).requires_grad_(True)
    wt = w.permute(0, 3, 1, 2).requires_grad_(True)
    for _ in range(3):
        y = torch.nn.functional.conv2d(xt, wt, None, 1, 1)
        gx, gw = torch.autograd.grad(y, [xt, wt], gy.permute(0, 3, 1, 2))
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            y = torch.nn.functional.conv2d(xt, wt, None, 1, 1)
            gx, gw = torch.autograd.grad(y, [xt, wt], gy.permute(0, 3, 1, 2))
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
        print("torch: %-80s %8.1f us x%d" % (ev.key[:80], ev.device_time_total / max(ev.count, 1), ev.count))


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("probe wall %.0f s" % (time.time() - t0), file=sys.stderr)

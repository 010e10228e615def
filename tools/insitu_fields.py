"""What do the flow fields of the benchmarked step look like, and how fast are the warp kernels on exactly those fields?

Runs a few eager chairs_uflow steps (random-init PWCFlow on U[0,1) noise pairs, as bench.py does), captures the
(source, field) pair of every warp call of the last step, prints field statistics and times arf_warp_fwd / arf_warp_bwd on
the captured tensors (L2-cold, CUDA-graph timed, tools/microbench.py's method).

    python tools/insitu_fields.py [--save gpurun_out/fields.pt]
"""
import argparse
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--save", default=None)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    from arflow_b200 import _lib, warp_utils
    from arflow_b200.train_step import UFlowTrainStep
    from arflow_b200.uflow_loss import UFlowLoss
    from arflow_b200.uflow_model import PWCFlow
    from microbench import time_graph
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = PWCFlow(types.SimpleNamespace(level_dropout=0.1, feature_norm=True)).to(dev)
    model.init_weights()
    model.train()
    loss_fn = UFlowLoss(types.SimpleNamespace(edge_constant=150, w_smooth=4.0, w_census=1.0, with_bk=True, smooth_order=1))
    step = UFlowTrainStep(model, loss_fn, lr=1e-4, use_graph=False, world_size=1)
    gen = torch.Generator(device="cpu").manual_seed(0)
    captured = []
    orig = warp_utils._WarpFunction.forward

    def spy(ctx, x, field, *a):
        captured.append((x.detach().clone(), field.detach().clone(), a))
        return orig(ctx, x, field, *a)
    for i in range(args.steps):
        batch = torch.rand(8, 6, 384, 512, generator=gen).to(dev)
        if i == args.steps - 1:
            warp_utils._WarpFunction.forward = staticmethod(spy)
        out = step(batch)
    warp_utils._WarpFunction.forward = staticmethod(orig)
    torch.cuda.synchronize()
    print("step result", [float(v) for v in out])
    # the in-situ event timing bench.py uses, call by call (one eager step)
    _lib.profile_start()
    for i in range(2):
        step(torch.rand(8, 6, 384, 512, generator=gen).to(dev))
    rec = _lib.profile_stop()
    half = len(rec) // 2
    for name, a, ms in rec[half:]:
        if name.endswith(("_out_dims", "_num_partials", "_workspace")):
            continue
        ints = [v for v in a if isinstance(v, int) and 0 <= v < 100000][:7]
        print("  insitu %-28s %8.1f us  %s" % (name, ms * 1e3, ints))
    lib = _lib.load()
    cs = lambda: torch.cuda.current_stream().cuda_stream
    saved = []
    for (x, field, a) in captured:
        B, C, Hs, Ws = x.shape
        Ho, Wo = field.shape[2:]
        nW1, nH1, kind, interp, pad, align = a
        fl = field if kind == 0 else field - torch.stack(torch.meshgrid(torch.arange(Ho, device=dev), torch.arange(Wo, device=dev),
                                                                        indexing="ij")[::-1]).float()[None]
        gx_ = (fl[..., 1:] - fl[..., :-1]).abs().mean().item()
        gy_ = (fl[..., 1:, :] - fl[..., :-1, :]).abs().mean().item()
        jj = torch.arange(Wo, device=dev).float()[None, None, :]
        ii = torch.arange(Ho, device=dev).float()[None, :, None]
        X, Y = jj + fl[:, 0], ii + fl[:, 1]
        inside = ((X >= 0) & (X <= Ws - 1) & (Y >= 0) & (Y <= Hs - 1)).float().mean().item()
        print("warp call B%d C%d %dx%d kind %d: mean|flow| %.2f  mean|dflow/dx| %.3f  mean|dflow/dy| %.3f  inside %.3f"
              % (B, C, Ho, Wo, kind, fl.abs().mean().item(), gx_, gy_, inside), flush=True)
        wa = (B, C, Hs, Ws, Ho, Wo, float(nW1), float(nH1), kind, interp, pad, int(align))
        px = B * Ho * Wo

        def mk(which):
            def make():
                xx, ff = x.clone(), field.clone()
                y = torch.empty(B, C, Ho, Wo, device=dev)
                gy = torch.randn(B, C, Ho, Wo, device=dev)
                gx, gf = torch.empty_like(xx), torch.empty_like(ff)
                if which == "fwd":
                    return lambda: lib.arf_warp_fwd(xx.data_ptr(), ff.data_ptr(), y.data_ptr(), *wa, cs())
                if which == "bwd":
                    return lambda: lib.arf_warp_bwd(xx.data_ptr(), ff.data_ptr(), gy.data_ptr(), gx.data_ptr(), gf.data_ptr(), *wa, cs())
                return lambda: lib.arf_warp_bwd(xx.data_ptr(), ff.data_ptr(), gy.data_ptr(), None, gf.data_ptr(), *wa, cs())
            return make
        for which, nb in (("fwd", px * (8 * C + 8)), ("bwd", px * (12 * C + 16)), ("bwdF", px * (8 * C + 16))):
            med, best = time_graph(mk(which), nb)
            print("   %-5s %8.1f us   %7.1f GB/s algorithmic" % (which, med * 1e6, nb / med / 1e9), flush=True)
        saved.append({"x_shape": tuple(x.shape), "field": field.cpu(), "args": a})
    if args.save:
        torch.save(saved, args.save)


if __name__ == "__main__":
    main()

"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Each fixture holds seeded float32 inputs and the reference's outputs computed twice: in float64
("truth", suffix _f64) and in float32 (reference precision, suffix _f32).  tests/ never imports the
reference; they read these files.
"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ARFLOW_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
# easydict is only needed by modules we do not exercise; a 3-line stand-in keeps imports working
if "easydict" not in sys.modules:
    m = types.ModuleType("easydict")
    class EasyDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    m.EasyDict = EasyDict
    sys.modules["easydict"] = m

OUT = os.path.dirname(os.path.abspath(__file__))


def rnd(gen, *shape, scale=1.0, uniform=False):
    t = torch.rand(*shape, generator=gen) if uniform else torch.randn(*shape, generator=gen)
    return (t * scale).float()


def both(fn, *inputs, grads=None):
    """Run fn on float64 and float32 copies; returns {name_f64/f32: ndarray}.  If `grads` is given
    (indices of inputs to differentiate), also returns d(sum(out * w))/d(input) for a fixed w."""
    res = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        ins = [i.to(dt).clone().requires_grad_(grads is not None and k in grads) if torch.is_tensor(i) else i
               for k, i in enumerate(inputs)]
        out = fn(*ins)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        for k, o in enumerate(outs):
            res["out%d_%s" % (k, tag)] = o.detach().numpy()
        if grads is not None:
            g = torch.Generator().manual_seed(1234)
            loss = 0
            for o in outs:
                if o.requires_grad:
                    w = torch.randn(o.shape, generator=g).to(dt)
                    loss = loss + (o * w).sum()
            gs = torch.autograd.grad(loss, [ins[k] for k in grads], allow_unused=True)
            for k, gk in zip(grads, gs):
                if gk is not None:
                    res["grad%d_%s" % (k, tag)] = gk.numpy()
    return res


def save(name, inputs, res):
    d = {"in%d" % k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in enumerate(inputs)}
    d.update(res)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("%-28s %6.1f KB" % (name, os.path.getsize(os.path.join(OUT, name + ".npz")) / 1024))


def main():
    gen = torch.Generator().manual_seed(0)

    # ---- correlation: correlation_native.Correlation == uflow_model.compute_cost_volume ----
    from models.correlation_native import Correlation
    from models.uflow_model import compute_cost_volume, normalize_features
    corr = Correlation(max_displacement=4, kernel_size=1, stride1=1, stride2=1, corr_multiply=1)
    for name, shape in (("corr_b2c5_9x11", (2, 5, 9, 11)), ("corr_b1c32_12x16", (1, 32, 12, 16)),
                        ("corr_b1c3_6x40", (1, 3, 6, 40))):
        f1, f2 = rnd(gen, *shape), rnd(gen, *shape)
        res = both(lambda a, b: corr(a, b), f1, f2, grads=(0, 1))
        chk = both(lambda a, b: compute_cost_volume(a, b, 4), f1, f2)
        assert np.array_equal(res["out0_f64"], chk["out0_f64"])
        save(name, (f1, f2), res)

    # ---- warp: flow_warp and resample ----
    from utils.warp_utils import flow_warp
    from utils import uflow_utils as uu
    x, flow = rnd(gen, 2, 3, 7, 9), rnd(gen, 2, 2, 7, 9, scale=2.0)
    for pad in ("zeros", "border", "reflection"):
        for align in (True, False):
            res = both(lambda a, f: flow_warp(a, f, pad=pad, align_corners=align), x, flow, grads=(0, 1))
            save("flow_warp_%s_%d" % (pad, align), (x, flow), res)
    res = both(lambda a, f: flow_warp(a, f, mode="nearest"), x, flow)
    save("flow_warp_nearest", (x, flow), res)
    xs = rnd(gen, 2, 4, 6, 10)   # source size != flow size
    res = both(lambda a, f: flow_warp(a, f), xs, flow, grads=(0, 1))
    save("flow_warp_othersize", (xs, flow), res)
    x, flow = rnd(gen, 2, 5, 8, 12), rnd(gen, 2, 2, 8, 12, scale=3.0)
    res = both(lambda a, f: uu.resample(a, uu.flow_to_warp(f)), x, flow, grads=(0, 1))
    save("resample", (x, flow), res)

    # ---- masks / range map / occlusion ----
    from utils import warp_utils as wu
    flow = rnd(gen, 2, 2, 12, 16, scale=3.0)
    flow_b = rnd(gen, 2, 2, 12, 16, scale=3.0)
    res = {}
    for k, v in both(lambda f: uu.mask_invalid(uu.flow_to_warp(f)), flow).items(): res["invalid_" + k] = v
    for k, v in both(lambda f: uu.compute_range_map(f), flow).items(): res["range_" + k] = v
    for k, v in both(lambda f: wu.compute_range_map(f), flow).items(): res["range_wu_" + k] = v
    for k, v in both(lambda f: wu.get_corresponding_map(uu.flow_to_warp(f)), flow).items(): res["corrmap_" + k] = v
    for k, v in both(lambda f: wu.get_occu_mask_backward(f, th=0.2), flow).items(): res["occ_bw_" + k] = v
    for k, v in both(lambda f: wu.get_occu_mask_backward(f, th=0.0), flow).items(): res["occ_bw0_" + k] = v
    for k, v in both(lambda f: wu.border_mask(f), flow).items(): res["border_" + k] = v
    small, small_b = flow * 0.3, flow_b * 0.3
    for k, v in both(lambda a, b: wu.get_occu_mask_bidirection(a, b), small, small_b).items(): res["occ_bi_" + k] = v
    save("masks", (flow, flow_b), res)

    # ---- resize ----
    img = rnd(gen, 2, 3, 6, 10)
    res = {}
    for s in (2.0, 4.0):
        for is_flow in (False, True):
            for k, v in both(lambda a: uu.upsample(a, is_flow, scale_factor=s), img, grads=(0,)).items():
                res["up%d_%d_%s" % (s, is_flow, k)] = v
    big = rnd(gen, 2, 3, 16, 24)
    for s in (2.0, 4.0):
        for is_flow in (False, True):
            for k, v in both(lambda a: uu.downsample(a, is_flow, scale_factor=s), big, grads=(0,)).items():
                res["down%d_%d_%s" % (s, is_flow, k)] = v
    save("resize", (img, big), res)

    # ---- census / ternary ----
    from losses import loss_blocks as lb
    a, b = rnd(gen, 2, 3, 20, 24, uniform=True), rnd(gen, 2, 3, 20, 24, uniform=True)
    mask = (rnd(gen, 2, 1, 20, 24, uniform=True) > 0.3).float() * rnd(gen, 2, 1, 20, 24, uniform=True)
    res = {}
    for k, v in both(lambda x, y, m: uu.census_loss(x, y, m), a, b, mask, grads=(0, 1)).items(): res["loss_" + k] = v
    for k, v in both(lambda x, y, m: uu.census_loss_no_penalty(x, y, m), a, b, mask, grads=(0, 1)).items(): res["nopen_" + k] = v
    for k, v in both(lambda x, y: lb.TernaryLoss(x, y, max_distance=1)[0], a, b, grads=(1,)).items(): res["tern1_" + k] = v
    for k, v in both(lambda x, y: lb.TernaryLoss(x, y, max_distance=3, sum_dist=True)[0], a, b).items(): res["tern3s_" + k] = v
    for k, v in both(lambda x, y: lb.TernaryLoss(x, y, max_distance=2)[1], a, b).items(): res["tern2mask_" + k] = v
    save("census", (a, b, mask), res)

    # ---- smoothness blocks of loss_blocks ----
    flo, image = rnd(gen, 2, 2, 10, 14), rnd(gen, 2, 3, 10, 14, uniform=True)
    res = {}
    for k, v in both(lambda f, i: lb.smooth_grad_1st(f, i, 10.0), flo, image, grads=(0,)).items(): res["s1abs_" + k] = v
    for k, v in both(lambda f, i: lb.smooth_grad_1st(f, i, 10.0, penalty="uflow"), flo, image, grads=(0,)).items(): res["s1uf_" + k] = v
    for k, v in both(lambda f, i: lb.smooth_grad_2nd(f, i, 10.0), flo, image, grads=(0,)).items(): res["s2_" + k] = v
    save("smooth_blocks", (flo, image), res)

    # ---- UFlowLoss end to end (both smoothness orders) ----
    from losses.uflow_loss import UFlowLoss
    from easydict import EasyDict
    H, W = 32, 40
    out0 = rnd(gen, 2, 4, H, W, scale=2.0)
    out1 = rnd(gen, 2, 4, H // 2, W // 2)
    out2 = rnd(gen, 2, 4, H // 4, W // 4, scale=0.7)
    target = rnd(gen, 2, 6, H, W, uniform=True)
    for order in (1, 2):
        cfg = EasyDict(edge_constant=150, w_smooth=4.0, w_census=1.0, with_bk=True, smooth_order=order)
        res = both(lambda o0, o1, o2, t: UFlowLoss(cfg)([o0, o1, o2], t), out0, out1, out2, target, grads=(0, 2))
        save("uflow_loss_order%d" % order, (out0, out1, out2, target), res)

    # ---- SSIM blocks and the NHWC resampler ----
    from utils import uflow_resampler as ur
    a, b = rnd(gen, 2, 3, 14, 18, uniform=True), rnd(gen, 2, 3, 14, 18, uniform=True)
    mask = rnd(gen, 2, 1, 14, 18, uniform=True)
    res = {}
    for kk, v in both(lambda x, y, m: tuple(uu.ssim_loss(x, y, m)[0]) + (uu.ssim_loss(x, y, m)[1],), a, b, mask,
                      grads=(0, 1)).items(): res["ssimloss_" + kk] = v
    for kk, v in both(lambda x, y: lb.SSIM(x, y, md=1), a, b, grads=(0, 1)).items(): res["ssim1_" + kk] = v
    for kk, v in both(lambda x, y: lb.SSIM(x, y, md=2), a, b).items(): res["ssim2_" + kk] = v
    data = rnd(gen, 2, 6, 7, 5)
    wxy = torch.stack([rnd(gen, 2, 4, 9, uniform=True) * 8 - 1, rnd(gen, 2, 4, 9, uniform=True) * 7 - 1], -1)
    wxy[0, 0, 0] = torch.tensor([2.0, 3.0])      # exact integer coordinates: floor == ceil
    for kk, v in both(lambda d, w: ur.resampler(d, w), data, wxy, grads=(0, 1)).items(): res["resampler_" + kk] = v
    res["data"], res["warp"] = data.numpy(), wxy.numpy()
    save("ssim_resampler", (a, b, mask), res)

    # ---- UFlowElboLoss (noise injected so the fixture is reproducible) ----
    from losses.uflow_elbo_loss import UFlowElboLoss
    He, We, Be = 32, 40, 2
    base = dict(edge_constant=150, edge_asymp=0.01, w_smooth=4.0, penalty_smooth="charbonnier", data_loss=["census"],
                data_weight=[1.0], data_penalty=["abs_robust_loss"], w_entropy=0.1, with_bk=True, n_components=1,
                inv_cov=False, approx_entropy=False, natural_grad=False, isotropic_smooth=False)
    cases = {
        "elbo_sparse": dict(base, approx="sparse", cov_supp=3, n_samples=2, occ_type="sample", closed_form_smooth=False,
                            w_oof=0.0, w_occ=0.0, offdiag_reg=0.01),
        "elbo_diag": dict(base, approx="diag", cov_supp=0, n_samples=1, occ_type="mean", closed_form_smooth=True,
                          order_smooth=1, w_oof=0.5, w_occ=0.3, offdiag_reg=0.0),
    }
    for name, c in cases.items():
        cfg = EasyDict(c)
        nch = 4 if c["approx"] == "diag" else 4 + 2 * ((c["cov_supp"] + 1) ** 2 - 1)
        fw2 = rnd(gen, Be, nch, He // 4, We // 4, scale=0.5)
        bw2 = rnd(gen, Be, nch, He // 4, We // 4, scale=0.5)
        im1, im2 = rnd(gen, Be, 3, He, We, uniform=True), rnd(gen, Be, 3, He, We, uniform=True)
        eps = [rnd(gen, c["n_samples"] * Be, 2, He // 4, We // 4) for _ in range(2)]

        def run(f, b, i1, i2):
            loss = UFlowElboLoss(cfg)
            it = iter(eps)
            loss.Normal.sample = lambda size: next(it).to(f.dtype)
            out = loss({"flows_fw": [None, None, f], "flows_bw": [None, None, b]}, i1, i2)
            return out[:5] if not isinstance(out[4], (int, float)) else out[:4] + (torch.zeros((), dtype=f.dtype),)
        res = both(run, fw2, bw2, im1, im2, grads=(0, 1))
        res["eps0"], res["eps1"] = eps[0].numpy(), eps[1].numpy()
        res["cfg"] = np.asarray(json.dumps(c))
        save(name, (fw2, bw2, im1, im2), res)

    # ---- stencil-triangular algebra ----
    from utils import triag_solve as ts
    res = {}
    for k in (1, 3):
        A = rnd(gen, 2, 2 * (k + 1) ** 2, 7, 9)
        X = rnd(gen, 2, 2, 7, 9)
        for key, fn in (("mv", ts.matrix_vector_product_general), ("mvT", ts.matrix_vector_product_T_general)):
            r = both(lambda a, x: fn(a, x, k=k), A, X, grads=(0, 1))
            for kk, v in r.items(): res["%s%d_%s" % (key, k, kk)] = v
        res["A%d" % k], res["X%d" % k] = A.numpy(), X.numpy()
    a4 = 1.0 + rnd(gen, 2, 2, 6, 7, uniform=True)
    b4, c4, d4 = rnd(gen, 2, 2, 6, 6, scale=0.4), rnd(gen, 2, 2, 5, 7, scale=0.4), rnd(gen, 2, 2, 5, 6, scale=0.4)
    x4 = rnd(gen, 2, 2, 6, 7)
    for kk, v in both(ts.forward_substitution, a4, b4, c4, d4, x4).items(): res["fsub_" + kk] = v
    for kk, v in both(ts.backward_substitution, a4, b4, c4, d4, x4).items(): res["bsub_" + kk] = v
    for kk, v in both(ts.matrix_vector_product, a4, b4, c4, d4, x4).items(): res["mv4_" + kk] = v
    for kk, v in both(ts.matrix_vector_product_T, a4, b4, c4, d4, x4).items(): res["mv4T_" + kk] = v
    a5, b5, c5 = a4[:, :, :4, :5].contiguous(), b4[:, :, :4, :4].contiguous(), c4[:, :, :3, :5].contiguous()
    def marginal(a, b, c):
        # marginal_variances (triag_solve.py:205-218) as intended; as written it calls the 5-argument
        # solver with 4 arguments, and marginal_variances_fast fails on a shape error (:263)
        Hh = torch.zeros_like(a)
        d0 = torch.zeros_like(a[:, :, :-1, :-1])
        for i in range(a.shape[2]):
            for j in range(a.shape[3]):
                e = torch.zeros_like(a)
                e[:, :, i, j] = 1
                y = ts.forward_substitution(a, b, c, d0, e)
                Hh[:, :, i, j] = torch.sum(y * y, dim=(2, 3))
        return Hh
    for kk, v in both(marginal, a5, b5, c5).items(): res["invdiag_" + kk] = v
    save("triag", (a4, b4, c4, d4, x4), res)

    # ---- PWCFlow network (structure check of the caller): eval mode, seeded Xavier init ----
    from models.uflow_model import PWCFlow
    cfg = EasyDict(level_dropout=0.1, feature_norm=True)
    torch.manual_seed(123)          # weights = PyTorch default init in construction order
    net = PWCFlow(cfg)
    net.init_weights()              # a no-op in the reference (iterates (name, module) tuples)
    net.eval()
    pair = torch.rand(1, 6, 192, 256, generator=torch.Generator().manual_seed(77))   # regenerated by the test
    with torch.no_grad():
        r = net(pair, with_bk=True)
    n_par = sum(p.numel() for p in net.parameters())
    save("pwcflow_eval", (np.asarray(77),), {"fw2": r["flows_fw"][2].numpy(), "bw2": r["flows_bw"][2].numpy(),
                                   "fw0_mean": np.asarray(r["flows_fw"][0].abs().mean().item()),
                                   "n_params": np.asarray(n_par),
                                   "keys": np.asarray(list(net.state_dict().keys()))})


def prob_model():
    """PWCProbFlow (config 3's network): eval mode, PyTorch default init in construction order."""
    from easydict import EasyDict
    from models.uflow_prob_model import PWCProbFlow
    for tag, cfg, seed in (("nondiag", EasyDict(out_channels=[2, 2, 30], inv_cov=False, n_pyramids=1, mixture_weights=False,
                                                feature_norm=True, level_dropout=0.1), 321),
                           ("diag2pyr", EasyDict(out_channels=[2, 2, 0], inv_cov=True, n_pyramids=2, mixture_weights=False,
                                                 feature_norm=True, level_dropout=0.1), 322)):
        torch.manual_seed(seed)
        net = PWCProbFlow(cfg)
        net.init_weights()          # a no-op in the reference (iterates (name, module) tuples)
        net.eval()
        gen = torch.Generator().manual_seed(seed + 1000)    # regenerated by the test
        im1, im2 = torch.rand(1, 3, 192, 256, generator=gen), torch.rand(1, 3, 192, 256, generator=gen)
        with torch.no_grad():
            r = net(im1, im2, with_bk=True)
        save("pwcprobflow_" + tag, (np.asarray(seed),),
             {"fw2": r["flows_fw"][2].numpy(), "bw2": r["flows_bw"][2].numpy(),
              "fw0_absmean": np.asarray(r["flows_fw"][0].abs().mean(dim=(0, 2, 3)).numpy()),
              "fw4": r["flows_fw"][4].numpy(),
              "n_params": np.asarray(sum(p.numel() for p in net.parameters())),
              "keys": np.asarray(list(net.state_dict().keys()))})


def round2():
    """Fixtures added in round 2 (own generators, so the files above are unchanged): normalize_features,
    align_corners=True flow up-sampling, a larger inverse-diagonal case, the PWC-Lite network."""
    from easydict import EasyDict
    gen = torch.Generator().manual_seed(2024)

    # ---- normalize_features (uflow_model.py:8-50), the setting PWCFlow uses + two other switch combinations ----
    from models.uflow_model import normalize_features
    f1, f2 = rnd(gen, 2, 6, 9, 11, scale=1.7) + 0.3, rnd(gen, 2, 6, 9, 11, scale=0.6) - 0.2
    res = {}
    for tag, kw in (("all", dict(normalize=True, center=True, moments_across_channels=True, moments_across_images=True)),
                    ("perimg", dict(normalize=True, center=True, moments_across_channels=True, moments_across_images=False)),
                    ("perch", dict(normalize=True, center=False, moments_across_channels=False, moments_across_images=True))):
        for k, v in both(lambda a, b: tuple(normalize_features([a, b], **kw)), f1, f2, grads=(0, 1)).items():
            res["%s_%s" % (tag, k)] = v
    save("normalize_features", (f1, f2), res)

    # ---- F.interpolate(flow * s, scale_factor=s, bilinear, align_corners=True) (pwclite.py:178-179, 203) ----
    import torch.nn.functional as F
    flow = rnd(gen, 2, 2, 6, 10, scale=2.0)
    res = {}
    for s in (2, 4):
        for k, v in both(lambda a: F.interpolate(a * s, scale_factor=s, mode='bilinear', align_corners=True), flow,
                         grads=(0,)).items():
            res["up%d_%s" % (s, k)] = v
    save("resize_align_corners", (flow,), res)

    # ---- inverse diagonal, larger than the 4x5 case of `triag` (marginal variances, triag_solve.py:205-218) ----
    from utils import triag_solve as ts
    a = 1.0 + rnd(gen, 2, 2, 12, 17, uniform=True)
    b, c = rnd(gen, 2, 2, 12, 16, scale=0.4), rnd(gen, 2, 2, 11, 17, scale=0.4)

    def marginal(a, b, c):
        Hh = torch.zeros_like(a)
        d0 = torch.zeros_like(a[:, :, :-1, :-1])
        for i in range(a.shape[2]):
            for j in range(a.shape[3]):
                e = torch.zeros_like(a)
                e[:, :, i, j] = 1
                y = ts.forward_substitution(a, b, c, d0, e)
                Hh[:, :, i, j] = torch.sum(y * y, dim=(2, 3))
        return Hh
    save("invdiag_12x17", (a, b, c), both(marginal, a, b, c))

    # ---- PWCLite (config 1's network): eval mode, PyTorch default init in construction order ----
    from models.pwclite import PWCLite
    for tag, cfg, seed, shape in (("pwclite_eval", EasyDict(upsample=True, n_frames=2, reduce_dense=True), 511, (1, 6, 128, 192)),
                                  ("pwclite3_eval", EasyDict(upsample=True, n_frames=3, reduce_dense=False), 512, (1, 9, 128, 128))):
        torch.manual_seed(seed)
        net = PWCLite(cfg)
        net.init_weights()          # a no-op in the reference (iterates (name, module) tuples)
        net.eval()
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(seed + 1000))   # regenerated by the test
        with torch.no_grad():
            r = net(x, with_bk=True)
        out = {"n_params": np.asarray(sum(p.numel() for p in net.parameters())),
               "keys": np.asarray(list(net.state_dict().keys())), "shape": np.asarray(shape)}
        for d in ("flows_fw", "flows_bw"):
            for i in (1, 3, len(r[d]) - 1):      # two-frame: 1/4, 1/16, 1/64 resolution
                out["%s%d" % (d[6:], i)] = r[d][i].numpy()
            out["%s0_absmean" % d[6:]] = np.asarray(r[d][0].abs().mean().item())
        save(tag, (np.asarray(seed),), out)


if __name__ == "__main__":
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "probflow":
        prob_model()
    elif len(sys.argv) > 1 and sys.argv[1] == "round2":
        round2()
    else:
        main()
        prob_model()
        round2()

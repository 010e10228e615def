// Backward bilinear warp, forward and gradients.
//
// One kernel serves flow_warp (utils/warp_utils.py:83-90) and resample (utils/uflow_utils.py:53-77).
// The coordinate arithmetic reproduces, in fp32 and operation by operation, what the reference
// does: base grid + flow, normalise to [-1,1], and ATen's grid_sampler un-normalise / padding /
// tap weights.  Layout stays NCHW: a warp walks 32 consecutive output pixels, so each tap of each
// channel plane is a (nearly) contiguous 128-byte gather for smooth flow fields.
//
// Forward: thread <-> (pixel, channel group of 4); coordinates are computed once per thread and
// the 16 tap loads of the group are issued together.
// Backward: block = 32 pixels x G channel groups; every thread accumulates d/dX, d/dY over its
// channels, the G partials are reduced through shared memory in a fixed order (deterministic
// flow gradient, no atomics); the optional source gradient uses red.global.add.f32.
#include "common.cuh"

namespace {

struct WarpGeom {
    int B, C, Hs, Ws, Ho, Wo;
    float nW1, nH1;
    int field_kind, interp, pad_mode, align;
};

__device__ __forceinline__ float unnormalize(float g, int size, int align) {
    // ATen grid_sampler_unnormalize
    if (align) return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), (float)(size - 1));
    return __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), -1.f), 2.f);
}

__device__ __forceinline__ float reflect_coord(float in, int twice_low, int twice_high, float& gmul) {
    if (twice_low == twice_high) { gmul = 0.f; return 0.f; }
    float mn = (float)twice_low / 2.f;
    float span = (float)(twice_high - twice_low) / 2.f;
    in = in - mn;
    float s = 1.f;
    if (in < 0.f) { s = -1.f; in = -in; }
    float extra = fmodf(in, span);
    int flips = (int)floorf(in / span);
    if ((flips & 1) == 0) { gmul = s; return extra + mn; }
    gmul = -s;
    return span - extra + mn;
}

__device__ __forceinline__ float clip_coord(float in, int size, float& gmul) {
    if (in <= 0.f) { gmul = 0.f; return 0.f; }
    float mx = (float)(size - 1);
    if (in >= mx) { gmul = 0.f; return mx; }
    gmul = 1.f;
    return in;
}

// source index along one axis + d(index)/d(pixel coordinate before normalisation)
__device__ __forceinline__ float source_index(float p, float n1, int size, int pad_mode, int align,
                                              float& dmul) {
    float g = __fadd_rn(__fdiv_rn(__fmul_rn(2.0f, p), n1), -1.0f);
    float c = unnormalize(g, size, align);
    float m = align ? (float)(size - 1) / 2.f : (float)size / 2.f;
    if (pad_mode == ARF_PAD_BORDER) {
        float gm;
        c = clip_coord(c, size, gm);
        m *= gm;
    } else if (pad_mode == ARF_PAD_REFLECTION) {
        float g1, g2;
        if (align) c = reflect_coord(c, 0, 2 * (size - 1), g1);
        else       c = reflect_coord(c, -1, 2 * size - 1, g1);
        c = clip_coord(c, size, g2);
        m *= g1 * g2;
    }
    dmul = m * (2.0f / n1);
    return c;
}

__device__ __forceinline__ void pixel_coords(const float* __restrict__ field, const WarpGeom& g, int b,
                                             int i, int j, float& X, float& Y, float& dX, float& dY) {
    size_t hw = (size_t)g.Ho * g.Wo;
    const float* fb = field + (size_t)b * 2 * hw + (size_t)i * g.Wo + j;
    float px = __ldg(fb), py = __ldg(fb + hw);
    if (g.field_kind == ARF_FIELD_FLOW) {
        px = __fadd_rn((float)j, px);
        py = __fadd_rn((float)i, py);
    }
    X = source_index(px, g.nW1, g.Ws, g.pad_mode, g.align, dX);
    Y = source_index(py, g.nH1, g.Hs, g.pad_mode, g.align, dY);
}

constexpr int kFwdCG = 4;  // channels per thread in the forward

__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ field, float* __restrict__ y,
                WarpGeom g) {
    const size_t hwo = (size_t)g.Ho * g.Wo, hws = (size_t)g.Hs * g.Ws;
    const int ngroups = (g.C + kFwdCG - 1) / kFwdCG;
    const long long total = (long long)g.B * ngroups * hwo;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int pix = idx % hwo;
        long long t = idx / hwo;
        int cg = t % ngroups;
        int b = t / ngroups;
        int i = pix / g.Wo, j = pix - i * g.Wo;
        float X, Y, dX, dY;
        pixel_coords(field, g, b, i, j, X, Y, dX, dY);
        const int c0 = cg * kFwdCG;
        const float* xb = x + ((size_t)b * g.C + c0) * hws;
        float* yb = y + ((size_t)b * g.C + c0) * hwo + pix;
        const int nc = min(kFwdCG, g.C - c0);
        if (g.interp == ARF_INTERP_NEAREST) {
            int xn = (int)nearbyintf(X), yn = (int)nearbyintf(Y);
            bool in = xn >= 0 && xn < g.Ws && yn >= 0 && yn < g.Hs;
            for (int c = 0; c < nc; ++c)
                yb[c * hwo] = in ? __ldg(xb + c * hws + (size_t)yn * g.Ws + xn) : 0.f;
            continue;
        }
        float xf = floorf(X), yf = floorf(Y);
        int xw = (int)xf, yn = (int)yf, xe = xw + 1, ys = yn + 1;
        float wnw = ((float)xe - X) * ((float)ys - Y);
        float wne = (X - (float)xw) * ((float)ys - Y);
        float wsw = ((float)xe - X) * (Y - (float)yn);
        float wse = (X - (float)xw) * (Y - (float)yn);
        bool inw = xw >= 0 && xw < g.Ws, ine = xe >= 0 && xe < g.Ws;
        bool inn = yn >= 0 && yn < g.Hs, ins = ys >= 0 && ys < g.Hs;
        // clamp addresses so all loads are legal, mask through the weights
        int xwc = min(max(xw, 0), g.Ws - 1), xec = min(max(xe, 0), g.Ws - 1);
        int ync = min(max(yn, 0), g.Hs - 1), ysc = min(max(ys, 0), g.Hs - 1);
        if (!(inn && inw)) wnw = 0.f;
        if (!(inn && ine)) wne = 0.f;
        if (!(ins && inw)) wsw = 0.f;
        if (!(ins && ine)) wse = 0.f;
        size_t onw = (size_t)ync * g.Ws + xwc, one = (size_t)ync * g.Ws + xec;
        size_t osw = (size_t)ysc * g.Ws + xwc, ose = (size_t)ysc * g.Ws + xec;
        float v[kFwdCG][4];
#pragma unroll
        for (int c = 0; c < kFwdCG; ++c) {
            if (c < nc) {
                const float* p = xb + c * hws;
                v[c][0] = __ldg(p + onw); v[c][1] = __ldg(p + one);
                v[c][2] = __ldg(p + osw); v[c][3] = __ldg(p + ose);
            }
        }
#pragma unroll
        for (int c = 0; c < kFwdCG; ++c) {
            if (c < nc) {
                float o = v[c][0] * wnw;
                o = fmaf(v[c][1], wne, o);
                o = fmaf(v[c][2], wsw, o);
                o = fmaf(v[c][3], wse, o);
                yb[c * hwo] = o;
            }
        }
    }
}

// block = (32 pixels, G channel groups); thread (px, gy) handles channels gy, gy+G, ...
template <int G>
__global__ void __launch_bounds__(32 * G)
warp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ field,
                const float* __restrict__ gy, float* __restrict__ gx, float* __restrict__ gfield,
                WarpGeom g) {
    __shared__ float red[2][G][32];
    const size_t hwo = (size_t)g.Ho * g.Wo, hws = (size_t)g.Hs * g.Ws;
    const int lane = threadIdx.x, grp = threadIdx.y;
    const long long nrun = ((long long)hwo + 31) / 32;  // 32-pixel runs per image
    const long long total = (long long)g.B * nrun;
    for (long long run = blockIdx.x; run < total; run += gridDim.x) {
        int b = run / nrun;
        long long pix = (run - (long long)b * nrun) * 32 + lane;
        bool live = pix < (long long)hwo;
        float ax = 0.f, ay = 0.f, dX = 0.f, dY = 0.f;
        if (live) {
            int i = pix / g.Wo, j = pix - (long long)i * g.Wo;
            float X, Y;
            pixel_coords(field, g, b, i, j, X, Y, dX, dY);
            if (g.interp == ARF_INTERP_BILINEAR) {
                float xf = floorf(X), yf = floorf(Y);
                int xw = (int)xf, yn = (int)yf, xe = xw + 1, ys = yn + 1;
                float fxe = (float)xe - X, fxw = X - (float)xw, fys = (float)ys - Y, fyn = Y - (float)yn;
                bool inw = xw >= 0 && xw < g.Ws, ine = xe >= 0 && xe < g.Ws;
                bool inn = yn >= 0 && yn < g.Hs, ins = ys >= 0 && ys < g.Hs;
                bool bnw = inn && inw, bne = inn && ine, bsw = ins && inw, bse = ins && ine;
                int xwc = min(max(xw, 0), g.Ws - 1), xec = min(max(xe, 0), g.Ws - 1);
                int ync = min(max(yn, 0), g.Hs - 1), ysc = min(max(ys, 0), g.Hs - 1);
                size_t onw = (size_t)ync * g.Ws + xwc, one = (size_t)ync * g.Ws + xec;
                size_t osw = (size_t)ysc * g.Ws + xwc, ose = (size_t)ysc * g.Ws + xec;
                for (int c = grp; c < g.C; c += G) {
                    float go = __ldg(gy + ((size_t)b * g.C + c) * hwo + pix);
                    const float* p = x + ((size_t)b * g.C + c) * hws;
                    float vnw = bnw ? __ldg(p + onw) : 0.f, vne = bne ? __ldg(p + one) : 0.f;
                    float vsw = bsw ? __ldg(p + osw) : 0.f, vse = bse ? __ldg(p + ose) : 0.f;
                    ax += go * ((vne - vnw) * fys + (vse - vsw) * fyn);
                    ay += go * ((vsw - vnw) * fxe + (vse - vne) * fxw);
                    if (gx) {
                        float* q = gx + ((size_t)b * g.C + c) * hws;
                        if (bnw) atomicAdd(q + onw, go * fxe * fys);
                        if (bne) atomicAdd(q + one, go * fxw * fys);
                        if (bsw) atomicAdd(q + osw, go * fxe * fyn);
                        if (bse) atomicAdd(q + ose, go * fxw * fyn);
                    }
                }
            } else if (gx) {  // nearest: no field gradient, source gradient is a plain scatter
                int xn = (int)nearbyintf(X), yn = (int)nearbyintf(Y);
                if (xn >= 0 && xn < g.Ws && yn >= 0 && yn < g.Hs)
                    for (int c = grp; c < g.C; c += G)
                        atomicAdd(gx + ((size_t)b * g.C + c) * hws + (size_t)yn * g.Ws + xn,
                                  __ldg(gy + ((size_t)b * g.C + c) * hwo + pix));
            }
        }
        if (gfield) {
            if (G > 1) {
                red[0][grp][lane] = ax;
                red[1][grp][lane] = ay;
                __syncthreads();
                if (grp == 0) {
                    ax = 0.f; ay = 0.f;
#pragma unroll
                    for (int k = 0; k < G; ++k) { ax += red[0][k][lane]; ay += red[1][k][lane]; }
                }
                __syncthreads();
            }
            if (grp == 0 && live) {
                float* gf = gfield + (size_t)b * 2 * hwo + pix;
                gf[0] = ax * dX;
                gf[hwo] = ay * dY;
            }
        }
    }
}

int make_geom(WarpGeom& g, int B, int C, int Hs, int Ws, int Ho, int Wo, float nW1, float nH1,
              int field_kind, int interp, int pad_mode, int align) {
    if (B <= 0 || C <= 0 || Hs <= 0 || Ws <= 0 || Ho <= 0 || Wo <= 0) return ARF_EINVAL;
    if (field_kind != ARF_FIELD_FLOW && field_kind != ARF_FIELD_COORDS) return ARF_EINVAL;
    if (interp != ARF_INTERP_BILINEAR && interp != ARF_INTERP_NEAREST) return ARF_EUNSUPPORTED;
    if (pad_mode < ARF_PAD_ZEROS || pad_mode > ARF_PAD_REFLECTION) return ARF_EINVAL;
    g.B = B; g.C = C; g.Hs = Hs; g.Ws = Ws; g.Ho = Ho; g.Wo = Wo;
    g.nW1 = nW1; g.nH1 = nH1;
    g.field_kind = field_kind; g.interp = interp; g.pad_mode = pad_mode; g.align = align ? 1 : 0;
    return ARF_OK;
}

}  // namespace

extern "C" int arf_warp_fwd(const float* x, const float* field, float* y, int B, int C, int Hs, int Ws,
                            int Ho, int Wo, float nW1, float nH1, int field_kind, int interp,
                            int pad_mode, int align_corners, void* stream) {
    ARF_REQUIRE(x && field && y);
    WarpGeom g;
    int rc = make_geom(g, B, C, Hs, Ws, Ho, Wo, nW1, nH1, field_kind, interp, pad_mode, align_corners);
    if (rc) return rc;
    long long total = (long long)B * ((C + kFwdCG - 1) / kFwdCG) * Ho * Wo;
    warp_fwd_kernel<<<arf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(x, field, y, g);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_warp_bwd(const float* x, const float* field, const float* gy, float* gx, float* gfield,
                            int B, int C, int Hs, int Ws, int Ho, int Wo, float nW1, float nH1,
                            int field_kind, int interp, int pad_mode, int align_corners, void* stream) {
    ARF_REQUIRE(x && field && gy);
    WarpGeom g;
    int rc = make_geom(g, B, C, Hs, Ws, Ho, Wo, nW1, nH1, field_kind, interp, pad_mode, align_corners);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (gx) {
        cudaError_t e = cudaMemsetAsync(gx, 0, (size_t)B * C * Hs * Ws * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    if (!gx && !gfield) return ARF_OK;
    long long runs = (long long)B * (((long long)Ho * Wo + 31) / 32);
    int grid = (int)(runs < (long long)ARF_NUM_SMS * 16 ? runs : (long long)ARF_NUM_SMS * 16);
    if (C <= 4) {
        warp_bwd_kernel<1><<<grid, dim3(32, 1), 0, st>>>(x, field, gy, gx, gfield, g);
    } else {
        warp_bwd_kernel<8><<<grid, dim3(32, 8), 0, st>>>(x, field, gy, gx, gfield, g);
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

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

// Bilinear tap geometry of one output pixel (ATen grid_sampler_2d): weights already zeroed for taps
// outside the source, offsets clamped so every load is legal.
struct Taps {
    size_t o[4];       // nw, ne, sw, se element offsets inside one source plane
    float w[4];        // matching weights (0 for out-of-range taps)
    float fxe, fxw, fys, fyn;
    bool in[4];
};

__device__ __forceinline__ void make_taps(float X, float Y, const WarpGeom& g, Taps& t) {
    float xf = floorf(X), yf = floorf(Y);
    int xw = (int)xf, yn = (int)yf, xe = xw + 1, ys = yn + 1;
    t.fxe = (float)xe - X; t.fxw = X - (float)xw; t.fys = (float)ys - Y; t.fyn = Y - (float)yn;
    bool inw = xw >= 0 && xw < g.Ws, ine = xe >= 0 && xe < g.Ws;
    bool inn = yn >= 0 && yn < g.Hs, ins = ys >= 0 && ys < g.Hs;
    t.in[0] = inn && inw; t.in[1] = inn && ine; t.in[2] = ins && inw; t.in[3] = ins && ine;
    int xwc = min(max(xw, 0), g.Ws - 1), xec = min(max(xe, 0), g.Ws - 1);
    int ync = min(max(yn, 0), g.Hs - 1), ysc = min(max(ys, 0), g.Hs - 1);
    t.o[0] = (size_t)ync * g.Ws + xwc; t.o[1] = (size_t)ync * g.Ws + xec;
    t.o[2] = (size_t)ysc * g.Ws + xwc; t.o[3] = (size_t)ysc * g.Ws + xec;
    t.w[0] = t.in[0] ? t.fxe * t.fys : 0.f;
    t.w[1] = t.in[1] ? t.fxw * t.fys : 0.f;
    t.w[2] = t.in[2] ? t.fxe * t.fyn : 0.f;
    t.w[3] = t.in[3] ? t.fxw * t.fyn : 0.f;
}

constexpr int kCU = 4;   // channels whose loads are issued together

// grid = (pixel blocks, channel splits, batch); a thread computes its pixel's coordinates once and walks
// its channel range four channels (16 independent tap loads) at a time.
__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ field, float* __restrict__ y,
                WarpGeom g, int ch_per_split) {
    const size_t hwo = (size_t)g.Ho * g.Wo, hws = (size_t)g.Hs * g.Ws;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= hwo) return;
    const int b = blockIdx.z;
    const int c_lo = blockIdx.y * ch_per_split, c_hi = min(g.C, c_lo + ch_per_split);
    const int i = (int)(pix / g.Wo), j = (int)(pix - (size_t)i * g.Wo);
    float X, Y, dX, dY;
    pixel_coords(field, g, b, i, j, X, Y, dX, dY);
    const float* xb = x + (size_t)b * g.C * hws;
    float* yb = y + (size_t)b * g.C * hwo + pix;
    if (g.interp == ARF_INTERP_NEAREST) {
        int xn = (int)nearbyintf(X), yn = (int)nearbyintf(Y);
        bool in = xn >= 0 && xn < g.Ws && yn >= 0 && yn < g.Hs;
        for (int c = c_lo; c < c_hi; ++c)
            yb[(size_t)c * hwo] = in ? __ldg(xb + (size_t)c * hws + (size_t)yn * g.Ws + xn) : 0.f;
        return;
    }
    Taps t;
    make_taps(X, Y, g, t);
    for (int c0 = c_lo; c0 < c_hi; c0 += kCU) {
        float v[kCU][4];
#pragma unroll
        for (int u = 0; u < kCU; ++u)
            if (c0 + u < c_hi) {
                const float* p = xb + (size_t)(c0 + u) * hws;
#pragma unroll
                for (int k = 0; k < 4; ++k) v[u][k] = __ldg(p + t.o[k]);
            }
#pragma unroll
        for (int u = 0; u < kCU; ++u)
            if (c0 + u < c_hi) {
                float o = v[u][0] * t.w[0];
                o = fmaf(v[u][1], t.w[1], o);
                o = fmaf(v[u][2], t.w[2], o);
                o = fmaf(v[u][3], t.w[3], o);
                __stcs(yb + (size_t)(c0 + u) * hwo, o);
            }
    }
}

// block = (32 pixels, G channel groups); thread (px, grp) handles channels grp, grp+G, ... four at a time.
template <int G>
__global__ void __launch_bounds__(32 * G)
warp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ field,
                const float* __restrict__ gy, float* __restrict__ gx, float* __restrict__ gfield,
                WarpGeom g) {
    __shared__ float red[2][G][32];
    const size_t hwo = (size_t)g.Ho * g.Wo, hws = (size_t)g.Hs * g.Ws;
    const int lane = threadIdx.x, grp = threadIdx.y;
    const int b = blockIdx.y;
    const long long nrun = ((long long)hwo + 31) / 32;  // 32-pixel runs per image
    for (long long run = blockIdx.x; run < nrun; run += gridDim.x) {
        const long long pix = run * 32 + lane;
        const bool live = pix < (long long)hwo;
        float ax = 0.f, ay = 0.f, dX = 0.f, dY = 0.f;
        if (live) {
            int i = (int)(pix / g.Wo), j = (int)(pix - (long long)i * g.Wo);
            float X, Y;
            pixel_coords(field, g, b, i, j, X, Y, dX, dY);
            const float* xb = x + (size_t)b * g.C * hws;
            const float* gb = gy + (size_t)b * g.C * hwo + pix;
            if (g.interp == ARF_INTERP_BILINEAR) {
                Taps t;
                make_taps(X, Y, g, t);
                for (int c0 = grp; c0 < g.C; c0 += kCU * G) {
                    float go[kCU], v[kCU][4];
#pragma unroll
                    for (int u = 0; u < kCU; ++u) {
                        const int c = c0 + u * G;
                        if (c < g.C) {
                            go[u] = __ldg(gb + (size_t)c * hwo);
                            const float* p = xb + (size_t)c * hws;
#pragma unroll
                            for (int k = 0; k < 4; ++k) v[u][k] = t.in[k] ? __ldg(p + t.o[k]) : 0.f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kCU; ++u) {
                        const int c = c0 + u * G;
                        if (c < g.C) {
                            ax = fmaf(go[u], (v[u][1] - v[u][0]) * t.fys + (v[u][3] - v[u][2]) * t.fyn, ax);
                            ay = fmaf(go[u], (v[u][2] - v[u][0]) * t.fxe + (v[u][3] - v[u][1]) * t.fxw, ay);
                            if (gx) {
                                float* q = gx + ((size_t)b * g.C + c) * hws;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (t.in[k]) atomicAdd(q + t.o[k], go[u] * t.w[k]);
                            }
                        }
                    }
                }
            } else if (gx) {  // nearest: no field gradient, source gradient is a plain scatter
                int xn = (int)nearbyintf(X), yn = (int)nearbyintf(Y);
                if (xn >= 0 && xn < g.Ws && yn >= 0 && yn < g.Hs)
                    for (int c = grp; c < g.C; c += G)
                        atomicAdd(gx + ((size_t)b * g.C + c) * hws + (size_t)yn * g.Ws + xn, __ldg(gb + (size_t)c * hwo));
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
    if (B > 65535) return ARF_EINVAL;
    // enough threads to fill the machine (~2 waves of 2048 threads/SM) before splitting channels further
    const long long px = (long long)B * Ho * Wo;
    long long want = (2LL * ARF_NUM_SMS * 2048 + px - 1) / px;
    int groups = (C + kCU - 1) / kCU;
    int nsplit = (int)(want < 1 ? 1 : (want > groups ? groups : want));
    int ch_per_split = ((groups + nsplit - 1) / nsplit) * kCU;
    nsplit = (C + ch_per_split - 1) / ch_per_split;
    dim3 grid(arf_cdiv((long long)Ho * Wo, 256), nsplit, B);
    warp_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, field, y, g, ch_per_split);
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
    if (B > 65535) return ARF_EINVAL;
    const long long runs = ((long long)Ho * Wo + 31) / 32;
    const long long cap = ((long long)ARF_NUM_SMS * 32 + B - 1) / B;
    dim3 grid((unsigned)(runs < cap ? runs : cap), B);
    if (C <= 4) {
        warp_bwd_kernel<1><<<grid, dim3(32, 1), 0, st>>>(x, field, gy, gx, gfield, g);
    } else if (C <= 16) {
        warp_bwd_kernel<4><<<grid, dim3(32, 4), 0, st>>>(x, field, gy, gx, gfield, g);
    } else {
        warp_bwd_kernel<8><<<grid, dim3(32, 8), 0, st>>>(x, field, gy, gx, gfield, g);
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

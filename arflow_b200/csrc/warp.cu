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
// Backward, direct kernels (small problems): block = 32 pixels x G channel groups; every thread accumulates
// d/dX, d/dY over its channels, the G partials are reduced through shared memory in a fixed order (deterministic
// flow gradient, no atomics); the optional source gradient uses red.global.add.f32.
// Backward, window kernels (large problems, see "window kernels" below): the source window of a 64 x 4P pixel
// tile is staged in shared memory with 16-byte cp.async (flow gradient), and the scatter pattern of the tile is
// sorted once into a CSR table so the source gradient needs one coalesced red.global.add per touched element
// and channel instead of four scattered ones per pixel and channel.
//
// Tuning record (B200, tools/microbench.py warp): the forward gains nothing from shared-memory staging (74 vs 76 us at
// 64x32x96x128) and runs direct; fp32 shared-memory atomics are CAS loops (~1 lane-add/clk/SM), which is why the
// source gradient sorts its scatter into a CSR table instead; aggregating a lane's east taps into its neighbour's
// west taps by shuffle ("warp-aggregated atomics") halves the red.global.add count for smooth flows but was SLOWER
// in the direct kernel (16x32x96x128 smooth: 93 vs 87 us) - the extra registers cost more occupancy than the
// atomics saved; 32-bit tap offsets with advancing plane pointers: slower (84 vs 74 us forward).
#include "common.cuh"

thread_local int g_warp_variant = 0;   // test hook, per calling thread (arf_debug_set key 3): 1 = force the direct kernels, 2 = force the window kernels

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
// kCUF = channels whose 4 taps are in flight together: 8 for feature maps (53% of the HBM roof for a smooth flow at
// 64x32x96x128 vs 48% with 4), 4 for images (3 channels: 26 vs 30 us at 8x3x384x512)
template <int kCUF>
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
    for (int c0 = c_lo; c0 < c_hi; c0 += kCUF) {
        float v[kCUF][4];
#pragma unroll
        for (int u = 0; u < kCUF; ++u)
            if (c0 + u < c_hi) {
                const float* p = xb + (size_t)(c0 + u) * hws;
#pragma unroll
                for (int k = 0; k < 4; ++k) v[u][k] = __ldg(p + t.o[k]);
            }
#pragma unroll
        for (int u = 0; u < kCUF; ++u)
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
__global__ void __launch_bounds__(G == 1 ? 128 : 32 * G)
warp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ field,
                const float* __restrict__ gy, float* __restrict__ gx, float* __restrict__ gfield,
                WarpGeom g) {
    __shared__ float red[2][G][32];
    const size_t hwo = (size_t)g.Ho * g.Wo, hws = (size_t)g.Hs * g.Ws;
    // G == 1 (few channels): threadIdx.y indexes independent 32-pixel runs instead of channel groups
    const int lane = threadIdx.x, grp = G == 1 ? 0 : threadIdx.y;
    const int rsub = G == 1 ? threadIdx.y : 0, rpb = G == 1 ? blockDim.y : 1;
    const int b = blockIdx.y;
    const long long nrun = ((long long)hwo + 31) / 32;  // 32-pixel runs per image
    for (long long run = (long long)blockIdx.x * rpb + rsub; run < nrun; run += (long long)gridDim.x * rpb) {
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
                            for (int k = 0; k < 4; ++k) v[u][k] = __ldg(p + t.o[k]);   // offsets are clamped: always legal
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kCU; ++u) {
                        const int c = c0 + u * G;
                        if (c < g.C) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) v[u][k] = t.in[k] ? v[u][k] : 0.f;   // taps outside the source read as 0
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

// ------------------------------------------------------------------ window kernels --------
// Large images (bilinear): a CTA owns a 64 x (4P) tile of output pixels of one batch item.  It computes the tap
// geometry of its pixels once, reduces their bounding box over the source image, and stages that WINDOW of the
// source (channel chunk by channel chunk, 16-byte cp.async, double-buffered) in shared memory.  Every HBM/L2
// access is then a full coalesced row segment, whatever the flow looks like; the data-dependent 4-tap gathers
// (and, in the backward, the scatter of the source gradient) run against shared memory.  Pixels whose taps fall
// outside the (size-capped) window take the direct global path, so any flow field is handled; a warp whose 32
// pixels are all inside (the normal case) runs a branch-free, fully unrolled loop.
constexpr int kWTW = 64;          // tile width
constexpr int kWThreads = 256;    // thread <-> (x = tid % 64, rows tid / 64 + 4p)
constexpr int kWMaxW = 96;        // window cap (floats per row, multiple of 4)
constexpr int kWMaxH = 48;        // window cap (rows)
constexpr int kWCap = 12288;      // floats per staging buffer (48 KB); a chunk of CC channels uses stride kWCap/CC

struct TapGeo {
    int xwc, xec, ync, ysc;        // clamped tap columns / rows
    float fxe, fxw, fys, fyn;
    unsigned in;                   // bit k: tap k (nw, ne, sw, se) lies inside the source
};

__device__ __forceinline__ void make_tap_geo(float X, float Y, const WarpGeom& g, TapGeo& t) {
    float xf = floorf(X), yf = floorf(Y);
    int xw = (int)xf, yn = (int)yf, xe = xw + 1, ys = yn + 1;
    t.fxe = (float)xe - X; t.fxw = X - (float)xw; t.fys = (float)ys - Y; t.fyn = Y - (float)yn;
    bool inw = xw >= 0 && xw < g.Ws, ine = xe >= 0 && xe < g.Ws;
    bool inn = yn >= 0 && yn < g.Hs, ins = ys >= 0 && ys < g.Hs;
    t.in = (inn && inw ? 1u : 0u) | (inn && ine ? 2u : 0u) | (ins && inw ? 4u : 0u) | (ins && ine ? 8u : 0u);
    t.xwc = min(max(xw, 0), g.Ws - 1); t.xec = min(max(xe, 0), g.Ws - 1);
    t.ync = min(max(yn, 0), g.Hs - 1); t.ysc = min(max(ys, 0), g.Hs - 1);
}

struct Window { int x0, y0, pw, rows; };   // origin in the source, row pitch (floats), number of rows

// Bounding box of the clamped taps of all live pixels of the CTA -> window (size-capped, centred on the box
// when it is too large); result broadcast through shared memory.  vec: rows are 16-byte copyable.
__device__ __forceinline__ Window block_window(int mnx, int mxx, int mny, int mxy, bool vec, int (*sred)[4], int* swin) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sred[warp][0] = mnx; sred[warp][1] = mxx; sred[warp][2] = mny; sred[warp][3] = mxy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kWThreads / 32; ++w) {
            mnx = min(mnx, sred[w][0]); mxx = max(mxx, sred[w][1]);
            mny = min(mny, sred[w][2]); mxy = max(mxy, sred[w][3]);
        }
        int x0 = 0, y0 = 0, pw = 0, rows = 0;
        if (mxx >= mnx && mxy >= mny) {
            x0 = (mxx - mnx + 1 > kWMaxW) ? max((mnx + mxx) / 2 - kWMaxW / 2, 0) : mnx;
            if (vec) x0 &= ~3;
            int x1 = min(mxx, x0 + kWMaxW - 1);
            y0 = (mxy - mny + 1 > kWMaxH) ? max((mny + mxy) / 2 - kWMaxH / 2, 0) : mny;
            int y1 = min(mxy, y0 + kWMaxH - 1);
            pw = x1 - x0 + 1;
            if (vec) pw = (pw + 3) & ~3;
            rows = y1 - y0 + 1;
        }
        swin[0] = x0; swin[1] = y0; swin[2] = pw; swin[3] = rows;
    }
    __syncthreads();
    Window w;
    w.x0 = swin[0]; w.y0 = swin[1]; w.pw = swin[2]; w.rows = swin[3];
    return w;
}

__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Staging plan of one thread: the window plane is cut into 16-byte vectors, vector e = tid + 256*i goes to
// offset 4e of the staged plane (rows are packed with pitch w.pw) and comes from element soff[i] of the source
// window plane.  Computed once per tile, reused for every channel (4 instructions per vector in the loop).
constexpr int kWSlots = (kWMaxW / 4 * kWMaxH + kWThreads - 1) / kWThreads;   // 5
struct StagePlan { int soff[kWSlots]; };

__device__ __forceinline__ void make_plan(StagePlan& sp, const Window& w, int Ws, bool vec) {
    const int pw4 = w.pw >> 2, nvec = pw4 * w.rows;
#pragma unroll
    for (int i = 0; i < kWSlots; ++i) {
        const int e = threadIdx.x + i * kWThreads;
        sp.soff[i] = -1;
        if (vec && e < nvec) {
            const int row = e / pw4, v = e - row * pw4;
            sp.soff[i] = row * Ws + 4 * v;
        }
    }
}

// stage channels [c_lo, c_lo+n) of the window: channel c at buf + c*S, rows of pitch w.pw
template <int S>
__device__ __forceinline__ void stage_window(float* buf, const float* __restrict__ xb, const Window& w, const StagePlan& sp,
                                             int c_lo, int n, int Hs, int Ws, bool vec) {
    if (vec) {
        const float* src_c = xb + ((size_t)c_lo * Hs + w.y0) * Ws + w.x0;
        float* dst_c = buf + 4 * threadIdx.x;
        const size_t hws = (size_t)Hs * Ws;
        for (int c = 0; c < n; ++c, src_c += hws, dst_c += S) {
#pragma unroll
            for (int i = 0; i < kWSlots; ++i)
                if (sp.soff[i] >= 0) cp_async_16(dst_c + 4 * i * kWThreads, src_c + sp.soff[i]);
        }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < n; ++c) {
        const float* src_c = xb + ((size_t)(c_lo + c) * Hs + w.y0) * Ws + w.x0;
        float* dst_c = buf + c * S;
        for (int row = warp; row < w.rows; row += kWThreads / 32) {
            const float* src = src_c + (size_t)row * Ws;
            float* dst = dst_c + row * w.pw;
            for (int v = lane; v < w.pw; v += 32) cp_async_4(dst + v, src + v);
        }
    }
}

// Per-pixel state kept in registers across the channel loop.
//   fl bits 0-3: taps in range, 4: east step, 5: south step, 6: taps staged in the window, 7: live pixel
template <int P>
struct PixSet {
    int gbase[P];       // element offset of the (clamped) nw tap inside one source plane
    unsigned fl[P];
    int o[P][4];        // offsets of the four taps inside one staged window plane
    float f[P][4];      // fxe, fxw, fys, fyn
    float dm[P][2];     // d(source index)/d(field) along x, y
    bool fast;          // warp-uniform: all 32 x P pixels of the warp are live and staged
    Window w;
};

template <int P>
__device__ __forceinline__ void setup_pixels(const float* __restrict__ field, const WarpGeom& g, int b, int x0, int y0,
                                             bool vec, PixSet<P>& ps, int (*sred)[4], int* swin) {
    const int lx = threadIdx.x & (kWTW - 1), ly = threadIdx.x / kWTW;
    int mnx = 0x7fffffff, mxx = -1, mny = 0x7fffffff, mxy = -1;
    int xwc[P], xec[P], ync[P], ysc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int i = y0 + ly + 4 * p, j = x0 + lx;
        ps.fl[p] = 0; ps.gbase[p] = 0;
        xwc[p] = xec[p] = ync[p] = ysc[p] = 0;
        ps.f[p][0] = ps.f[p][1] = ps.f[p][2] = ps.f[p][3] = 0.f;
        ps.dm[p][0] = ps.dm[p][1] = 0.f;
        if (i < g.Ho && j < g.Wo) {
            float X, Y, dX, dY;
            pixel_coords(field, g, b, i, j, X, Y, dX, dY);
            TapGeo t;
            make_tap_geo(X, Y, g, t);
            ps.gbase[p] = t.ync * g.Ws + t.xwc;
            ps.fl[p] = t.in | (t.xec > t.xwc ? 16u : 0u) | (t.ysc > t.ync ? 32u : 0u) | 128u;
            xwc[p] = t.xwc; xec[p] = t.xec; ync[p] = t.ync; ysc[p] = t.ysc;
            ps.f[p][0] = t.fxe; ps.f[p][1] = t.fxw; ps.f[p][2] = t.fys; ps.f[p][3] = t.fyn;
            ps.dm[p][0] = dX; ps.dm[p][1] = dY;
            if (t.in) {
                mnx = min(mnx, t.xwc); mxx = max(mxx, t.xec);
                mny = min(mny, t.ync); mxy = max(mxy, t.ysc);
            }
        }
    }
    const Window w = block_window(mnx, mxx, mny, mxy, vec, sred, swin);
    ps.w = w;
    bool ok = w.rows > 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        // a pixel without a single tap in range contributes nothing: park its (zero-weight) taps on the window origin
        const bool none = (ps.fl[p] & 15u) == 0;
        const bool inwin = w.rows > 0 && (none || (xwc[p] >= w.x0 && xec[p] < w.x0 + w.pw && ync[p] >= w.y0 &&
                                                   ysc[p] < w.y0 + w.rows));
        const int o0 = (inwin && !none) ? (ync[p] - w.y0) * w.pw + (xwc[p] - w.x0) : 0;
        const int dx = (inwin && !none) ? (xec[p] - xwc[p]) : 0, dy = (inwin && !none) ? (ysc[p] - ync[p]) * w.pw : 0;
        ps.o[p][0] = o0; ps.o[p][1] = o0 + dx; ps.o[p][2] = o0 + dy; ps.o[p][3] = o0 + dy + dx;
        if (inwin) ps.fl[p] |= 64u;
        ok = ok && inwin && (ps.fl[p] & 128u);
    }
    ps.fast = __all_sync(0xffffffffu, ok);
}

// the four tap values of pixel p in one channel: staged copy `win` if the pixel is staged, else global `plane`
template <int P>
__device__ __forceinline__ void load_taps_any(const PixSet<P>& ps, int p, const float* win, const float* __restrict__ plane,
                                              int Ws, float (&v)[4]) {
    if (ps.fl[p] & 64u) {
        v[0] = win[ps.o[p][0]]; v[1] = win[ps.o[p][1]]; v[2] = win[ps.o[p][2]]; v[3] = win[ps.o[p][3]];
    } else {
        const float* q = plane + ps.gbase[p];
        const int dx = (ps.fl[p] >> 4) & 1, dy = (ps.fl[p] & 32u) ? Ws : 0;
        v[0] = __ldg(q); v[1] = __ldg(q + dx); v[2] = __ldg(q + dy); v[3] = __ldg(q + dy + dx);
    }
}

// Backward, flow/coordinate gradient: every thread sums over all channels for its own pixels (fixed order,
// deterministic).   d out/dX = sum_k v_k cx_k,  d out/dY = sum_k v_k cy_k   (cx, cy masked by "tap in range").
// The source gradient is a separate kernel (warp_gx_csr below): it does not need the source at all.
template <int P, int CC>
__device__ __forceinline__ void warp_bwd_body(float* wbuf, const float* __restrict__ xb, const float* __restrict__ gyb,
                                              const WarpGeom& g, const PixSet<P>& ps, const float (&cx)[P][4],
                                              const float (&cy)[P][4], float (&ax)[P], float (&ay)[P], bool vec) {
    constexpr int S = kWCap / CC;
    const Window& w = ps.w;
    const size_t hwo = (size_t)g.Ho * g.Wo, hws = (size_t)g.Hs * g.Ws;
    const int nchunks = (g.C + CC - 1) / CC;
    const int rowstep = 4 * g.Wo;
    StagePlan sp;
    make_plan(sp, w, g.Ws, vec);
    stage_window<S>(wbuf, xb, w, sp, 0, min(CC, g.C), g.Hs, g.Ws, vec);
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int c_lo = ch * CC, n = min(CC, g.C - c_lo);
        const float* cur = wbuf + (ch & 1) * kWCap;
        if (ch + 1 < nchunks)
            stage_window<S>(wbuf + ((ch + 1) & 1) * kWCap, xb, w, sp, c_lo + CC, min(CC, g.C - c_lo - CC), g.Hs, g.Ws, vec);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (ps.fast) {
            // all gy loads of the chunk are issued before the first use (CC x P independent global loads in flight)
            float go[CC][P];
#pragma unroll
            for (int c = 0; c < CC; ++c) {
                const float* gc = gyb + (size_t)(c_lo + (c < n ? c : 0)) * hwo;
#pragma unroll
                for (int p = 0; p < P; ++p) go[c][p] = __ldg(gc + p * rowstep);
            }
#pragma unroll
            for (int c = 0; c < CC; ++c) {
                if (c < n) {
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const float* q = cur + c * S;
                        const float v0 = q[ps.o[p][0]], v1 = q[ps.o[p][1]], v2 = q[ps.o[p][2]], v3 = q[ps.o[p][3]];
                        ax[p] = fmaf(go[c][p], fmaf(v3, cx[p][3], fmaf(v2, cx[p][2], fmaf(v1, cx[p][1], v0 * cx[p][0]))), ax[p]);
                        ay[p] = fmaf(go[c][p], fmaf(v3, cy[p][3], fmaf(v2, cy[p][2], fmaf(v1, cy[p][1], v0 * cy[p][0]))), ay[p]);
                    }
                }
            }
        } else {
            for (int c = 0; c < n; ++c) {
                const float* plane = xb + (size_t)(c_lo + c) * hws;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (ps.fl[p] & 128u) {
                        const float go = __ldg(gyb + (size_t)(c_lo + c) * hwo + p * rowstep);
                        float v[4];
                        load_taps_any<P>(ps, p, cur + c * S, plane, g.Ws, v);
                        ax[p] = fmaf(go, fmaf(v[3], cx[p][3], fmaf(v[2], cx[p][2], fmaf(v[1], cx[p][1], v[0] * cx[p][0]))), ax[p]);
                        ay[p] = fmaf(go, fmaf(v[3], cy[p][3], fmaf(v[2], cy[p][2], fmaf(v[1], cy[p][1], v[0] * cy[p][0]))), ay[p]);
                    }
                }
            }
        }
        __syncthreads();
    }
}

template <int P>
__global__ void __launch_bounds__(kWThreads, 2)
warp_gfield_win(const float* __restrict__ x, const float* __restrict__ field, const float* __restrict__ gy,
                float* __restrict__ gfield, WarpGeom g, int tiles_x, int tiles_y, int vec) {
    extern __shared__ __align__(16) float wbuf[];       // two source-window buffers
    __shared__ int sred[kWThreads / 32][4];
    __shared__ int swin[4];
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * kWTW, y0 = ty * 4 * P;
    const size_t hwo = (size_t)g.Ho * g.Wo, hws = (size_t)g.Hs * g.Ws;
    const float* xb = x + (size_t)b * g.C * hws;
    const int lx = threadIdx.x & (kWTW - 1), ly = threadIdx.x / kWTW;
    const float* gyb = gy + (size_t)b * g.C * hwo + (size_t)(y0 + ly) * g.Wo + (x0 + lx);

    PixSet<P> ps;
    setup_pixels<P>(field, g, b, x0, y0, vec != 0, ps, sred, swin);
    float cx[P][4], cy[P][4], ax[P], ay[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const float m0 = (ps.fl[p] & 1u) ? 1.f : 0.f, m1 = (ps.fl[p] & 2u) ? 1.f : 0.f;
        const float m2 = (ps.fl[p] & 4u) ? 1.f : 0.f, m3 = (ps.fl[p] & 8u) ? 1.f : 0.f;
        cx[p][0] = -ps.f[p][2] * m0; cx[p][1] = ps.f[p][2] * m1; cx[p][2] = -ps.f[p][3] * m2; cx[p][3] = ps.f[p][3] * m3;
        cy[p][0] = -ps.f[p][0] * m0; cy[p][1] = -ps.f[p][1] * m1; cy[p][2] = ps.f[p][0] * m2; cy[p][3] = ps.f[p][1] * m3;
        ax[p] = 0.f; ay[p] = 0.f;
    }
    const int area = ps.w.pw * ps.w.rows;
    if (area > 0) {
        if (area <= kWCap / 8) warp_bwd_body<P, 8>(wbuf, xb, gyb, g, ps, cx, cy, ax, ay, vec != 0);
        else if (area <= kWCap / 4) warp_bwd_body<P, 4>(wbuf, xb, gyb, g, ps, cx, cy, ax, ay, vec != 0);
        else warp_bwd_body<P, 2>(wbuf, xb, gyb, g, ps, cx, cy, ax, ay, vec != 0);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (ps.fl[p] & 128u) {
            float* gf = gfield + (size_t)b * 2 * hwo + (size_t)(y0 + ly + 4 * p) * g.Wo + (x0 + lx);
            gf[0] = ax[p] * ps.dm[p][0];
            gf[hwo] = ay[p] * ps.dm[p][1];
        }
}

// Source gradient without floating-point atomics in shared memory (they are CAS loops on this architecture,
// ~1 lane per clock per SM).  The scatter pattern of a tile - which output pixel adds how much to which window
// element - does not depend on the channel, so it is sorted ONCE per tile into a CSR table in shared memory
// (integer shared atomics only: count, scan, fill).  Per channel chunk the CTA stages its gy tile, and a thread
// that owns a window element sums weight * gy over that element's short entry list in registers for all channels
// of the chunk, then issues one coalesced red.global.add per element and channel.  Taps outside the capped window
// fall back to direct global atomics.
constexpr int kGCc = 8;                         // channels per chunk of the CSR kernel (8: 292 -> 245 us for both gradients; a red.v4 flush of 4 elements per thread: slower)
constexpr int kWArea = kWMaxW * kWMaxH;         // 4608 window elements at most
template <int P>
struct GxSmem {
    int start[kWArea + 4];                      // CSR offsets (see the fill step for the convention)
    float ew[4 * kWTW * 4 * P];                 // entry weights
    unsigned short eidx[4 * kWTW * 4 * P];      // entry -> pixel index inside the tile (row * 64 + x)
    float gos[2][kGCc][4 * P][kWTW];            // staged gy tile, double-buffered
    int sred[kWThreads / 32][4];
    int swin[4];
    int scan[kWThreads / 32];
};

template <int P>
__device__ __forceinline__ void stage_gy(float (*dst)[4 * P][kWTW], const float* __restrict__ gyt, int n, size_t hwo, int Wo,
                                         int rows_live, int cols_live, bool vec) {
    // gyt -> first pixel of the tile in channel c_lo; rows of 64 floats
    constexpr int kVecPerRow = kWTW / 4;
    if (vec && cols_live == kWTW) {
        for (int e = threadIdx.x; e < n * 4 * P * kVecPerRow; e += kWThreads) {
            const int v = e % kVecPerRow, r = (e / kVecPerRow) % (4 * P), c = e / (kVecPerRow * 4 * P);
            if (r < rows_live) cp_async_16(&dst[c][r][4 * v], gyt + (size_t)c * hwo + (size_t)r * Wo + 4 * v);
        }
    } else {
        for (int e = threadIdx.x; e < n * 4 * P * kWTW; e += kWThreads) {
            const int xx = e % kWTW, r = (e / kWTW) % (4 * P), c = e / (kWTW * 4 * P);
            if (r < rows_live && xx < cols_live) cp_async_4(&dst[c][r][xx], gyt + (size_t)c * hwo + (size_t)r * Wo + xx);
        }
    }
}

template <int P>
__global__ void __launch_bounds__(kWThreads, 2)
warp_gx_csr(const float* __restrict__ field, const float* __restrict__ gy, float* __restrict__ gx, WarpGeom g,
            int tiles_x, int tiles_y, int vec_src, int vec_out) {
    extern __shared__ __align__(16) unsigned char gx_raw[];
    GxSmem<P>& sm = *reinterpret_cast<GxSmem<P>*>(gx_raw);
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * kWTW, y0 = ty * 4 * P;
    const size_t hwo = (size_t)g.Ho * g.Wo, hws = (size_t)g.Hs * g.Ws;
    const int lx = threadIdx.x & (kWTW - 1), ly = threadIdx.x / kWTW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* gyb = gy + (size_t)b * g.C * hwo;
    float* gxb = gx + (size_t)b * g.C * hws;

    PixSet<P> ps;
    setup_pixels<P>(field, g, b, x0, y0, vec_src != 0, ps, sm.sred, sm.swin);
    const Window w = ps.w;
    const int area = w.pw * w.rows;
    float wt[P][4];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        wt[p][0] = (ps.fl[p] & 1u) ? ps.f[p][0] * ps.f[p][2] : 0.f;
        wt[p][1] = (ps.fl[p] & 2u) ? ps.f[p][1] * ps.f[p][2] : 0.f;
        wt[p][2] = (ps.fl[p] & 4u) ? ps.f[p][0] * ps.f[p][3] : 0.f;
        wt[p][3] = (ps.fl[p] & 8u) ? ps.f[p][1] * ps.f[p][3] : 0.f;
    }
    // taps that are not staged (outside the capped window): direct global atomics, all channels
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if ((ps.fl[p] & 128u) && !(ps.fl[p] & 64u) && (ps.fl[p] & 15u)) {
            const int dx = (ps.fl[p] >> 4) & 1, dy = (ps.fl[p] & 32u) ? g.Ws : 0;
            const int og[4] = {0, dx, dy, dy + dx};
            const float* gp = gyb + (size_t)(y0 + ly + 4 * p) * g.Wo + (x0 + lx);
            for (int c = 0; c < g.C; ++c) {
                const float go = __ldg(gp + (size_t)c * hwo);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (wt[p][k] != 0.f) atomicAdd(gxb + (size_t)c * hws + ps.gbase[p] + og[k], go * wt[p][k]);
            }
        }
    }
    if (area == 0) return;

    // ---- CSR build: count -> scan -> fill -------------------------------------------------------------------
    for (int e = threadIdx.x; e <= area; e += kWThreads) sm.start[e] = 0;
    __syncthreads();
#pragma unroll
    for (int p = 0; p < P; ++p)
        if ((ps.fl[p] & 192u) == 192u) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (wt[p][k] != 0.f) atomicAdd(&sm.start[ps.o[p][k] + 1], 1);
        }
    __syncthreads();
    {   // inclusive scan of start[0..area]: thread t owns the run [t*run, t*run+run)
        const int run = (area + 1 + kWThreads - 1) / kWThreads;
        const int lo = threadIdx.x * run, hi = min(lo + run, area + 1);
        int sum = 0;
        for (int e = lo; e < hi; ++e) sum += sm.start[e];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.scan[warp] = incl;
        __syncthreads();
        int base = incl - sum;
        for (int wv = 0; wv < warp; ++wv) base += sm.scan[wv];
        for (int e = lo; e < hi; ++e) { base += sm.start[e]; sm.start[e] = base; }
    }
    __syncthreads();
    // now start[q] = first entry of element q, start[q+1] = one past its last.  The fill advances start[q] as a
    // cursor, so afterwards element q owns [q ? start[q-1] : 0, start[q]).
#pragma unroll
    for (int p = 0; p < P; ++p)
        if ((ps.fl[p] & 192u) == 192u) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (wt[p][k] != 0.f) {
                    const int slot = atomicAdd(&sm.start[ps.o[p][k]], 1);
                    sm.ew[slot] = wt[p][k];
                    sm.eidx[slot] = (unsigned short)((ly + 4 * p) * kWTW + lx);
                }
        }
    // ---- channel chunks --------------------------------------------------------------------------------------
    const int rows_live = min(4 * P, g.Ho - y0), cols_live = min(kWTW, g.Wo - x0);
    const float* gyt = gyb + (size_t)y0 * g.Wo + x0;
    const int nchunks = (g.C + kGCc - 1) / kGCc;
    stage_gy<P>(sm.gos[0], gyt, min(kGCc, g.C), hwo, g.Wo, rows_live, cols_live, vec_out != 0);
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int c_lo = ch * kGCc, n = min(kGCc, g.C - c_lo);
        if (ch + 1 < nchunks)
            stage_gy<P>(sm.gos[(ch + 1) & 1], gyt + (size_t)(c_lo + kGCc) * hwo, min(kGCc, g.C - c_lo - kGCc), hwo, g.Wo,
                        rows_live, cols_live, vec_out != 0);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();    // also orders the CSR fill before its first use
        const float* gs = &sm.gos[ch & 1][0][0][0];
        constexpr int kPlane = 4 * P * kWTW;
        for (int row = warp; row < w.rows; row += kWThreads / 32) {
            float* dst = gxb + ((size_t)c_lo * g.Hs + (w.y0 + row)) * g.Ws + w.x0;
            for (int col = lane; col < w.pw; col += 32) {
                const int q = row * w.pw + col;
                const int e0 = q ? sm.start[q - 1] : 0, e1 = sm.start[q];
                if (e1 > e0) {
                    float acc[kGCc];
#pragma unroll
                    for (int c = 0; c < kGCc; ++c) acc[c] = 0.f;
                    for (int j = e0; j < e1; ++j) {
                        const float wj = sm.ew[j];
                        const int pj = sm.eidx[j];
#pragma unroll
                        for (int c = 0; c < kGCc; ++c) acc[c] = fmaf(wj, gs[c * kPlane + pj], acc[c]);
                    }
#pragma unroll
                    for (int c = 0; c < kGCc; ++c)
                        if (c < n) atomicAdd(dst + (size_t)c * hws + col, acc[c]);
                }
            }
        }
        __syncthreads();
    }
}

constexpr size_t kWinSmem = 2 * kWCap * sizeof(float);

// Routing, measured on B200 (tools/microbench.py warp --warp-variant 1|2, smooth flow, C = 32):
//   flow gradient only:   window wins at every size that fills tiles (16x32x48x64: 10.9 vs 13.5 us, 8x32x96x128: 19.7 vs
//                         24.4, 16x32x96x128: 30.2 vs 44.9, 64x32x96x128: 75 vs 206);
//   both gradients:       window + CSR needs a large problem to pay for its per-tile sort (32x32x96x128: 126 vs 133 us,
//                         64x32x96x128: 222 vs 314; 16x32x96x128: 81 vs 71, 8x32x96x128: 53 vs 38 -> direct).
// The forward gains nothing from staging (74 vs 76 us) and always runs direct.
inline bool use_window(const WarpGeom& g, int variant, bool want_gx) {
    if (g.interp != ARF_INTERP_BILINEAR || (size_t)g.Hs * g.Ws >= 0x7fffffffu) return false;
    if (variant == 2) return true;                        // test hook: force
    if (!(g.C >= 8 && g.Wo >= 48 && g.Ho >= 12)) return false;
    return !want_gx || (long long)g.B * g.Ho * g.Wo >= 350000LL;
}
// rows per tile = 4P: the tallest tile that still gives every SM a few CTAs
inline int window_p(const WarpGeom& g) {
    for (int P = 4; P > 1; P >>= 1)
        if ((long long)arf_cdiv(g.Wo, kWTW) * arf_cdiv(g.Ho, 4 * P) * g.B >= 3LL * ARF_NUM_SMS) return P;
    return 1;
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
    const int kCUF = C >= 8 ? 8 : 4;
    int groups = (C + kCUF - 1) / kCUF;
    int nsplit = (int)(want < 1 ? 1 : (want > groups ? groups : want));
    int ch_per_split = ((groups + nsplit - 1) / nsplit) * kCUF;
    nsplit = (C + ch_per_split - 1) / ch_per_split;
    dim3 grid(arf_cdiv((long long)Ho * Wo, 256), nsplit, B);
    if (kCUF == 8) warp_fwd_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(x, field, y, g, ch_per_split);
    else warp_fwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(x, field, y, g, ch_per_split);
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
    if (g_warp_variant != 1 && use_window(g, g_warp_variant, gx != nullptr)) {
        const int P = window_p(g);
        const int tiles_x = arf_cdiv(Wo, kWTW), tiles_y = arf_cdiv(Ho, 4 * P);
        const long long ntiles = (long long)tiles_x * tiles_y * B;
        if (ntiles <= 0x7fffffffLL) {
#define ARF_WIN_ATTR(PP)                                                                \
    do {                                                                                \
        ARF_ENSURE_SMEM(warp_gfield_win<PP>, kWinSmem);                                 \
        ARF_ENSURE_SMEM(warp_gx_csr<PP>, sizeof(GxSmem<PP>));                           \
    } while (0)
            if (P == 4) ARF_WIN_ATTR(4);
            else if (P == 2) ARF_WIN_ATTR(2);
            else ARF_WIN_ATTR(1);
#undef ARF_WIN_ATTR
            const int vec = ((uintptr_t)x % 16 == 0) && (Ws % 4 == 0);
            const int vec_out = ((uintptr_t)gy % 16 == 0) && (Wo % 4 == 0);
#define ARF_WIN_LAUNCH(PP)                                                                                             \
    do {                                                                                                               \
        if (gfield) {                                                                                                  \
            warp_gfield_win<PP><<<(unsigned)ntiles, kWThreads, kWinSmem, st>>>(x, field, gy, gfield, g, tiles_x,        \
                                                                              tiles_y, vec);                           \
            ARF_CHECK_LAUNCH();                                                                                        \
        }                                                                                                              \
        if (gx) {                                                                                                      \
            warp_gx_csr<PP><<<(unsigned)ntiles, kWThreads, sizeof(GxSmem<PP>), st>>>(field, gy, gx, g, tiles_x,         \
                                                                                    tiles_y, vec, vec_out);            \
            ARF_CHECK_LAUNCH();                                                                                        \
        }                                                                                                              \
    } while (0)
            if (P == 4) ARF_WIN_LAUNCH(4);
            else if (P == 2) ARF_WIN_LAUNCH(2);
            else ARF_WIN_LAUNCH(1);
#undef ARF_WIN_LAUNCH
            return ARF_OK;
        }
    }
    const long long runs = ((long long)Ho * Wo + 31) / 32;
    const long long cap = ((long long)ARF_NUM_SMS * 32 + B - 1) / B;
    dim3 grid((unsigned)(runs < cap ? runs : cap), B);
    if (C <= 4) {
        const long long blocks = (runs + 3) / 4;
        grid.x = (unsigned)(blocks < 2 * cap ? blocks : 2 * cap);
        warp_bwd_kernel<1><<<grid, dim3(32, 4), 0, st>>>(x, field, gy, gx, gfield, g);
    } else if (C <= 16) {
        warp_bwd_kernel<4><<<grid, dim3(32, 4), 0, st>>>(x, field, gy, gx, gfield, g);
    } else {
        warp_bwd_kernel<8><<<grid, dim3(32, 8), 0, st>>>(x, field, gy, gx, gfield, g);
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

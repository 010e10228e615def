// Backward bilinear warp, forward and gradients.
//
// One kernel serves flow_warp (utils/warp_utils.py:83-90) and resample (utils/uflow_utils.py:53-77).
// The coordinate arithmetic reproduces, in fp32 and operation by operation, what the reference
// does: base grid + flow, normalise to [-1,1], and ATen's grid_sampler un-normalise / padding /
// tap weights.  Layout stays NCHW: a warp walks 32 consecutive output pixels, so each tap of each
// channel plane is a (nearly) contiguous 128-byte gather for smooth flow fields.
//
// Forward: thread <-> pixel; coordinates are computed once per thread and the tap loads of 8 channels (32 loads) are
// issued together.
// Backward: thread <-> (pixel, channel group); the flow gradient is summed over channels in registers and, when a block
// carries several channel groups, through shared memory in a fixed order (deterministic, no atomics); the optional
// source gradient is scattered with red.global.add.f32, east/west taps of neighbouring lanes merged by shuffle first.
//
// Tuning record (B200, tools/microbench.py warp):
//  * round 1's kernels were bound by their own address arithmetic (ncu: 80 instructions per output element, ALU pipe
//    47 %, issue 73 %, DRAM 14 %): 64-bit element offsets rebuilt for every tap and channel.  32-bit pixel / tap offsets
//    and integer byte addresses that advance by one plane per channel leave ~19 instructions per element: forward
//    60 -> 53 us at 64x32x96x128 (60 % of the HBM roof), 21 -> 16 us at 16x32x96x128, image warp 26 -> 19 us;
//    flow-gradient-only backward 75 -> 58 us at 64x32x96x128, 30 -> 20 us at 16x, 47 -> 31 us for the 8x3x384x512 image.
//  * with that, the shared-memory "window" kernels of round 1 (source window staged with cp.async; source gradient
//    through a per-tile CSR sort and one red per window element) lose everywhere (flow gradient 75 vs 58 us, both
//    gradients 221 vs 199 us at 64x32x96x128) and were deleted.
//  * the source gradient is bound by the L2's reduction throughput (~460 G lane-adds/s; 4 per element and channel):
//    merging lane L's east taps into lane L+1's west taps halves the lanes but not the requests: 220 -> 199 us.
//    fp32 shared-memory atomics run at ~8 lanes/clk/SM (tools/peaks.cu), 4.7x the global rate in aggregate.
//  * round 2, measured and dropped: (a) software-pipelining the channel chunks (two chunks in flight): 64 vs 56 us at
//    16x32x96x128; (b) 64 registers x 4 CTAs/SM, or 8 channels in flight: 56-62 us - the kernel is not occupancy-bound;
//    (c) a source-gradient-only kernel whose threads own 4 vertically adjacent pixels and add coinciding south / north
//    and east / west taps in registers before the reduction (1.25 reductions per pixel and channel on a regular flow):
//    correct, but 73 vs 56 us on the benchmark's smooth flow (gradient ~0.35 px/px: only ~45 % of the row pairs and
//    ~65 % of the lane pairs line up, ~2.3 reductions per pixel and channel) - register pre-aggregation pays only on
//    flows far smoother than the ones a network under training produces.  tools/red_probe.cu: the L2 retires ~630 G
//    fp32 reductions per second, scalar or .v4, coalesced or not, so the remaining lever is a shared-memory window per
//    output tile (fp32 shared atomics: 2.2 T/s) flushed with one reduction per touched element (~1.3-1.6).
#include "common.cuh"

ARF_HOOK g_warp_variant = 0;   // test hook slot (arf_debug_set key 3), unused since the window kernels were removed

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

__device__ __forceinline__ float ldg_at(unsigned long long byte_addr) {
    return __ldg(reinterpret_cast<const float*>(byte_addr));
}

// fire-and-forget global float add at an integer byte address (a cast pointer would be treated as generic)
__device__ __forceinline__ void red_add_at(unsigned long long byte_addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(byte_addr), "f"(v) : "memory");
}

// Tap geometry with 32-bit offsets: nw tap at element o0 of a source plane, the east / south neighbours one step of
// de / ds elements further (0 when clamped onto the same column / row), weights zero for taps outside the source.
struct LeanTaps {
    int o0, de, ds;
    float w[4];
    float fxe, fxw, fys, fyn;
    unsigned in;       // bit k: tap k (nw, ne, sw, se) inside the source
};

__device__ __forceinline__ void make_lean_taps(float X, float Y, const WarpGeom& g, LeanTaps& t) {
    float xf = floorf(X), yf = floorf(Y);
    int xw = (int)xf, yn = (int)yf, xe = xw + 1, ys = yn + 1;
    t.fxe = (float)xe - X; t.fxw = X - (float)xw; t.fys = (float)ys - Y; t.fyn = Y - (float)yn;
    const bool inw = xw >= 0 && xw < g.Ws, ine = xe >= 0 && xe < g.Ws;
    const bool inn = yn >= 0 && yn < g.Hs, ins = ys >= 0 && ys < g.Hs;
    t.in = (inn && inw ? 1u : 0u) | (inn && ine ? 2u : 0u) | (ins && inw ? 4u : 0u) | (ins && ine ? 8u : 0u);
    const int xwc = min(max(xw, 0), g.Ws - 1), xec = min(max(xe, 0), g.Ws - 1);
    const int ync = min(max(yn, 0), g.Hs - 1), ysc = min(max(ys, 0), g.Hs - 1);
    t.o0 = ync * g.Ws + xwc;
    t.de = xec - xwc;
    t.ds = (ysc - ync) * g.Ws;
    t.w[0] = (t.in & 1u) ? t.fxe * t.fys : 0.f;
    t.w[1] = (t.in & 2u) ? t.fxw * t.fys : 0.f;
    t.w[2] = (t.in & 4u) ? t.fxe * t.fyn : 0.f;
    t.w[3] = (t.in & 8u) ? t.fxw * t.fyn : 0.f;
}

// grid = (pixel blocks, channel splits, batch); a thread computes its pixel's coordinates once and walks its channel
// range kCUF channels (4 * kCUF independent tap loads) at a time.  The first version of this kernel was bound by its
// own address arithmetic (ncu: 80 instructions per output element, ALU pipe 47 %, issue 73 %, DRAM 14 %): 64-bit
// element offsets rebuilt for every tap and channel.  Here the pixel index and the tap offsets are 32-bit, the four tap
// pointers are formed once per thread and advance by one source plane per channel: ~19 instructions per element.
// kCUF = channels whose 4 taps are in flight together: 8 for feature maps, 4 for images (3 channels).
template <int kCUF>
__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ field, float* __restrict__ y,
                WarpGeom g, int ch_per_split) {
    const unsigned hwo = (unsigned)g.Ho * (unsigned)g.Wo;
    const unsigned pix = blockIdx.x * 256u + threadIdx.x;
    if (pix >= hwo) return;
    const size_t hws = (size_t)g.Hs * g.Ws;
    const int b = blockIdx.z;
    const int c_lo = blockIdx.y * ch_per_split, c_hi = min(g.C, c_lo + ch_per_split);
    const unsigned iu = pix / (unsigned)g.Wo;
    const int i = (int)iu, j = (int)(pix - iu * (unsigned)g.Wo);
    float X, Y, dX, dY;
    pixel_coords(field, g, b, i, j, X, Y, dX, dY);
    const float* xb = x + ((size_t)b * g.C + c_lo) * hws;
    float* yb = y + ((size_t)b * g.C + c_lo) * hwo + pix;
    if (g.interp == ARF_INTERP_NEAREST) {
        int xn = (int)nearbyintf(X), yn = (int)nearbyintf(Y);
        bool in = xn >= 0 && xn < g.Ws && yn >= 0 && yn < g.Hs;
        for (int c = c_lo; c < c_hi; ++c, xb += hws, yb += hwo)
            *yb = in ? __ldg(xb + (size_t)yn * g.Ws + xn) : 0.f;
        return;
    }
    LeanTaps t;
    make_lean_taps(X, Y, g, t);
    // Byte addresses kept as integers: left as typed pointers, the compiler rebuilds base + 4 * index (LEA, LEA.HI.X)
    // for every load and advances 64-bit indices beside them - five integer instructions per tap.
    unsigned long long a0 = (unsigned long long)(xb + t.o0);      // nw
    unsigned long long a1 = a0 + 4ll * t.de;                      // ne
    unsigned long long a2 = a0 + 4ll * t.ds;                      // sw
    unsigned long long a3 = a2 + 4ll * t.de;                      // se
    unsigned long long ay = (unsigned long long)yb;
    const unsigned long long sstep = 4ull * hws, ostep = 4ull * hwo;
    int c = c_lo;
    for (; c + kCUF <= c_hi; c += kCUF) {
        float v[kCUF][4];
#pragma unroll
        for (int u = 0; u < kCUF; ++u) {
            v[u][0] = ldg_at(a0); v[u][1] = ldg_at(a1); v[u][2] = ldg_at(a2); v[u][3] = ldg_at(a3);
            a0 += sstep; a1 += sstep; a2 += sstep; a3 += sstep;
        }
#pragma unroll
        for (int u = 0; u < kCUF; ++u) {
            float o = v[u][0] * t.w[0];
            o = fmaf(v[u][1], t.w[1], o);
            o = fmaf(v[u][2], t.w[2], o);
            o = fmaf(v[u][3], t.w[3], o);
            __stcs((float*)ay, o);
            ay += ostep;
        }
    }
    for (; c < c_hi; ++c) {
        float o = ldg_at(a0) * t.w[0];
        o = fmaf(ldg_at(a1), t.w[1], o);
        o = fmaf(ldg_at(a2), t.w[2], o);
        o = fmaf(ldg_at(a3), t.w[3], o);
        __stcs((float*)ay, o);
        a0 += sstep; a1 += sstep; a2 += sstep; a3 += sstep;
        ay += ostep;
    }
}

// Backward, direct, lean: thread <-> (pixel, channel group).  A block is 256 / G pixels x G channel groups; G > 1 only
// when the image alone cannot fill the machine, and the G partial flow gradients of a pixel are then summed through
// shared memory in a fixed order (deterministic, no atomics on gfield).  Same address scheme as the forward.
// kGx: also scatter the source gradient (red.global.add).
template <int kCUB, bool kGx, int G, int kMinB = 1>
__global__ void __launch_bounds__(256, kMinB)
warp_bwd_lean(const float* __restrict__ x, const float* __restrict__ field, const float* __restrict__ gy,
              float* __restrict__ gx, float* __restrict__ gfield, WarpGeom g, int ch_per_split) {
    constexpr int P = 256 / G;
    __shared__ float red[G > 1 ? 2 * G * P : 1];
    const unsigned hwo = (unsigned)g.Ho * (unsigned)g.Wo;
    const int lp = threadIdx.x % P, grp = threadIdx.x / P;
    const unsigned pix_raw = blockIdx.x * (unsigned)P + lp;
    const bool live = pix_raw < hwo;
    const unsigned pix = live ? pix_raw : hwo - 1;          // dead threads shadow the last pixel, write nothing
    const size_t hws = (size_t)g.Hs * g.Ws;
    const int b = blockIdx.z;
    // dead threads keep the trip count of their warp when the source gradient shuffles (kGx), else skip the loop
    const int c_lo = min(g.C, grp * ch_per_split), c_hi = (live || kGx) ? min(g.C, c_lo + ch_per_split) : c_lo;
    const unsigned iu = pix / (unsigned)g.Wo;
    const int i = (int)iu, j = (int)(pix - iu * (unsigned)g.Wo);
    float X, Y, dX, dY;
    pixel_coords(field, g, b, i, j, X, Y, dX, dY);
    const float* gp = gy + ((size_t)b * g.C + c_lo) * hwo + pix;
    const size_t xoff = ((size_t)b * g.C + c_lo) * hws;
    if (g.interp == ARF_INTERP_NEAREST) {   // no field gradient; the source gradient is a plain scatter
        if (kGx) {
            int xn = (int)nearbyintf(X), yn = (int)nearbyintf(Y);
            if (xn >= 0 && xn < g.Ws && yn >= 0 && yn < g.Hs) {
                float* q = gx + xoff + (size_t)yn * g.Ws + xn;
                for (int c = c_lo; c < c_hi; ++c, q += hws, gp += hwo) atomicAdd(q, __ldg(gp));
            }
        }
        if (gfield && grp == 0 && live) {
            gfield[(size_t)b * 2 * hwo + pix] = 0.f;
            gfield[(size_t)b * 2 * hwo + hwo + pix] = 0.f;
        }
        return;      // interp is uniform over the grid: the whole block leaves
    }
    LeanTaps t;
    make_lean_taps(X, Y, g, t);
    // d out / dX = sum_k v_k cx_k, d out / dY = sum_k v_k cy_k (taps outside the source read as 0)
    const float m0 = (t.in & 1u) ? 1.f : 0.f, m1 = (t.in & 2u) ? 1.f : 0.f, m2 = (t.in & 4u) ? 1.f : 0.f, m3 = (t.in & 8u) ? 1.f : 0.f;
    const float cx0 = -t.fys * m0, cx1 = t.fys * m1, cx2 = -t.fyn * m2, cx3 = t.fyn * m3;
    const float cy0 = -t.fxe * m0, cy1 = -t.fxw * m1, cy2 = t.fxe * m2, cy3 = t.fxw * m3;
    // Source gradient: where lane L's east taps fall on lane L+1's west taps (the rule for a smooth flow) the two
    // contributions are added by shuffle first, so a touched element costs one red lane instead of two (the kernel is
    // bound by the L2's reduction rate, ~460 G lane-adds/s measured).  Decided per lane pair; other taps go out alone.
    bool give = false, take = false;
    if (kGx) {
        if (!live) t.in = 0;
        const int lane = threadIdx.x & 31;
        const int o0_next = __shfl_down_sync(0xffffffffu, t.o0, 1);
        const unsigned in_next = __shfl_down_sync(0xffffffffu, t.in, 1);
        const int ds_next = __shfl_down_sync(0xffffffffu, t.ds, 1);
        // both east taps of L inside, both west taps of L+1 inside, same two rows, adjacent columns
        give = lane < 31 && (t.in & 10u) == 10u && (in_next & 5u) == 5u && t.de == 1 && ds_next == t.ds && o0_next == t.o0 + 1;
        take = __shfl_up_sync(0xffffffffu, give, 1) && lane > 0;
    }
    unsigned long long a0 = (unsigned long long)(x + xoff + t.o0);     // byte addresses, see warp_fwd_kernel
    unsigned long long a1 = a0 + 4ll * t.de, a2 = a0 + 4ll * t.ds, a3 = a2 + 4ll * t.de;
    unsigned long long ag = (unsigned long long)gp;
    unsigned long long q0 = kGx ? (unsigned long long)(gx + xoff + t.o0) : 0ull;
    const long long qe = 4ll * t.de, qs = 4ll * t.ds;
    const unsigned long long sstep = 4ull * hws, ostep = 4ull * hwo;
    float ax = 0.f, ay = 0.f;
    int c = c_lo;
    for (; c + kCUB <= c_hi; c += kCUB) {
        float go[kCUB], v[kCUB][4];
#pragma unroll
        for (int u = 0; u < kCUB; ++u) {
            go[u] = ldg_at(ag);
            ag += ostep;
            if (gfield) {
                v[u][0] = ldg_at(a0); v[u][1] = ldg_at(a1); v[u][2] = ldg_at(a2); v[u][3] = ldg_at(a3);
                a0 += sstep; a1 += sstep; a2 += sstep; a3 += sstep;
            }
        }
#pragma unroll
        for (int u = 0; u < kCUB; ++u) {
            if (gfield) {
                ax = fmaf(go[u], fmaf(v[u][3], cx3, fmaf(v[u][2], cx2, fmaf(v[u][1], cx1, v[u][0] * cx0))), ax);
                ay = fmaf(go[u], fmaf(v[u][3], cy3, fmaf(v[u][2], cy2, fmaf(v[u][1], cy1, v[u][0] * cy0))), ay);
            }
            if (kGx) {
                const float e_n = go[u] * t.w[1], e_s = go[u] * t.w[3];
                const float pn = __shfl_up_sync(0xffffffffu, e_n, 1), ps = __shfl_up_sync(0xffffffffu, e_s, 1);
                if (t.in & 1u) red_add_at(q0, fmaf(go[u], t.w[0], take ? pn : 0.f));
                if (t.in & 4u) red_add_at(q0 + qs, fmaf(go[u], t.w[2], take ? ps : 0.f));
                if (!give) {
                    if (t.in & 2u) red_add_at(q0 + qe, e_n);
                    if (t.in & 8u) red_add_at(q0 + qs + qe, e_s);
                }
                q0 += sstep;
            }
        }
    }
    for (; c < c_hi; ++c) {
        const float go = ldg_at(ag);
        ag += ostep;
        if (gfield) {
            const float v0 = ldg_at(a0), v1 = ldg_at(a1), v2 = ldg_at(a2), v3 = ldg_at(a3);
            a0 += sstep; a1 += sstep; a2 += sstep; a3 += sstep;
            ax = fmaf(go, fmaf(v3, cx3, fmaf(v2, cx2, fmaf(v1, cx1, v0 * cx0))), ax);
            ay = fmaf(go, fmaf(v3, cy3, fmaf(v2, cy2, fmaf(v1, cy1, v0 * cy0))), ay);
        }
        if (kGx) {
            const float e_n = go * t.w[1], e_s = go * t.w[3];
            const float pn = __shfl_up_sync(0xffffffffu, e_n, 1), ps = __shfl_up_sync(0xffffffffu, e_s, 1);
            if (t.in & 1u) red_add_at(q0, fmaf(go, t.w[0], take ? pn : 0.f));
            if (t.in & 4u) red_add_at(q0 + qs, fmaf(go, t.w[2], take ? ps : 0.f));
            if (!give) {
                if (t.in & 2u) red_add_at(q0 + qe, e_n);
                if (t.in & 8u) red_add_at(q0 + qs + qe, e_s);
            }
            q0 += sstep;
        }
    }
    if (gfield) {
        if (G > 1) {
            red[grp * P + lp] = ax;
            red[(G + grp) * P + lp] = ay;
            __syncthreads();
            if (grp == 0) {
                ax = 0.f; ay = 0.f;
#pragma unroll
                for (int k = 0; k < G; ++k) { ax += red[k * P + lp]; ay += red[(G + k) * P + lp]; }
            }
        }
        if (grp == 0 && live) {
            float* gf = gfield + (size_t)b * 2 * hwo + pix;
            gf[0] = ax * dX;
            gf[hwo] = ay * dY;
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
    if ((long long)Ho * Wo >= 0x7fffff00LL) return ARF_EINVAL;
    const long long px = (long long)B * Ho * Wo;
    long long want = (2LL * ARF_NUM_SMS * 2048 + px - 1) / px;
    const int kCUF = C >= 8 ? 8 : (C == 3 ? 3 : 4);
    int groups = (C + kCUF - 1) / kCUF;
    int nsplit = (int)(want < 1 ? 1 : (want > groups ? groups : want));
    int ch_per_split = ((groups + nsplit - 1) / nsplit) * kCUF;
    nsplit = (C + ch_per_split - 1) / ch_per_split;
    dim3 grid(arf_cdiv((long long)Ho * Wo, 256), nsplit, B);
    if (kCUF == 8) warp_fwd_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(x, field, y, g, ch_per_split);
    else if (kCUF == 3) warp_fwd_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(x, field, y, g, ch_per_split);
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
    if ((long long)Ho * Wo >= 0x7fffff00LL) return ARF_EINVAL;
    // direct kernels: one thread per (pixel, channel group); channel groups only when the image alone leaves the
    // machine mostly empty
    const long long px = (long long)B * Ho * Wo;
    const int kCUB = gx ? 4 : (C == 3 ? 3 : (C >= 8 ? 8 : 4));
    const long long want = ((long long)ARF_NUM_SMS * 2048 + px - 1) / px;
    const int groups = (C + kCUB - 1) / kCUB;
    int G = 1;
    while (G < 8 && G < want && 2 * G <= groups) G *= 2;
    const int ch_per_split = ((groups + G - 1) / G) * kCUB;
    dim3 grid(arf_cdiv((long long)Ho * Wo, 256 / G), 1, B);
    // MINB = CTAs per SM the register allocation aims for.  The image warp (3 channels) is latency-bound - one round of
    // flow loads, one of tap loads per thread - and gains from 5 CTAs of 48 registers (8x3x384x512: 29.7 -> 26.8 us;
    // 6 or 8 CTAs spill: 30.8 / 40.9 us); the feature kernels keep their registers.
#define ARF_BWD_LEAN(CU, GX, MINB)                                                                                        \
    do {                                                                                                                  \
        if (G == 1) warp_bwd_lean<CU, GX, 1, MINB><<<grid, 256, 0, st>>>(x, field, gy, gx, gfield, g, ch_per_split);       \
        else if (G == 2) warp_bwd_lean<CU, GX, 2, MINB><<<grid, 256, 0, st>>>(x, field, gy, gx, gfield, g, ch_per_split);  \
        else if (G == 4) warp_bwd_lean<CU, GX, 4, MINB><<<grid, 256, 0, st>>>(x, field, gy, gx, gfield, g, ch_per_split);  \
        else warp_bwd_lean<CU, GX, 8, MINB><<<grid, 256, 0, st>>>(x, field, gy, gx, gfield, g, ch_per_split);              \
    } while (0)
    if (gx) ARF_BWD_LEAN(4, true, 3);
    else if (kCUB == 3) ARF_BWD_LEAN(3, false, 5);
    else if (kCUB == 8) ARF_BWD_LEAN(8, false, 3);      // 80 registers; left alone ptxas takes 89 (2 CTAs): 24.7 vs 20.5 us
    else ARF_BWD_LEAN(4, false, 4);
#undef ARF_BWD_LEAN
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// Bilinear resize: upsample / downsample of utils/uflow_utils.py:163-204 (align_corners=False; F.interpolate with
// scale_factor, so the source step is exactly 1/scale_factor, ATen area_pixel_compute_scale) and the flow
// up-sampling of the PWC-Lite family (models/pwclite.py:178-179, 203: align_corners=True, source step
// (in-1)/(out-1)); optional value scaling for flow fields.
#include "common.cuh"

namespace {

struct ResizeGeom {
    int N, Hi, Wi, Ho, Wo;
    float rh, rw, mul;
    int align;
};

__device__ __forceinline__ void src_index(int dst, float r, int in_size, int align, int& i0, int& i1, float& l0,
                                          float& l1) {
    // ATen area_pixel_compute_source_index (non-cubic): align_corners=True -> r*dst; False -> clamp below at 0
    float s;
    if (align) {
        s = r * (float)dst;
    } else {
        s = r * ((float)dst + 0.5f) - 0.5f;
        s = s < 0.f ? 0.f : s;
    }
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}

// grid = (pixel blocks of one plane, planes): 32-bit index math only (a 64-bit div/mod per pixel used to cost more
// than the four taps)
__global__ void __launch_bounds__(256) resize_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, ResizeGeom g) {
    const unsigned npix = (unsigned)g.Ho * (unsigned)g.Wo;
    for (unsigned n = blockIdx.y; n < (unsigned)g.N; n += gridDim.y) {
        const float* p = in + (size_t)n * g.Hi * g.Wi;
        float* q = out + (size_t)n * npix;
        for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < npix; idx += gridDim.x * blockDim.x) {
            const unsigned oy = idx / (unsigned)g.Wo, ox = idx - oy * (unsigned)g.Wo;
            int y0, y1, x0, x1;
            float hy0, hy1, wx0, wx1;
            src_index((int)oy, g.rh, g.Hi, g.align, y0, y1, hy0, hy1);
            src_index((int)ox, g.rw, g.Wi, g.align, x0, x1, wx0, wx1);
            const float* r0 = p + (size_t)y0 * g.Wi;
            const float* r1 = p + (size_t)y1 * g.Wi;
            float v = hy0 * (wx0 * __ldg(r0 + x0) + wx1 * __ldg(r0 + x1)) + hy1 * (wx0 * __ldg(r1 + x0) + wx1 * __ldg(r1 + x1));
            q[idx] = v * g.mul;
        }
    }
}

// Gather-form backward: each input pixel sums the weights with which the output pixels of its
// neighbourhood referenced it (deterministic, no atomics).
__global__ void __launch_bounds__(256) resize_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, ResizeGeom g) {
    const unsigned npix = (unsigned)g.Hi * (unsigned)g.Wi;
    // output rows whose taps can touch input row iy: source coordinate in (iy-1, iy+1)
    const float inv_rh = 1.f / g.rh, inv_rw = 1.f / g.rw;
    for (unsigned n = blockIdx.y; n < (unsigned)g.N; n += gridDim.y) {
        const float* go = gout + (size_t)n * g.Ho * g.Wo;
        float* gi = gin + (size_t)n * npix;
        for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < npix; idx += gridDim.x * blockDim.x) {
            const int iy = (int)(idx / (unsigned)g.Wi), ix = (int)(idx - (unsigned)iy * (unsigned)g.Wi);
            // exact range is [ceil(lo), ceil(hi) - 1]; floor/ceil leave one candidate of slack on each side against
            // rounding, candidates that do not reference the pixel get weight 0 below
            // (align_corners=True: source = r*dst, so dst in ((iy-1)/r, (iy+1)/r); same slack)
            const float sh = g.align ? 0.f : 0.5f;
            int oy_lo = max(0, (int)floorf(((float)iy - 1.f + sh) * inv_rh - sh) - (g.align ? 1 : 0));
            int oy_hi = min(g.Ho - 1, (int)ceilf(((float)iy + 1.f + sh) * inv_rh - sh) + (g.align ? 1 : 0));
            int ox_lo = max(0, (int)floorf(((float)ix - 1.f + sh) * inv_rw - sh) - (g.align ? 1 : 0));
            int ox_hi = min(g.Wo - 1, (int)ceilf(((float)ix + 1.f + sh) * inv_rw - sh) + (g.align ? 1 : 0));
            if (iy == 0) oy_lo = 0;            // rows clamped at the top edge all read row 0
            if (ix == 0) ox_lo = 0;
            if (iy == g.Hi - 1) oy_hi = g.Ho - 1;
            if (ix == g.Wi - 1) ox_hi = g.Wo - 1;
            float acc = 0.f;
            for (int oy = oy_lo; oy <= oy_hi; ++oy) {
                int y0, y1;
                float hy0, hy1;
                src_index(oy, g.rh, g.Hi, g.align, y0, y1, hy0, hy1);
                float wy = (y0 == iy ? hy0 : 0.f) + (y1 == iy ? hy1 : 0.f);
                if (wy == 0.f) continue;
                float row = 0.f;
                for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                    int x0, x1;
                    float wx0, wx1;
                    src_index(ox, g.rw, g.Wi, g.align, x0, x1, wx0, wx1);
                    float wx = (x0 == ix ? wx0 : 0.f) + (x1 == ix ? wx1 : 0.f);
                    if (wx != 0.f) row = fmaf(wx, __ldg(go + (size_t)oy * g.Wo + ox), row);
                }
                acc = fmaf(wy, row, acc);
            }
            gi[idx] = acc * g.mul;
        }
    }
}

// ------------------------------------------------------------------ exact x2 / x4 up-sampling (align_corners = False) ---
// upsample(flow, is_flow=True) of every pyramid level and the two x2 steps to full resolution (uflow_model.py:220,
// 343-344) are all the x2 geometry, the occlusion mask of the loss (uflow_loss.py:41) is up-sampled x4 (forward only).  Same arithmetic as the general kernels - the tap indices and weights still come
// from src_index, and the backward adds its taps in the same order, so results are bit-identical - but the work is
// laid out for the geometry: forward, a thread produces the two outputs above one source pixel from 2 x 3 loads and
// stores them as one float2 (no per-pixel div / mod); backward, a thread gathers its fixed 4 x 4 candidate window
// instead of deriving a candidate range with floor / ceil and walking it (the general kernel: ~400 instructions per
// input pixel, 59 us for 32 planes of 192 x 256 -> 384 x 512; the window form: 16 loads and FMAs).
template <int S>     // integer up-sampling factor: a thread writes the S outputs above source column k as one vector
__global__ void __launch_bounds__(256) resize_ups_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, ResizeGeom g) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (k >= g.Wi) return;
    int y0, y1, xa[S], xb[S];
    float hy0, hy1, wa[S], wb[S];
    src_index(oy, g.rh, g.Hi, 0, y0, y1, hy0, hy1);
#pragma unroll
    for (int i = 0; i < S; ++i) src_index(S * k + i, g.rw, g.Wi, 0, xa[i], xb[i], wa[i], wb[i]);
    const size_t ip = (size_t)g.Hi * g.Wi, op = (size_t)g.Ho * g.Wo;
    for (unsigned n = blockIdx.z; n < (unsigned)g.N; n += gridDim.z) {
        const float* r0 = in + n * ip + (size_t)y0 * g.Wi;
        const float* r1 = in + n * ip + (size_t)y1 * g.Wi;
        float v[S];
#pragma unroll
        for (int i = 0; i < S; ++i)
            v[i] = (hy0 * (wa[i] * __ldg(r0 + xa[i]) + wb[i] * __ldg(r0 + xb[i])) +
                    hy1 * (wa[i] * __ldg(r1 + xa[i]) + wb[i] * __ldg(r1 + xb[i]))) * g.mul;
        float* o = out + n * op + (size_t)oy * g.Wo + S * k;
        if (S == 2) *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
        else *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__global__ void __launch_bounds__(256) resize_up2_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, ResizeGeom g) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y;
    if (ix >= g.Wi) return;
    // candidate outputs 2i-1 .. 2i+2 per axis; weight with which candidate j referenced source index i (0 if it did not)
    float wy[4], wx[4];
    int oys[4], oxs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int a0, a1;
        float l0, l1;
        const int oy = 2 * iy - 1 + j, ox = 2 * ix - 1 + j;
        const bool vy = oy >= 0 && oy < g.Ho, vx = ox >= 0 && ox < g.Wo;
        oys[j] = vy ? oy : 0;
        oxs[j] = vx ? ox : 0;
        src_index(oys[j], g.rh, g.Hi, 0, a0, a1, l0, l1);
        wy[j] = vy ? (a0 == iy ? l0 : 0.f) + (a1 == iy ? l1 : 0.f) : 0.f;
        src_index(oxs[j], g.rw, g.Wi, 0, a0, a1, l0, l1);
        wx[j] = vx ? (a0 == ix ? l0 : 0.f) + (a1 == ix ? l1 : 0.f) : 0.f;
    }
    const size_t ip = (size_t)g.Hi * g.Wi, op = (size_t)g.Ho * g.Wo;
    for (unsigned n = blockIdx.z; n < (unsigned)g.N; n += gridDim.z) {
        const float* go = gout + n * op;
        float v[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) v[j][i] = __ldg(go + (size_t)oys[j] * g.Wo + oxs[i]);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (wy[j] == 0.f) continue;
            float row = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (wx[i] != 0.f) row = fmaf(wx[i], v[j][i], row);
            acc = fmaf(wy[j], row, acc);
        }
        gin[n * ip + (size_t)iy * g.Wi + ix] = acc * g.mul;
    }
}

static inline bool is_ups(const ResizeGeom& g, int S) {
    return !g.align && g.Ho == S * g.Hi && g.Wo == S * g.Wi && g.rh == 1.0f / S && g.rw == 1.0f / S && g.Ho <= 65535 &&
           g.Hi >= 2 && g.Wi >= 2;
}
static inline bool is_up2(const ResizeGeom& g) { return is_ups(g, 2); }
// rows on grid.y, planes strided over grid.z
static inline dim3 up2_grid(int cols, int rows, long long planes) {
    const long long bx = (cols + 255) / 256;
    long long bz = (16LL * ARF_NUM_SMS + bx * rows - 1) / (bx * rows);
    if (bz > planes) bz = planes;
    if (bz > 65535) bz = 65535;
    if (bz < 1) bz = 1;
    return dim3((unsigned)bx, (unsigned)rows, (unsigned)bz);
}

// grid for the two kernels: x covers one plane, y the planes (capped; both loops are strided)
static inline dim3 resize_grid(long long pix_per_plane, long long planes) {
    long long bx = (pix_per_plane + 255) / 256;
    if (bx > 4096) bx = 4096;
    long long by = planes < 65535 ? planes : 65535;
    const long long cap = 16LL * ARF_NUM_SMS;
    if (bx * by > cap) by = cap / bx > 0 ? cap / bx : 1;
    return dim3((unsigned)bx, (unsigned)by);
}

int make_geom(ResizeGeom& g, long long N, int Hi, int Wi, int Ho, int Wo, float rh, float rw, float mul, int align) {
    // align_corners=True with a single output row / column has source step 0 (ATen); the gather backward divides by
    // the step, so that degenerate case is refused
    if (N <= 0 || N > 0x7fffffffLL || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || !(rh > 0.f) || !(rw > 0.f))
        return ARF_EINVAL;
    g.align = align ? 1 : 0;
    if ((long long)Hi * Wi > 0x7fffffffLL || (long long)Ho * Wo > 0x7fffffffLL) return ARF_EINVAL;
    g.N = (int)N; g.Hi = Hi; g.Wi = Wi; g.Ho = Ho; g.Wo = Wo; g.rh = rh; g.rw = rw; g.mul = mul;
    return ARF_OK;
}

}  // namespace

extern "C" int arf_resize_bilinear_fwd(const float* in, float* out, long long planes, int Hi, int Wi, int Ho, int Wo,
                                       float rh, float rw, float mul, int align_corners, void* stream) {
    ARF_REQUIRE(in && out);
    ResizeGeom g;
    int rc = make_geom(g, planes, Hi, Wi, Ho, Wo, rh, rw, mul, align_corners);
    if (rc) return rc;
    if (is_up2(g) && ((uintptr_t)out % 8 == 0))
        resize_ups_fwd_kernel<2><<<up2_grid(Wi, Ho, planes), 256, 0, (cudaStream_t)stream>>>(in, out, g);
    else if (is_ups(g, 4) && ((uintptr_t)out % 16 == 0))
        resize_ups_fwd_kernel<4><<<up2_grid(Wi, Ho, planes), 256, 0, (cudaStream_t)stream>>>(in, out, g);
    else
        resize_fwd_kernel<<<resize_grid((long long)Ho * Wo, planes), 256, 0, (cudaStream_t)stream>>>(in, out, g);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_resize_bilinear_bwd(const float* gout, float* gin, long long planes, int Hi, int Wi, int Ho,
                                       int Wo, float rh, float rw, float mul, int align_corners, void* stream) {
    ARF_REQUIRE(gout && gin);
    ResizeGeom g;
    int rc = make_geom(g, planes, Hi, Wi, Ho, Wo, rh, rw, mul, align_corners);
    if (rc) return rc;
    if (is_up2(g))
        resize_up2_bwd_kernel<<<up2_grid(Wi, Hi, planes), 256, 0, (cudaStream_t)stream>>>(gout, gin, g);
    else
        resize_bwd_kernel<<<resize_grid((long long)Hi * Wi, planes), 256, 0, (cudaStream_t)stream>>>(gout, gin, g);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// Bilinear resize with align_corners=False: upsample / downsample of utils/uflow_utils.py:163-204
// (F.interpolate with scale_factor, so the source step is exactly 1/scale_factor, ATen
// area_pixel_compute_scale), optional value scaling for flow fields.
#include "common.cuh"

namespace {

struct ResizeGeom {
    int N, Hi, Wi, Ho, Wo;
    float rh, rw, mul;
};

__device__ __forceinline__ void src_index(int dst, float r, int in_size, int& i0, int& i1, float& l0, float& l1) {
    // ATen area_pixel_compute_source_index (align_corners=False, non-cubic): clamp below at 0
    float s = r * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}

__global__ void resize_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, ResizeGeom g) {
    long long total = (long long)g.N * g.Ho * g.Wo;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int ox = idx % g.Wo;
        long long t = idx / g.Wo;
        int oy = t % g.Ho;
        long long n = t / g.Ho;
        int y0, y1, x0, x1;
        float hy0, hy1, wx0, wx1;
        src_index(oy, g.rh, g.Hi, y0, y1, hy0, hy1);
        src_index(ox, g.rw, g.Wi, x0, x1, wx0, wx1);
        const float* p = in + n * (long long)g.Hi * g.Wi;
        float v = hy0 * (wx0 * __ldg(p + (size_t)y0 * g.Wi + x0) + wx1 * __ldg(p + (size_t)y0 * g.Wi + x1)) +
                  hy1 * (wx0 * __ldg(p + (size_t)y1 * g.Wi + x0) + wx1 * __ldg(p + (size_t)y1 * g.Wi + x1));
        out[idx] = v * g.mul;
    }
}

// Gather-form backward: each input pixel sums the weights with which the output pixels of its
// neighbourhood referenced it (deterministic, no atomics).
__global__ void resize_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, ResizeGeom g) {
    long long total = (long long)g.N * g.Hi * g.Wi;
    // output rows whose taps can touch input row iy: source coordinate in (iy-1, iy+1)
    const float inv_rh = 1.f / g.rh, inv_rw = 1.f / g.rw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int ix = idx % g.Wi;
        long long t = idx / g.Wi;
        int iy = t % g.Hi;
        long long n = t / g.Hi;
        // exact range is [ceil(lo), ceil(hi) - 1]; floor/ceil leave one candidate of slack on each side against
        // rounding, candidates that do not reference the pixel get weight 0 below
        int oy_lo = max(0, (int)floorf(((float)iy - 0.5f) * inv_rh - 0.5f));
        int oy_hi = min(g.Ho - 1, (int)ceilf(((float)iy + 1.5f) * inv_rh - 0.5f));
        int ox_lo = max(0, (int)floorf(((float)ix - 0.5f) * inv_rw - 0.5f));
        int ox_hi = min(g.Wo - 1, (int)ceilf(((float)ix + 1.5f) * inv_rw - 0.5f));
        if (iy == 0) oy_lo = 0;            // rows clamped at the top edge all read row 0
        if (ix == 0) ox_lo = 0;
        if (iy == g.Hi - 1) oy_hi = g.Ho - 1;
        if (ix == g.Wi - 1) ox_hi = g.Wo - 1;
        const float* go = gout + n * (long long)g.Ho * g.Wo;
        float acc = 0.f;
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
            int y0, y1;
            float hy0, hy1;
            src_index(oy, g.rh, g.Hi, y0, y1, hy0, hy1);
            float wy = (y0 == iy ? hy0 : 0.f) + (y1 == iy ? hy1 : 0.f);
            if (wy == 0.f) continue;
            float row = 0.f;
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                int x0, x1;
                float wx0, wx1;
                src_index(ox, g.rw, g.Wi, x0, x1, wx0, wx1);
                float wx = (x0 == ix ? wx0 : 0.f) + (x1 == ix ? wx1 : 0.f);
                if (wx != 0.f) row = fmaf(wx, __ldg(go + (size_t)oy * g.Wo + ox), row);
            }
            acc = fmaf(wy, row, acc);
        }
        gin[idx] = acc * g.mul;
    }
}

int make_geom(ResizeGeom& g, long long N, int Hi, int Wi, int Ho, int Wo, float rh, float rw, float mul) {
    if (N <= 0 || N > 0x7fffffffLL || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || !(rh > 0.f) || !(rw > 0.f))
        return ARF_EINVAL;
    g.N = (int)N; g.Hi = Hi; g.Wi = Wi; g.Ho = Ho; g.Wo = Wo; g.rh = rh; g.rw = rw; g.mul = mul;
    return ARF_OK;
}

}  // namespace

extern "C" int arf_resize_bilinear_fwd(const float* in, float* out, long long planes, int Hi, int Wi, int Ho, int Wo,
                                       float rh, float rw, float mul, void* stream) {
    ARF_REQUIRE(in && out);
    ResizeGeom g;
    int rc = make_geom(g, planes, Hi, Wi, Ho, Wo, rh, rw, mul);
    if (rc) return rc;
    long long total = planes * Ho * Wo;
    resize_fwd_kernel<<<arf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, g);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_resize_bilinear_bwd(const float* gout, float* gin, long long planes, int Hi, int Wi, int Ho,
                                       int Wo, float rh, float rw, float mul, void* stream) {
    ARF_REQUIRE(gout && gin);
    ResizeGeom g;
    int rc = make_geom(g, planes, Hi, Wi, Ho, Wo, rh, rw, mul);
    if (rc) return rc;
    long long total = planes * Hi * Wi;
    resize_bwd_kernel<<<arf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(gout, gin, g);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// NHWC bilinear resampler: utils/uflow_resampler.py:137-241, the TF `resampler` port.
// Taps are floor and CEIL of the coordinate (not floor+1), weights x - floor(x); with safe=True a tap outside
// the image contributes zero (safe_gather_nd, :104-134).  The reference materialises index lists on the host
// (`.tolist()`, :99); here one thread handles one (sample point, 4-channel group) with float4 gathers when C % 4 == 0.
#include "common.cuh"

namespace {

struct RsGeom { int B, H, W, C; long long P; };  // P sample points per batch item

__device__ __forceinline__ void taps(float wx, float wy, int W, int H, int& x0, int& x1, int& y0, int& y1, float& fx,
                                     float& fy, bool ok[4]) {
    float flx = floorf(wx), fly = floorf(wy);
    fx = wx - flx; fy = wy - fly;
    x0 = (int)flx; y0 = (int)fly;
    x1 = (int)ceilf(wx); y1 = (int)ceilf(wy);
    bool x0ok = x0 >= 0 && x0 < W, x1ok = x1 >= 0 && x1 < W, y0ok = y0 >= 0 && y0 < H, y1ok = y1 >= 0 && y1 < H;
    ok[0] = y0ok && x0ok; ok[1] = y0ok && x1ok; ok[2] = y1ok && x0ok; ok[3] = y1ok && x1ok;
}

__global__ void __launch_bounds__(256)
resampler_fwd_kernel(const float* __restrict__ data, const float* __restrict__ wxp, const float* __restrict__ wyp,
                     long long wstride, float* __restrict__ out, RsGeom g) {
    const long long total = (long long)g.B * g.P * g.C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int c = idx % g.C;
        long long pt = idx / g.C;
        int b = pt / g.P;
        float wx = __ldg(wxp + pt * wstride), wy = __ldg(wyp + pt * wstride);
        int x0, x1, y0, y1; float fx, fy; bool ok[4];
        taps(wx, wy, g.W, g.H, x0, x1, y0, y1, fx, fy, ok);
        const float* d = data + (size_t)b * g.H * g.W * g.C + c;
        float v00 = ok[0] ? __ldg(d + ((size_t)y0 * g.W + x0) * g.C) : 0.f;
        float v01 = ok[1] ? __ldg(d + ((size_t)y0 * g.W + x1) * g.C) : 0.f;
        float v10 = ok[2] ? __ldg(d + ((size_t)y1 * g.W + x0) * g.C) : 0.f;
        float v11 = ok[3] ? __ldg(d + ((size_t)y1 * g.W + x1) * g.C) : 0.f;
        out[idx] = (v00 * (1.f - fx) + v01 * fx) * (1.f - fy) + (v10 * (1.f - fx) + v11 * fx) * fy;
    }
}

// one warp per sample point: lanes stride the channels, d/dwarp reduced with shuffles, d/ddata with atomics
__global__ void __launch_bounds__(256)
resampler_bwd_kernel(const float* __restrict__ data, const float* __restrict__ wxp, const float* __restrict__ wyp,
                     long long wstride, const float* __restrict__ gout, float* __restrict__ gdata,
                     float* __restrict__ gwx, float* __restrict__ gwy, long long gwstride, RsGeom g) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long npts = (long long)g.B * g.P;
    for (long long pt = warp; pt < npts; pt += nwarps) {
        int b = pt / g.P;
        float wx = __ldg(wxp + pt * wstride), wy = __ldg(wyp + pt * wstride);
        int x0, x1, y0, y1; float fx, fy; bool ok[4];
        taps(wx, wy, g.W, g.H, x0, x1, y0, y1, fx, fy, ok);
        const size_t base = (size_t)b * g.H * g.W * g.C;
        const size_t o00 = base + ((size_t)y0 * g.W + x0) * g.C, o01 = base + ((size_t)y0 * g.W + x1) * g.C;
        const size_t o10 = base + ((size_t)y1 * g.W + x0) * g.C, o11 = base + ((size_t)y1 * g.W + x1) * g.C;
        float ax = 0.f, ay = 0.f;
        for (int c = lane; c < g.C; c += 32) {
            float go = __ldg(gout + pt * g.C + c);
            float v00 = ok[0] ? __ldg(data + o00 + c) : 0.f, v01 = ok[1] ? __ldg(data + o01 + c) : 0.f;
            float v10 = ok[2] ? __ldg(data + o10 + c) : 0.f, v11 = ok[3] ? __ldg(data + o11 + c) : 0.f;
            // out = (v00 (1-fx) + v01 fx)(1-fy) + (v10 (1-fx) + v11 fx) fy ;  d fx/d wx = 1 (floor is piecewise constant)
            ax += go * ((v01 - v00) * (1.f - fy) + (v11 - v10) * fy);
            ay += go * ((v10 * (1.f - fx) + v11 * fx) - (v00 * (1.f - fx) + v01 * fx));
            if (gdata) {
                if (ok[0]) atomicAdd(gdata + o00 + c, go * (1.f - fx) * (1.f - fy));
                if (ok[1]) atomicAdd(gdata + o01 + c, go * fx * (1.f - fy));
                if (ok[2]) atomicAdd(gdata + o10 + c, go * (1.f - fx) * fy);
                if (ok[3]) atomicAdd(gdata + o11 + c, go * fx * fy);
            }
        }
        ax = arf_warp_sum(ax);
        ay = arf_warp_sum(ay);
        if (lane == 0) {
            if (gwx) gwx[pt * gwstride] = ax;
            if (gwy) gwy[pt * gwstride] = ay;
        }
    }
}

// Few channels (images: C = 3): a warp per sample point leaves 29 lanes idle (600 us for 8 x 384 x 512 x 3); here a THREAD
// owns the point and walks its channels, the coordinate gradients stay in registers.
__global__ void __launch_bounds__(256)
resampler_bwd_thread_kernel(const float* __restrict__ data, const float* __restrict__ wxp, const float* __restrict__ wyp,
                            long long wstride, const float* __restrict__ gout, float* __restrict__ gdata,
                            float* __restrict__ gwx, float* __restrict__ gwy, long long gwstride, RsGeom g) {
    const long long npts = (long long)g.B * g.P;
    for (long long pt = blockIdx.x * (long long)blockDim.x + threadIdx.x; pt < npts; pt += (long long)gridDim.x * blockDim.x) {
        const int b = pt / g.P;
        const float wx = __ldg(wxp + pt * wstride), wy = __ldg(wyp + pt * wstride);
        int x0, x1, y0, y1; float fx, fy; bool ok[4];
        taps(wx, wy, g.W, g.H, x0, x1, y0, y1, fx, fy, ok);
        const size_t base = (size_t)b * g.H * g.W * g.C;
        const size_t o00 = base + ((size_t)y0 * g.W + x0) * g.C, o01 = base + ((size_t)y0 * g.W + x1) * g.C;
        const size_t o10 = base + ((size_t)y1 * g.W + x0) * g.C, o11 = base + ((size_t)y1 * g.W + x1) * g.C;
        float ax = 0.f, ay = 0.f;
        for (int c = 0; c < g.C; ++c) {
            const float go = __ldg(gout + pt * g.C + c);
            const float v00 = ok[0] ? __ldg(data + o00 + c) : 0.f, v01 = ok[1] ? __ldg(data + o01 + c) : 0.f;
            const float v10 = ok[2] ? __ldg(data + o10 + c) : 0.f, v11 = ok[3] ? __ldg(data + o11 + c) : 0.f;
            ax += go * ((v01 - v00) * (1.f - fy) + (v11 - v10) * fy);
            ay += go * ((v10 * (1.f - fx) + v11 * fx) - (v00 * (1.f - fx) + v01 * fx));
            if (gdata) {
                if (ok[0]) atomicAdd(gdata + o00 + c, go * (1.f - fx) * (1.f - fy));
                if (ok[1]) atomicAdd(gdata + o01 + c, go * fx * (1.f - fy));
                if (ok[2]) atomicAdd(gdata + o10 + c, go * (1.f - fx) * fy);
                if (ok[3]) atomicAdd(gdata + o11 + c, go * fx * fy);
            }
        }
        if (gwx) gwx[pt * gwstride] = ax;
        if (gwy) gwy[pt * gwstride] = ay;
    }
}

}  // namespace

/* data: (B,H,W,C) NHWC; warp_x, warp_y: B*P coordinates read at stride `wstride` elements (1 for separate
 * tensors, 2 for an interleaved (..,2) warp tensor); out: (B,P,C). */
extern "C" int arf_resampler_fwd(const float* data, const float* warp_x, const float* warp_y, long long wstride,
                                 float* out, int B, int H, int W, int C, long long P, void* stream) {
    ARF_REQUIRE(data && warp_x && warp_y && out && B > 0 && H > 0 && W > 0 && C > 0 && P > 0 && wstride > 0);
    RsGeom g{B, H, W, C, P};
    long long total = (long long)B * P * C;
    resampler_fwd_kernel<<<arf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(data, warp_x, warp_y, wstride, out, g);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_resampler_bwd(const float* data, const float* warp_x, const float* warp_y, long long wstride,
                                 const float* gout, float* gdata, float* gwx, float* gwy, long long gwstride, int B,
                                 int H, int W, int C, long long P, void* stream) {
    ARF_REQUIRE(data && warp_x && warp_y && gout && B > 0 && H > 0 && W > 0 && C > 0 && P > 0 && wstride > 0);
    if (!gdata && !gwx && !gwy) return ARF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (gdata) {
        cudaError_t e = cudaMemsetAsync(gdata, 0, (size_t)B * H * W * C * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    RsGeom g{B, H, W, C, P};
    if (C <= 8) {
        resampler_bwd_thread_kernel<<<arf_grid_1d((long long)B * P, 256), 256, 0, st>>>(
            data, warp_x, warp_y, wstride, gout, gdata, gwx, gwy, gwstride > 0 ? gwstride : 1, g);
    } else {
        long long threads = (long long)B * P * 32;
        resampler_bwd_kernel<<<arf_grid_1d(threads, 256), 256, 0, st>>>(data, warp_x, warp_y, wstride, gout, gdata, gwx,
                                                                       gwy, gwstride > 0 ? gwstride : 1, g);
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

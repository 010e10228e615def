// Edge-aware flow smoothness, first and second order.
//
// One parametrised kernel pair covers
//   UFlowLoss smoothness, order 1 / 2        losses/uflow_loss.py:58-102
//   smooth_grad_1st ("abs" | "uflow"), 2nd   losses/loss_blocks.py:93-124
// term_x(b,ch,y,x) = w_x(b,y,x) * pen(d_x),   x in [0, W-order)
//   d_x  = f[x+1]-f[x]                (order 1)      f[x+2]-2f[x+1]+f[x]   (order 2)
//   w_x  = exp(-edge * mean_c |I[x+woff+wstride] - I[x+woff]|)
//   pen  = sqrt(d^2 + eps2)  (robust_l1 on the squared gradient / penalty_uflow)   or   |d|
// loss = final * ( mean(term_x) + mean(term_y) ), each mean over its own (B,2,H,W-order) / (B,2,H-order,W) box.
// The image is data (detached in the reference, uflow_loss.py:65); only d(loss)/d(flow) exists.
#include "common.cuh"

namespace {

struct SmoothGeom {
    int B, Ci, H, W, order, wstride, woff, penalty;
    float edge, eps2, final_scale;
};

__device__ __forceinline__ float pen(float d, const SmoothGeom& g) {
    return g.penalty == 0 ? sqrtf(fmaf(d, d, g.eps2)) : fabsf(d);
}
__device__ __forceinline__ float dpen(float d, const SmoothGeom& g) {
    if (g.penalty == 0) return d * rsqrtf(fmaf(d, d, g.eps2));
    return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
}

// "Loads first": a pixel's whole neighbourhood (flow and image, clamped addresses instead of branches) is fetched before
// anything is computed, so a thread has 10-45 independent loads in flight instead of a chain of dependent ones behind the
// range tests (the first version: 44 / 55-75 us forward / backward at 64x3x80x256, latency-bound at 9 % of the HBM roof).
// ORDER, WSTRIDE, WOFF are compile-time so that the neighbourhood lives in registers; CI = 3 unrolls the channel loop
// (CI = 0: run-time channel count).
template <int ORDER, int WSTRIDE, int WOFF, int CI>
__global__ void __launch_bounds__(256)
smooth_fwd_kernel(const float* __restrict__ img, const float* __restrict__ flow, float* __restrict__ partials,
                  SmoothGeom g) {
    __shared__ float red[32];
    const size_t hw = (size_t)g.H * g.W;
    long long total = (long long)g.B * g.H * g.W;
    const int nci = CI ? CI : g.Ci;
    const float inv_ci = 1.f / (float)nci;
    float sx = 0.f, sy = 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int x, y, b;
        arf_split3(idx, g.W, g.H, x, y, b);
        // offsets of the pixels x .. x+ORDER of this row and y .. y+ORDER of this column (clamped: the terms that would
        // use a clamped pixel are masked below)
        int ox[ORDER + 1], oy[ORDER + 1];
#pragma unroll
        for (int j = 0; j <= ORDER; ++j) {
            ox[j] = y * g.W + min(x + j, g.W - 1);
            oy[j] = min(y + j, g.H - 1) * g.W + x;
        }
        const float* fb = flow + (size_t)b * 2 * hw;
        float fx[2][ORDER + 1], fy[2][ORDER + 1];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int j = 0; j <= ORDER; ++j) {
                fx[c][j] = __ldg(fb + c * hw + ox[j]);
                fy[c][j] = __ldg(fb + c * hw + oy[j]);
            }
        const float* ib = img + (size_t)b * nci * hw;
        float ex = 0.f, ey = 0.f;
#pragma unroll
        for (int c = 0; c < (CI ? CI : 1); ++c) {
            for (int cc = c; cc < nci; cc += (CI ? CI : 1)) {
                const float* p = ib + (size_t)cc * hw;
                ex += fabsf(__ldg(p + ox[WOFF + WSTRIDE]) - __ldg(p + ox[WOFF]));
                ey += fabsf(__ldg(p + oy[WOFF + WSTRIDE]) - __ldg(p + oy[WOFF]));
                if (CI) break;
            }
        }
        const float wx = expf(-g.edge * (ex * inv_ci)), wy = expf(-g.edge * (ey * inv_ci));
        float tx = 0.f, ty = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float dx = ORDER == 1 ? fx[c][1] - fx[c][0] : (fx[c][2] - fx[c][1]) - (fx[c][1] - fx[c][0]);
            const float dy = ORDER == 1 ? fy[c][1] - fy[c][0] : (fy[c][2] - fy[c][1]) - (fy[c][1] - fy[c][0]);
            tx += pen(dx, g);
            ty += pen(dy, g);
        }
        if (x < g.W - ORDER) sx += wx * tx;
        if (y < g.H - ORDER) sy += wy * ty;
    }
    float a = arf_block_sum(sx, red);
    float c = arf_block_sum(sy, red);
    if (threadIdx.x == 0) {
        partials[2 * (size_t)blockIdx.x] = a;
        partials[2 * (size_t)blockIdx.x + 1] = c;
    }
}

__global__ void smooth_finalize_kernel(const float* __restrict__ partials, int n, float* __restrict__ out,
                                       double inv_nx, double inv_ny, float final_scale) {
    __shared__ double s0[256], s1[256];
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a += (double)partials[2 * (size_t)i];
        c += (double)partials[2 * (size_t)i + 1];
    }
    s0[threadIdx.x] = a;
    s1[threadIdx.x] = c;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            s0[threadIdx.x] += s0[threadIdx.x + s];
            s1[threadIdx.x] += s1[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)((double)final_scale * (s0[0] * inv_nx + s1[0] * inv_ny));
}

template <int ORDER, int WSTRIDE, int WOFF, int CI>
__global__ void __launch_bounds__(256)
smooth_bwd_kernel(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ gloss,
                  float* __restrict__ gflow, SmoothGeom g, float inv_nx, float inv_ny) {
    constexpr int N = 2 * ORDER + 1;             // pixels x-ORDER .. x+ORDER of the row, y-ORDER .. y+ORDER of the column
    const size_t hw = (size_t)g.H * g.W;
    long long total = (long long)g.B * g.H * g.W;
    const int nci = CI ? CI : g.Ci;
    const float inv_ci = 1.f / (float)nci;
    const float gl = __ldg(gloss) * g.final_scale;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int x, y, b;
        arf_split3(idx, g.W, g.H, x, y, b);
        int ox[N], oy[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            ox[j] = y * g.W + min(max(x + j - ORDER, 0), g.W - 1);
            oy[j] = min(max(y + j - ORDER, 0), g.H - 1) * g.W + x;
        }
        const float* fb = flow + (size_t)b * 2 * hw;
        float fx[2][N], fy[2][N];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                fx[c][j] = __ldg(fb + c * hw + ox[j]);
                fy[c][j] = __ldg(fb + c * hw + oy[j]);
            }
        // edge sums of the ORDER+1 terms that contain this pixel: the term starting k pixels before it
        const float* ib = img + (size_t)b * nci * hw;
        float ex[ORDER + 1], ey[ORDER + 1];
#pragma unroll
        for (int k = 0; k <= ORDER; ++k) { ex[k] = 0.f; ey[k] = 0.f; }
#pragma unroll
        for (int c = 0; c < (CI ? CI : 1); ++c) {
            for (int cc = c; cc < nci; cc += (CI ? CI : 1)) {
                const float* p = ib + (size_t)cc * hw;
                float ix[N], iy[N];
#pragma unroll
                for (int j = 0; j < N; ++j) { ix[j] = __ldg(p + ox[j]); iy[j] = __ldg(p + oy[j]); }
#pragma unroll
                for (int k = 0; k <= ORDER; ++k) {
                    ex[k] += fabsf(ix[ORDER - k + WOFF + WSTRIDE] - ix[ORDER - k + WOFF]);
                    ey[k] += fabsf(iy[ORDER - k + WOFF + WSTRIDE] - iy[ORDER - k + WOFF]);
                }
                if (CI) break;
            }
        }
        float gx0 = 0.f, gx1 = 0.f, gy0 = 0.f, gy1 = 0.f;
        // term starting at xt = x-k has coefficient coef[k] on f[x]:  order 1: {-1,+1}   order 2: {+1,-2,+1}
#pragma unroll
        for (int k = 0; k <= ORDER; ++k) {
            const float coef = ORDER == 1 ? (k == 0 ? -1.f : 1.f) : (k == 1 ? -2.f : 1.f);
            const int s = ORDER - k;             // index of the term's first pixel in the neighbourhood arrays
            const int xt = x - k, yt = y - k;
            const bool vx = xt >= 0 && xt < g.W - ORDER, vy = yt >= 0 && yt < g.H - ORDER;
            const float wx = vx ? expf(-g.edge * (ex[k] * inv_ci)) * coef : 0.f;
            const float wy = vy ? expf(-g.edge * (ey[k] * inv_ci)) * coef : 0.f;
            const float dx0 = ORDER == 1 ? fx[0][s + 1] - fx[0][s] : (fx[0][s + 2] - fx[0][s + 1]) - (fx[0][s + 1] - fx[0][s]);
            const float dx1 = ORDER == 1 ? fx[1][s + 1] - fx[1][s] : (fx[1][s + 2] - fx[1][s + 1]) - (fx[1][s + 1] - fx[1][s]);
            const float dy0 = ORDER == 1 ? fy[0][s + 1] - fy[0][s] : (fy[0][s + 2] - fy[0][s + 1]) - (fy[0][s + 1] - fy[0][s]);
            const float dy1 = ORDER == 1 ? fy[1][s + 1] - fy[1][s] : (fy[1][s + 2] - fy[1][s + 1]) - (fy[1][s + 1] - fy[1][s]);
            gx0 += wx * dpen(dx0, g);
            gx1 += wx * dpen(dx1, g);
            gy0 += wy * dpen(dy0, g);
            gy1 += wy * dpen(dy1, g);
        }
        float* go = gflow + (size_t)b * 2 * hw + (size_t)y * g.W + x;
        go[0] = gl * (gx0 * inv_nx + gy0 * inv_ny);
        go[hw] = gl * (gx1 * inv_nx + gy1 * inv_ny);
    }
}

int make_geom(SmoothGeom& g, int B, int Ci, int H, int W, int order, int wstride, int woff, int penalty, float edge,
              float eps2, float final_scale) {
    if (B <= 0 || Ci <= 0 || H <= 0 || W <= 0) return ARF_EINVAL;
    if (order < 1 || order > 2 || wstride < 1 || woff < 0 || woff + wstride > order) return ARF_EINVAL;
    if (penalty < 0 || penalty > 1) return ARF_EINVAL;
    if (H <= order || W <= order) return ARF_EINVAL;
    g.B = B; g.Ci = Ci; g.H = H; g.W = W; g.order = order; g.wstride = wstride; g.woff = woff; g.penalty = penalty;
    g.edge = edge; g.eps2 = eps2; g.final_scale = final_scale;
    return ARF_OK;
}

// 16 CTAs per SM: the kernels are chains of dependent loads (edge weights, differences), so they want many warps in flight
// rather than many pixels per thread (4 per SM: 86 us backward at 64x3x80x256)
int smooth_grid(long long total) { return arf_grid_1d(total, 256, 16); }

}  // namespace

extern "C" int arf_smooth_num_partials(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return ARF_EINVAL;
    return smooth_grid((long long)B * H * W);
}

extern "C" int arf_smooth_fwd(const float* img, const float* flow, float* out, float* partials, int B, int Ci, int H,
                              int W, int order, int wstride, int woff, int penalty, float edge, float eps2,
                              float final_scale, void* stream) {
    ARF_REQUIRE(img && flow && out && partials);
    SmoothGeom g;
    int rc = make_geom(g, B, Ci, H, W, order, wstride, woff, penalty, edge, eps2, final_scale);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = smooth_grid((long long)B * H * W);
#define ARF_SMOOTH_FWD(O, S, F)                                                                      \
    do {                                                                                               \
        if (Ci == 3) smooth_fwd_kernel<O, S, F, 3><<<grid, 256, 0, st>>>(img, flow, partials, g);      \
        else smooth_fwd_kernel<O, S, F, 0><<<grid, 256, 0, st>>>(img, flow, partials, g);              \
    } while (0)
    if (order == 1) ARF_SMOOTH_FWD(1, 1, 0);
    else if (wstride == 2) ARF_SMOOTH_FWD(2, 2, 0);
    else if (woff == 0) ARF_SMOOTH_FWD(2, 1, 0);
    else ARF_SMOOTH_FWD(2, 1, 1);
#undef ARF_SMOOTH_FWD
    ARF_CHECK_LAUNCH();
    double nx = (double)B * 2 * H * (W - order), ny = (double)B * 2 * (H - order) * W;
    smooth_finalize_kernel<<<1, 256, 0, st>>>(partials, grid, out, 1.0 / nx, 1.0 / ny, final_scale);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_smooth_bwd(const float* img, const float* flow, const float* gloss, float* gflow, int B, int Ci,
                              int H, int W, int order, int wstride, int woff, int penalty, float edge, float eps2,
                              float final_scale, void* stream) {
    ARF_REQUIRE(img && flow && gloss && gflow);
    SmoothGeom g;
    int rc = make_geom(g, B, Ci, H, W, order, wstride, woff, penalty, edge, eps2, final_scale);
    if (rc) return rc;
    double nx = (double)B * 2 * H * (W - order), ny = (double)B * 2 * (H - order) * W;
    const int grid = smooth_grid((long long)B * H * W);
    cudaStream_t st = (cudaStream_t)stream;
    const float inx = (float)(1.0 / nx), iny = (float)(1.0 / ny);
#define ARF_SMOOTH_BWD(O, S, F)                                                                                    \
    do {                                                                                                             \
        if (Ci == 3) smooth_bwd_kernel<O, S, F, 3><<<grid, 256, 0, st>>>(img, flow, gloss, gflow, g, inx, iny);      \
        else smooth_bwd_kernel<O, S, F, 0><<<grid, 256, 0, st>>>(img, flow, gloss, gflow, g, inx, iny);              \
    } while (0)
    if (order == 1) ARF_SMOOTH_BWD(1, 1, 0);
    else if (wstride == 2) ARF_SMOOTH_BWD(2, 2, 0);
    else if (woff == 0) ARF_SMOOTH_BWD(2, 1, 0);
    else ARF_SMOOTH_BWD(2, 1, 1);
#undef ARF_SMOOTH_BWD
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

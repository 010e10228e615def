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

__device__ __forceinline__ float edge_weight(const float* __restrict__ img, const SmoothGeom& g, int b, size_t o_a,
                                             size_t o_b) {
    size_t hw = (size_t)g.H * g.W;
    const float* p = img + (size_t)b * g.Ci * hw;
    float s = 0.f;
    for (int c = 0; c < g.Ci; ++c) s += fabsf(__ldg(p + c * hw + o_b) - __ldg(p + c * hw + o_a));
    return expf(-g.edge * (s / (float)g.Ci));
}

__device__ __forceinline__ float pen(float d, const SmoothGeom& g) {
    return g.penalty == 0 ? sqrtf(fmaf(d, d, g.eps2)) : fabsf(d);
}
__device__ __forceinline__ float dpen(float d, const SmoothGeom& g) {
    if (g.penalty == 0) return d * rsqrtf(fmaf(d, d, g.eps2));
    return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
}

// difference of order `order` starting at element offset o with element stride s
__device__ __forceinline__ float fdiff(const float* __restrict__ f, size_t o, size_t s, int order) {
    if (order == 1) return __ldg(f + o + s) - __ldg(f + o);
    return (__ldg(f + o + 2 * s) - __ldg(f + o + s)) - (__ldg(f + o + s) - __ldg(f + o));
}

__global__ void __launch_bounds__(256)
smooth_fwd_kernel(const float* __restrict__ img, const float* __restrict__ flow, float* __restrict__ partials,
                  SmoothGeom g) {
    __shared__ float red[32];
    const size_t hw = (size_t)g.H * g.W;
    long long total = (long long)g.B * g.H * g.W;
    float sx = 0.f, sy = 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int x, y, b;
        arf_split3(idx, g.W, g.H, x, y, b);
        size_t o = (size_t)y * g.W + x;
        const float* fb = flow + (size_t)b * 2 * hw;
        if (x < g.W - g.order) {
            float w = edge_weight(img, g, b, o + g.woff, o + g.woff + g.wstride);
            sx += w * (pen(fdiff(fb, o, 1, g.order), g) + pen(fdiff(fb + hw, o, 1, g.order), g));
        }
        if (y < g.H - g.order) {
            float w = edge_weight(img, g, b, o + (size_t)g.woff * g.W, o + (size_t)(g.woff + g.wstride) * g.W);
            sy += w * (pen(fdiff(fb, o, g.W, g.order), g) + pen(fdiff(fb + hw, o, g.W, g.order), g));
        }
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

__global__ void __launch_bounds__(256)
smooth_bwd_kernel(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ gloss,
                  float* __restrict__ gflow, SmoothGeom g, float inv_nx, float inv_ny) {
    const size_t hw = (size_t)g.H * g.W;
    long long total = (long long)g.B * g.H * g.W;
    const float gl = __ldg(gloss) * g.final_scale;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int x, y, b;
        arf_split3(idx, g.W, g.H, x, y, b);
        const float* fb = flow + (size_t)b * 2 * hw;
        float gx0 = 0.f, gx1 = 0.f, gy0 = 0.f, gy1 = 0.f;
        // term starting at xt = x-k has coefficient coef[k] on f[x]:  order 1: {-1,+1}   order 2: {+1,-2,+1}
        for (int k = 0; k <= g.order; ++k) {
            float coef = g.order == 1 ? (k == 0 ? -1.f : 1.f) : (k == 1 ? -2.f : 1.f);
            int xt = x - k;
            if (xt >= 0 && xt < g.W - g.order) {
                size_t o = (size_t)y * g.W + xt;
                float w = edge_weight(img, g, b, o + g.woff, o + g.woff + g.wstride) * coef;
                gx0 += w * dpen(fdiff(fb, o, 1, g.order), g);
                gx1 += w * dpen(fdiff(fb + hw, o, 1, g.order), g);
            }
            int yt = y - k;
            if (yt >= 0 && yt < g.H - g.order) {
                size_t o = (size_t)yt * g.W + x;
                float w = edge_weight(img, g, b, o + (size_t)g.woff * g.W, o + (size_t)(g.woff + g.wstride) * g.W) * coef;
                gy0 += w * dpen(fdiff(fb, o, g.W, g.order), g);
                gy1 += w * dpen(fdiff(fb + hw, o, g.W, g.order), g);
            }
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

int smooth_grid(long long total) { return arf_grid_1d(total, 256, 4); }

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
    smooth_fwd_kernel<<<grid, 256, 0, st>>>(img, flow, partials, g);
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
    smooth_bwd_kernel<<<smooth_grid((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(
        img, flow, gloss, gflow, g, (float)(1.0 / nx), (float)(1.0 / ny));
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

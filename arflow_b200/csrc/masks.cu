// Validity / occlusion masks and the forward-splat range map.
#include "common.cuh"

namespace {

__device__ __forceinline__ void field_xy(const float* __restrict__ field, int kind, int b, int i, int j, int H,
                                         int W, float& x, float& y) {
    size_t hw = (size_t)H * W;
    const float* p = field + (size_t)b * 2 * hw + (size_t)i * W + j;
    x = __ldg(p);
    y = __ldg(p + hw);
    if (kind == ARF_FIELD_FLOW) {
        x = __fadd_rn((float)j, x);
        y = __fadd_rn((float)i, y);
    }
}

// mode 0: mask_invalid  (uflow_utils.py:35-50)   1[0 <= x <= W-1 and 0 <= y <= H-1]
// mode 1: border_mask   (warp_utils.py:119-134)  1[0 <  x <  W-1 and 0 <  y <  H-1]
__global__ void inside_mask_kernel(const float* __restrict__ field, float* __restrict__ mask, int B, int H, int W,
                                   int kind, int mode) {
    long long total = (long long)B * H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int j, i, b;
        arf_split3(idx, W, H, j, i, b);
        float x, y;
        field_xy(field, kind, b, i, j, H, W, x, y);
        float mw = (float)(W - 1), mh = (float)(H - 1);
        bool ok = mode == 0 ? (x >= 0.f && x <= mw && y >= 0.f && y <= mh) : (x > 0.f && x < mw && y > 0.f && y < mh);
        mask[idx] = ok ? 1.f : 0.f;
    }
}

// four pixels of a row per thread (W % 4 == 0, 16-byte aligned planes): the scalar kernel keeps two 4-byte loads per thread
// in flight and is latency-bound (102 us for 64 x 320 x 1024 = 2.5 TB/s)
__global__ void __launch_bounds__(256)
inside_mask_vec4_kernel(const float* __restrict__ field, float* __restrict__ mask, int B, int H, int W, int kind, int mode) {
    const unsigned wq = (unsigned)W / 4u;
    const long long total = (long long)B * H * wq;
    const size_t hw = (size_t)H * W;
    const float mw = (float)(W - 1), mh = (float)(H - 1);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int q, i, b;
        arf_split3(idx, (int)wq, H, q, i, b);
        const size_t o = (size_t)b * 2 * hw + (size_t)i * W + 4 * (size_t)q;
        const float4 fx = __ldg(reinterpret_cast<const float4*>(field + o));
        const float4 fy = __ldg(reinterpret_cast<const float4*>(field + o + hw));
        float xs[4] = {fx.x, fx.y, fx.z, fx.w}, ys[4] = {fy.x, fy.y, fy.z, fy.w}, m[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float x = xs[t], y = ys[t];
            if (kind == ARF_FIELD_FLOW) {
                x = __fadd_rn((float)(4 * q + t), x);
                y = __fadd_rn((float)i, y);
            }
            const bool ok = mode == 0 ? (x >= 0.f && x <= mw && y >= 0.f && y <= mh) : (x > 0.f && x < mw && y > 0.f && y < mh);
            m[t] = ok ? 1.f : 0.f;
        }
        *reinterpret_cast<float4*>(mask + (size_t)b * hw + (size_t)i * W + 4 * (size_t)q) = make_float4(m[0], m[1], m[2], m[3]);
    }
}

// compute_range_map (uflow_utils.py:80-160 == warp_utils.py:158-239) and get_corresponding_map
// (warp_utils.py:26-80): every pixel splats the bilinear weights of its target onto the 4 integer
// neighbours that lie inside the image.  count must be zero on entry.
__global__ void range_map_kernel(const float* __restrict__ field, float* __restrict__ count, int B, int H, int W,
                                 int kind) {
    long long total = (long long)B * H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int j, i, b;
        arf_split3(idx, W, H, j, i, b);
        float x, y;
        field_xy(field, kind, b, i, j, H, W, x, y);
        float xf = floorf(x), yf = floorf(y);
        float ox = x - xf, oy = y - yf;
        // guard the float->int conversion for far-away targets (they hit no pixel anyway)
        if (!(xf >= -2.f && xf <= (float)W && yf >= -2.f && yf <= (float)H)) continue;
        int x0 = (int)xf, y0 = (int)yf;
        float* cb = count + (size_t)b * H * W;
#pragma unroll
        for (int di = 0; di < 2; ++di)
#pragma unroll
            for (int dj = 0; dj < 2; ++dj) {
                int yy = y0 + di, xx = x0 + dj;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    float wi = di ? oy : 1.f - oy;
                    float wj = dj ? ox : 1.f - ox;
                    atomicAdd(cb + (size_t)yy * W + xx, wi * wj);
                }
            }
    }
}

// d(count)/d(field): the splat weights are (1-ox | ox) * (1-oy | oy) with ox = x - floor(x), so each pixel
// gathers +-(other-axis weight) * upstream gradient from the in-image taps it wrote to.
__global__ void range_map_bwd_kernel(const float* __restrict__ field, const float* __restrict__ gcount,
                                     float* __restrict__ gfield, int B, int H, int W, int kind) {
    long long total = (long long)B * H * W;
    const size_t hw = (size_t)H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int j, i, b;
        arf_split3(idx, W, H, j, i, b);
        float x, y;
        field_xy(field, kind, b, i, j, H, W, x, y);
        float xf = floorf(x), yf = floorf(y);
        float ox = x - xf, oy = y - yf;
        float gx = 0.f, gy = 0.f;
        if (xf >= -2.f && xf <= (float)W && yf >= -2.f && yf <= (float)H) {
            int x0 = (int)xf, y0 = (int)yf;
            const float* gb = gcount + (size_t)b * hw;
#pragma unroll
            for (int di = 0; di < 2; ++di)
#pragma unroll
                for (int dj = 0; dj < 2; ++dj) {
                    int yy = y0 + di, xx = x0 + dj;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                        float g = __ldg(gb + (size_t)yy * W + xx);
                        float wi = di ? oy : 1.f - oy, wj = dj ? ox : 1.f - ox;
                        gx += g * wi * (dj ? 1.f : -1.f);
                        gy += g * wj * (di ? 1.f : -1.f);
                    }
                }
        }
        float* go = gfield + (size_t)b * 2 * hw + (size_t)i * W + j;
        go[0] = gx;
        go[hw] = gy;
    }
}

// mode 0: clamp(c,0,1)   mode 1: clamp(c,0,1) < th   mode 2: 1 - clamp(c,0,1)
__global__ void count_to_mask_kernel(const float* __restrict__ count, float* __restrict__ out, long long n, int mode,
                                     float th) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        float c = fminf(fmaxf(__ldg(count + idx), 0.f), 1.f);
        out[idx] = mode == 0 ? c : (mode == 1 ? (c < th ? 1.f : 0.f) : 1.f - c);
    }
}

// get_occu_mask_bidirection tail (warp_utils.py:93-100): |f12 + f21w|^2 > scale*(|f12|^2+|f21w|^2) + bias
__global__ void occ_bidir_kernel(const float* __restrict__ f12, const float* __restrict__ f21w, float* __restrict__ out,
                                 int B, long long hw, float scale, float bias) {
    long long total = (long long)B * hw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long b = idx / hw, p = idx - b * hw;
        const float* a = f12 + b * 2 * hw + p;
        const float* w = f21w + b * 2 * hw + p;
        float ax = __ldg(a), ay = __ldg(a + hw), wx = __ldg(w), wy = __ldg(w + hw);
        float dx = ax + wx, dy = ay + wy;
        float mag = (ax * ax + ay * ay) + (wx * wx + wy * wy);
        float diff = dx * dx + dy * dy;
        out[idx] = diff > scale * mag + bias ? 1.f : 0.f;
    }
}

}  // namespace

extern "C" int arf_inside_mask(const float* field, float* mask, int B, int H, int W, int field_kind, int strict,
                               void* stream) {
    ARF_REQUIRE(field && mask && B > 0 && H > 0 && W > 0);
    long long total = (long long)B * H * W;
    if (W % 4 == 0 && (uintptr_t)field % 16 == 0 && (uintptr_t)mask % 16 == 0) {
        inside_mask_vec4_kernel<<<arf_grid_1d(total / 4, 256, 16), 256, 0, (cudaStream_t)stream>>>(field, mask, B, H, W,
                                                                                                   field_kind, strict);
        ARF_CHECK_LAUNCH();
        return ARF_OK;
    }
    inside_mask_kernel<<<arf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(field, mask, B, H, W, field_kind,
                                                                                   strict ? 1 : 0);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_range_map(const float* field, float* count, int B, int H, int W, int field_kind, void* stream) {
    ARF_REQUIRE(field && count && B > 0 && H > 0 && W > 0);
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)B * H * W;
    cudaError_t e = cudaMemsetAsync(count, 0, (size_t)total * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    range_map_kernel<<<arf_grid_1d(total, 256), 256, 0, st>>>(field, count, B, H, W, field_kind);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_range_map_bwd(const float* field, const float* gcount, float* gfield, int B, int H, int W,
                                 int field_kind, void* stream) {
    ARF_REQUIRE(field && gcount && gfield && B > 0 && H > 0 && W > 0);
    long long total = (long long)B * H * W;
    range_map_bwd_kernel<<<arf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(field, gcount, gfield, B, H, W,
                                                                                     field_kind);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_count_to_mask(const float* count, float* out, long long n, int mode, float th, void* stream) {
    ARF_REQUIRE(count && out && n > 0 && mode >= 0 && mode <= 2);
    count_to_mask_kernel<<<arf_grid_1d(n, 256), 256, 0, (cudaStream_t)stream>>>(count, out, n, mode, th);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_occ_bidir(const float* flow12, const float* flow21_warped, float* out, int B, int H, int W,
                             float scale, float bias, void* stream) {
    ARF_REQUIRE(flow12 && flow21_warped && out && B > 0 && H > 0 && W > 0);
    long long hw = (long long)H * W;
    occ_bidir_kernel<<<arf_grid_1d(B * hw, 256), 256, 0, (cudaStream_t)stream>>>(flow12, flow21_warped, out, B, hw,
                                                                                  scale, bias);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

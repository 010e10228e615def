// Fused bias + leaky-ReLU epilogue of the PWC decoder / pyramid convolutions, forward and backward.
//
// The convolutions themselves stay on cuDNN (out of scope).  What the reference spends around every
// one of them is: an elementwise bias add, an elementwise leaky ReLU (models/uflow_model.py:134-135, 427-436),
// and in the backward pass an elementwise leaky-ReLU gradient followed by a separate full-tensor reduction for
// the bias gradient.  Here the forward is ONE in-place pass over the convolution output and the backward is ONE
// pass that writes the pre-activation gradient and reduces the bias gradient on the fly (two-stage, fixed
// summation order).  Pure streaming: 8 B/element forward, 12 B/element backward.
//   y   = leaky(conv + bias)                      (slope > 0, so sign(y) = sign(conv + bias))
//   g   = gy * (y > 0 ? 1 : slope)                (ATen leaky_relu_backward: x > 0 ? g : g * slope)
//   db  = sum over batch and pixels of g
#include "common.cuh"

namespace {

constexpr int kEThreads = 256;
constexpr int kEChunk = 4096;      // elements of one (b, c) plane handled by one CTA

__global__ void __launch_bounds__(kEThreads)
bias_leaky_fwd_kernel(float* __restrict__ y, const float* __restrict__ bias, int C, long long HW, float slope, int vec) {
    const long long plane = blockIdx.x;
    const float bv = bias ? __ldg(bias + (int)(plane % C)) : 0.f;
    float* p = y + plane * HW;
    const long long lo = (long long)blockIdx.y * kEChunk;
    const long long hi = lo + kEChunk < HW ? lo + kEChunk : HW;
    if (vec) {
        for (long long e = lo + 4 * threadIdx.x; e < hi; e += 4 * kEThreads) {
            float4 v = *reinterpret_cast<float4*>(p + e);
            v.x += bv; v.y += bv; v.z += bv; v.w += bv;
            v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
            v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
            *reinterpret_cast<float4*>(p + e) = v;
        }
    } else {
        for (long long e = lo + threadIdx.x; e < hi; e += kEThreads) {
            float v = p[e] + bv;
            p[e] = v > 0.f ? v : v * slope;
        }
    }
}

__global__ void __launch_bounds__(kEThreads)
bias_leaky_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, float* __restrict__ g,
                      float* __restrict__ partials, long long HW, float slope, int vec) {
    __shared__ float red[32];
    const long long plane = blockIdx.x;
    const float* pg = gy + plane * HW;
    const float* py = y + plane * HW;
    float* po = g + plane * HW;
    const long long lo = (long long)blockIdx.y * kEChunk;
    const long long hi = lo + kEChunk < HW ? lo + kEChunk : HW;
    float acc = 0.f;
    if (vec) {
        for (long long e = lo + 4 * threadIdx.x; e < hi; e += 4 * kEThreads) {
            const float4 a = *reinterpret_cast<const float4*>(pg + e);
            const float4 b = *reinterpret_cast<const float4*>(py + e);
            float4 o;
            o.x = b.x > 0.f ? a.x : a.x * slope; o.y = b.y > 0.f ? a.y : a.y * slope;
            o.z = b.z > 0.f ? a.z : a.z * slope; o.w = b.w > 0.f ? a.w : a.w * slope;
            *reinterpret_cast<float4*>(po + e) = o;
            acc += (o.x + o.y) + (o.z + o.w);
        }
    } else {
        for (long long e = lo + threadIdx.x; e < hi; e += kEThreads) {
            const float a = pg[e];
            const float o = py[e] > 0.f ? a : a * slope;
            po[e] = o;
            acc += o;
        }
    }
    if (partials) {
        const float s = arf_block_sum(acc, red);
        if (threadIdx.x == 0) partials[plane * gridDim.y + blockIdx.y] = s;
    }
}

// dbias[c] = sum over b and chunks of partials[(b*C + c)*nchunks + chunk]; one warp per channel, fixed order
__global__ void bias_grad_finalize_kernel(const float* __restrict__ partials, float* __restrict__ dbias, long long B, int C,
                                          int nchunks) {
    const int c = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    const long long n = B * nchunks;
    double acc = 0.0;
    for (long long i = lane; i < n; i += 32) {
        const long long b = i / nchunks, k = i - b * nchunks;
        acc += (double)partials[(b * C + c) * nchunks + k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dbias[c] = (float)acc;
}

inline int nchunks_of(long long HW) { return (int)((HW + kEChunk - 1) / kEChunk); }

// ---- NHWC (channels-last) variants: y is a (rows = N*H*W) x C row-major matrix --------------------------------
constexpr int kERows = 256;       // rows of the matrix one CTA of the backward handles by default

// Rows per CTA of the backward: 256 when that still gives every SM four CTAs, else halved down to 32.  The coarse
// pyramid levels (49152, 12288, 3072 rows) ran on 192, 48 and 12 CTAs with the fixed 256 and were latency-bound
// (22, 13 and 9 us per launch in the step's launch list).
static int nhwc_rows_per_cta(long long rows) {
    int r = kERows;
    while (r > 32 && rows / r < 4LL * ARF_NUM_SMS) r >>= 1;
    // very tall matrices (786432 rows x 32 channels at pyramid level 0): fewer, longer CTAs as long as eight per SM
    // remain - the single-CTA-per-32-channels finalize took 16 us for 3072 partial rows
    while (r < 1024 && rows / (2 * r) >= 8LL * ARF_NUM_SMS) r <<= 1;
    return r;
}

__global__ void __launch_bounds__(kEThreads)
bias_leaky_nhwc_fwd_kernel(const float* src, float* dst, long long dst_ld, const float* __restrict__ bias, long long total,
                           int C, float slope, int vec) {
    // src: packed rows x C (the convolution output); dst: the same matrix (in place) or a column slice of a wider
    // row-major matrix with row stride dst_ld (the next dense-block input: no separate concat copy of this part)
    if (vec) {
        for (long long e = 4 * (blockIdx.x * (long long)kEThreads + threadIdx.x); e < total; e += 4LL * gridDim.x * kEThreads) {
            float4 v = *reinterpret_cast<const float4*>(src + e);
            const long long r = e / C;
            const int c = (int)(e - r * C);
            if (bias) {
                const float4 b = *reinterpret_cast<const float4*>(bias + c);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
            v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
            *reinterpret_cast<float4*>(dst + r * dst_ld + c) = v;
        }
    } else {
        for (long long e = blockIdx.x * (long long)kEThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kEThreads) {
            const long long r = e / C;
            const int c = (int)(e - r * C);
            float v = src[e] + (bias ? __ldg(bias + c) : 0.f);
            dst[r * dst_ld + c] = v > 0.f ? v : v * slope;
        }
    }
}

// One CTA: rpc rows x all C columns.  Thread t owns column group (t % G) (4 columns when vec, else 1) and walks
// rows t / G, t / G + 256 / G, ...; the per-thread column sums are reduced through shared memory, one partial row
// of C sums per CTA.
__global__ void __launch_bounds__(kEThreads)
bias_leaky_nhwc_bwd_kernel(const float* __restrict__ gy, long long gy_ld, const float* __restrict__ y, long long y_ld,
                           float* __restrict__ g, float* __restrict__ partials, long long rows, int C, float slope, int vec,
                           int rpc) {
    extern __shared__ float sacc[];          // kEThreads * 4 floats
    const long long r0 = (long long)blockIdx.x * rpc;
    const long long r1 = r0 + rpc < rows ? r0 + rpc : rows;
    const int w = vec ? 4 : 1;
    const int G = C / w;                     // column groups per row
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // a thread keeps ONE column group (t % G) for all its rows, so its sums stay in registers
    const bool fixed = G <= kEThreads;
    if (fixed) {
        const int cg = threadIdx.x % G, rstep = kEThreads / G;      // threads >= rstep * G stay idle (G = 24: 240 of 256)
        for (long long r = r0 + threadIdx.x / G; r < r1 && threadIdx.x < rstep * G; r += rstep) {
            const long long e = r * C + (long long)cg * w;
            const long long eg = r * gy_ld + (long long)cg * w;     // gy and y may be column slices of wider matrices
            const long long ey = r * y_ld + (long long)cg * w;
            if (vec) {
                const float4 a = *reinterpret_cast<const float4*>(gy + eg);
                const float4 b = *reinterpret_cast<const float4*>(y + ey);
                float4 o;
                o.x = b.x > 0.f ? a.x : a.x * slope; o.y = b.y > 0.f ? a.y : a.y * slope;
                o.z = b.z > 0.f ? a.z : a.z * slope; o.w = b.w > 0.f ? a.w : a.w * slope;
                *reinterpret_cast<float4*>(g + e) = o;
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            } else {
                const float a = gy[eg];
                const float o = y[ey] > 0.f ? a : a * slope;
                g[e] = o;
                acc.x += o;
            }
        }
        if (partials) {
            sacc[threadIdx.x * 4 + 0] = acc.x; sacc[threadIdx.x * 4 + 1] = acc.y;
            sacc[threadIdx.x * 4 + 2] = acc.z; sacc[threadIdx.x * 4 + 3] = acc.w;
            __syncthreads();
            for (int c = threadIdx.x; c < C; c += kEThreads) {
                const int cg2 = c / w, k = c - cg2 * w;
                float s = 0.f;
                for (int t = cg2; t < rstep * G; t += G) s += sacc[t * 4 + k];
                partials[(long long)blockIdx.x * C + c] = s;
            }
        }
    } else {
        // generic widths (not used by the PWC networks): elementwise pass, then each thread sums whole columns
        for (long long e = r0 * C + threadIdx.x; e < r1 * C; e += kEThreads) {
            const long long r = e / C;
            const float a = gy[r * gy_ld + (e - r * C)];
            g[e] = y[r * y_ld + (e - r * C)] > 0.f ? a : a * slope;
        }
        if (partials) {
            __syncthreads();
            for (int c = threadIdx.x; c < C; c += kEThreads) {
                float s = 0.f;
                for (long long r = r0; r < r1; ++r) s += g[r * C + c];
                partials[(long long)blockIdx.x * C + c] = s;
            }
        }
    }
}

// dbias[c] = sum over CTAs of partials[cta*C + c].  block = (32 channels, 32 stripes over the CTAs): coalesced
// 128-byte rows, 32-way parallel over the partial rows, fixed summation order.
__global__ void __launch_bounds__(1024)
bias_grad_nhwc_finalize_kernel(const float* __restrict__ partials, float* __restrict__ dbias, long long nblk, int C) {
    __shared__ double red[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double acc = 0.0;
    if (c < C)
        for (long long i = threadIdx.y; i < nblk; i += 32) acc += (double)partials[i * C + c];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) s += red[k][threadIdx.x];
        dbias[c] = (float)s;
    }
}

}  // namespace

extern "C" long long arf_bias_leaky_nhwc_num_partials(long long rows, int C) {
    if (rows <= 0 || C <= 0) return ARF_EINVAL;
    const int rpc = nhwc_rows_per_cta(rows);
    return ((rows + rpc - 1) / rpc) * C;
}

extern "C" int arf_bias_leaky_nhwc_fwd(float* y, const float* bias, long long rows, int C, float slope, void* stream) {
    return arf_bias_leaky_nhwc_fwd_ld(y, y, C, bias, rows, C, slope, stream);
}

extern "C" int arf_bias_leaky_nhwc_fwd_ld(const float* src, float* dst, long long dst_ld, const float* bias, long long rows,
                                          int C, float slope, void* stream) {
    ARF_REQUIRE(src && dst);
    ARF_REQUIRE(rows > 0 && C > 0 && dst_ld >= C);
    const int vec = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0) && (C % 4 == 0) && (dst_ld % 4 == 0) &&
                    (!bias || (uintptr_t)bias % 16 == 0);
    const long long total = rows * C;
    bias_leaky_nhwc_fwd_kernel<<<arf_grid_1d(vec ? total / 4 : total, kEThreads, 16), kEThreads, 0, (cudaStream_t)stream>>>(
        src, dst, dst_ld, bias, total, C, slope, vec);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_bias_leaky_nhwc_bwd(const float* gy, const float* y, float* g, float* partials, float* dbias,
                                       long long rows, int C, float slope, void* stream) {
    return arf_bias_leaky_nhwc_bwd_ld(gy, C, y, C, g, partials, dbias, rows, C, slope, stream);
}

extern "C" int arf_bias_leaky_nhwc_bwd_ld(const float* gy, long long gy_ld, const float* y, long long y_ld, float* g,
                                          float* partials, float* dbias, long long rows, int C, float slope, void* stream) {
    ARF_REQUIRE(gy && y && g);
    ARF_REQUIRE(rows > 0 && C > 0 && gy_ld >= C && y_ld >= C);
    if (dbias) ARF_REQUIRE(partials != nullptr);
    const int rpc = nhwc_rows_per_cta(rows);
    const long long nblk = (rows + rpc - 1) / rpc;
    if (nblk > 0x7fffffffLL) return ARF_EINVAL;
    const int vec = ((uintptr_t)gy % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)g % 16 == 0) && (C % 4 == 0) &&
                    (gy_ld % 4 == 0) && (y_ld % 4 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    bias_leaky_nhwc_bwd_kernel<<<(unsigned)nblk, kEThreads, kEThreads * 4 * sizeof(float), st>>>(
        gy, gy_ld, y, y_ld, g, dbias ? partials : nullptr, rows, C, slope, vec, rpc);
    ARF_CHECK_LAUNCH();
    if (dbias) {
        bias_grad_nhwc_finalize_kernel<<<arf_cdiv(C, 32), dim3(32, 32), 0, st>>>(partials, dbias, nblk, C);
        ARF_CHECK_LAUNCH();
    }
    return ARF_OK;
}

namespace {
}  // namespace

extern "C" long long arf_bias_leaky_num_partials(long long B, int C, long long HW) {
    if (B <= 0 || C <= 0 || HW <= 0) return ARF_EINVAL;
    return B * C * nchunks_of(HW);
}

extern "C" int arf_bias_leaky_fwd(float* y, const float* bias, long long B, int C, long long HW, float slope, void* stream) {
    ARF_REQUIRE(y);
    ARF_REQUIRE(B > 0 && C > 0 && HW > 0 && B * C <= 0x7fffffffLL && nchunks_of(HW) <= 65535);
    const int vec = ((uintptr_t)y % 16 == 0) && (HW % 4 == 0);
    dim3 grid((unsigned)(B * C), (unsigned)nchunks_of(HW));
    bias_leaky_fwd_kernel<<<grid, kEThreads, 0, (cudaStream_t)stream>>>(y, bias, C, HW, slope, vec);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_bias_leaky_bwd(const float* gy, const float* y, float* g, float* partials, float* dbias, long long B,
                                  int C, long long HW, float slope, void* stream) {
    ARF_REQUIRE(gy && y && g);
    ARF_REQUIRE(B > 0 && C > 0 && HW > 0 && B * C <= 0x7fffffffLL && nchunks_of(HW) <= 65535);
    if (dbias) ARF_REQUIRE(partials != nullptr);
    const int vec = ((uintptr_t)gy % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)g % 16 == 0) && (HW % 4 == 0);
    dim3 grid((unsigned)(B * C), (unsigned)nchunks_of(HW));
    cudaStream_t st = (cudaStream_t)stream;
    bias_leaky_bwd_kernel<<<grid, kEThreads, 0, st>>>(gy, y, g, dbias ? partials : nullptr, HW, slope, vec);
    ARF_CHECK_LAUNCH();
    if (dbias) {
        bias_grad_finalize_kernel<<<arf_cdiv(C, 8), 256, 0, st>>>(partials, dbias, B, C, nchunks_of(HW));
        ARF_CHECK_LAUNCH();
    }
    return ARF_OK;
}

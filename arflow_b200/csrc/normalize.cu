// normalize_features of the PWC networks (models/uflow_model.py:8-50) for the setting every
// model uses: normalize=True, center=True, moments_across_channels=True, moments_across_images=True, two feature
// maps.  Per sample b (n = C*H*W elements per map):
//   mu_k = mean(f_k),  v_k = sum (f_k - mu_k)^2 / (n - 1)         (torch.var is unbiased)
//   mu = (mu_1 + mu_2) / 2,  s = sqrt((v_1 + v_2) / 2 + 1e-16)
//   y_k = (f_k - mu) / s
// The reference spends two var_mean reductions and ~8 elementwise kernels forward and ~25 kernels backward on this;
// here: one reduction + one elementwise pass each way (two-stage double-precision sums, fixed order).
// Backward, with S1 = sum_k sum g_k and S2 = sum_k sum g_k (f_k - mu):
//   df_k = g_k / s  -  S1 / (2 n s)  -  S2 (f_k - mu_k) / (2 s^3 (n - 1))
// The maps are only required to be dense and of equal size (any layout: nothing here depends on the element order).
#include "common.cuh"

namespace {

constexpr int kNThreads = 256;
constexpr int kNChunk = 8192;      // elements of one sample handled by one CTA (16x32x96x128 fwd / bwd: 4096 -> 41 / 52 us, 8192 -> 37 / 51, 16384 -> 36 / 57)

__device__ __forceinline__ double block_sum_d(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kNThreads / 32; ++i) r += red[i];
    return r;
}

// partials[(b * nchunks + chunk) * 4 + {0,1,2,3}] = sum f1, sum f1^2, sum f2, sum f2^2 of the chunk
__global__ void __launch_bounds__(kNThreads)
featnorm_moments_kernel(const float* __restrict__ f1, const float* __restrict__ f2, double* __restrict__ partials,
                        long long n, int vec) {
    __shared__ double red[kNThreads / 32];
    const long long b = blockIdx.y;
    const long long lo = (long long)blockIdx.x * kNChunk, hi = lo + kNChunk < n ? lo + kNChunk : n;
    const float* p1 = f1 + b * n;
    const float* p2 = f2 + b * n;
    double s1 = 0, q1 = 0, s2 = 0, q2 = 0;
    if (vec) {
        for (long long e = lo + 4 * threadIdx.x; e < hi; e += 4 * kNThreads) {
            const float4 a = *reinterpret_cast<const float4*>(p1 + e);
            const float4 c = *reinterpret_cast<const float4*>(p2 + e);
            s1 += (double)a.x + (double)a.y + (double)a.z + (double)a.w;
            q1 += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
            s2 += (double)c.x + (double)c.y + (double)c.z + (double)c.w;
            q2 += (double)c.x * c.x + (double)c.y * c.y + (double)c.z * c.z + (double)c.w * c.w;
        }
    } else {
        for (long long e = lo + threadIdx.x; e < hi; e += kNThreads) {
            const double a = p1[e], c = p2[e];
            s1 += a; q1 += a * a; s2 += c; q2 += c * c;
        }
    }
    double* out = partials + (b * gridDim.x + blockIdx.x) * 4;
    double r;
    r = block_sum_d(s1, red); if (threadIdx.x == 0) out[0] = r;
    r = block_sum_d(q1, red); if (threadIdx.x == 0) out[1] = r;
    r = block_sum_d(s2, red); if (threadIdx.x == 0) out[2] = r;
    r = block_sum_d(q2, red); if (threadIdx.x == 0) out[3] = r;
}

// stats[b*4 + {0,1,2,3}] = mu, s, mu_1, mu_2
__global__ void featnorm_stats_kernel(const double* __restrict__ partials, float* __restrict__ stats, int nchunks, long long n) {
    const long long b = blockIdx.x;
    if (threadIdx.x != 0) return;
    double s1 = 0, q1 = 0, s2 = 0, q2 = 0;
    for (int i = 0; i < nchunks; ++i) {
        const double* p = partials + (b * nchunks + i) * 4;
        s1 += p[0]; q1 += p[1]; s2 += p[2]; q2 += p[3];
    }
    const double m1 = s1 / n, m2 = s2 / n;
    const double dof = n > 1 ? (double)(n - 1) : 1.0;          // torch.var: unbiased
    double v1 = (q1 - n * m1 * m1) / dof, v2 = (q2 - n * m2 * m2) / dof;
    v1 = v1 < 0 ? 0 : v1; v2 = v2 < 0 ? 0 : v2;
    // the reference forms the moments in fp32: mean_all = (m1+m2)/2, var_all = (v1+v2)/2, std = sqrt(var_all + 1e-16)
    const float mu = __fdiv_rn(__fadd_rn((float)m1, (float)m2), 2.f);
    const float var = __fdiv_rn(__fadd_rn((float)v1, (float)v2), 2.f);
    stats[b * 4 + 0] = mu;
    stats[b * 4 + 1] = __fsqrt_rn(__fadd_rn(var, 1e-16f));
    stats[b * 4 + 2] = (float)m1;
    stats[b * 4 + 3] = (float)m2;
}

__global__ void __launch_bounds__(kNThreads)
featnorm_apply_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ stats,
                      float* __restrict__ y1, float* __restrict__ y2, long long n, int vec) {
    const long long b = blockIdx.y;
    const float mu = __ldg(stats + b * 4), s = __ldg(stats + b * 4 + 1);
    const long long lo = (long long)blockIdx.x * kNChunk, hi = lo + kNChunk < n ? lo + kNChunk : n;
    const long long base = b * n;
    if (vec) {
        for (long long e = lo + 4 * threadIdx.x; e < hi; e += 4 * kNThreads) {
            float4 a = *reinterpret_cast<const float4*>(f1 + base + e);
            float4 c = *reinterpret_cast<const float4*>(f2 + base + e);
            a.x = __fdiv_rn(a.x - mu, s); a.y = __fdiv_rn(a.y - mu, s); a.z = __fdiv_rn(a.z - mu, s); a.w = __fdiv_rn(a.w - mu, s);
            c.x = __fdiv_rn(c.x - mu, s); c.y = __fdiv_rn(c.y - mu, s); c.z = __fdiv_rn(c.z - mu, s); c.w = __fdiv_rn(c.w - mu, s);
            *reinterpret_cast<float4*>(y1 + base + e) = a;
            *reinterpret_cast<float4*>(y2 + base + e) = c;
        }
    } else {
        for (long long e = lo + threadIdx.x; e < hi; e += kNThreads) {
            y1[base + e] = __fdiv_rn(f1[base + e] - mu, s);
            y2[base + e] = __fdiv_rn(f2[base + e] - mu, s);
        }
    }
}

// partials[(b * nchunks + chunk) * 2 + {0,1}] = sum_k sum g_k, sum_k sum g_k (f_k - mu) of the chunk
__global__ void __launch_bounds__(kNThreads)
featnorm_bwd_reduce_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ g1,
                           const float* __restrict__ g2, const float* __restrict__ stats, double* __restrict__ partials,
                           long long n, int vec) {
    __shared__ double red[kNThreads / 32];
    const long long b = blockIdx.y;
    const float mu = __ldg(stats + b * 4);
    const long long lo = (long long)blockIdx.x * kNChunk, hi = lo + kNChunk < n ? lo + kNChunk : n;
    const long long base = b * n;
    double S1 = 0, S2 = 0;
    if (vec) {
        for (long long e = lo + 4 * threadIdx.x; e < hi; e += 4 * kNThreads) {
            const float4 a = *reinterpret_cast<const float4*>(f1 + base + e), ga = *reinterpret_cast<const float4*>(g1 + base + e);
            const float4 c = *reinterpret_cast<const float4*>(f2 + base + e), gc = *reinterpret_cast<const float4*>(g2 + base + e);
            S1 += (double)ga.x + (double)ga.y + (double)ga.z + (double)ga.w + (double)gc.x + (double)gc.y + (double)gc.z + (double)gc.w;
            S2 += (double)ga.x * (a.x - mu) + (double)ga.y * (a.y - mu) + (double)ga.z * (a.z - mu) + (double)ga.w * (a.w - mu) +
                  (double)gc.x * (c.x - mu) + (double)gc.y * (c.y - mu) + (double)gc.z * (c.z - mu) + (double)gc.w * (c.w - mu);
        }
    } else {
        for (long long e = lo + threadIdx.x; e < hi; e += kNThreads) {
            const double ga = g1[base + e], gc = g2[base + e];
            S1 += ga + gc;
            S2 += ga * (f1[base + e] - mu) + gc * (f2[base + e] - mu);
        }
    }
    double* out = partials + (b * gridDim.x + blockIdx.x) * 2;
    double r;
    r = block_sum_d(S1, red); if (threadIdx.x == 0) out[0] = r;
    r = block_sum_d(S2, red); if (threadIdx.x == 0) out[1] = r;
}

// coef[b*2 + {0,1}] = A = -S1 / (2 n s),  Bv = -S2 / (2 s^3 (n-1))
__global__ void featnorm_bwd_coef_kernel(const double* __restrict__ partials, const float* __restrict__ stats,
                                         float* __restrict__ coef, int nchunks, long long n) {
    const long long b = blockIdx.x;
    if (threadIdx.x != 0) return;
    double S1 = 0, S2 = 0;
    for (int i = 0; i < nchunks; ++i) {
        S1 += partials[(b * nchunks + i) * 2];
        S2 += partials[(b * nchunks + i) * 2 + 1];
    }
    const double s = stats[b * 4 + 1];
    const double dof = n > 1 ? (double)(n - 1) : 1.0;
    coef[b * 2 + 0] = (float)(-S1 / (2.0 * n * s));
    coef[b * 2 + 1] = (float)(-S2 / (2.0 * s * s * s * dof));
}

__global__ void __launch_bounds__(kNThreads)
featnorm_bwd_apply_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ g1,
                          const float* __restrict__ g2, const float* __restrict__ stats, const float* __restrict__ coef,
                          float* __restrict__ d1, float* __restrict__ d2, long long n, int vec) {
    const long long b = blockIdx.y;
    const float s = __ldg(stats + b * 4 + 1), m1 = __ldg(stats + b * 4 + 2), m2 = __ldg(stats + b * 4 + 3);
    const float A = __ldg(coef + b * 2), Bv = __ldg(coef + b * 2 + 1);
    const float inv_s = 1.f / s;
    const long long lo = (long long)blockIdx.x * kNChunk, hi = lo + kNChunk < n ? lo + kNChunk : n;
    const long long base = b * n;
    if (vec) {
        for (long long e = lo + 4 * threadIdx.x; e < hi; e += 4 * kNThreads) {
            const float4 a = *reinterpret_cast<const float4*>(f1 + base + e), ga = *reinterpret_cast<const float4*>(g1 + base + e);
            const float4 c = *reinterpret_cast<const float4*>(f2 + base + e), gc = *reinterpret_cast<const float4*>(g2 + base + e);
            float4 o1, o2;
            o1.x = fmaf(ga.x, inv_s, fmaf(Bv, a.x - m1, A)); o1.y = fmaf(ga.y, inv_s, fmaf(Bv, a.y - m1, A));
            o1.z = fmaf(ga.z, inv_s, fmaf(Bv, a.z - m1, A)); o1.w = fmaf(ga.w, inv_s, fmaf(Bv, a.w - m1, A));
            o2.x = fmaf(gc.x, inv_s, fmaf(Bv, c.x - m2, A)); o2.y = fmaf(gc.y, inv_s, fmaf(Bv, c.y - m2, A));
            o2.z = fmaf(gc.z, inv_s, fmaf(Bv, c.z - m2, A)); o2.w = fmaf(gc.w, inv_s, fmaf(Bv, c.w - m2, A));
            if (d1) *reinterpret_cast<float4*>(d1 + base + e) = o1;
            if (d2) *reinterpret_cast<float4*>(d2 + base + e) = o2;
        }
    } else {
        for (long long e = lo + threadIdx.x; e < hi; e += kNThreads) {
            if (d1) d1[base + e] = fmaf(g1[base + e], inv_s, fmaf(Bv, f1[base + e] - m1, A));
            if (d2) d2[base + e] = fmaf(g2[base + e], inv_s, fmaf(Bv, f2[base + e] - m2, A));
        }
    }
}

// ------------------------------------------------------------------ small maps: one launch ---------------------
// The coarse pyramid levels (n = C*H*W <= 8 k elements per map and sample; at 24 k one CTA per sample already loses: 41 vs 25 us backward) are pure launch latency for the three-kernel
// chain above.  Here ONE CTA of 1024 threads owns a sample: pass 1 forms the sums, thread 0 the statistics (same
// formulas), pass 2 re-reads the maps (a few hundred KB: L1 / L2 hits) and applies them.
constexpr int kSmallThreads = 1024;
constexpr long long kSmallMax = 8192;

// sum over the CTA of up to four doubles per thread; result broadcast to every thread through `out`
template <int K>
__device__ __forceinline__ void cta_sum_d(double (&v)[K], double (*red)[32], double* out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) red[k][w] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double r = 0.0;
        for (int i = 0; i < kSmallThreads / 32; ++i) r += red[threadIdx.x][i];      // fixed order
        out[threadIdx.x] = r;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kSmallThreads)
featnorm_fwd_small_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ y1,
                          float* __restrict__ y2, float* __restrict__ stats, long long n) {
    __shared__ double red[4][32];
    __shared__ double tot[4];
    __shared__ float ms[2];
    const long long base = (long long)blockIdx.x * n;
    double v[4] = {0, 0, 0, 0};
    for (long long e = threadIdx.x; e < n; e += kSmallThreads) {
        const double a = f1[base + e], c = f2[base + e];
        v[0] += a; v[1] += a * a; v[2] += c; v[3] += c * c;
    }
    cta_sum_d<4>(v, red, tot);
    if (threadIdx.x == 0) {
        const double m1 = tot[0] / n, m2 = tot[2] / n;
        const double dof = n > 1 ? (double)(n - 1) : 1.0;
        double v1 = (tot[1] - n * m1 * m1) / dof, v2 = (tot[3] - n * m2 * m2) / dof;
        v1 = v1 < 0 ? 0 : v1; v2 = v2 < 0 ? 0 : v2;
        const float mu = __fdiv_rn(__fadd_rn((float)m1, (float)m2), 2.f);
        const float var = __fdiv_rn(__fadd_rn((float)v1, (float)v2), 2.f);
        const float sd = __fsqrt_rn(__fadd_rn(var, 1e-16f));
        stats[blockIdx.x * 4 + 0] = mu;
        stats[blockIdx.x * 4 + 1] = sd;
        stats[blockIdx.x * 4 + 2] = (float)m1;
        stats[blockIdx.x * 4 + 3] = (float)m2;
        ms[0] = mu; ms[1] = sd;
    }
    __syncthreads();
    const float mu = ms[0], sd = ms[1];
    for (long long e = threadIdx.x; e < n; e += kSmallThreads) {
        y1[base + e] = __fdiv_rn(f1[base + e] - mu, sd);
        y2[base + e] = __fdiv_rn(f2[base + e] - mu, sd);
    }
}

__global__ void __launch_bounds__(kSmallThreads)
featnorm_bwd_small_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ g1,
                          const float* __restrict__ g2, const float* __restrict__ stats, float* __restrict__ d1,
                          float* __restrict__ d2, float* __restrict__ coef, long long n) {
    __shared__ double red[2][32];
    __shared__ double tot[2];
    __shared__ float ab[2];
    const long long base = (long long)blockIdx.x * n;
    const float mu = __ldg(stats + blockIdx.x * 4), sd = __ldg(stats + blockIdx.x * 4 + 1);
    const float m1 = __ldg(stats + blockIdx.x * 4 + 2), m2 = __ldg(stats + blockIdx.x * 4 + 3);
    double v[2] = {0, 0};
    for (long long e = threadIdx.x; e < n; e += kSmallThreads) {
        const double ga = g1[base + e], gc = g2[base + e];
        v[0] += ga + gc;
        v[1] += ga * (f1[base + e] - mu) + gc * (f2[base + e] - mu);
    }
    cta_sum_d<2>(v, red, tot);
    if (threadIdx.x == 0) {
        const double s = sd;
        const double dof = n > 1 ? (double)(n - 1) : 1.0;
        ab[0] = (float)(-tot[0] / (2.0 * n * s));
        ab[1] = (float)(-tot[1] / (2.0 * s * s * s * dof));
        coef[blockIdx.x * 2 + 0] = ab[0];
        coef[blockIdx.x * 2 + 1] = ab[1];
    }
    __syncthreads();
    const float A = ab[0], Bv = ab[1], inv_s = 1.f / sd;
    for (long long e = threadIdx.x; e < n; e += kSmallThreads) {
        if (d1) d1[base + e] = fmaf(g1[base + e], inv_s, fmaf(Bv, f1[base + e] - m1, A));
        if (d2) d2[base + e] = fmaf(g2[base + e], inv_s, fmaf(Bv, f2[base + e] - m2, A));
    }
}

inline int nchunks_n(long long n) { return (int)((n + kNChunk - 1) / kNChunk); }
inline bool al16(const void* p) { return ((uintptr_t)p % 16) == 0; }

}  // namespace

// workspace (doubles): 4 per (sample, chunk)
extern "C" long long arf_featnorm_workspace(long long B, long long n) {
    if (B <= 0 || n <= 0) return ARF_EINVAL;
    return B * nchunks_n(n) * 4 * (long long)sizeof(double);
}

extern "C" int arf_featnorm_fwd(const float* f1, const float* f2, float* y1, float* y2, float* stats, void* ws,
                                long long B, long long n, void* stream) {
    ARF_REQUIRE(f1 && f2 && y1 && y2 && stats && ws);
    ARF_REQUIRE(B > 0 && n > 0 && B <= 65535 && nchunks_n(n) <= 0x7fffffff);
    const int nc = nchunks_n(n);
    const int vec = (n % 4 == 0) && al16(f1) && al16(f2) && al16(y1) && al16(y2);
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= kSmallMax) {
        featnorm_fwd_small_kernel<<<(unsigned)B, kSmallThreads, 0, st>>>(f1, f2, y1, y2, stats, n);
        ARF_CHECK_LAUNCH();
        return ARF_OK;
    }
    dim3 grid((unsigned)nc, (unsigned)B);
    featnorm_moments_kernel<<<grid, kNThreads, 0, st>>>(f1, f2, (double*)ws, n, vec);
    ARF_CHECK_LAUNCH();
    featnorm_stats_kernel<<<(unsigned)B, 32, 0, st>>>((const double*)ws, stats, nc, n);
    ARF_CHECK_LAUNCH();
    featnorm_apply_kernel<<<grid, kNThreads, 0, st>>>(f1, f2, stats, y1, y2, n, vec);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_featnorm_bwd(const float* f1, const float* f2, const float* g1, const float* g2, const float* stats,
                                float* d1, float* d2, float* coef, void* ws, long long B, long long n, void* stream) {
    ARF_REQUIRE(f1 && f2 && g1 && g2 && stats && coef && ws);
    ARF_REQUIRE(B > 0 && n > 0 && B <= 65535);
    if (!d1 && !d2) return ARF_OK;
    const int nc = nchunks_n(n);
    const int vec = (n % 4 == 0) && al16(f1) && al16(f2) && al16(g1) && al16(g2) && (!d1 || al16(d1)) && (!d2 || al16(d2));
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= kSmallMax) {
        featnorm_bwd_small_kernel<<<(unsigned)B, kSmallThreads, 0, st>>>(f1, f2, g1, g2, stats, d1, d2, coef, n);
        ARF_CHECK_LAUNCH();
        return ARF_OK;
    }
    dim3 grid((unsigned)nc, (unsigned)B);
    featnorm_bwd_reduce_kernel<<<grid, kNThreads, 0, st>>>(f1, f2, g1, g2, stats, (double*)ws, n, vec);
    ARF_CHECK_LAUNCH();
    featnorm_bwd_coef_kernel<<<(unsigned)B, 32, 0, st>>>((const double*)ws, stats, coef, nc, n);
    ARF_CHECK_LAUNCH();
    featnorm_bwd_apply_kernel<<<grid, kNThreads, 0, st>>>(f1, f2, g1, g2, stats, coef, d1, d2, n, vec);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// Cost-volume correlation, forward and backward.
//
// Semantics follow the reference CUDA package (correlation_cuda_kernel.cu:41-300) for every
// (pad, ks, md, s1, s2); the setting all models use (pad=md=4, ks=1, s1=s2=1; pwclite.py:124-126,
// uflow_model.py:175) runs on the tiled kernels below, everything else on the literal ones.
//
// Tiled forward ("column thread" layout):
//   CTA tile = 32 x 8 output pixels of one batch item, 9 warps, warp w <-> horizontal displacement
//   dx = w-4, lane <-> x.  A thread keeps the 8 rows x 9 vertical displacements of its column in
//   registers (72 accumulators).  Per channel it reads 8 f1 values and the 16-row f2 halo column
//   at x+dx from shared memory (24 conflict-free LDS.32) and issues 72 FFMA.  f1/f2 never go
//   through an NHWC staging copy (the reference's channels_first pass, .cu:15-39, is gone):
//   tiles are staged straight from NCHW, zero padding is produced while staging.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ literal kernels ------
struct CorrGeom {
    int B, C, H, W, pad, ks, md, s1, s2;
    int kr, dr, D, oH, oW;
};

__device__ __forceinline__ float ld_padded(const float* __restrict__ f, int b, int c, int yp, int xp,
                                           const CorrGeom& g) {
    int y = yp - g.pad, x = xp - g.pad;
    if (y < 0 || y >= g.H || x < 0 || x >= g.W) return 0.f;
    return __ldg(f + (((size_t)b * g.C + c) * g.H + y) * g.W + x);
}

__global__ void corr_fwd_literal(const float* __restrict__ f1, const float* __restrict__ f2,
                                 float* __restrict__ out, CorrGeom g) {
    long long total = (long long)g.B * g.D * g.D * g.oH * g.oW;
    float nelems = (float)(g.ks * g.ks * g.C);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int ox = idx % g.oW;
        long long t = idx / g.oW;
        int oy = t % g.oH; t /= g.oH;
        int tc = t % (g.D * g.D);
        int b = t / (g.D * g.D);
        int ti = tc % g.D - g.dr, tj = tc / g.D - g.dr;
        int y1 = oy * g.s1 + g.md, x1 = ox * g.s1 + g.md;
        int y2 = y1 + tj * g.s2, x2 = x1 + ti * g.s2;
        float acc = 0.f;
        for (int j = -g.kr; j <= g.kr; ++j)
            for (int i = -g.kr; i <= g.kr; ++i)
                for (int c = 0; c < g.C; ++c)
                    acc = fmaf(ld_padded(f1, b, c, y1 + j, x1 + i, g),
                               ld_padded(f2, b, c, y2 + j, x2 + i, g), acc);
        out[idx] = acc / nelems;
    }
}

// which = 0: gradient w.r.t. f1 (other = f2); which = 1: gradient w.r.t. f2 (other = f1).
// Window arithmetic (C integer division, truncating) is the reference's, .cu:141-159 and 255-277.
__global__ void corr_bwd_literal(const float* __restrict__ other, const float* __restrict__ gout,
                                 float* __restrict__ gin, CorrGeom g, int which) {
    long long total = (long long)g.B * g.C * g.H * g.W;
    float nelems = (float)(g.ks * g.ks * g.C);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int xu = idx % g.W;
        long long t = idx / g.W;
        int yu = t % g.H; t /= g.H;
        int c = t % g.C;
        int b = t / g.C;
        int y = yu + g.pad, x = xu + g.pad;
        float acc = 0.f;
        for (int tc = 0; tc < g.D * g.D; ++tc) {
            int i2 = (tc % g.D - g.dr) * g.s2;
            int j2 = (tc / g.D - g.dr) * g.s2;
            int sx = which ? i2 : 0, sy = which ? j2 : 0;
            int xmin = (x - g.kr - g.md - sx) / g.s1, ymin = (y - g.kr - g.md - sy) / g.s1;
            int xmax = (x + g.kr - g.md - sx) / g.s1, ymax = (y + g.kr - g.md - sy) / g.s1;
            if (xmax < 0 || ymax < 0 || xmin >= g.oW || ymin >= g.oH) continue;
            if (xmin > xmax || ymin > ymax) continue;
            xmin = max(0, xmin); xmax = min(g.oW - 1, xmax);
            ymin = max(0, ymin); ymax = min(g.oH - 1, ymax);
            float v = which ? ld_padded(other, b, c, y - j2, x - i2, g)
                            : ld_padded(other, b, c, y + j2, x + i2, g);
            const float* go = gout + ((size_t)b * g.D * g.D + tc) * g.oH * g.oW;
            for (int j = ymin; j <= ymax; ++j)
                for (int i = xmin; i <= xmax; ++i) acc = fmaf(__ldg(go + (size_t)j * g.oW + i), v, acc);
        }
        gin[idx] = acc / nelems;
    }
}

// ------------------------------------------------------------------ tiled forward, md=4 --
constexpr int kTW = 32;            // tile width  (lane <-> x)
constexpr int kTH = 8;             // tile height (rows per thread)
constexpr int kMD = 4;
constexpr int kD = 2 * kMD + 1;    // 9
constexpr int kHW = kTW + 2 * kMD; // 40 halo width
constexpr int kHH = kTH + 2 * kMD; // 16 halo height
constexpr int kCc = 8;             // channels staged per chunk
constexpr int kFwdThreads = 32 * kD;

__global__ void __launch_bounds__(kFwdThreads, 2)
corr_fwd_md4(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out,
             int C, int H, int W, float inv_c, int use_div) {
    __shared__ float s1[kCc][kTH][kTW];
    __shared__ float s2[kCc][kHH][kHW];

    const int lane = threadIdx.x & 31;
    const int wdx = threadIdx.x >> 5;  // 0..8  -> dx = wdx-4
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, b = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* f1b = f1 + (size_t)b * C * plane;
    const float* f2b = f2 + (size_t)b * C * plane;

    float acc[kTH][kD];
#pragma unroll
    for (int r = 0; r < kTH; ++r)
#pragma unroll
        for (int d = 0; d < kD; ++d) acc[r][d] = 0.f;

    for (int c0 = 0; c0 < C; c0 += kCc) {
        // stage f1 tile
        for (int e = threadIdx.x; e < kCc * kTH * kTW; e += kFwdThreads) {
            int xx = e % kTW, rr = (e / kTW) % kTH, cc = e / (kTW * kTH);
            int gx = x0 + xx, gy = y0 + rr, gc = c0 + cc;
            float v = 0.f;
            if (gc < C && gy < H && gx < W) v = __ldg(f1b + gc * plane + (size_t)gy * W + gx);
            s1[cc][rr][xx] = v;
        }
        // stage f2 halo tile (zero outside the image == the reference's zero padding)
        for (int e = threadIdx.x; e < kCc * kHH * kHW; e += kFwdThreads) {
            int xx = e % kHW, rr = (e / kHW) % kHH, cc = e / (kHW * kHH);
            int gx = x0 + xx - kMD, gy = y0 + rr - kMD, gc = c0 + cc;
            float v = 0.f;
            if (gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W)
                v = __ldg(f2b + gc * plane + (size_t)gy * W + gx);
            s2[cc][rr][xx] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int cc = 0; cc < kCc; ++cc) {
            float a[kTH], bb[kHH];
#pragma unroll
            for (int r = 0; r < kTH; ++r) a[r] = s1[cc][r][lane];
#pragma unroll
            for (int k = 0; k < kHH; ++k) bb[k] = s2[cc][k][lane + wdx];
#pragma unroll
            for (int r = 0; r < kTH; ++r)
#pragma unroll
                for (int d = 0; d < kD; ++d) acc[r][d] = fmaf(a[r], bb[r + d], acc[r][d]);
        }
        __syncthreads();
    }

    const int gx = x0 + lane;
    if (gx < W) {
        float* ob = out + (size_t)b * (kD * kD) * plane;
#pragma unroll
        for (int d = 0; d < kD; ++d) {
            float* op = ob + (size_t)(d * kD + wdx) * plane;
#pragma unroll
            for (int r = 0; r < kTH; ++r) {
                int gy = y0 + r;
                if (gy < H) {
                    float v = use_div ? acc[r][d] / (float)C : acc[r][d] * inv_c;
                    __stcs(op + (size_t)gy * W + gx, v);
                }
            }
        }
    }
}

int make_geom(CorrGeom& g, int B, int C, int H, int W, int pad, int ks, int md, int s1, int s2) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return ARF_EINVAL;
    if (pad < 0 || ks < 1 || (ks & 1) == 0 || md < 0 || s1 < 1 || s2 < 1) return ARF_EINVAL;
    g.B = B; g.C = C; g.H = H; g.W = W; g.pad = pad; g.ks = ks; g.md = md; g.s1 = s1; g.s2 = s2;
    g.kr = (ks - 1) / 2;
    g.dr = md / s2;
    g.D = 2 * g.dr + 1;
    int br = g.kr + md;
    int ph = H + 2 * pad - 2 * br, pw = W + 2 * pad - 2 * br;
    if (ph <= 0 || pw <= 0) return ARF_EINVAL;
    g.oH = (ph + s1 - 1) / s1;
    g.oW = (pw + s1 - 1) / s1;
    return ARF_OK;
}

inline bool is_fast(const CorrGeom& g) {
    return g.ks == 1 && g.s1 == 1 && g.s2 == 1 && g.md == kMD && g.pad == kMD;
}

}  // namespace

extern "C" int arf_corr_out_dims(int H, int W, int pad, int ks, int md, int s1, int s2,
                                 int* D2, int* oH, int* oW) {
    CorrGeom g;
    int rc = make_geom(g, 1, 1, H, W, pad, ks, md, s1, s2);
    if (rc) return rc;
    if (D2) *D2 = g.D * g.D;
    if (oH) *oH = g.oH;
    if (oW) *oW = g.oW;
    return ARF_OK;
}

extern "C" int arf_corr_fwd(const float* f1, const float* f2, float* out, int B, int C, int H, int W,
                            int pad, int ks, int md, int s1, int s2, void* stream) {
    ARF_REQUIRE(f1 && f2 && out);
    CorrGeom g;
    int rc = make_geom(g, B, C, H, W, pad, ks, md, s1, s2);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (is_fast(g) && B <= 65535) {
        dim3 grid(arf_cdiv(W, kTW), arf_cdiv(H, kTH), B);
        int pow2 = (C & (C - 1)) == 0;
        corr_fwd_md4<<<grid, kFwdThreads, 0, st>>>(f1, f2, out, C, H, W, 1.0f / (float)C, !pow2);
    } else {
        long long total = (long long)B * g.D * g.D * g.oH * g.oW;
        corr_fwd_literal<<<arf_grid_1d(total, 256), 256, 0, st>>>(f1, f2, out, g);
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_corr_bwd(const float* f1, const float* f2, const float* gout, float* g1, float* g2,
                            int B, int C, int H, int W, int pad, int ks, int md, int s1, int s2,
                            void* stream) {
    ARF_REQUIRE(f1 && f2 && gout);
    CorrGeom g;
    int rc = make_geom(g, B, C, H, W, pad, ks, md, s1, s2);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)B * C * H * W;
    if (g1) {
        corr_bwd_literal<<<arf_grid_1d(total, 256), 256, 0, st>>>(f2, gout, g1, g, 0);
        ARF_CHECK_LAUNCH();
    }
    if (g2) {
        corr_bwd_literal<<<arf_grid_1d(total, 256), 256, 0, st>>>(f1, gout, g2, g, 1);
        ARF_CHECK_LAUNCH();
    }
    return ARF_OK;
}

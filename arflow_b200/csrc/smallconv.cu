// Weight and bias gradient of the flow-output convolutions of the PWC decoders (sm_100a).
//
// Every pyramid level ends in Conv2d(32, 2, 3, padding=1) on the dense block's context features, and the refinement
// network ends in the same shape (models/uflow_model.py:139-143, 232-249).  As an implicit GEMM that weight gradient
// is 2 x 288 outputs over K = N*H*W pixels: cuDNN runs `wgrad_alg0_engine_NHWC` for it, 145 us at 16 x 96 x 128 (plus
// 12 us for ATen's bias reduction), although the operand is 25 MB (4 us of HBM time) and the arithmetic is 0.2 GFLOP.
// This kernel streams the channels-last input once in fp32 (no TF32 rounding):
//   dW[co][kh][kw][ci] = sum_{n,y,x} gy[n][co][y][x] * X[n][y+kh-1][x+kw-1][ci]
//   db[co]             = sum_{n,y,x} gy[n][co][y][x]
// lane <-> input channel (32 per channel group), a warp walks a 32-pixel run of one image row with the 3 x 3 x Cout
// neighbourhood of gy in registers (sliding window, 3 * Cout broadcast loads per pixel), loads of four pixels issued
// before their use.  Per-CTA partial sums, then a fixed-order finalize: deterministic.
#include "common.cuh"

namespace {

constexpr int kSWarps = 8;      // warps (pixel runs) per CTA
constexpr int kSRun = 32;       // pixels per run

template <int kCout>
__global__ void __launch_bounds__(kSWarps * 32)
conv3x3_small_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partials, int N,
                           int H, int W, int Cin, int runs_per_row) {
    __shared__ float red[kSWarps][kCout * 9 + kCout][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cg = blockIdx.y;                                   // channel group of 32
    const long long run = (long long)blockIdx.x * kSWarps + w;
    const long long nrun = (long long)N * H * runs_per_row;
    float acc[kCout][9];
    float sb[kCout];
#pragma unroll
    for (int co = 0; co < kCout; ++co) {
        sb[co] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[co][t] = 0.f;
    }
    if (run < nrun) {
        const int rx = (int)(run % runs_per_row);
        const long long t = run / runs_per_row;
        const int yy = (int)(t % H), n = (int)(t / H);
        const int x0 = rx * kSRun, x1 = x0 + kSRun < W ? x0 + kSRun : W;
        const float* xr = x + ((long long)n * H + yy) * W * Cin + cg * 32 + lane;
        const float* gyn = gy + (long long)n * kCout * H * W;
        // gy[n][co][yy + 1 - kh][cx], zero outside the image (the convolution's zero padding seen from the input side)
        auto ld = [&](int co, int kh, int cx) {
            const int ry = yy + 1 - kh;
            const bool ok = ry >= 0 && ry < H && cx >= 0 && cx < W;
            return ok ? __ldg(gyn + ((long long)co * H + ry) * W + cx) : 0.f;
        };
        // window[co][kh][j]: column xx - 1 + j of gy, i.e. kw = 2 - j
        float win[kCout][3][3];
#pragma unroll
        for (int co = 0; co < kCout; ++co)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                win[co][kh][0] = ld(co, kh, x0 - 1);
                win[co][kh][1] = ld(co, kh, x0);
            }
        for (int xx = x0; xx < x1; xx += 4) {
            float xv[4], gn[4][kCout][3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int px = xx + i;
                xv[i] = px < x1 ? __ldg(xr + (long long)px * Cin) : 0.f;     // a zero input adds nothing
#pragma unroll
                for (int co = 0; co < kCout; ++co)
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) gn[i][co][kh] = ld(co, kh, px + 1);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int co = 0; co < kCout; ++co) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        win[co][kh][2] = gn[i][co][kh];
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) acc[co][kh * 3 + kw] = fmaf(xv[i], win[co][kh][2 - kw], acc[co][kh * 3 + kw]);
                    }
                    if (xx + i < x1) sb[co] += win[co][1][1];               // gy at the pixel itself
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        win[co][kh][0] = win[co][kh][1];
                        win[co][kh][1] = win[co][kh][2];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int co = 0; co < kCout; ++co) {
#pragma unroll
        for (int t = 0; t < 9; ++t) red[w][co * 9 + t][lane] = acc[co][t];
        red[w][kCout * 9 + co][lane] = sb[co];
    }
    __syncthreads();
    // partials[cta][c], c = (co*9 + tap) * Cin + ci for the weights, then kCout bias sums (written by channel group 0)
    const int ctot = kCout * 9 * Cin + kCout;
    float* out = partials + (long long)blockIdx.x * ctot;
    for (int e = threadIdx.x; e < kCout * 9 * 32; e += kSWarps * 32) {
        const int r = e >> 5, l = e & 31;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kSWarps; ++k) s += red[k][r][l];
        out[r * Cin + cg * 32 + l] = s;
    }
    if (cg == 0 && threadIdx.x < kCout) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kSWarps; ++k) s += red[k][kCout * 9 + threadIdx.x][0];
        out[kCout * 9 * Cin + threadIdx.x] = s;
    }
}

// out[c] = sum over CTAs of partials[cta * C + c]; block (32 columns, 32 stripes over the CTAs), fixed order, doubles
__global__ void __launch_bounds__(1024)
column_sum_kernel(const float* __restrict__ partials, float* __restrict__ out, long long nblk, int C) {
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
        out[c] = (float)s;
    }
}

long long small_wgrad_ctas(int N, int H, int W) {
    const long long runs = (long long)N * H * arf_cdiv(W, kSRun);
    return (runs + kSWarps - 1) / kSWarps;
}

}  // namespace

extern "C" long long arf_conv3x3_small_wgrad_workspace(int N, int H, int W, int Cin, int Cout) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cin % 32 || Cout != 2) return ARF_EINVAL;
    return small_wgrad_ctas(N, H, W) * (Cout * 9LL * Cin + Cout);
}

extern "C" int arf_conv3x3_small_wgrad(const float* x, const float* gy, float* out, float* partials, int N, int H, int W,
                                       int Cin, int Cout, void* stream) {
    ARF_REQUIRE(x && gy && out && partials && N > 0 && H > 0 && W > 0);
    if (Cout != 2 || Cin % 32 != 0 || Cin <= 0) return ARF_EUNSUPPORTED;
    const long long ctas = small_wgrad_ctas(N, H, W);
    ARF_REQUIRE(ctas <= 0x7fffffffLL && Cin / 32 <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ctas, Cin / 32);
    conv3x3_small_wgrad_kernel<2><<<grid, kSWarps * 32, 0, st>>>(x, gy, partials, N, H, W, Cin, arf_cdiv(W, kSRun));
    ARF_CHECK_LAUNCH();
    const int ctot = Cout * 9 * Cin + Cout;
    column_sum_kernel<<<arf_cdiv(ctot, 32), dim3(32, 32), 0, st>>>(partials, out, ctas, ctot);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

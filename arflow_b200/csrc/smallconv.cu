// The flow-output convolutions of the PWC decoders, forward and backward (sm_100a).
//
// Every pyramid level ends in Conv2d(32, 2, 3, padding=1) on the dense block's context features, and the refinement
// network ends in the same shape (models/uflow_model.py:139-143, 232-249).  As implicit GEMMs these are 2-wide: cuDNN
// pads the channels and runs 256-wide tiles - at 16 x 96 x 128 the forward costs 67 us (two padding kernels, fprop,
// ATen bias add, NHWC -> NCHW copy), the input gradient 33 us and `wgrad_alg0_engine_NHWC` 145 us (plus 12 us for
// ATen's bias reduction), although the operand is 25 MB (4 us of HBM time) and the arithmetic is 0.2 GFLOP per pass.
// These kernels stream the channels-last operand once, in fp32 (no TF32 rounding):
//   y[n][co][y][x]     = b[co] + sum_{kh,kw,ci} w[co][kh][kw][ci] * X[n][y+kh-1][x+kw-1][ci]          (NCHW out)
//   dX[n][y][x][ci]    = sum_{co,kh,kw} gy[n][co][y-kh+1][x-kw+1] * w[co][kh][kw][ci]
//   dW[co][kh][kw][ci] = sum_{n,y,x} gy[n][co][y][x] * X[n][y+kh-1][x+kw-1][ci]
//   db[co]             = sum_{n,y,x} gy[n][co][y][x]
// lane <-> input channel (32 per channel group), a warp walks a 32-pixel run of one image row.  Backward: the 3 x 3
// x Cout neighbourhood of gy slides along in registers (3 * Cout broadcast loads per pixel) and feeds both gradients;
// per-CTA partial sums of dW / db, then a fixed-order finalize (deterministic).  Forward: the 3 x 3 neighbourhood of
// X slides along, per-lane partial dot products of a run go through shared memory so that the sum over the 32
// channels and the NCHW store are done with lane <-> pixel.  Loads of four pixels are issued before their use.
#include "common.cuh"

namespace {

constexpr int kSWarps = 8;      // warps (pixel runs) per CTA
constexpr int kSRun = 32;       // pixels per run

template <int kCout, bool kDgrad>
__global__ void __launch_bounds__(kSWarps * 32)
conv3x3_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ wt,
                         float* __restrict__ gx, float* __restrict__ partials, int N, int H, int W, int Cin,
                         int runs_per_row) {
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
    float wr[kCout][9];                                          // w[co][tap][this lane's channel]
    if (kDgrad) {
#pragma unroll
        for (int co = 0; co < kCout; ++co)
#pragma unroll
            for (int t = 0; t < 9; ++t) wr[co][t] = __ldg(wt + ((long long)co * 9 + t) * Cin + cg * 32 + lane);
    }
    if (run < nrun) {
        const int rx = (int)(run % runs_per_row);
        const long long t = run / runs_per_row;
        const int yy = (int)(t % H), n = (int)(t / H);
        const int x0 = rx * kSRun, x1 = x0 + kSRun < W ? x0 + kSRun : W;
        const long long row_off = ((long long)n * H + yy) * W * Cin + cg * 32 + lane;
        const float* xr = x + row_off;
        float* gxr = kDgrad ? gx + row_off : nullptr;
        const float* gyn = gy + (long long)n * kCout * H * W;
        // gy[n][co][yy + 1 - kh][cx], zero outside the image (the convolution's zero padding seen from the input side).
        // (Per-(co, kh) row pointers instead of this index arithmetic were slower: 71 vs 59 us at 16 x 96 x 128.)
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
                float dxv = 0.f;
#pragma unroll
                for (int co = 0; co < kCout; ++co) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        win[co][kh][2] = gn[i][co][kh];
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            acc[co][kh * 3 + kw] = fmaf(xv[i], win[co][kh][2 - kw], acc[co][kh * 3 + kw]);
                            if (kDgrad) dxv = fmaf(win[co][kh][2 - kw], wr[co][kh * 3 + kw], dxv);
                        }
                    }
                    if (xx + i < x1) sb[co] += win[co][1][1];               // gy at the pixel itself
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        win[co][kh][0] = win[co][kh][1];
                        win[co][kh][1] = win[co][kh][2];
                    }
                }
                if (kDgrad && xx + i < x1) gxr[(long long)(xx + i) * Cin] = dxv;
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

constexpr int kFWarps = 4;      // forward: warps per CTA (each owns kCout x 32 x 33 floats of shared memory)

template <int kCout>
__global__ void __launch_bounds__(kFWarps * 32)
conv3x3_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                         float* __restrict__ y, int N, int H, int W, int Cin, int runs_per_row) {
    __shared__ float part[kFWarps][kCout][kSRun][33];            // [pixel of the run][channel lane], padded
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long run = (long long)blockIdx.x * kFWarps + w;
    const long long nrun = (long long)N * H * runs_per_row;
    if (run >= nrun) return;                                     // no block-wide barrier below
    const int rx = (int)(run % runs_per_row);
    const long long t = run / runs_per_row;
    const int yy = (int)(t % H), n = (int)(t / H);
    const int x0 = rx * kSRun, x1 = x0 + kSRun < W ? x0 + kSRun : W;
    float acc[kCout][kSRun / 32];                                // this lane's output pixel (lane <-> pixel at the end)
#pragma unroll
    for (int co = 0; co < kCout; ++co) acc[co][0] = 0.f;
    for (int cg = 0; cg < Cin / 32; ++cg) {
        float wr[kCout][9];
#pragma unroll
        for (int co = 0; co < kCout; ++co)
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) wr[co][tp] = __ldg(wt + ((long long)co * 9 + tp) * Cin + cg * 32 + lane);
        const float* xn = x + (long long)n * H * W * Cin + cg * 32 + lane;
        // X[n][yy - 1 + r][cx][lane's channel], zero outside the image
        auto ld = [&](int r, int cx) {
            const int ry = yy - 1 + r;
            const bool ok = ry >= 0 && ry < H && cx >= 0 && cx < W;
            return ok ? __ldg(xn + ((long long)ry * W + cx) * Cin) : 0.f;
        };
        float win[3][3];                                         // [kh][kw]: column xx - 1 + kw
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            win[r][0] = ld(r, x0 - 1);
            win[r][1] = ld(r, x0);
        }
        for (int xx = x0; xx < x1; xx += 4) {
            float xn4[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int r = 0; r < 3; ++r) xn4[i][r] = ld(r, xx + i + 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int r = 0; r < 3; ++r) win[r][2] = xn4[i][r];
#pragma unroll
                for (int co = 0; co < kCout; ++co) {
                    float s = 0.f;
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) s = fmaf(win[r][kw], wr[co][r * 3 + kw], s);
                    if (xx + i < x1) part[w][co][xx + i - x0][lane] = s;
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    win[r][0] = win[r][1];
                    win[r][1] = win[r][2];
                }
            }
        }
        __syncwarp();
        if (x0 + lane < x1) {
#pragma unroll
            for (int co = 0; co < kCout; ++co) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) s += part[w][co][lane][c];
                acc[co][0] += s;
            }
        }
        __syncwarp();
    }
    if (x0 + lane < x1) {
#pragma unroll
        for (int co = 0; co < kCout; ++co)
            y[(((long long)n * kCout + co) * H + yy) * W + x0 + lane] = acc[co][0] + (bias ? __ldg(bias + co) : 0.f);
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

// ---- weight gradient of the first pyramid convolution --------------------------------------------------------------
// Conv2d(3, 32, 3, stride 2, padding 1) on the image (models/uflow_model.py:427-436), which the channels-last path
// stores as 8-channel NHWC pixels (3 real channels, 5 zeros).  2304 outputs over K = N*Ho*Wo = 786432 pixels at
// chairs_uflow: cuDNN's 64 x 64 wgrad tile takes 197 us, the operands are 200 MB (31 us of HBM time).  lane <-> output
// channel; a warp walks 32-pixel runs of output rows with the 3 x 3 x 3 input neighbourhood in registers (stride 2:
// one column carried over, two loaded per pixel as broadcast 16-byte loads).  Only the real input channels get a
// gradient.  partials[cta][((kh*3 + kw)*3 + ci)*32 + co], then column_sum_kernel.
constexpr int kFirstCin = 3, kFirstCout = 32, kFirstK = 9 * kFirstCin;

__global__ void __launch_bounds__(kSWarps * 32, 3)
conv3x3s2_first_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ partials, int Hi,
                             int Wi, int Ho, int Wo, int runs_per_row, long long nrun) {
    __shared__ float red[kSWarps][kFirstK][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float acc[9][kFirstCin];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < kFirstCin; ++c) acc[t][c] = 0.f;
    for (long long run = (long long)blockIdx.x * kSWarps + w; run < nrun; run += (long long)gridDim.x * kSWarps) {
        const int rx = (int)(run % runs_per_row);
        const long long t = run / runs_per_row;
        const int oy = (int)(t % Ho);
        const long long n = t / Ho;
        const int ox0 = rx * kSRun, ox1 = ox0 + kSRun < Wo ? ox0 + kSRun : Wo;
        const float* gr = g + ((n * Ho + oy) * Wo) * kFirstCout + lane;
        // channels 0..3 of input pixel (2*oy - 1 + r, ix), zero outside the image: one row pointer and one validity
        // flag per kernel row, 32-bit column arithmetic per load
        const float4* rp[3];
        bool rok[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int iy = 2 * oy - 1 + r;
            rok[r] = iy >= 0 && iy < Hi;
            rp[r] = reinterpret_cast<const float4*>(x + (n * Hi + (rok[r] ? iy : 0)) * Wi * 8);
        }
        auto ld = [&](int r, int ix) {
            return (rok[r] && (unsigned)ix < (unsigned)Wi) ? __ldg(rp[r] + 2 * ix) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        float4 prev[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) prev[r] = ld(r, 2 * ox0 - 1);
        for (int ox = ox0; ox < ox1; ox += 2) {
            float gv[2];
            float4 ca[2][3], cb[2][3];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int px = ox + i;
                gv[i] = px < ox1 ? __ldg(gr + (long long)px * kFirstCout) : 0.f;    // a zero gradient adds nothing
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    ca[i][r] = ld(r, 2 * px);
                    cb[i][r] = ld(r, 2 * px + 1);
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float4 col[3] = {prev[r], ca[i][r], cb[i][r]};
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        acc[r * 3 + kw][0] = fmaf(gv[i], col[kw].x, acc[r * 3 + kw][0]);
                        acc[r * 3 + kw][1] = fmaf(gv[i], col[kw].y, acc[r * 3 + kw][1]);
                        acc[r * 3 + kw][2] = fmaf(gv[i], col[kw].z, acc[r * 3 + kw][2]);
                    }
                    prev[r] = cb[i][r];
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < kFirstCin; ++c) red[w][t * kFirstCin + c][lane] = acc[t][c];
    __syncthreads();
    float* out = partials + (long long)blockIdx.x * (kFirstK * 32);
    for (int e = threadIdx.x; e < kFirstK * 32; e += kSWarps * 32) {
        const int r = e >> 5, l = e & 31;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kSWarps; ++k) s += red[k][r][l];
        out[e] = s;
    }
}

long long first_wgrad_ctas(int N, int Ho, int Wo) {
    const long long runs = (long long)N * Ho * arf_cdiv(Wo, kSRun);
    const long long need = (runs + kSWarps - 1) / kSWarps;
    return need < 3LL * ARF_NUM_SMS ? need : 3LL * ARF_NUM_SMS;   // one resident wave (3 CTAs per SM), grid-stride over runs
}

long long small_wgrad_ctas(int N, int H, int W) {
    const long long runs = (long long)N * H * arf_cdiv(W, kSRun);
    return (runs + kSWarps - 1) / kSWarps;
}

}  // namespace

extern "C" int arf_conv3x3_small_fwd(const float* x, const float* w, const float* bias, float* y, int N, int H, int W,
                                     int Cin, int Cout, void* stream) {
    ARF_REQUIRE(x && w && y && N > 0 && H > 0 && W > 0);
    if (Cout != 2 || Cin % 32 != 0 || Cin <= 0) return ARF_EUNSUPPORTED;
    const int rpr = arf_cdiv(W, kSRun);
    const long long ctas = ((long long)N * H * rpr + kFWarps - 1) / kFWarps;
    ARF_REQUIRE(ctas <= 0x7fffffffLL);
    conv3x3_small_fwd_kernel<2><<<(unsigned)ctas, kFWarps * 32, 0, (cudaStream_t)stream>>>(x, w, bias, y, N, H, W, Cin, rpr);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" long long arf_conv3x3_small_bwd_workspace(int N, int H, int W, int Cin, int Cout) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cin % 32 || Cout != 2) return ARF_EINVAL;
    return small_wgrad_ctas(N, H, W) * (Cout * 9LL * Cin + Cout);
}

extern "C" int arf_conv3x3_small_bwd(const float* x, const float* gy, const float* w, float* gx, float* out, float* partials,
                                     int N, int H, int W, int Cin, int Cout, void* stream) {
    ARF_REQUIRE(x && gy && out && partials && N > 0 && H > 0 && W > 0);
    ARF_REQUIRE(!gx || w);
    if (Cout != 2 || Cin % 32 != 0 || Cin <= 0) return ARF_EUNSUPPORTED;
    const long long ctas = small_wgrad_ctas(N, H, W);
    ARF_REQUIRE(ctas <= 0x7fffffffLL && Cin / 32 <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ctas, Cin / 32);
    if (gx)
        conv3x3_small_bwd_kernel<2, true><<<grid, kSWarps * 32, 0, st>>>(x, gy, w, gx, partials, N, H, W, Cin,
                                                                          arf_cdiv(W, kSRun));
    else
        conv3x3_small_bwd_kernel<2, false><<<grid, kSWarps * 32, 0, st>>>(x, gy, nullptr, nullptr, partials, N, H, W, Cin,
                                                                           arf_cdiv(W, kSRun));
    ARF_CHECK_LAUNCH();
    const int ctot = Cout * 9 * Cin + Cout;
    column_sum_kernel<<<arf_cdiv(ctot, 32), dim3(32, 32), 0, st>>>(partials, out, ctas, ctot);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" long long arf_conv3x3s2_first_wgrad_workspace(int N, int Hi, int Wi) {
    if (N <= 0 || Hi <= 0 || Wi <= 0) return ARF_EINVAL;
    return first_wgrad_ctas(N, (Hi - 1) / 2 + 1, (Wi - 1) / 2 + 1) * (kFirstK * 32);
}

extern "C" int arf_conv3x3s2_first_wgrad(const float* x, const float* g, float* out, float* partials, int N, int Hi, int Wi,
                                         int Cin_real, int Cout, void* stream) {
    ARF_REQUIRE(x && g && out && partials && N > 0 && Hi > 0 && Wi > 0);
    if (Cin_real != kFirstCin || Cout != kFirstCout) return ARF_EUNSUPPORTED;
    ARF_REQUIRE((uintptr_t)x % 16 == 0);
    const int Ho = (Hi - 1) / 2 + 1, Wo = (Wi - 1) / 2 + 1;
    const int rpr = arf_cdiv(Wo, kSRun);
    const long long ctas = first_wgrad_ctas(N, Ho, Wo);
    cudaStream_t st = (cudaStream_t)stream;
    conv3x3s2_first_wgrad_kernel<<<(unsigned)ctas, kSWarps * 32, 0, st>>>(x, g, partials, Hi, Wi, Ho, Wo, rpr,
                                                                          (long long)N * Ho * rpr);
    ARF_CHECK_LAUNCH();
    column_sum_kernel<<<arf_cdiv(kFirstK * 32, 32), dim3(32, 32), 0, st>>>(partials, out, ctas, kFirstK * 32);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

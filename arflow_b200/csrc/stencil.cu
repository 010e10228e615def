// Stencil-triangular algebra of the non-diagonal ELBO losses.
//
// T1/T2  y = L x and y = L^T x for the lower-triangular stencil matrix with (k+1)^2 taps per flow
//        channel (utils/triag_solve.py:29-43, 59-73): HBM-bound streaming kernels, every coefficient
//        is read exactly once; the backward pass produces dX and all tap gradients in one sweep.
// T3     forward / backward substitution for the k=1 stencil (triag_solve_cuda.cu:7-69).  The reference
//        runs one THREAD per system through M*N serial global-memory updates.  Here one CTA per system
//        walks the M+N-1 anti-diagonals: thread <-> row, the three previous values a cell needs come
//        from the thread's own register (left) and its upper neighbour's last two values (shared
//        memory), coefficients for the next diagonal are prefetched while the current one is solved.
// T4     diag((L L^T)^-1) (triag_solve_cuda.cu:72-139): one CTA per (system, start pixel) runs the same
//        wavefront on the sub-rectangle below/right of the start pixel and reduces the squares.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------- mat-vec -----
// transposed = 0:  Y[p] = sum_t A[t][p - o_t] * X[p - o_t]        (o_t = (i,j), inside the image)
// transposed = 1:  Y[p] = sum_t A[t][p]       * X[p + o_t]
// grid = (x blocks, y, n*2 + ch): no integer division; K1 = k+1 is a template parameter for the common
// supports (k = 1, 3) so the (k+1)^2 coefficient loads of a pixel are all in flight together.
template <int K1T>
__global__ void __launch_bounds__(128)
stencil_mv_kernel(const float* __restrict__ A, const float* __restrict__ X, float* __restrict__ Y, int H, int W,
                  int k1_rt, int transposed) {
    const int k1 = K1T > 0 ? K1T : k1_rt;
    const size_t hw = (size_t)H * W;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t n = blockIdx.z >> 1;
    const int ch = blockIdx.z & 1;
    const float* An = A + (n * 2 * k1 * k1 + ch) * hw;
    const float* Xn = X + (n * 2 + ch) * hw;
    const size_t p = (size_t)y * W + x;
    float acc = 0.f;
    // Interior warps (no tap of any lane leaves the image; all but the first k rows / first warp of a row) take a
    // path with 32-bit offsets and no bounds logic: the checked path below spends ~500 instructions per pixel on
    // 64-bit index arithmetic and selects (ncu: ALU pipe 78%, issue 68%, HBM 36%).
    if (K1T > 0 && 2ull * K1T * K1T * hw < 0x7fffffffull) {
        constexpr int T = K1T > 0 ? K1T * K1T : 1;
        constexpr int K = K1T > 0 ? K1T : 1;
        const bool interior = transposed ? (y + K - 1 < H && x + K - 1 < W) : (y >= K - 1 && x >= K - 1);
        if (__all_sync(__activemask(), interior)) {
            const int ihw2 = 2 * (int)hw, ip = (int)p;
            float a[T], xv[T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int d = (t / K) * W + (t % K);
                a[t] = arf_ldg_stream(An + (t * ihw2 + (transposed ? ip : ip - d)));
                xv[t] = __ldg(Xn + (transposed ? ip + d : ip - d));
            }
#pragma unroll
            for (int t = 0; t < T; ++t) acc = fmaf(a[t], xv[t], acc);
            Y[(n * 2 + ch) * hw + p] = acc;
            return;
        }
    }
    if (K1T > 0) {
        // every load of the pixel is issued before the first FMA: taps outside the image read the pixel's own
        // (valid) address and are discarded, so no branch sits between the loads (the first version serialised
        // 16 dependent round trips per thread - ncu: 86% long-scoreboard stalls, 37% of the HBM roof)
        constexpr int T = K1T > 0 ? K1T * K1T : 1;
        float a[T], xv[T];
        bool ok[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int i = t / (K1T > 0 ? K1T : 1), j = t % (K1T > 0 ? K1T : 1);
            const size_t to = (size_t)t * 2 * hw;
            if (!transposed) {
                ok[t] = (y - i >= 0) && (x - j >= 0);
                const size_t o = ok[t] ? p - (size_t)i * W - j : p;
                a[t] = arf_ldg_stream(An + to + o);
                xv[t] = __ldg(Xn + o);
            } else {
                ok[t] = (y + i < H) && (x + j < W);
                a[t] = arf_ldg_stream(An + to + p);
                xv[t] = __ldg(Xn + (ok[t] ? p + (size_t)i * W + j : p));
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t) acc = ok[t] ? fmaf(a[t], xv[t], acc) : acc;
    } else {
        for (int i = 0; i < k1; ++i)
            for (int j = 0; j < k1; ++j) {
                const size_t to = (size_t)(i * k1 + j) * 2 * hw;
                if (!transposed) {
                    if (y - i >= 0 && x - j >= 0) {
                        size_t o = p - (size_t)i * W - j;
                        acc = fmaf(arf_ldg_stream(An + to + o), __ldg(Xn + o), acc);
                    }
                } else {
                    // the reference slices A[..., 0:-i, 0:-j]: the tap exists where p + o stays inside
                    if (y + i < H && x + j < W)
                        acc = fmaf(arf_ldg_stream(An + to + p), __ldg(Xn + p + (size_t)i * W + j), acc);
                }
            }
    }
    Y[(n * 2 + ch) * hw + p] = acc;
}

// Backward of both products in one sweep over A:
//   transposed = 0 (y = L x):    dX[p] = sum_t A[t][p] * gY[p + o_t],    dA[t][p] = X[p] * gY[p + o_t]
//   transposed = 1 (y = L^T x):  dX[p] = sum_t A[t][p - o_t] * gY[p - o_t],  dA[t][p] = gY[p] * X[p + o_t]
template <int K1T>
__global__ void __launch_bounds__(128)
stencil_mv_bwd_kernel(const float* __restrict__ A, const float* __restrict__ X, const float* __restrict__ gY,
                      float* __restrict__ dA, float* __restrict__ dX, int H, int W, int k1_rt, int transposed) {
    const int k1 = K1T > 0 ? K1T : k1_rt;
    const size_t hw = (size_t)H * W;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t n = blockIdx.z >> 1;
    const int ch = blockIdx.z & 1;
    const size_t p = (size_t)y * W + x;
    const float* An = A + (n * 2 * k1 * k1 + ch) * hw;
    float* dAn = dA ? dA + (n * 2 * k1 * k1 + ch) * hw : nullptr;
    const float* Xn = X + (n * 2 + ch) * hw;
    const float* Gn = gY + (n * 2 + ch) * hw;
    const float xp = __ldg(Xn + p), gp = __ldg(Gn + p);
    float acc = 0.f;
    // interior warps: 32-bit offsets, no bounds logic (see stencil_mv_kernel)
    if (K1T > 0 && 2ull * K1T * K1T * hw < 0x7fffffffull) {
        constexpr int T = K1T > 0 ? K1T * K1T : 1;
        constexpr int K = K1T > 0 ? K1T : 1;
        const bool interior = (y + K - 1 < H) && (x + K - 1 < W) && (y >= K - 1) && (x >= K - 1);
        if (__all_sync(__activemask(), interior)) {
            const int ihw2 = 2 * (int)hw, ip = (int)p;
            float a[T], v[T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int d = (t / K) * W + (t % K);
                if (!transposed) {
                    v[t] = __ldg(Gn + (ip + d));
                    a[t] = dX ? arf_ldg_stream(An + (t * ihw2 + ip)) : 0.f;
                } else {
                    v[t] = __ldg(Xn + (ip + d));
                    a[t] = dX ? arf_ldg_stream(An + (t * ihw2 + ip - d)) * __ldg(Gn + (ip - d)) : 0.f;
                }
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (!transposed) {
                    acc = fmaf(a[t], v[t], acc);
                    if (dAn) __stcs(dAn + (t * ihw2 + ip), xp * v[t]);
                } else {
                    if (dAn) __stcs(dAn + (t * ihw2 + ip), gp * v[t]);
                    acc += a[t];
                }
            }
            if (dX) dX[(n * 2 + ch) * hw + p] = acc;
            return;
        }
    }
    if (K1T > 0) {
        // loads first (clamped addresses, no branches between them), then the FMAs and the streaming stores
        constexpr int T = K1T > 0 ? K1T * K1T : 1;
        float a[T], v[T];
        bool in[T], inb[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int i = t / (K1T > 0 ? K1T : 1), j = t % (K1T > 0 ? K1T : 1);
            const size_t to = (size_t)t * 2 * hw;
            in[t] = (y + i < H) && (x + j < W);
            inb[t] = (y - i >= 0) && (x - j >= 0);
            const size_t fw = in[t] ? p + (size_t)i * W + j : p;
            if (!transposed) {
                v[t] = __ldg(Gn + fw);
                a[t] = dX ? arf_ldg_stream(An + to + p) : 0.f;
            } else {
                v[t] = __ldg(Xn + fw);
                const size_t o = inb[t] ? p - (size_t)i * W - j : p;
                a[t] = dX ? arf_ldg_stream(An + to + o) * __ldg(Gn + o) : 0.f;
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const size_t to = (size_t)t * 2 * hw;
            if (!transposed) {
                const float g = in[t] ? v[t] : 0.f;
                acc = fmaf(a[t], g, acc);
                if (dAn) __stcs(dAn + to + p, xp * g);
            } else {
                if (dAn) __stcs(dAn + to + p, in[t] ? gp * v[t] : 0.f);
                acc += inb[t] ? a[t] : 0.f;
            }
        }
    } else {
        for (int i = 0; i < k1; ++i)
            for (int j = 0; j < k1; ++j) {
                const size_t to = (size_t)(i * k1 + j) * 2 * hw;
                const bool in = (y + i < H) && (x + j < W);
                if (!transposed) {
                    float g = in ? __ldg(Gn + p + (size_t)i * W + j) : 0.f;
                    if (dX) acc = fmaf(arf_ldg_stream(An + to + p), g, acc);
                    if (dAn) __stcs(dAn + to + p, xp * g);
                } else {
                    if (dAn) __stcs(dAn + to + p, in ? gp * __ldg(Xn + p + (size_t)i * W + j) : 0.f);
                    if (dX && y - i >= 0 && x - j >= 0) {
                        size_t o = p - (size_t)i * W - j;
                        acc = fmaf(arf_ldg_stream(An + to + o), __ldg(Gn + o), acc);
                    }
                }
            }
    }
    if (dX) dX[(n * 2 + ch) * hw + p] = acc;
}

// ---------------------------------------------------------------------------- wavefront ---
// Solves the stencil system on the rectangle rows [r_lo, M), cols [c_lo, N) of one (M x N) system,
// sweeping anti-diagonals.  Lower (forward substitution, triag_solve.py:76-94):
//   Y[i,j] = (X[i,j] - C[i-1,j] Y[i-1,j] - B[i,j-1] Y[i,j-1] - D[i-1,j-1] Y[i-1,j-1]) / A[i,j]
// Upper (back substitution, :97-115) is the same recurrence on the point-reflected grid with the
// coefficients taken at (i,j) instead of the predecessor:
//   Y[i,j] = (X[i,j] - C[i,j] Y[i+1,j] - B[i,j] Y[i,j+1] - D[i,j] Y[i+1,j+1]) / A[i,j]
// Thread t owns logical row t (rows are processed in logical coordinates li = i - r_lo or reflected).
// unit_rhs: X = e_(r_lo,c_lo) (inverse-diagonal mode), Y is not stored, the sum of squares is returned.
// Block size = rows rounded up to a warp; sY holds the last two diagonals: sY[2][blockDim.x + 1].
// Measured (B200, 64 systems of 112x256): 0.6 us per anti-diagonal.  The limiter is not DRAM latency but L1
// wavefronts: thread <-> row makes every coefficient load touch 32 different sectors (5 arrays x 4 warps x 32
// = 640 L1 wavefronts per step); a deeper register prefetch ring was tried and is slower (274 vs 219 us).
// Next step (round 2): thread <-> column with a per-row affine scan, which makes all loads row-contiguous.
template <bool kUpper, bool kUnit>
__device__ float wavefront(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ C,
                           const float* __restrict__ D, const float* __restrict__ X, float* __restrict__ Y, int M,
                           int N, int r_lo, int c_lo, float* sY) {
    const int rows = M - r_lo, cols = N - c_lo;
    const int t = threadIdx.x;                 // logical row
    const int stride = blockDim.x + 1;
    // physical coordinates of logical (li, lj)
    auto pi = [&](int li) { return kUpper ? M - 1 - li : r_lo + li; };
    auto pj = [&](int lj) { return kUpper ? N - 1 - lj : c_lo + lj; };
    // upper solves always cover the whole grid in this library (r_lo = c_lo = 0)
    float left = 0.f;                          // Y[li][lj-1], own previous value
    float sumsq = 0.f;
    sY[t + 1] = 0.f;
    sY[stride + t + 1] = 0.f;
    if (t == 0) { sY[0] = 0.f; sY[stride] = 0.f; }
    __syncthreads();

    // prefetch for the first cell of this row
    float a = 1.f, b = 0.f, c = 0.f, d = 0.f, x = 0.f;
    auto fetch = [&](int lj) {
        if (t < rows && lj >= 0 && lj < cols) {
            const int i = pi(t), j = pj(lj);
            a = __ldg(A + (size_t)i * N + j);
            x = kUnit ? ((t == 0 && lj == 0) ? 1.f : 0.f) : __ldg(X + (size_t)i * N + j);
            if (!kUpper) {
                b = lj > 0 ? __ldg(B + (size_t)i * (N - 1) + (j - 1)) : 0.f;
                c = t > 0 ? __ldg(C + (size_t)(i - 1) * N + j) : 0.f;
                d = (t > 0 && lj > 0 && D) ? __ldg(D + (size_t)(i - 1) * (N - 1) + (j - 1)) : 0.f;
            } else {
                b = lj > 0 ? __ldg(B + (size_t)i * (N - 1) + j) : 0.f;
                c = t > 0 ? __ldg(C + (size_t)i * N + j) : 0.f;
                d = (t > 0 && lj > 0 && D) ? __ldg(D + (size_t)i * (N - 1) + j) : 0.f;
            }
        }
    };
    fetch(0 - t);
    const int ndiag = rows + cols - 1;
    for (int dg = 0; dg < ndiag; ++dg) {
        const int lj = dg - t;
        const bool live = t < rows && lj >= 0 && lj < cols;
        float* cur = sY + (dg & 1) * stride;           // receives diagonal dg
        const float* prev = sY + ((dg + 1) & 1) * stride;  // diagonal dg-1 ... and, before overwrite, dg-2 lives in cur
        float y = 0.f;
        float up = 0.f, upleft = 0.f;
        if (live) {
            up = prev[t];          // row t-1 on diagonal dg-1  -> Y[li-1][lj]      (slot t holds row t-1)
            upleft = cur[t];       // row t-1 on diagonal dg-2  -> Y[li-1][lj-1]
        }
        const float ca = a, cb = b, cc = c, cd = d, cx = x;
        fetch(lj + 1);             // coefficients of the next diagonal: latency overlaps the solve below
        __syncthreads();           // everyone has read diagonal dg-2 from `cur`
        if (live) {
            y = (((cx - cc * up) - cb * left) - cd * upleft) / ca;
            left = y;
            if (!kUnit) Y[(size_t)pi(t) * N + pj(lj)] = y;
            sumsq = fmaf(y, y, sumsq);
        }
        cur[t + 1] = live ? y : 0.f;
        __syncthreads();
    }
    return sumsq;
}

template <bool kUpper>
__global__ void trisolve_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ C,
                                const float* __restrict__ D, const float* __restrict__ X, float* __restrict__ Y, int M,
                                int N) {
    extern __shared__ float sY[];
    const size_t s = blockIdx.x;
    wavefront<kUpper, false>(A + s * M * N, B + s * M * (N - 1), C + s * (M - 1) * N,
                             D ? D + s * (M - 1) * (N - 1) : nullptr, X + s * M * N, Y + s * M * N, M, N, 0, 0, sY);
}

__global__ void inv_diag_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ C,
                                float* __restrict__ Hout, int M, int N) {
    extern __shared__ float sY[];
    __shared__ float red[32];
    const size_t cell = blockIdx.x;            // (system, k, l)
    const size_t s = cell / ((size_t)M * N);
    const int kl = cell % ((size_t)M * N);
    const int k = kl / N, l = kl % N;
    float v = wavefront<false, true>(A + s * M * N, B + s * M * (N - 1), C + s * (M - 1) * N, nullptr, nullptr, nullptr,
                                     M, N, k, l, sY);
    v = arf_block_sum(v, red);
    if (threadIdx.x == 0) Hout[cell] = v;
}

}  // namespace

extern "C" int arf_stencil_mv_fwd(const float* A, const float* X, float* Y, int N, int H, int W, int k,
                                  int transposed, void* stream) {
    ARF_REQUIRE(A && X && Y && N > 0 && H > 0 && W > 0 && k >= 0 && k <= 15);
    ARF_REQUIRE(H <= 65535 && 2LL * N <= 65535);
    dim3 grid(arf_cdiv(W, 128), H, 2 * N);
    cudaStream_t st = (cudaStream_t)stream;
    const int tr = transposed ? 1 : 0;
    if (k == 1) stencil_mv_kernel<2><<<grid, 128, 0, st>>>(A, X, Y, H, W, 2, tr);
    else if (k == 3) stencil_mv_kernel<4><<<grid, 128, 0, st>>>(A, X, Y, H, W, 4, tr);
    else stencil_mv_kernel<0><<<grid, 128, 0, st>>>(A, X, Y, H, W, k + 1, tr);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_stencil_mv_bwd(const float* A, const float* X, const float* gY, float* dA, float* dX, int N, int H,
                                  int W, int k, int transposed, void* stream) {
    ARF_REQUIRE(A && X && gY && N > 0 && H > 0 && W > 0 && k >= 0 && k <= 15);
    ARF_REQUIRE(H <= 65535 && 2LL * N <= 65535);
    if (!dA && !dX) return ARF_OK;
    dim3 grid(arf_cdiv(W, 128), H, 2 * N);
    cudaStream_t st = (cudaStream_t)stream;
    const int tr = transposed ? 1 : 0;
    if (k == 1) stencil_mv_bwd_kernel<2><<<grid, 128, 0, st>>>(A, X, gY, dA, dX, H, W, 2, tr);
    else if (k == 3) stencil_mv_bwd_kernel<4><<<grid, 128, 0, st>>>(A, X, gY, dA, dX, H, W, 4, tr);
    else stencil_mv_bwd_kernel<0><<<grid, 128, 0, st>>>(A, X, gY, dA, dX, H, W, k + 1, tr);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_trisolve(const float* A, const float* B, const float* C, const float* D, const float* X, float* Y,
                            long long systems, int M, int N, int upper, void* stream) {
    ARF_REQUIRE(A && B && C && X && Y && systems > 0 && systems <= 0x7fffffffLL && M > 0 && N > 0);
    if (M > 1024) return ARF_EUNSUPPORTED;     // thread <-> row
    const int threads = (M + 31) / 32 * 32;
    const size_t smem = 2 * (threads + 1) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (upper) trisolve_kernel<true><<<(int)systems, threads, smem, st>>>(A, B, C, D, X, Y, M, N);
    else trisolve_kernel<false><<<(int)systems, threads, smem, st>>>(A, B, C, D, X, Y, M, N);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_inv_diag(const float* A, const float* B, const float* C, float* Hout, long long systems, int M,
                            int N, void* stream) {
    ARF_REQUIRE(A && B && C && Hout && systems > 0 && M > 0 && N > 0);
    if (M > 1024) return ARF_EUNSUPPORTED;
    const long long cells = systems * M * N;
    if (cells > 0x7fffffffLL) return ARF_EUNSUPPORTED;
    const int threads = (M + 31) / 32 * 32;
    const size_t smem = 2 * (threads + 1) * sizeof(float);
    inv_diag_kernel<<<(int)cells, threads, smem, (cudaStream_t)stream>>>(A, B, C, Hout, M, N);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

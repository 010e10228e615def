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

ARF_HOOK g_trisolve_variant = 0;   // test hook, per calling thread (arf_debug_set key 4): 1 = force the wavefront solve

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
// arf_trisolve therefore uses trisolve_scan_kernel below whenever a row fits one block; the wavefront stays
// for the inverse-diagonal mode (sub-rectangles, no stores) and for rows wider than 1024.
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

// ---------------------------------------------------------------------------- row scan ---
// Same recurrence, thread <-> column, rows in sequence.  Inside one row everything above is known, so
//   y_j = p_j + q_j y_(j-1),   p_j = (X - C up - D upleft) / A,   q_j = -B / A,
// a first-order linear recurrence: affine maps compose associatively ((p,q) after (p',q') = (p + q p', q q')), so
// a warp's 32 columns are an inclusive scan of 5 shuffle rounds, and the only thing a warp needs from its left
// neighbour is one number per row, the y of the neighbour's last column (the carry).  Warps therefore run as a
// pipeline, not in lock step: warp w publishes (row, carry) as one 8-byte shared store per row, warp w+1 polls
// for the row tag.  Warp w runs about one hop ahead of warp w+1, a system costs M row-steps plus one hop per
// warp (the wavefront: M + N - 1 steps), and there is no block barrier on the row-to-row dependent chain (the
// first version had one per row plus a serial fold over the warp totals and ran 78 us against this one's 39).
// The carry is also the left neighbour's y across the warp seam, which the next row needs as `upleft`.  Every load and store is row-contiguous (the wavefront's were 32 sectors per request).
//
// Coefficients come through a shared-memory ring filled by 4-byte cp.async (LDGSTS) kScanRing-1 rows ahead; each
// thread reads back only what it copied itself, so cp.async.wait_group is the only synchronisation.  A register
// ring does not work: loads retire through six counting scoreboards per warp, so waiting for the oldest row
// also waits for the newest one, and the kernel ran at one DRAM latency per row (0.9 us) with one row or eight
// rows of lookahead alike.  cp.async groups complete in order and are waited by count.  1/A and the products
// with it are taken one row ahead, off the dependent chain.
//
// Carry slots are reused every kScanSlots rows; a __syncthreads at those rows keeps a fast producer from
// overwriting a slot its consumer has not read (tags are row numbers, so a stale slot never matches).
// Rounding: the scan reassociates the j-chain and multiplies by 1/A instead of dividing (fp32, fma); against
// the f64 oracle the difference is a few ulp of the solution for the diagonally dominant systems this operator
// is built for.  (q products overflow only where the sequential solve has already left fp32 range.)

// 4-byte cp.async of `bytes` (4 or 0) source bytes, the rest zero-filled; src stays a mapped address either way
__device__ __forceinline__ void cp4(uint32_t dst, const float* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

constexpr int kScanRing = 8;     // rows of coefficients in flight per block (power of two)
constexpr int kScanSlots = 64;   // rows of carries between block barriers (power of two)

static size_t scan_smem_bytes(int threads) {
    return (size_t)kScanRing * 5 * threads * sizeof(float) + (size_t)kScanSlots * (threads / 32) * 8;
}

// Hides a loop invariant from the optimiser: left alone, nvcc rematerialises the per-thread bases, predicates and
// strides of trisolve_scan_kernel inside the row loop (about 110 of its 200 instructions per row).
template <class V> __device__ __forceinline__ V keep32(V v) {
    asm volatile("" : "+r"(v));
    return v;
}
template <class V> __device__ __forceinline__ V* keep64(V* p) {
    asm volatile("" : "+l"(p));
    return p;
}

template <bool kUpper>
__global__ void __launch_bounds__(1024)
trisolve_scan_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ C,
                     const float* __restrict__ D, const float* __restrict__ X, float* __restrict__ Y, int M, int N) {
    const size_t s = blockIdx.x;
    const bool has_d = D != nullptr;
    const int lj = threadIdx.x, lane = lj & 31, w = lj >> 5;
    const int T = blockDim.x, nw = T >> 5;
    const bool col = lj < N;

    extern __shared__ __align__(16) unsigned char scan_smem[];
    const uint32_t smem_u32 = (uint32_t)__cvta_generic_to_shared(scan_smem);
    // ring: [kScanRing][T][5] floats (a, x, b, c, d); slots: [kScanSlots][nw] of (row << 32) | carry
    const uint32_t slots_u32 = smem_u32 + kScanRing * 20 * T;
    {
        volatile unsigned long long* sl = reinterpret_cast<volatile unsigned long long*>(scan_smem + kScanRing * 20 * T);
        for (int k = lj; k < kScanSlots * nw; k += T) sl[k] = ~0ull;
    }

    // A warp alone issues one instruction every few cycles, so the row time is the instruction count of the loop
    // below (the first pipelined version spent 265 instructions per row, mostly 64-bit address arithmetic, and
    // was no faster than the barrier version).  Addresses are therefore two 32-bit element offsets per thread
    // (A/X/C rows are N wide, B/D rows N-1), stepped by one row per copy, on bases shifted so that C shares A's
    // offset and D shares B's: lower solves read C and D one row up, (i-1, j) = offset - N, (i-1, j-1) = offset
    // - (N-1); upper solves read them at the cell itself.  Threads outside the grid keep clamped, mapped offsets
    // and copy zero bytes.
    const int jc = col ? (kUpper ? N - 1 - lj : lj) : 0;                     // physical column, clamped
    const bool hb = col && lj > 0 && N > 1;
    const int jb = hb ? (kUpper ? jc : jc - 1) : 0;
    const int stepA = keep32(kUpper ? -N : N), stepB = keep32(kUpper ? -(N - 1) : N - 1);
    int oa = (kUpper ? (M - 1) * N : 0) + jc;                                // offsets of the next row to copy
    int ob = (kUpper ? (M - 1) * (N - 1) : 0) + jb;
    int oy = oa;                                                             // offset of the row being solved
    const float* Ab = keep64(A + s * M * N);
    const float* Xb = keep64(X + s * M * N);
    const float* Bb = keep64(B + s * M * (N - 1));
    const float* Cb = keep64(C + s * (M - 1) * N - (kUpper ? 0 : N));
    const float* Db = keep64(has_d ? D + s * (M - 1) * (N - 1) - (kUpper ? 0 : N - 1) : Ab);
    float* Yb = keep64(Y + s * M * N);
    const int na = keep32(col ? 4 : 0), nb = keep32(hb ? 4 : 0), nd = keep32((hb && has_d) ? 4 : 0);
    const uint32_t ring_me = keep32(smem_u32 + lj * 20), ring_stride = keep32((uint32_t)T * 20);
    const uint32_t slot_me = keep32(slots_u32 + w * 8), slot_stride = keep32((uint32_t)nw * 8);
    const bool publish = lane == 31 && w + 1 < nw;

    auto issue = [&](int li, bool first) {
        if (li < M) {
            const uint32_t dst = ring_me + (li & (kScanRing - 1)) * ring_stride;
            cp4(dst, Ab + oa, na);
            cp4(dst + 4, Xb + oa, na);
            cp4(dst + 8, Bb + ob, nb);
            // row 0 has nothing above it; its shifted C / D addresses would lie before the arrays
            cp4(dst + 12, first ? Ab : Cb + oa, first ? 0 : na);
            cp4(dst + 16, first ? Ab : Db + ob, first ? 0 : nd);
            oa += stepA;
            ob += stepB;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // row li of the ring, scaled by 1/A: xi = X/A, q = -B/A, ci = C/A, di = D/A
    struct Row { float xi, q, ci, di; };
    auto scaled = [&](int li) {
        const uint32_t src = ring_me + (li & (kScanRing - 1)) * ring_stride;
        float a, x, b, c, d;
        asm volatile("ld.shared.f32 %0, [%5];\n\tld.shared.f32 %1, [%5+4];\n\tld.shared.f32 %2, [%5+8];\n\t"
                     "ld.shared.f32 %3, [%5+12];\n\tld.shared.f32 %4, [%5+16];"
                     : "=f"(a), "=f"(x), "=f"(b), "=f"(c), "=f"(d)
                     : "r"(src)
                     : "memory");
        a = na ? a : 1.f;
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(a));
        inv = fmaf(inv, fmaf(-a, inv, 1.f), inv);  // one Newton step: correctly rounded but for rare last-bit ties
        return Row{x * inv, -b * inv, c * inv, d * inv};
    };
    auto peek = [&](uint32_t addr) {
        unsigned long long v;
        asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
        return v;
    };

    issue(0, true);
#pragma unroll
    for (int r = 1; r < kScanRing - 1; ++r) issue(r, false);
    asm volatile("cp.async.wait_group %0;" ::"n"(kScanRing - 2) : "memory");   // row 0 has landed
    Row nxt = scaled(0);
    __syncthreads();                               // slot tags initialised
    float up = 0.f, upleft = 0.f;
    // Measured, 64 systems of 112 x 256: 39 us (0.35 us per row, ~125 instructions) against the wavefront's 219 us.
    // Placing the independent work (next row's 1/A, the copies) by hand between the shuffle rounds was slower
    // (46 us); the order below, independent work first, is what ptxas schedules best.
    for (int li = 0; li < M; ++li) {
        const Row k = nxt;
        float p = fmaf(-k.ci, up, fmaf(-k.di, upleft, k.xi));
        float q = k.q;
        const uint32_t in = slot_me + (li & (kScanSlots - 1)) * slot_stride;
        unsigned long long cv = w > 0 ? peek(in) : 0ull;   // first look at the carry, overlapped with the scan
        issue(li + kScanRing - 1, false);          // into the slot row li-1 was read from
        asm volatile("cp.async.wait_group %0;" ::"n"(kScanRing - 2) : "memory");   // row li+1 has landed
        nxt = scaled(li + 1);                      // independent of this row's chain (past M: an unused old slot)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float pp = __shfl_up_sync(0xffffffffu, p, o), qq = __shfl_up_sync(0xffffffffu, q, o);
            if (lane >= o) {
                p = fmaf(q, pp, p);
                q *= qq;
            }
        }
        // the left neighbour's map, so that its y needs no further shuffle once the carry is known
        const float pl = __shfl_up_sync(0xffffffffu, p, 1), ql = __shfl_up_sync(0xffffffffu, q, 1);
        float carry = 0.f;                         // y of the last column of the warp to the left
        if (w > 0) {
            while ((unsigned)(cv >> 32) != (unsigned)li) cv = peek(in);
            carry = __uint_as_float((unsigned)cv);
        }
        const float y = fmaf(q, carry, p);
        if (publish) {
            const unsigned long long v = ((unsigned long long)(unsigned)li << 32) | __float_as_uint(y);
            asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(in + 8), "l"(v) : "memory");
        }
        if (na) Yb[oy] = y;
        oy += stepA;
        upleft = lane == 0 ? carry : fmaf(ql, carry, pl);
        up = y;
        if ((li & (kScanSlots - 1)) == kScanSlots - 1) __syncthreads();
    }
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
    ARF_REQUIRE(A && (B || N == 1) && (C || M == 1) && X && Y && systems > 0 && systems <= 0x7fffffffLL && M > 0 &&
                N > 0);   // B / C are empty (NULL allowed) for single-column / single-row systems
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 1024 && (long long)M * N < 0x7fffffffLL && g_trisolve_variant != 1) {  // thread <-> column, row scan
        const int threads = (N + 31) / 32 * 32;
        const size_t smem = scan_smem_bytes(threads);   // 44 KB at N = 256, 176 KB at 1024
        auto kern = upper ? trisolve_scan_kernel<true> : trisolve_scan_kernel<false>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        kern<<<(int)systems, threads, smem, st>>>(A, B, C, D, X, Y, M, N);
        ARF_CHECK_LAUNCH();
        return ARF_OK;
    }
    if (M > 1024) return ARF_EUNSUPPORTED;       // thread <-> row wavefront
    const int threads = (M + 31) / 32 * 32;
    const size_t smem = 2 * (threads + 1) * sizeof(float);
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

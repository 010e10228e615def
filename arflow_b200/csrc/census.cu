// Fused census / ternary photometric block.
//
// Reference chain (utils/uflow_utils.py:241-306, losses/loss_blocks.py:12-62): RGB -> gray*255 ->
// identity-kernel conv2d to P*P channels -> diff/sqrt(.81+diff^2) for both images -> squared difference ->
// sq/(.1+sq) -> channel sum (or mean) -> (|h|+.01)^.4 -> * border-zeroed mask -> sum / (sum(mask)+1e-6).
// About twenty full-resolution 49-channel temporaries.  Here: ONE pass.  A CTA stages the two gray
// tiles (with a (P-1)/2 halo, zeros outside the image == the conv's zero padding) in shared memory,
// each thread walks the P*P offsets of its pixel in registers, writes the soft Hamming distance and,
// when a mask is given, block-reduces numerator and denominator of the masked robust mean.
// Algorithmic traffic: 2 RGB images + mask in, one map out = 32 B/px; the work is MUFU-bound
// (2 rsqrt + 1 rcp per offset).
//
// Backward is a gather (no atomics): the gradient of gray pixel q collects, for every offset k, the term
// of pixel q itself (q is the centre) and the term of pixel q-k (q is the neighbour).
#include "common.cuh"
#include <math.h>

ARF_HOOK g_census_variant = 0;   // test hook, per calling thread (arf_debug_set key 5): 1 = force the per-pixel kernels, 8..64 = strip height

namespace {

constexpr int kCTW = 32;   // tile width
constexpr int kCTH = 16;   // tile height (2 pixels per thread, 256 threads)
constexpr int kCThreads = 256;

__device__ __forceinline__ float gray255(const float* __restrict__ im, size_t plane, size_t off) {
    // ((R*0.2989 + G*0.5870) + B*0.1140) * 255, the reference's evaluation order (uflow_utils.py:227-231,252)
    float g = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(im + off), 0.2989f), __fmul_rn(__ldg(im + plane + off), 0.5870f)),
                        __fmul_rn(__ldg(im + 2 * plane + off), 0.1140f));
    return __fmul_rn(g, 255.f);
}

template <int R>
__device__ __forceinline__ void load_gray_tile(float (*tile)[kCTW + 2 * R], const float* __restrict__ im, int b,
                                               int x0, int y0, int H, int W) {
    const size_t plane = (size_t)H * W;
    const float* ib = im + (size_t)b * 3 * plane;
    constexpr int TW = kCTW + 2 * R, THh = kCTH + 2 * R;
    for (int e = threadIdx.x; e < TW * THh; e += kCThreads) {
        int xx = e % TW, yy = e / TW;
        int gx = x0 + xx - R, gy = y0 + yy - R;
        float v = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = gray255(ib, plane, (size_t)gy * W + gx);
        tile[yy][xx] = v;
    }
}

// Raw MUFU.RSQ / MUFU.RCP: every argument here is >= 0.1, so the denormal/zero fix-up code rsqrtf() and
// __fdividef() carry (FSETP/FMUL/FSEL per call, ~40% of the instruction stream in the first ncu capture)
// is dead weight.  Relative error <= 2^-22.
__device__ __forceinline__ float mufu_rsq(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float ctransform(float d, float& r) {
    r = mufu_rsq(fmaf(d, d, 0.81f));
    return d * r;
}

// term of one pixel pair: sq / (0.1 + sq), sq = (ta - tb)^2 - exactly the per-pixel kernel's arithmetic (identical
// patches give exactly 0).  Measured and dropped: (a) summing 1 / (0.1 + sq) and forming h = n - 0.1 * sum per pixel
// (two instructions fewer per pair; no time gained, and identical patches no longer give 0: loss off by 1e-4 there);
// (b) one rsqrt of (0.81 + da^2)(0.81 + db^2) instead of two (2 instead of 3 MUFU per pair: 57 -> 54 us, but the
// cancellation leaves ~2e-7 absolute in sq, 8e-6 relative on the loss of near-identical images - too close to the bar).
__device__ __forceinline__ float pair_term(float da, float db) {
    float r_;
    const float df = ctransform(da, r_) - ctransform(db, r_);
    const float sq = df * df;
    return sq * mufu_rcp(0.1f + sq);
}

// mask value with the patch/2 border zeroed (zero_mask_border, uflow_utils.py:234-238); mask may be NULL (= ones)
__device__ __forceinline__ float border_mask(const float* __restrict__ mask, int b, int y, int x, int H, int W, int R) {
    if (x < R || x >= W - R || y < R || y >= H - R) return 0.f;
    return mask ? __ldg(mask + ((size_t)b * H + y) * W + x) : 1.f;
}

template <int R>
__global__ void __launch_bounds__(kCThreads)
census_fwd_kernel(const float* __restrict__ im_a, const float* __restrict__ im_b, const float* __restrict__ mask,
                  float* __restrict__ hamming, float* __restrict__ partials, int B, int H, int W, int tiles_x,
                  int tiles_y, float scale, int want_sums, float eps, float q) {
    __shared__ float ga[kCTH + 2 * R][kCTW + 2 * R];
    __shared__ float gb[kCTH + 2 * R][kCTW + 2 * R];
    __shared__ float red[32];
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * kCTW, y0 = ty * kCTH;
    load_gray_tile<R>(ga, im_a, b, x0, y0, H, W);
    load_gray_tile<R>(gb, im_b, b, x0, y0, H, W);
    __syncthreads();

    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;  // rows ly0 and ly0+8
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int ly = ly0 + half * 8;
        const int x = x0 + lx, y = y0 + ly;
        const float ca = ga[ly + R][lx + R], cb = gb[ly + R][lx + R];
        float h = 0.f;
#pragma unroll
        for (int dy = 0; dy <= 2 * R; ++dy)
#pragma unroll
            for (int dx = 0; dx <= 2 * R; ++dx) {
                if (dy == R && dx == R) continue;   // the centre's own term is exactly 0
                float ra, rb;
                float ta = ctransform(ga[ly + dy][lx + dx] - ca, ra);
                float tb = ctransform(gb[ly + dy][lx + dx] - cb, rb);
                float df = ta - tb;
                float sq = df * df;
                h = fmaf(sq, mufu_rcp(0.1f + sq), h);
            }
        h *= scale;
        if (x < W && y < H) {
            hamming[((size_t)b * H + y) * W + x] = h;
            if (want_sums) {
                float pm = border_mask(mask, b, y, x, H, W, R);
                num += __powf(fabsf(h) + eps, q) * pm;
                den += pm;
            }
        }
    }
    if (want_sums) {
        float n = arf_block_sum(num, red);
        float d = arf_block_sum(den, red);
        if (threadIdx.x == 0) {
            partials[2 * (size_t)blockIdx.x] = n;
            partials[2 * (size_t)blockIdx.x + 1] = d;
        }
    }
}


// ------------------------------------------------------------------ pair-symmetric kernels -----------------------
// The census term that couples pixel p with its neighbour p+k is the same seen from either end: swapping the roles
// flips the sign of both transformed differences, so (ta - tb)^2 and everything after it is unchanged
// (term(p, k) == term(p+k, -k), also when one end lies in the zero padding).  The kernels below therefore evaluate
// every UNORDERED pair once - the (2R+1)^2/2 offsets k of the half plane dy > 0 or (dy == 0 and dx > 0) - and hand the
// value to both ends: 72 instead of 144 MUFU operations per pixel for the 7x7 patch.
//
// A warp owns a strip of pixels and walks it top to bottom.  Lane l evaluates the SITES of two columns
// (strip column l and 32 + l); the value for the far end of the pair goes to the lane that owns column x + dx by
// shuffle, and waits there in a rolling register accumulator until the walk reaches row y + dy.  No shared-memory
// exchange, no atomics, no block barrier; sums are formed in a fixed order (deterministic).  The outermost R strip
// columns and the R rows above the strip are evaluated only for what they send into the strip (for the 7x7 patch:
// 64 site columns for 58 owned ones, Hs + 3 site rows for Hs owned ones).
constexpr int kSymCols = 64;        // site columns per strip (two per lane)
constexpr int kSymWarps = 4;        // strips per CTA (independent; they share only the final block reduction)

template <int R>
struct SymGeo {
    static constexpr int kOwn = kSymCols - 2 * R;        // owned columns per strip
    static constexpr int kTW = kSymCols + 2 * R;         // staged tile width
};

// strip -> geometry; returns false for a padding strip (warp stays idle)
struct SymStrip { int b, x_site0, y_own0, y_own1; };

template <int R>
__device__ __forceinline__ bool sym_strip(long long strip, long long nstrips, int nsx, int nsy, int Hs, int H, SymStrip& g) {
    if (strip >= nstrips) return false;
    const int sx = (int)(strip % nsx);
    const long long t = strip / nsx;
    const int sy = (int)(t % nsy);
    g.b = (int)(t / nsy);
    g.x_site0 = sx * SymGeo<R>::kOwn - R;
    g.y_own0 = sy * Hs;
    g.y_own1 = min(H, g.y_own0 + Hs);
    return true;
}

// A warp keeps only a RING of kRing staged rows per plane in shared memory (static, 2.2 KB per plane and warp): the walk
// needs rows ys .. ys + R, and the row after them is fetched into registers at the top of an iteration and written to
// the ring at its end, so its global latency hides behind the ~1000-cycle evaluation of a site row and the shared-memory
// footprint no longer grows with the strip height (it capped residency at 8-16 warps per SM).
constexpr int kRing = 8;
static_assert(kRing >= 3 + 2, "ring must hold R + 1 rows in use and the one being written");

// raw RGB of one staged row: columns lane, lane + 32, lane + 64 (< kTW) of the tile, clamped addresses (branch-free)
struct RowFetch { float c[3][3]; unsigned ok; };

template <int R>
__device__ __forceinline__ void sym_fetch_rgb(RowFetch& f, const float* __restrict__ ib, size_t plane, int gy, int x0,
                                              int H, int W, int lane) {
    constexpr int TW = SymGeo<R>::kTW;
    const bool rowok = gy >= 0 && gy < H;
    const size_t rowoff = (size_t)min(max(gy, 0), H - 1) * W;
    f.ok = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int xx = lane + 32 * j, gx = x0 + xx;
        if (rowok && xx < TW && gx >= 0 && gx < W) f.ok |= 1u << j;
        const size_t off = rowoff + min(max(gx, 0), W - 1);
        f.c[j][0] = __ldg(ib + off);
        f.c[j][1] = __ldg(ib + plane + off);
        f.c[j][2] = __ldg(ib + 2 * plane + off);
    }
}

// gray255 of a fetched row -> ring row (zeros outside the image == the reference conv's zero padding)
template <int R>
__device__ __forceinline__ void sym_store_gray(float* row, const RowFetch& f, int lane) {
    constexpr int TW = SymGeo<R>::kTW;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        // ((R*0.2989 + G*0.5870) + B*0.1140) * 255 in the reference's evaluation order (gray255)
        const float gv = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(f.c[j][0], 0.2989f), __fmul_rn(f.c[j][1], 0.5870f)),
                                             __fmul_rn(f.c[j][2], 0.1140f)), 255.f);
        const int xx = lane + 32 * j;
        if (xx < TW) row[xx] = ((f.ok >> j) & 1u) ? gv : 0.f;
    }
}

// Deliver the pair values t0 (column l) and t1 (column 32 + l) to the owners of column + DX.  Lanes whose source falls
// outside the strip receive a meaningless value; those are exactly the halo columns whose results are never written.
template <int DX>
__device__ __forceinline__ void sym_send(float t0, float t1, int lane, float& r0, float& r1) {
    if (DX == 0) { r0 = t0; r1 = t1; return; }
    const int src = (lane - DX) & 31;
    const float a = __shfl_sync(0xffffffffu, t0, src);
    const float b = __shfl_sync(0xffffffffu, t1, src);
    if (DX > 0) { r0 = a; r1 = lane >= DX ? b : a; }
    else        { r0 = lane - DX < 32 ? a : b; r1 = b; }
}

// dx is a compile-time constant after unrolling; the switch folds away
__device__ __forceinline__ void sym_send_dx(int dx, float t0, float t1, int lane, float& r0, float& r1) {
    switch (dx) {
        case -3: sym_send<-3>(t0, t1, lane, r0, r1); break;
        case -2: sym_send<-2>(t0, t1, lane, r0, r1); break;
        case -1: sym_send<-1>(t0, t1, lane, r0, r1); break;
        case 0:  sym_send<0>(t0, t1, lane, r0, r1); break;
        case 1:  sym_send<1>(t0, t1, lane, r0, r1); break;
        case 2:  sym_send<2>(t0, t1, lane, r0, r1); break;
        default: sym_send<3>(t0, t1, lane, r0, r1); break;
    }
}

template <int R>
__global__ void __launch_bounds__(32 * kSymWarps)
census_fwd_sym(const float* __restrict__ im_a, const float* __restrict__ im_b, const float* __restrict__ mask,
               float* __restrict__ hamming, float* __restrict__ partials, int B, int H, int W, int nsx, int nsy, int Hs,
               long long nstrips, float scale, int want_sums, float eps, float q) {
    constexpr int TW = SymGeo<R>::kTW;
    __shared__ float ring[kSymWarps][2][kRing][TW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float (*ring_a)[TW] = ring[warp][0];
    float (*ring_b)[TW] = ring[warp][1];
    float num = 0.f, den = 0.f;
    SymStrip g;
    if (sym_strip<R>((long long)blockIdx.x * kSymWarps + warp, nstrips, nsx, nsy, Hs, H, g)) {
        const int ys0 = g.y_own0 - R;                       // first site row
        const int nsite = g.y_own1 - ys0;                   // site rows; staged rows = nsite + R
        const size_t plane = (size_t)H * W;
        const float* ia = im_a + (size_t)g.b * 3 * plane;
        const float* ib = im_b + (size_t)g.b * 3 * plane;
        const int xt0 = g.x_site0 - R;                      // image x of tile column 0
        {   // rows 0 .. R of the tile
            RowFetch fa[R + 1], fb[R + 1];
#pragma unroll
            for (int j = 0; j <= R; ++j) {
                sym_fetch_rgb<R>(fa[j], ia, plane, ys0 + j, xt0, H, W, lane);
                sym_fetch_rgb<R>(fb[j], ib, plane, ys0 + j, xt0, H, W, lane);
            }
#pragma unroll
            for (int j = 0; j <= R; ++j) {
                sym_store_gray<R>(ring_a[j], fa[j], lane);
                sym_store_gray<R>(ring_b[j], fb[j], lane);
            }
        }
        __syncwarp();
        float acc0[R + 1], acc1[R + 1];
#pragma unroll
        for (int i = 0; i <= R; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
        const int x0 = g.x_site0 + lane, x1 = x0 + 32;
        const bool own0 = lane >= R && x0 < W, own1 = lane < 32 - R && x1 < W;   // x0 >= 0 for owned columns
        // mask with the patch/2 border zeroed (zero_mask_border); column part of the test is per lane
        const bool in0 = own0 && x0 >= R && x0 < W - R, in1 = own1 && x1 >= R && x1 < W - R;
        for (int i = 0; i < nsite; ++i) {
            const int ys = ys0 + i;
            const bool have_next = i + R + 1 < nsite + R;
            RowFetch fa, fb;
            if (have_next) {
                sym_fetch_rgb<R>(fa, ia, plane, ys + R + 1, xt0, H, W, lane);
                sym_fetch_rgb<R>(fb, ib, plane, ys + R + 1, xt0, H, W, lane);
            }
            float pm0 = 0.f, pm1 = 0.f;
            if (want_sums && ys >= g.y_own0 && ys >= R && ys < H - R) {
                const float* mrow = mask ? mask + ((size_t)g.b * H + ys) * W : nullptr;
                pm0 = in0 ? (mrow ? __ldg(mrow + x0) : 1.f) : 0.f;
                pm1 = in1 ? (mrow ? __ldg(mrow + x1) : 1.f) : 0.f;
            }
            const float ca0 = ring_a[i & (kRing - 1)][R + lane], cb0 = ring_b[i & (kRing - 1)][R + lane];
            const float ca1 = ring_a[i & (kRing - 1)][R + lane + 32], cb1 = ring_b[i & (kRing - 1)][R + lane + 32];
            // Per dy: first every pair value of the row of offsets (independent chains of LDS / FMA / MUFU the scheduler
            // can interleave), then the shuffles - a shuffle is a scheduling fence, one per pair serialised the chains.
#pragma unroll
            for (int dy = 0; dy <= R; ++dy) {
                const float* ra = &ring_a[(i + dy) & (kRing - 1)][R + lane];
                const float* rb = &ring_b[(i + dy) & (kRing - 1)][R + lane];
                float t0[2 * R + 1], t1[2 * R + 1];
#pragma unroll
                for (int dx = -R; dx <= R; ++dx) {
                    if (dy == 0 && dx <= 0) continue;
                    t0[dx + R] = pair_term(ra[dx] - ca0, rb[dx] - cb0);
                    t1[dx + R] = pair_term(ra[dx + 32] - ca1, rb[dx + 32] - cb1);
                    acc0[0] += t0[dx + R];
                    acc1[0] += t1[dx + R];
                }
#pragma unroll
                for (int dx = -R; dx <= R; ++dx) {
                    if (dy == 0 && dx <= 0) continue;
                    float r0, r1;
                    sym_send_dx(dx, t0[dx + R], t1[dx + R], lane, r0, r1);
                    acc0[dy] += r0;
                    acc1[dy] += r1;
                }
            }
            // row ys is complete: every pair that touches it has been evaluated
            if (ys >= g.y_own0) {
                const float h0 = acc0[0] * scale, h1 = acc1[0] * scale;
                const size_t o = ((size_t)g.b * H + ys) * W;
                if (own0) hamming[o + x0] = h0;
                if (own1) hamming[o + x1] = h1;
                if (want_sums) {
                    num = fmaf(__powf(fabsf(h0) + eps, q), pm0, num);
                    num = fmaf(__powf(fabsf(h1) + eps, q), pm1, num);
                    den += pm0 + pm1;
                }
            }
#pragma unroll
            for (int k = 0; k < R; ++k) { acc0[k] = acc0[k + 1]; acc1[k] = acc1[k + 1]; }
            acc0[R] = 0.f;
            acc1[R] = 0.f;
            if (have_next) {
                sym_store_gray<R>(ring_a[(i + R + 1) & (kRing - 1)], fa, lane);
                sym_store_gray<R>(ring_b[(i + R + 1) & (kRing - 1)], fb, lane);
            }
            __syncwarp();
        }
    }
    if (want_sums) {
        // one partial pair per STRIP (strips are ordered by batch item, so a group of the batch is a contiguous range)
        num = arf_warp_sum(num);
        den = arf_warp_sum(den);
        const long long strip = (long long)blockIdx.x * kSymWarps + warp;
        if (lane == 0 && strip < nstrips) {
            partials[2 * (size_t)strip] = num;
            partials[2 * (size_t)strip + 1] = den;
        }
    }
}

// Block g: out[3g + {0,1,2}] = num, den, num / (den + 1e-6) of group g = partial entries [g * n, (g + 1) * n) (the partials
// are ordered by batch item, a group is a contiguous range of the batch); fixed summation order (deterministic)
__global__ void census_finalize_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
    __shared__ double sn[256], sd[256];
    const float* p = partials + 2 * (size_t)blockIdx.x * n;
    double a = 0.0, d = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a += (double)p[2 * (size_t)i];
        d += (double)p[2 * (size_t)i + 1];
    }
    sn[threadIdx.x] = a;
    sd[threadIdx.x] = d;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sn[threadIdx.x] += sn[threadIdx.x + s];
            sd[threadIdx.x] += sd[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float num = (float)sn[0], den = (float)sd[0];
        out[3 * blockIdx.x + 0] = num;
        out[3 * blockIdx.x + 1] = den;
        out[3 * blockIdx.x + 2] = num / (den + 1e-6f);
    }
}

// Gradient of the two census terms that couple pixel q and its neighbour q+k, in one evaluation.
// Term "q is the centre, neighbour q+k" has diff d = I[q+k]-I[q]; term "q is the neighbour of centre q+k" has
// diff -d.  Both share rsqrt(0.81+d^2) and 1/(0.1+df^2) (ta, tb, df only flip sign), so
//   dL/dI_a[q] += -(gh[q] + gh[q+k]) * df * inv^2 * ra^3          (and the mirror image for I_b)
// which halves the MUFU and FMA work of evaluating the two terms separately.
// gsum = gh[q] + gh[q+k]; gh already carries the constant factor 0.2 * 0.81 (applied when the tile is staged).
template <bool kA, bool kB>
__device__ __forceinline__ void pair_grads(float da, float db, float gsum, float& acc_a, float& acc_b) {
    float ra, rb;
    float ta = ctransform(da, ra);
    float tb = ctransform(db, rb);
    float df = ta - tb;
    float inv = mufu_rcp(fmaf(df, df, 0.1f));
    float common = gsum * df * inv * inv;               // d/d(ta) [sq/(0.1+sq)] = 2*df*0.1/(0.1+sq)^2
    if (kA) acc_a = fmaf(-common, ra * ra * ra, acc_a);  // d(ta)/d(da) = 0.81/(0.81+da^2)^1.5
    if (kB) acc_b = fmaf(common, rb * rb * rb, acc_b);
}

template <int R, bool kA, bool kB>
__global__ void __launch_bounds__(kCThreads, 6)
census_bwd_kernel(const float* __restrict__ im_a, const float* __restrict__ im_b, const float* __restrict__ ghamming,
                  const float* __restrict__ hamming, const float* __restrict__ mask, const float* __restrict__ sums,
                  const float* __restrict__ gloss, float* __restrict__ g_a, float* __restrict__ g_b, int B, int H,
                  int W, int tiles_x, int tiles_y, float scale, float eps, float q, int bg) {
    __shared__ float ga[kCTH + 2 * R][kCTW + 2 * R];
    __shared__ float gb[kCTH + 2 * R][kCTW + 2 * R];
    __shared__ float gh[kCTH + 2 * R][kCTW + 2 * R];   // upstream d(loss)/d(hamming), zero outside the image
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * kCTW, y0 = ty * kCTH;
    load_gray_tile<R>(ga, im_a, b, x0, y0, H, W);
    load_gray_tile<R>(gb, im_b, b, x0, y0, H, W);
    {
        float gl = 0.f, idn = 0.f;
        if (!ghamming) {      // bg batch items per group: its own upstream gradient and normaliser
            gl = __ldg(gloss + b / bg);
            idn = 1.f / (__ldg(sums + 3 * (b / bg) + 1) + 1e-6f);
        }
        constexpr int TW = kCTW + 2 * R, THh = kCTH + 2 * R;
        for (int e = threadIdx.x; e < TW * THh; e += kCThreads) {
            int xx = e % TW, yy = e / TW;
            int gx = x0 + xx - R, gy = y0 + yy - R;
            float v = 0.f;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
                size_t o = ((size_t)b * H + gy) * W + gx;
                if (ghamming) {
                    v = __ldg(ghamming + o);
                } else {
                    // loss = sum(pow(|h|+eps, q) * pm) / (sum(pm) + 1e-6)
                    float pm = border_mask(mask, b, gy, gx, H, W, R);
                    float h = __ldg(hamming + o);
                    float s = h > 0.f ? 1.f : (h < 0.f ? -1.f : 0.f);
                    v = gl * pm * idn * q * __powf(fabsf(h) + eps, q - 1.f) * s;
                }
                v *= scale * (0.2f * 0.81f);
            }
            gh[yy][xx] = v;
        }
    }
    __syncthreads();

    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
    const size_t plane = (size_t)H * W;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int ly = ly0 + half * 8;
        const int x = x0 + lx, y = y0 + ly;
        const int cy = ly + R, cx = lx + R;
        const float ca = ga[cy][cx], cb = gb[cy][cx], ghc = gh[cy][cx];
        float acc_a = 0.f, acc_b = 0.f;
#pragma unroll 1
        for (int dy = -R; dy <= R; ++dy)
#pragma unroll
            for (int dx = -R; dx <= R; ++dx) {
                if (dy == 0 && dx == 0) continue;
                // gh is zero off-image, so a neighbour outside the image only contributes this pixel's own term
                pair_grads<kA, kB>(ga[cy + dy][cx + dx] - ca, gb[cy + dy][cx + dx] - cb, ghc + gh[cy + dy][cx + dx],
                                   acc_a, acc_b);
            }
        if (x < W && y < H) {
            size_t o = (size_t)b * 3 * plane + (size_t)y * W + x;
            if (g_a) {
                float v = acc_a * 255.f;
                g_a[o] = v * 0.2989f; g_a[o + plane] = v * 0.5870f; g_a[o + 2 * plane] = v * 0.1140f;
            }
            if (g_b) {
                float v = acc_b * 255.f;
                g_b[o] = v * 0.2989f; g_b[o + plane] = v * 0.5870f; g_b[o + 2 * plane] = v * 0.1140f;
            }
        }
    }
}


// Backward of the same pairing: L contains (gh[p] + gh[p+k]) * term(p, k) once per unordered pair, and
// d term / d I_a[p] = - d term / d I_a[p+k], so one evaluation yields the contribution X to one end and -X to the
// other (the per-pixel kernel above evaluates the pair twice, once from each end).
template <int R, bool kA, bool kB>
__global__ void __launch_bounds__(32 * kSymWarps)
census_bwd_sym(const float* __restrict__ im_a, const float* __restrict__ im_b, const float* __restrict__ ghamming,
               const float* __restrict__ hamming, const float* __restrict__ mask, const float* __restrict__ sums,
               const float* __restrict__ gloss, float* __restrict__ g_a, float* __restrict__ g_b, int B, int H, int W,
               int nsx, int nsy, int Hs, long long nstrips, float scale, float eps, float q, int bg) {
    constexpr int TW = SymGeo<R>::kTW;
    __shared__ float ring[kSymWarps][3][kRing][TW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float (*ring_a)[TW] = ring[warp][0];
    float (*ring_b)[TW] = ring[warp][1];
    float (*ring_g)[TW] = ring[warp][2];    // upstream d(loss)/d(hamming) * 0.2 * 0.81 * scale, zero outside the image
    SymStrip g;
    if (!sym_strip<R>((long long)blockIdx.x * kSymWarps + warp, nstrips, nsx, nsy, Hs, H, g)) return;
    const int ys0 = g.y_own0 - R;
    const int nsite = g.y_own1 - ys0;
    const size_t plane = (size_t)H * W;
    const float* ia = im_a + (size_t)g.b * 3 * plane;
    const float* ib = im_b + (size_t)g.b * 3 * plane;
    const int xt0 = g.x_site0 - R;
    float gl = 0.f, idn = 0.f;
    if (!ghamming) {          // bg batch items per group: its own upstream gradient and normaliser
        gl = __ldg(gloss + g.b / bg);
        idn = 1.f / (__ldg(sums + 3 * (g.b / bg) + 1) + 1e-6f);
    }
    // upstream gradient of one staged row: raw loads (clamped), finished by gh_store
    struct GhFetch { float h[3], m[3]; unsigned ok; };
    auto gh_fetch = [&](GhFetch& f, int gy) {
        const bool rowok = gy >= 0 && gy < H;
        const size_t rowoff = ((size_t)g.b * H + min(max(gy, 0), H - 1)) * W;
        const bool rowin = gy >= R && gy < H - R;
        f.ok = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int xx = lane + 32 * j, gx = xt0 + xx;
            const bool ok = rowok && xx < TW && gx >= 0 && gx < W;
            if (ok) f.ok |= 1u << j;
            const size_t off = rowoff + min(max(gx, 0), W - 1);
            if (ghamming) {
                f.h[j] = __ldg(ghamming + off);
                f.m[j] = 0.f;
            } else {
                f.h[j] = __ldg(hamming + off);
                const bool in = rowin && gx >= R && gx < W - R;       // zero_mask_border
                f.m[j] = in ? (mask ? __ldg(mask + off) : 1.f) : 0.f;
            }
        }
    };
    auto gh_store = [&](float* row, const GhFetch& f) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float v = 0.f;
            if ((f.ok >> j) & 1u) {
                if (ghamming) {
                    v = f.h[j];
                } else {
                    // loss = sum(pow(|h|+eps, q) * pm) / (sum(pm) + 1e-6)
                    const float h = f.h[j];
                    const float sg = h > 0.f ? 1.f : (h < 0.f ? -1.f : 0.f);
                    v = gl * f.m[j] * idn * q * __powf(fabsf(h) + eps, q - 1.f) * sg;
                }
                v *= scale * (0.2f * 0.81f);
            }
            const int xx = lane + 32 * j;
            if (xx < TW) row[xx] = v;
        }
    };
#pragma unroll
    for (int j = 0; j <= R; ++j) {       // rows 0 .. R of the tile
        RowFetch fa, fb;
        GhFetch fg;
        sym_fetch_rgb<R>(fa, ia, plane, ys0 + j, xt0, H, W, lane);
        sym_fetch_rgb<R>(fb, ib, plane, ys0 + j, xt0, H, W, lane);
        gh_fetch(fg, ys0 + j);
        sym_store_gray<R>(ring_a[j], fa, lane);
        sym_store_gray<R>(ring_b[j], fb, lane);
        gh_store(ring_g[j], fg);
    }
    __syncwarp();
    float aa0[R + 1], aa1[R + 1], ab0[R + 1], ab1[R + 1];
#pragma unroll
    for (int i = 0; i <= R; ++i) { aa0[i] = aa1[i] = ab0[i] = ab1[i] = 0.f; }
    const int x0 = g.x_site0 + lane, x1 = x0 + 32;
    const bool own0 = lane >= R && x0 < W, own1 = lane < 32 - R && x1 < W;
    for (int i = 0; i < nsite; ++i) {
        const int ys = ys0 + i;
        const bool have_next = i + R + 1 < nsite + R;
        RowFetch fa, fb;
        GhFetch fg;
        if (have_next) {
            sym_fetch_rgb<R>(fa, ia, plane, ys + R + 1, xt0, H, W, lane);
            sym_fetch_rgb<R>(fb, ib, plane, ys + R + 1, xt0, H, W, lane);
            gh_fetch(fg, ys + R + 1);
        }
        const int s0 = i & (kRing - 1);
        const float ca0 = ring_a[s0][R + lane], cb0 = ring_b[s0][R + lane], cg0 = ring_g[s0][R + lane];
        const float ca1 = ring_a[s0][R + lane + 32], cb1 = ring_b[s0][R + lane + 32], cg1 = ring_g[s0][R + lane + 32];
#pragma unroll
        for (int dy = 0; dy <= R; ++dy) {
            const float* ra = &ring_a[(i + dy) & (kRing - 1)][R + lane];
            const float* rb = &ring_b[(i + dy) & (kRing - 1)][R + lane];
            const float* rg = &ring_g[(i + dy) & (kRing - 1)][R + lane];
            float xa0[2 * R + 1], xa1[2 * R + 1], xb0[2 * R + 1], xb1[2 * R + 1];
#pragma unroll
            for (int dx = -R; dx <= R; ++dx) {
                if (dy == 0 && dx <= 0) continue;
                const int k = dx + R;
                xa0[k] = xb0[k] = xa1[k] = xb1[k] = 0.f;
                pair_grads<kA, kB>(ra[dx] - ca0, rb[dx] - cb0, cg0 + rg[dx], xa0[k], xb0[k]);            // xa = -X_a, xb = +X_b
                pair_grads<kA, kB>(ra[dx + 32] - ca1, rb[dx + 32] - cb1, cg1 + rg[dx + 32], xa1[k], xb1[k]);
                if (kA) { aa0[0] += xa0[k]; aa1[0] += xa1[k]; }
                if (kB) { ab0[0] += xb0[k]; ab1[0] += xb1[k]; }
            }
#pragma unroll
            for (int dx = -R; dx <= R; ++dx) {
                if (dy == 0 && dx <= 0) continue;
                const int k = dx + R;
                float r0, r1;
                if (kA) {
                    sym_send_dx(dx, xa0[k], xa1[k], lane, r0, r1);
                    aa0[dy] -= r0; aa1[dy] -= r1;
                }
                if (kB) {
                    sym_send_dx(dx, xb0[k], xb1[k], lane, r0, r1);
                    ab0[dy] -= r0; ab1[dy] -= r1;
                }
            }
        }
        if (ys >= g.y_own0) {
            const size_t o = (size_t)g.b * 3 * plane + (size_t)ys * W;
            if (kA && g_a) {
                if (own0) { const float v = aa0[0] * 255.f; g_a[o + x0] = v * 0.2989f; g_a[o + plane + x0] = v * 0.5870f; g_a[o + 2 * plane + x0] = v * 0.1140f; }
                if (own1) { const float v = aa1[0] * 255.f; g_a[o + x1] = v * 0.2989f; g_a[o + plane + x1] = v * 0.5870f; g_a[o + 2 * plane + x1] = v * 0.1140f; }
            }
            if (kB && g_b) {
                if (own0) { const float v = ab0[0] * 255.f; g_b[o + x0] = v * 0.2989f; g_b[o + plane + x0] = v * 0.5870f; g_b[o + 2 * plane + x0] = v * 0.1140f; }
                if (own1) { const float v = ab1[0] * 255.f; g_b[o + x1] = v * 0.2989f; g_b[o + plane + x1] = v * 0.5870f; g_b[o + 2 * plane + x1] = v * 0.1140f; }
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) { aa0[k] = aa0[k + 1]; aa1[k] = aa1[k + 1]; ab0[k] = ab0[k + 1]; ab1[k] = ab1[k + 1]; }
        aa0[R] = aa1[R] = ab0[R] = ab1[R] = 0.f;
        if (have_next) {
            const int sn = (i + R + 1) & (kRing - 1);
            sym_store_gray<R>(ring_a[sn], fa, lane);
            sym_store_gray<R>(ring_b[sn], fb, lane);
            gh_store(ring_g[sn], fg);
        }
        __syncwarp();
    }
}

// Strip height: a strip evaluates Hs + R site rows for Hs owned ones, so tall strips waste less, but every strip is one
// warp.  Measured on B200 (8x3x384x512 and 16x3x320x1024, Hs 8..64): an SM retires about 4.2 * (1 - exp(-w/4)) site
// rows per microsecond with w resident warps, and a launch that fits in one wave lasts as long as its busiest SM.
// Pick the height that minimises that estimate.
inline int sym_strip_height(int B, int H, int W, int own_cols, int R, int resident_ctas) {
    if (g_census_variant >= 8 && g_census_variant <= 64) return g_census_variant;
    const long long nsx = arf_cdiv(W, own_cols), sms = ARF_NUM_SMS;
    int best = 16;
    double best_t = -1.0;
    for (int hs = 8; hs <= 64; hs += 4) {
        const long long nblk = (nsx * arf_cdiv(H, hs) * B + kSymWarps - 1) / kSymWarps;
        double rows, w;
        if (nblk <= sms * resident_ctas) {
            const long long q = (nblk + sms - 1) / sms;
            rows = (double)q * kSymWarps * (hs + R);
            w = (double)q * kSymWarps;
        } else {
            rows = (double)nblk * kSymWarps * (hs + R) / (double)sms;
            w = (double)resident_ctas * kSymWarps;
        }
        const double t = rows / (1.0 - exp(-w / 4.0));
        if (best_t < 0 || t < best_t) { best = hs; best_t = t; }
    }
    return best;
}

}  // namespace

extern "C" int arf_census_num_partials(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return ARF_EINVAL;
    // upper bound over both kernel families: per-pixel tiles, and the pair-symmetric strips at their smallest
    long long n = (long long)arf_cdiv(W, kCTW) * arf_cdiv(H, kCTH) * B;
    long long m = (long long)arf_cdiv(W, kSymCols - 6) * arf_cdiv(H, 8) * B;
    if (m > n) n = m;
    return n > 0x7fffffffLL ? ARF_EINVAL : (int)n;
}

namespace {
// pair-symmetric kernels pay off once the image fills the machine with strips
inline bool use_sym(int B, int H, int W) {
    if (g_census_variant == 1) return false;
    if (g_census_variant >= 8) return true;
    return (long long)B * H * W >= 65536 && W >= 32 && H >= 16;
}
template <int R>
int launch_fwd_sym(const float* im_a, const float* im_b, const float* mask, float* hamming, float* partials, float* sums,
                   int B, int H, int W, int groups, float scale, float eps, float q, cudaStream_t st) {
    const int hs = sym_strip_height(B, H, W, SymGeo<R>::kOwn, R, 4);
    const int nsx = arf_cdiv(W, SymGeo<R>::kOwn), nsy = arf_cdiv(H, hs);
    const long long nstrips = (long long)nsx * nsy * B;
    const long long nblk = (nstrips + kSymWarps - 1) / kSymWarps;
    if (nblk > 0x7fffffffLL) return ARF_EINVAL;
    census_fwd_sym<R><<<(unsigned)nblk, 32 * kSymWarps, 0, st>>>(im_a, im_b, mask, hamming, partials, B, H, W, nsx, nsy, hs,
                                                                  nstrips, scale, sums != nullptr, eps, q);
    ARF_CHECK_LAUNCH();
    if (sums) {
        if (nstrips / groups > 0x7fffffffLL) return ARF_EINVAL;
        census_finalize_kernel<<<groups, 256, 0, st>>>(partials, (int)(nstrips / groups), sums);
        ARF_CHECK_LAUNCH();
    }
    return ARF_OK;
}
template <int R>
int launch_bwd_sym(const float* im_a, const float* im_b, const float* ghamming, const float* hamming, const float* mask,
                   const float* sums, const float* gloss, float* g_a, float* g_b, int B, int H, int W, int groups,
                   float scale, float eps, float q, cudaStream_t st) {
    const int hs = sym_strip_height(B, H, W, SymGeo<R>::kOwn, R, 4);
    const int nsx = arf_cdiv(W, SymGeo<R>::kOwn), nsy = arf_cdiv(H, hs);
    const long long nstrips = (long long)nsx * nsy * B;
    const long long nblk = (nstrips + kSymWarps - 1) / kSymWarps;
    if (nblk > 0x7fffffffLL) return ARF_EINVAL;
#define ARF_BWD_SYM(KA, KB)                                                                                            \
    do {                                                                                                               \
        census_bwd_sym<R, KA, KB><<<(unsigned)nblk, 32 * kSymWarps, 0, st>>>(im_a, im_b, ghamming, hamming, mask,    \
                                                                               sums, gloss, g_a, g_b, B, H, W, nsx,   \
                                                                               nsy, hs, nstrips, scale, eps, q,        \
                                                                               B / groups);                            \
    } while (0)
    if (g_a && g_b) ARF_BWD_SYM(true, true);
    else if (g_b) ARF_BWD_SYM(false, true);
    else ARF_BWD_SYM(true, false);
#undef ARF_BWD_SYM
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}
}  // namespace

extern "C" int arf_census_fwd_groups(const float* im_a, const float* im_b, const float* mask, float* hamming,
                                     float* partials, float* sums, int B, int H, int W, int groups, int patch,
                                     float scale, float eps, float q, void* stream) {
    ARF_REQUIRE(im_a && im_b && hamming);
    ARF_REQUIRE(B > 0 && H > 0 && W > 0 && patch >= 1 && (patch & 1) && groups >= 1 && B % groups == 0);
    const int want = sums != nullptr;
    if (want) ARF_REQUIRE(partials != nullptr);
    const int tiles_x = arf_cdiv(W, kCTW), tiles_y = arf_cdiv(H, kCTH);
    if (arf_census_num_partials(B, H, W) < 0) return ARF_EINVAL;
    const int n = tiles_x * tiles_y * B;      // per-pixel kernels: one CTA (and one partial pair) per tile
    cudaStream_t st = (cudaStream_t)stream;
    if (use_sym(B, H, W)) {
        switch (patch / 2) {
            case 1: return launch_fwd_sym<1>(im_a, im_b, mask, hamming, partials, sums, B, H, W, groups, scale, eps, q, st);
            case 2: return launch_fwd_sym<2>(im_a, im_b, mask, hamming, partials, sums, B, H, W, groups, scale, eps, q, st);
            case 3: return launch_fwd_sym<3>(im_a, im_b, mask, hamming, partials, sums, B, H, W, groups, scale, eps, q, st);
            default: return ARF_EUNSUPPORTED;
        }
    }
    switch (patch / 2) {
        case 1: census_fwd_kernel<1><<<n, kCThreads, 0, st>>>(im_a, im_b, mask, hamming, partials, B, H, W, tiles_x, tiles_y, scale, want, eps, q); break;
        case 2: census_fwd_kernel<2><<<n, kCThreads, 0, st>>>(im_a, im_b, mask, hamming, partials, B, H, W, tiles_x, tiles_y, scale, want, eps, q); break;
        case 3: census_fwd_kernel<3><<<n, kCThreads, 0, st>>>(im_a, im_b, mask, hamming, partials, B, H, W, tiles_x, tiles_y, scale, want, eps, q); break;
        default: return ARF_EUNSUPPORTED;
    }
    ARF_CHECK_LAUNCH();
    if (want) {
        census_finalize_kernel<<<groups, 256, 0, st>>>(partials, tiles_x * tiles_y * (B / groups), sums);
        ARF_CHECK_LAUNCH();
    }
    return ARF_OK;
}

extern "C" int arf_census_fwd(const float* im_a, const float* im_b, const float* mask, float* hamming,
                              float* partials, float* sums, int B, int H, int W, int patch, float scale,
                              float eps, float q, void* stream) {
    return arf_census_fwd_groups(im_a, im_b, mask, hamming, partials, sums, B, H, W, 1, patch, scale, eps, q, stream);
}

extern "C" int arf_census_bwd_groups(const float* im_a, const float* im_b, const float* ghamming, const float* hamming,
                                     const float* mask, const float* sums, const float* gloss, float* g_a, float* g_b,
                                     int B, int H, int W, int groups, int patch, float scale, float eps, float q,
                                     void* stream) {
    ARF_REQUIRE(im_a && im_b);
    ARF_REQUIRE(ghamming || (hamming && sums && gloss));
    ARF_REQUIRE(B > 0 && H > 0 && W > 0 && patch >= 1 && (patch & 1) && groups >= 1 && B % groups == 0);
    if (!g_a && !g_b) return ARF_OK;
    const int tiles_x = arf_cdiv(W, kCTW), tiles_y = arf_cdiv(H, kCTH);
    if (arf_census_num_partials(B, H, W) < 0) return ARF_EINVAL;
    const int n = tiles_x * tiles_y * B;
    const int bg = B / groups;
    cudaStream_t st = (cudaStream_t)stream;
    if (use_sym(B, H, W)) {
        switch (patch / 2) {
            case 1: return launch_bwd_sym<1>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, groups, scale, eps, q, st);
            case 2: return launch_bwd_sym<2>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, groups, scale, eps, q, st);
            case 3: return launch_bwd_sym<3>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, groups, scale, eps, q, st);
            default: return ARF_EUNSUPPORTED;
        }
    }
    switch (patch / 2) {
        case 1:
            if (g_a && g_b) census_bwd_kernel<1, true, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            else if (g_b) census_bwd_kernel<1, false, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            else census_bwd_kernel<1, true, false><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            break;
        case 2:
            if (g_a && g_b) census_bwd_kernel<2, true, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            else if (g_b) census_bwd_kernel<2, false, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            else census_bwd_kernel<2, true, false><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            break;
        case 3:
            if (g_a && g_b) census_bwd_kernel<3, true, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            else if (g_b) census_bwd_kernel<3, false, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            else census_bwd_kernel<3, true, false><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q, bg);
            break;
        default: return ARF_EUNSUPPORTED;
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_census_bwd(const float* im_a, const float* im_b, const float* ghamming, const float* hamming,
                              const float* mask, const float* sums, const float* gloss, float* g_a, float* g_b,
                              int B, int H, int W, int patch, float scale, float eps, float q, void* stream) {
    return arf_census_bwd_groups(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, 1, patch, scale, eps,
                                 q, stream);
}

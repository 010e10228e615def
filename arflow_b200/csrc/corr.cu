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
// Tiled forward, "pair-shared" layout (corr_fwd_md4_p2, the default wherever it fills the machine): 32 x 12 tiles, a
//   warp covers a PAIR of horizontal displacements and the two lanes of a lane pair read the same 8 bytes of the f2
//   halo row - one LDS.64 slot delivers two operands, and the register pair is the packed operand of FFMA2.
#include "common.cuh"
#include "tma.cuh"
#include <string.h>

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

// ------------------------------------------------------------------ small problems, md = 4 geometry --------
// The coarsest pyramid levels (PWC-Lite inference: 1 x 192 x 6 x 10, pwclite.py:113) have a few dozen pixels and a few
// hundred channels: one thread per output element (the literal kernels) walks 81 x 2 or C x 2 dependent loads and takes
// 13-33 us forward / ~100 us backward.  Here a WARP owns the output element and its lanes split the reduction
// (channels forward, the 81 displacements backward), then a shuffle tree: a few microseconds.
__global__ void __launch_bounds__(256)
corr_fwd_small_md4(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out, int B, int C,
                   int H, int W) {
    const int lane = threadIdx.x & 31;
    const long long nout = (long long)B * 81 * H * W;
    const size_t plane = (size_t)H * W;
    const float inv_c = 1.0f / (float)C;
    for (long long o = blockIdx.x * 8LL + (threadIdx.x >> 5); o < nout; o += gridDim.x * 8LL) {
        const int x = (int)(o % W);
        long long t = o / W;
        const int y = (int)(t % H); t /= H;
        const int tc = (int)(t % 81), b = (int)(t / 81);
        const int y2 = y + tc / 9 - 4, x2 = x + tc % 9 - 4;
        float acc = 0.f;
        if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
            const float* p1 = f1 + (size_t)b * C * plane + (size_t)y * W + x;
            const float* p2 = f2 + (size_t)b * C * plane + (size_t)y2 * W + x2;
            for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(p1 + c * plane), __ldg(p2 + c * plane), acc);
        }
        acc = arf_warp_sum(acc);
        if (lane == 0) out[o] = acc * inv_c;
    }
}

// which (blockIdx.y + y_base) = 0: gradient w.r.t. f1, 1: gradient w.r.t. f2 (formulas at the tiled backward below)
__global__ void __launch_bounds__(256)
corr_bwd_small_md4(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ gout,
                   float* __restrict__ g1, float* __restrict__ g2, int y_base, int B, int C, int H, int W) {
    const int lane = threadIdx.x & 31;
    const int which = blockIdx.y + y_base;
    const float* __restrict__ other = which ? f1 : f2;
    float* __restrict__ gin = which ? g2 : g1;
    const long long nel = (long long)B * C * H * W;
    const size_t plane = (size_t)H * W;
    const float inv_c = 1.0f / (float)C;
    for (long long e = blockIdx.x * 8LL + (threadIdx.x >> 5); e < nel; e += gridDim.x * 8LL) {
        const int x = (int)(e % W);
        long long t = e / W;
        const int y = (int)(t % H); t /= H;
        const int c = (int)(t % C), b = (int)(t / C);
        const float* ob = other + ((size_t)b * C + c) * plane;
        const float* gb = gout + (size_t)b * 81 * plane;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int tc = lane + 32 * k;
            if (tc < 81) {
                const int dy = tc / 9 - 4, dx = tc % 9 - 4;
                // first: gO at the pixel itself, f2 at the displaced pixel; second: both at the pixel displaced back
                const int yo = which ? y - dy : y + dy, xo = which ? x - dx : x + dx;
                if (yo >= 0 && yo < H && xo >= 0 && xo < W) {
                    const size_t go_off = which ? (size_t)yo * W + xo : (size_t)y * W + x;
                    acc = fmaf(__ldg(gb + tc * plane + go_off), __ldg(ob + (size_t)yo * W + xo), acc);
                }
            }
        }
        acc = arf_warp_sum(acc);
        if (lane == 0) gin[e] = acc * inv_c;
    }
}

// Packed FP32 FMA (FFMA2, sm_100): two IEEE fused multiply-adds per lane and instruction.  ptxas turns a pair
// built from the same scalar into the broadcast operand form (R.F32), so no MOV is spent on packing.
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void f2_fma(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// ------------------------------------------------------------------ tiled forward, md=4 --
constexpr int kTW = 32;            // tile width  (lane <-> x)
constexpr int kTH = 8;             // tile height (rows per thread)
constexpr int kMD = 4;
constexpr int kD = 2 * kMD + 1;    // 9
constexpr int kHW = kTW + 2 * kMD; // 40 halo width
constexpr int kHH = kTH + 2 * kMD; // 16 halo height
constexpr int kCcDefault = 8;      // channels per pipeline stage
// kRG row groups of 8 rows share one CTA tile (and its f2 halo): consumer warp w <-> (row group w/9, dx = w%9 - 4)
template <int kRG, int kCc>
struct __align__(128) FwdStageT {
    float s1[kCc][kTH * kRG][kTW];             // f1 tile                    8 KB per row group (8 channels)
    float s2[kCc][kTH * kRG + 2 * kMD][kHW];   // f2 tile with 4-px halo    20 KB (RG=1) / 30 KB (RG=2)
};
template <int kRG, int kCc>
constexpr size_t fwd_smem(int stages) { return stages * sizeof(FwdStageT<kRG, kCc>) + 2 * stages * sizeof(uint64_t); }

// Persistent, warp-specialised: warp 9 streams (f1 tile, f2 halo tile) channel chunks through a
// kStages-deep shared-memory ring (TMA with hardware zero fill when kTma, else zero-filling
// cp.async), warps 0..8 consume.  The ring keeps running across tile boundaries, so the loads of
// the next tile overlap the 72 output stores per thread of the current one.
template <bool kTma, int kRG, int kStg, int kMinBlocks, int kUnroll, bool kF2 = false, int kCc = kCcDefault>
__global__ void __launch_bounds__(32 * (kD * kRG + 1), kMinBlocks)
corr_fwd_md4(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
             const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out,
             int B, int C, int H, int W, int tiles_x, int tiles_y, float inv_c, int probe) {
    // probe (tuning only): 1 = consumers skip the FFMA loop and the stores (pure TMA streaming rate),
    //                      2 = skip only the stores, 3 = skip only the FFMA loop
    constexpr int kStages = kStg;
    constexpr int kConsumerWarps = kD * kRG;
    constexpr int kTileH = kTH * kRG;            // CTA tile height
    constexpr int kHaloH = kTileH + 2 * kMD;
    using FwdStage = FwdStageT<kRG, kCc>;
    constexpr uint32_t kFwdStageBytes = sizeof(FwdStage);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdStage* stg = reinterpret_cast<FwdStage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * sizeof(FwdStage));
    uint64_t* empty = full + kStages;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            arf::mbar_init(&full[s], kTma ? 1 : 32);
            arf::mbar_init(&empty[s], kConsumerWarps);
        }
        arf::mbar_fence_init();
    }
    __syncthreads();

    const int ntiles = tiles_x * tiles_y * B;
    const int nchunks = (C + kCc - 1) / kCc;
    const size_t plane = (size_t)H * W;

    if (warp == kConsumerWarps) {
        // ------------------------------------------------------------ producer warp
        if (kTma && lane == 0) {
            arf::tma_prefetch_desc(&map1);
            arf::tma_prefetch_desc(&map2);
        }
        uint32_t it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
            const int x0 = tx * kTW, y0 = ty * kTileH;
            for (int ch = 0; ch < nchunks; ++ch, ++it) {
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                arf::mbar_wait(&empty[s], ph ^ 1);
                const int c0 = ch * kCc;
                if (kTma) {
                    if (lane == 0) {
                        arf::mbar_arrive_expect_tx(&full[s], kFwdStageBytes);
                        arf::tma_load_4d(&stg[s].s1[0][0][0], &map1, &full[s], x0, y0, c0, b);
                        arf::tma_load_4d(&stg[s].s2[0][0][0], &map2, &full[s], x0 - kMD, y0 - kMD, c0, b);
                    }
                } else {
                    const float* f1b = f1 + (size_t)b * C * plane;
                    const float* f2b = f2 + (size_t)b * C * plane;
                    for (int e = lane; e < kCc * kTileH * kTW; e += 32) {
                        int rr = (e / kTW) % kTileH, cc = e / (kTW * kTileH);
                        int gx = x0 + lane, gy = y0 + rr, gc = c0 + cc;
                        bool ok = gc < C && gy < H && gx < W;
                        arf::cp_async_4_zfill(&stg[s].s1[cc][rr][lane],
                                              ok ? f1b + gc * plane + (size_t)gy * W + gx : f1b, ok);
                    }
                    for (int e = lane; e < kCc * kHaloH * kHW; e += 32) {
                        int xx = e % kHW, rr = (e / kHW) % kHaloH, cc = e / (kHW * kHaloH);
                        int gx = x0 + xx - kMD, gy = y0 + rr - kMD, gc = c0 + cc;
                        bool ok = gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
                        arf::cp_async_4_zfill(&stg[s].s2[cc][rr][xx],
                                              ok ? f2b + gc * plane + (size_t)gy * W + gx : f2b, ok);
                    }
                    arf::cp_async_mbar_arrive_noinc(&full[s]);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumer warps
    const int wdx = warp % kD;         // 0..8 -> dx = wdx-4
    const int r0 = (warp / kD) * kTH;  // first row of this warp's row group inside the CTA tile
    uint32_t it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const int x0 = tx * kTW, y0 = ty * kTileH + r0;

        float acc[kTH][kD];
#pragma unroll
        for (int r = 0; r < kTH; ++r)
#pragma unroll
            for (int d = 0; d < kD; ++d) acc[r][d] = 0.f;
        // kF2: rows (2m, 2m+1) share packed accumulators.  pk[m][e-1] = (acc[2m][e], acc[2m+1][e-1]) for e = 1..8:
        // one FFMA2 with the natural pair (a[2m], a[2m+1]) and the BROADCAST scalar bb[2m+e] updates both;
        // acc[2m][0] and acc[2m+1][8] stay scalar.  32 FFMA2 + 8 FFMA per channel instead of 72 FFMA, same
        // rounding (each accumulator sees the same fused multiply-adds in the same order).
        unsigned long long pk[kTH / 2][kD - 1];
        if (kF2) {
#pragma unroll
            for (int m = 0; m < kTH / 2; ++m)
#pragma unroll
                for (int e = 0; e < kD - 1; ++e) pk[m][e] = 0ull;
        }

        for (int ch = 0; ch < nchunks; ++ch, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            arf::mbar_wait(&full[s], ph);
            const FwdStage& S = stg[s];
            if (probe == 1 || probe == 3) {
                acc[0][0] += S.s1[0][r0][lane] + S.s2[0][r0][lane + wdx];
            } else if (kF2) {
#pragma unroll kUnroll
                for (int cc = 0; cc < kCc; ++cc) {
                    float a[kTH], bb[kHH];
#pragma unroll
                    for (int r = 0; r < kTH; ++r) a[r] = S.s1[cc][r0 + r][lane];
#pragma unroll
                    for (int k = 0; k < kHH; ++k) bb[k] = S.s2[cc][r0 + k][lane + wdx];
#pragma unroll
                    for (int m = 0; m < kTH / 2; ++m) {
                        const unsigned long long ap = f2_pack(a[2 * m], a[2 * m + 1]);
                        acc[2 * m][0] = fmaf(a[2 * m], bb[2 * m], acc[2 * m][0]);
#pragma unroll
                        for (int e = 1; e < kD; ++e) f2_fma(pk[m][e - 1], ap, f2_pack(bb[2 * m + e], bb[2 * m + e]));
                        acc[2 * m + 1][kD - 1] = fmaf(a[2 * m + 1], bb[2 * m + kD], acc[2 * m + 1][kD - 1]);
                    }
                }
            } else
#pragma unroll kUnroll
            for (int cc = 0; cc < kCc; ++cc) {
                float a[kTH], bb[kHH];
#pragma unroll
                for (int r = 0; r < kTH; ++r) a[r] = S.s1[cc][r0 + r][lane];
#pragma unroll
                for (int k = 0; k < kHH; ++k) bb[k] = S.s2[cc][r0 + k][lane + wdx];
#pragma unroll
                for (int r = 0; r < kTH; ++r)
#pragma unroll
                    for (int d = 0; d < kD; ++d) acc[r][d] = fmaf(a[r], bb[r + d], acc[r][d]);
            }
            __syncwarp();
            if (lane == 0) arf::mbar_arrive(&empty[s]);
        }
        if (kF2) {
#pragma unroll
            for (int m = 0; m < kTH / 2; ++m)
#pragma unroll
                for (int e = 1; e < kD; ++e) f2_unpack(pk[m][e - 1], acc[2 * m][e], acc[2 * m + 1][e - 1]);
        }

        if (probe == 1 || probe == 2) {
            if (acc[0][0] == 123.456f) out[0] = acc[1][1];   // keep the work alive
            continue;
        }
        // epilogue: 72 coalesced 128-byte row stores per warp; mean over channels = sum * (1/C)
        const int gx = x0 + lane;
        float* op = out + ((size_t)b * (kD * kD) + wdx) * plane + (size_t)y0 * W + gx;
        const size_t dstride = (size_t)kD * plane;
        if (x0 + kTW <= W && y0 + kTH <= H) {
#pragma unroll
            for (int d = 0; d < kD; ++d) {
#pragma unroll
                for (int r = 0; r < kTH; ++r) __stcs(op + r * W, acc[r][d] * inv_c);
                op += dstride;
            }
        } else if (gx < W) {
#pragma unroll
            for (int d = 0; d < kD; ++d) {
#pragma unroll
                for (int r = 0; r < kTH; ++r)
                    if (y0 + r < H) __stcs(op + r * W, acc[r][d] * inv_c);
                op += dstride;
            }
        }
    }
}

// Tuning record (B=64, C=32, 96x128; probes in corr_fwd_md4): memory side alone (TMA loads + 255 MB of row
// stores) 82 us, pure TMA streaming 33 us, LDS+FFMA loop alone 131 us, whole kernel 152 us.  ncu: no unit is
// saturated (shared-memory data pipe 50%, LSU 58%, FMA 46%, issue 69%); the loop is a latency/issue mix of 24 LDS
// feeding 72 FMAs per channel and warp.  Tried and measured:
//   * register tilings that read operands with LDS.128 along x (12 px x 9 dx per thread): 118-148 us for the loop;
//   * packed FFMA2 (32 FFMA2 + 8 FFMA instead of 72 FFMA per channel, broadcast operand, no MOVs, identical bits):
//     163 us with one channel in flight, 149 us with two (kUnroll = 2) - the default now, 2.5% over scalar;
//   * 6 stages of 4 channels instead of 3 of 8: 153-157 us; two row groups per CTA (1 CTA/SM): slower;
//   * the same FFMA2 packing in the backward kernels (40 FFMA2 per channel with zero-padded edge pairs, bit-identical):
//     398 vs 345 us for both gradients - dropped.
// Round 2: tools/lds_probe.cu shows what the loop is bound by - the LSU issues one warp-wide LDS.32 per ~1.45 cycles
// (not 1.0), so 24 LDS.32 per 72 FMAs cap the FMA pipe at ~50 %; the kernel ran at 80 % of that cap.  Also measured:
//   * one CTA per SM with ~170 registers and the next channel's 24 operands prefetched across stage and tile boundaries
//     (explicit double buffering, 9 consumer warps): 154 us - latency was not the limit, the LDS slot count is;
//   * corr_fwd_md4_p2 below (lane pairs share 64-bit operands): 127 us with 16 warps of 128 registers (R = 4 rows x
//     3 row groups); 144 us with R = 6 x 2 (11 warps, 168 registers), 213 us with R = 8 x 1 (6 warps); its loop without
//     the output stores 99 us; scalar FFMAs instead of FFMA2 165 us; starting the row groups one stage apart so that
//     their epilogues do not coincide 140 us (worse: the groups then compete for the LSU instead of sharing L1 lines).
//     The epilogue (127 - 99 us) is the SM -> L2 write path: 255 MB at 32 B/clk/SM is 28 us, and all 15 warps store at the
//     same time.  Making every store a full 128-byte line (the second accumulator's lines are half-filled, the other half
//     comes from the neighbouring warp) changes nothing: 125.9 vs 125.7 us with full-line stores to a wrong plane.

// ------------------------------------------------------------------ tiled forward, pair-shared 64-bit operands ---
// Measured on B200 (tools/lds_probe.cu): a warp-wide LDS.32 costs ~1.45 LSU cycles whatever its addresses, an LDS.64
// whose lane pairs (2p, 2p+1) read the SAME 8 bytes costs ~1.48 - the same instruction slot delivers two operands.  The
// kernels above are bound by exactly that slot (24 LDS.32 per 72 FMAs).  Here the two lanes of a pair share their f2
// operand: lane l <-> pixel x = l as before, but a warp covers a PAIR of horizontal displacements.  With d' even and
// xe = x & ~1, the 8 bytes at halo columns (xe + d', xe + d' + 1) are, for the even lane, its operands for dx = d' and
// d' + 1, and for the odd lane (x = xe + 1) those for dx = d' - 1 and d'.  Five warps (d' = -4, -2, 0, 2, 4) cover the nine
// displacements (dx = -5 / +5 of the outer warps are computed and dropped: 10 % waste).  The loaded register pair is
// directly the packed operand of FFMA2 (accumulator pair = the lane's two displacements), f1 is the broadcast scalar.
// Per channel and warp: R LDS.32 + (R + 8) LDS.64 feed 9R FFMA2 = 18R FMAs  (R = 6: 29 LSU cycles per 108 FMAs, the
// kernels above: 35 per 72).
template <int TH, int kCc>
struct __align__(128) FwdStageP {
    float s1[kCc][TH][kTW];
    float s2[kCc][TH + 2 * kMD][kHW];
};
template <int TH, int kCc>
constexpr size_t fwd_p2_smem(int stages) { return stages * sizeof(FwdStageP<TH, kCc>) + 2 * stages * sizeof(uint64_t); }

__device__ __forceinline__ unsigned long long lds_pair(const float* p) {
    unsigned long long v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(arf::smem_u32(p)));
    return v;
}

constexpr int kPW = 5;   // displacement-pair warps per row group

template <int R, int RG, int kStg, int kMinB = 1>
__global__ void __launch_bounds__(32 * (kPW * RG + 1), kMinB)
corr_fwd_md4_p2(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                float* __restrict__ out, int B, int C, int H, int W, int tiles_x, int tiles_y, float inv_c, int probe) {
    // probe (tuning only): 2 = skip the output stores
    constexpr int kCc = 8;
    constexpr int TH = R * RG;
    constexpr int kConsumers = kPW * RG;
    using Stage = FwdStageP<TH, kCc>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage* stg = reinterpret_cast<Stage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStg * sizeof(Stage));
    uint64_t* empty = full + kStg;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStg; ++s) {
            arf::mbar_init(&full[s], 1);
            arf::mbar_init(&empty[s], kConsumers);
        }
        arf::mbar_fence_init();
    }
    __syncthreads();
    const int ntiles = tiles_x * tiles_y * B;
    const int nchunks = (C + kCc - 1) / kCc;
    const size_t plane = (size_t)H * W;

    if (warp == kConsumers) {
        if (lane == 0) {
            arf::tma_prefetch_desc(&map1);
            arf::tma_prefetch_desc(&map2);
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
                const int x0 = tx * kTW, y0 = ty * TH;
                for (int ch = 0; ch < nchunks; ++ch) {
                    arf::mbar_wait(&empty[s], ph ^ 1);
                    arf::mbar_arrive_expect_tx(&full[s], (uint32_t)sizeof(Stage));
                    arf::tma_load_4d(&stg[s].s1[0][0][0], &map1, &full[s], x0, y0, ch * kCc, b);
                    arf::tma_load_4d(&stg[s].s2[0][0][0], &map2, &full[s], x0 - kMD, y0 - kMD, ch * kCc, b);
                    if (++s == kStg) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    const int rg = warp / kPW, wp = warp % kPW;
    const int dpr = 2 * wp;                    // d' + 4
    const int odd = lane & 1;
    const int col = (lane & ~1) + dpr;         // halo column of the pair's first operand
    const int r0 = rg * R;
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, b = t / (tiles_x * tiles_y);
        const int x0 = tx * kTW, y0 = ty * TH + r0;
        unsigned long long acc[R][kD];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int d = 0; d < kD; ++d) acc[r][d] = 0ull;
#pragma unroll 1
        for (int ch = 0; ch < nchunks; ++ch) {
            arf::mbar_wait(&full[s], ph);
            const float* s1c = &stg[s].s1[0][r0][lane];
            const float* s2c = &stg[s].s2[0][r0][col];
#pragma unroll
            for (int cc = 0; cc < kCc; ++cc) {
                float a[R];
                unsigned long long bb[R + 2 * kMD];
#pragma unroll
                for (int r = 0; r < R; ++r) a[r] = s1c[(cc * TH + r) * kTW];
#pragma unroll
                for (int k = 0; k < R + 2 * kMD; ++k) bb[k] = lds_pair(s2c + (cc * (TH + 2 * kMD) + k) * kHW);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned long long ap = f2_pack(a[r], a[r]);
#pragma unroll
                    for (int d = 0; d < kD; ++d) f2_fma(acc[r][d], ap, bb[r + d]);
                }
            }
            __syncwarp();
            if (lane == 0) arf::mbar_arrive(&empty[s]);
            if (++s == kStg) { s = 0; ph ^= 1; }
        }
        // epilogue.  Displacement plane dpr gets a full 128-byte row (even lanes: first accumulator, odd lanes: second);
        // the other accumulator goes to plane dpr + 1 (even lanes) / dpr - 1 (odd lanes), whose other half comes from
        // the neighbouring warp.
        const int gx = x0 + lane;
        if (probe == 2) {
            float va, vb;
            f2_unpack(acc[0][0], va, vb);
            if (va == 123.456f) out[0] = vb;   // keep the work alive
            continue;
        }
        if (gx < W) {
            const int p2 = odd ? dpr - 1 : dpr + 1;
            const bool ok2 = p2 >= 0 && p2 < kD;
            float* o1 = out + ((size_t)b * (kD * kD) + dpr) * plane + (size_t)y0 * W + gx;
            float* o2 = out + ((size_t)b * (kD * kD) + (ok2 ? p2 : dpr)) * plane + (size_t)y0 * W + gx;
            const size_t dstride = (size_t)kD * plane;
#pragma unroll
            for (int d = 0; d < kD; ++d) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float va, vb;
                    f2_unpack(acc[r][d], va, vb);
                    if (y0 + r < H) {
                        __stcs(o1 + r * W, (odd ? vb : va) * inv_c);
                        if (ok2) __stcs(o2 + r * W, (odd ? va : vb) * inv_c);
                    }
                }
                o1 += dstride;
                o2 += dstride;
            }
        }
    }
}

// ------------------------------------------------------------------ tiled backward, md=4 -
// g1[c,y,x] = 1/C sum_{dy,dx} gO[(dy,dx),y,x]       * f2[c,y+dy,x+dx]           (kSecond = false, F = f2)
// g2[c,y,x] = 1/C sum_{dy,dx} gO[(dy,dx),y-dy,x-dx] * f1[c,y-dy,x-dx]           (kSecond = true,  F = f1)
// Both are gathers (no atomics, nothing to zero-fill) and run in ONE launch: blockIdx.y selects the gradient.
// Work item = (32 x 8*RG pixel tile, 8*CG channels); RG*CG consumer warps = RG row groups (8 rows each) x CG channel
// groups (8 channels each); lane <-> x.  For each of the 9 horizontal displacements a warp pulls the 8 rows x 9
// vertical displacements of gO for its column into registers (from a TMA-streamed "slab" = the 9 planes of that dx)
// and reuses them for its 8 channels: per channel 16 LDS (halo column of F) feed 72 FFMA into 8 accumulators.
// Item shapes in use (arf_corr_bwd picks by rounds x round time): <2,4> = 32x16 px x 32 channels, 8 consumer warps, one
// CTA per SM, for problems that fill the machine with such items; <2,2> (16 channels) and <1,4> (32x8 px) for the coarse
// pyramid levels, where two launches of the big item left most SMs idle behind a fixed ~35 us (16x32x24x32: 36 -> 13 us).
// Round-2 measurements (ncu, profiles/r2_corr_bwd_cfg2_ncu.txt): FMA pipe 36 %, LSU 48 %, 9 resident warps.  The loop
// spends 72 + 8 x 16 LDS.32 slots (at ~1.45 cycles each) per 576 FMAs, which caps the FMA pipe near 50 %.  Tried: the
// same item with 4 rows per warp and 16 consumer warps at 96 registers (more LDS per FMA: 0.46 instead of 0.35) -
// 334 vs 320 us at 64x32x96x128, i.e. occupancy buys nothing, the LDS slot count is the bound here too.  The forward
// kernel's pair-shared operands do not carry over cheaply: the 2 x 36 gradient values of a displacement PAIR have to
// sit in registers next to the accumulators, which leaves 8 channels per warp (0.29 LDS per FMA, -16 %) at 10 warps.
constexpr int kBCg = 8;                  // channels per warp

// RT = rows per warp: 8 (72 + 8 x 16 LDS per 576 FMAs and dx step) for the big items, 4 for the small-problem items (more
// LDS per FMA, but half the latency of an item and twice as many items to spread over the SMs)
template <int RG, int CG, int STG, int RT = kTH>
struct BwdCfg {
    static constexpr int kTileH = RT * RG;            // tile height
    static constexpr int kHaloH = kTileH + 2 * kMD;
    static constexpr int kC = kBCg * CG;              // channels per work item
    static constexpr int kConsumers = RG * CG;
    static constexpr int kThreads = 32 * (kConsumers + 1);   // + producer warp
    static constexpr int kStages = STG;               // slab ring depth
};

template <bool kSecond, int RG, int CG, int STG, int RT = kTH>
struct BwdSmem {
    using Cfg = BwdCfg<RG, CG, STG, RT>;
    // The second gradient reads gO at (y-dy, x-dx): its slab carries the 4-px halo in both directions.
    // (A tiled TMA load needs a 16-byte aligned innermost coordinate — measured: x0-3 raises "illegal
    // instruction" — so the horizontal shift is applied when reading, not when loading.)
    static constexpr int kSlabRows = kSecond ? Cfg::kHaloH : Cfg::kTileH;
    static constexpr int kSlabW = kSecond ? kHW : kTW;
    float F[Cfg::kC][Cfg::kHaloH][kHW];                 // halo tile of the other feature map
    float slab[STG][kD][kSlabRows][kSlabW];
    uint64_t f_full[CG], f_empty, s_full[STG], s_empty[STG];
};

template <bool kSecond, bool kTma, int RG, int CG, int STG, int RT = kTH>
__device__ __forceinline__ void corr_bwd_body(unsigned char* smem_raw, const CUtensorMap* mapF, const CUtensorMap* mapG,
                                              const float* __restrict__ Fsrc, const float* __restrict__ gout,
                                              float* __restrict__ gin, int B, int C, int H, int W, int tiles_x,
                                              int tiles_y, int nsuper, float inv_c) {
    using Cfg = BwdCfg<RG, CG, STG, RT>;
    using Smem = BwdSmem<kSecond, RG, CG, STG, RT>;
    constexpr int kTH = RT, kHH = RT + 2 * kMD;      // shadow the forward kernel's 8-row constants
    constexpr int kSlabRows = Smem::kSlabRows;
    constexpr int kSlabW = Smem::kSlabW;
    constexpr int kBTH = Cfg::kTileH, kBHH = Cfg::kHaloH, kBC = Cfg::kC, kBConsumers = Cfg::kConsumers;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < CG; ++i) arf::mbar_init(&sm.f_full[i], kTma ? 1 : 32);
        arf::mbar_init(&sm.f_empty, kBConsumers);
        for (int s = 0; s < STG; ++s) {
            arf::mbar_init(&sm.s_full[s], kTma ? 1 : 32);
            arf::mbar_init(&sm.s_empty[s], kBConsumers);
        }
        arf::mbar_fence_init();
    }
    __syncthreads();

    const long long nitems = (long long)tiles_x * tiles_y * B * nsuper;
    const size_t plane = (size_t)H * W;

    if (warp == kBConsumers) {
        // ------------------------------------------------------------ producer warp
        if (kTma && lane == 0) {
            arf::tma_prefetch_desc(mapF);
            arf::tma_prefetch_desc(mapG);
        }
        uint32_t it = 0, item = 0;
        for (long long t = blockIdx.x; t < nitems; t += gridDim.x, ++item) {
            const int cs = t % nsuper;
            long long tt = t / nsuper;
            const int tx = tt % tiles_x, ty = (tt / tiles_x) % tiles_y, b = tt / ((long long)tiles_x * tiles_y);
            const int x0 = tx * kTW, y0 = ty * kBTH, c0 = cs * kBC;
            arf::mbar_wait(&sm.f_empty, (item & 1) ^ 1);
            if (kTma) {
                if (lane == 0) {
                    for (int gch = 0; gch < CG; ++gch) {
                        arf::mbar_arrive_expect_tx(&sm.f_full[gch], kBCg * kBHH * kHW * 4);
                        arf::tma_load_4d(&sm.F[gch * kBCg][0][0], mapF, &sm.f_full[gch], x0 - kMD, y0 - kMD,
                                         c0 + gch * kBCg, b);
                    }
                }
            } else {
                const float* fb = Fsrc + (size_t)b * C * plane;
                for (int gch = 0; gch < CG; ++gch) {
                    for (int e = lane; e < kBCg * kBHH * kHW; e += 32) {
                        int xx = e % kHW, rr = (e / kHW) % kBHH, cc = gch * kBCg + e / (kHW * kBHH);
                        int gx = x0 + xx - kMD, gy = y0 + rr - kMD, gc = c0 + cc;
                        bool ok = gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
                        arf::cp_async_4_zfill(&sm.F[cc][rr][xx], ok ? fb + gc * plane + (size_t)gy * W + gx : fb, ok);
                    }
                    arf::cp_async_mbar_arrive_noinc(&sm.f_full[gch]);
                }
            }
            for (int dx = 0; dx < kD; ++dx, ++it) {
                const int s = it % STG;
                arf::mbar_wait(&sm.s_empty[s], ((it / STG) & 1) ^ 1);
                // first: the tile itself; second: the tile with its 4-px halo
                const int sx = kSecond ? x0 - kMD : x0;
                const int sy = kSecond ? y0 - kMD : y0;
                if (kTma) {
                    if (lane == 0) {
                        arf::mbar_arrive_expect_tx(&sm.s_full[s], kD * kSlabRows * kSlabW * 4);
                        arf::tma_load_5d(&sm.slab[s][0][0][0], mapG, &sm.s_full[s], sx, sy, dx, 0, b);
                    }
                } else {
                    const float* gb = gout + (size_t)b * kD * kD * plane;
                    for (int e = lane; e < kD * kSlabRows * kSlabW; e += 32) {
                        int xx = e % kSlabW, rr = (e / kSlabW) % kSlabRows, dy = e / (kSlabW * kSlabRows);
                        int gx = sx + xx, gy = sy + rr;
                        bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
                        arf::cp_async_4_zfill(&sm.slab[s][dy][rr][xx],
                                              ok ? gb + (size_t)(dy * kD + dx) * plane + (size_t)gy * W + gx : gb, ok);
                    }
                    arf::cp_async_mbar_arrive_noinc(&sm.s_full[s]);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumer warps
    const int rg = warp / CG;        // row group: rows rg*8 .. rg*8+7 of the tile
    const int cgp = warp % CG;       // channel group: channels cgp*8 .. cgp*8+7 of the work item
    const int r0 = rg * kTH;
    uint32_t it = 0, item = 0;
    for (long long t = blockIdx.x; t < nitems; t += gridDim.x, ++item) {
        const int cs = t % nsuper;
        long long tt = t / nsuper;
        const int tx = tt % tiles_x, ty = (tt / tiles_x) % tiles_y, b = tt / ((long long)tiles_x * tiles_y);
        const int x0 = tx * kTW, y0 = ty * kBTH, c0 = cs * kBC;

        float o[kBCg][kTH];
#pragma unroll
        for (int c = 0; c < kBCg; ++c)
#pragma unroll
            for (int r = 0; r < kTH; ++r) o[c][r] = 0.f;

        arf::mbar_wait(&sm.f_full[cgp], item & 1);
#pragma unroll 1
        for (int dx = 0; dx < kD; ++dx, ++it) {
            const int s = it % STG;
            arf::mbar_wait(&sm.s_full[s], (it / STG) & 1);
            const int col = kSecond ? lane + 2 * kMD - dx : lane + dx;   // column in the halo frame
            float g[kTH][kD];
#pragma unroll
            for (int r = 0; r < kTH; ++r)
#pragma unroll
                for (int d = 0; d < kD; ++d)
                    g[r][d] = kSecond ? sm.slab[s][d][r0 + r + 2 * kMD - d][col] : sm.slab[s][d][r0 + r][lane];
            __syncwarp();
            if (lane == 0) arf::mbar_arrive(&sm.s_empty[s]);
#pragma unroll
            for (int c = 0; c < kBCg; ++c) {
                float bb[kHH];
#pragma unroll
                for (int k = 0; k < kHH; ++k) bb[k] = sm.F[cgp * kBCg + c][r0 + k][col];
#pragma unroll
                for (int r = 0; r < kTH; ++r)
#pragma unroll
                    for (int d = 0; d < kD; ++d)
                        o[c][r] = fmaf(g[r][d], kSecond ? bb[r + 2 * kMD - d] : bb[r + d], o[c][r]);
            }
        }
        __syncwarp();
        if (lane == 0) arf::mbar_arrive(&sm.f_empty);

        const int gx = x0 + lane;
        if (gx < W) {
#pragma unroll
            for (int c = 0; c < kBCg; ++c) {
                const int gc = c0 + cgp * kBCg + c;
                if (gc < C) {
                    float* op = gin + ((size_t)b * C + gc) * plane + (size_t)(y0 + r0) * W + gx;
#pragma unroll
                    for (int r = 0; r < kTH; ++r)
                        if (y0 + r0 + r < H) op[(size_t)r * W] = o[c][r] * inv_c;
                }
            }
        }
    }
}

// maps: [0] f2 halo tiles, [1] gO tile slabs (first gradient); [2] f1 halo tiles, [3] gO halo slabs (second).
struct BwdMaps { CUtensorMap m[4]; };

// blockIdx.y + y_base: 0 = gradient w.r.t. f1, 1 = gradient w.r.t. f2
template <bool kTma, int RG, int CG, int STG, int RT = kTH, int kMinB = 1>
__global__ void __launch_bounds__((BwdCfg<RG, CG, STG, RT>::kThreads), kMinB)
corr_bwd_md4(const __grid_constant__ BwdMaps maps, const float* __restrict__ f1, const float* __restrict__ f2,
             const float* __restrict__ gout, float* __restrict__ g1, float* __restrict__ g2, int y_base, int B, int C,
             int H, int W, int tiles_x, int tiles_y, int nsuper, float inv_c) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (blockIdx.y + y_base == 0)
        corr_bwd_body<false, kTma, RG, CG, STG, RT>(smem_raw, &maps.m[0], &maps.m[1], f2, gout, g1, B, C, H, W, tiles_x, tiles_y,
                                                nsuper, inv_c);
    else
        corr_bwd_body<true, kTma, RG, CG, STG, RT>(smem_raw, &maps.m[2], &maps.m[3], f1, gout, g2, B, C, H, W, tiles_x, tiles_y,
                                               nsuper, inv_c);
}

template <int RG, int CG, int STG, int RT = kTH, int kMinB = 1>
int launch_bwd_md4(const float* f1, const float* f2, const float* gout, float* g1, float* g2, int B, int C, int H, int W,
                   bool want_tma, int ctas_per_sm, cudaStream_t st) {
    using Cfg = BwdCfg<RG, CG, STG, RT>;
    const int tiles_x = arf_cdiv(W, kTW), tiles_y = arf_cdiv(H, Cfg::kTileH), nsuper = arf_cdiv(C, Cfg::kC);
    const long long nitems = (long long)tiles_x * tiles_y * B * nsuper;
    const long long cap = (long long)ARF_NUM_SMS * ctas_per_sm;
    const int ngrad = (g1 ? 1 : 0) + (g2 ? 1 : 0);
    // both gradients share the machine: each gets half of the resident CTAs
    const long long per = ngrad == 2 ? (cap + 1) / 2 : cap;
    dim3 grid((unsigned)(nitems < per ? nitems : per), ngrad);
    BwdMaps maps;
    memset(&maps, 0, sizeof(maps));
    bool tma = want_tma && arf::tma_ok_nchw(f1, W) && arf::tma_ok_nchw(f2, W) && arf::tma_ok_nchw(gout, W);
    if (tma && g1)
        tma = arf::make_map_nchw(&maps.m[0], f2, B, C, H, W, kHW, Cfg::kHaloH, kBCg) &&
              arf::make_map_costvol(&maps.m[1], gout, B, kD, H, W, kTW, Cfg::kTileH);
    if (tma && g2)
        tma = arf::make_map_nchw(&maps.m[2], f1, B, C, H, W, kHW, Cfg::kHaloH, kBCg) &&
              arf::make_map_costvol(&maps.m[3], gout, B, kD, H, W, kHW, Cfg::kHaloH);
    const float inv_c = 1.0f / (float)C;
    const size_t smem = g2 ? sizeof(BwdSmem<true, RG, CG, STG, RT>) : sizeof(BwdSmem<false, RG, CG, STG, RT>);
    const int y_base = g1 ? 0 : 1;
    if (tma) {
        ARF_ENSURE_SMEM((corr_bwd_md4<true, RG, CG, STG, RT, kMinB>), sizeof(BwdSmem<true, RG, CG, STG, RT>));
        corr_bwd_md4<true, RG, CG, STG, RT, kMinB><<<grid, Cfg::kThreads, smem, st>>>(maps, f1, f2, gout, g1, g2, y_base, B, C, H, W,
                                                                          tiles_x, tiles_y, nsuper, inv_c);
    } else {
        ARF_ENSURE_SMEM((corr_bwd_md4<false, RG, CG, STG, RT, kMinB>), sizeof(BwdSmem<true, RG, CG, STG, RT>));
        corr_bwd_md4<false, RG, CG, STG, RT, kMinB><<<grid, Cfg::kThreads, smem, st>>>(maps, f1, f2, gout, g1, g2, y_base, B, C, H, W,
                                                                           tiles_x, tiles_y, nsuper, inv_c);
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
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

// Test hooks (arf_debug_set): per calling thread, so concurrent callers (one Python thread per GPU under the
// reference's DataParallel) never see each other's settings; all default to 0 = production routing.
ARF_HOOK g_probe = 0;         // see corr_fwd_md4
ARF_HOOK g_variant = 0;       // kernel variant selection while tuning
ARF_HOOK g_force_no_tma = 0;  // exercise the cp.async producer on TMA-capable shapes

// The tiled kernels carry a fixed pipeline latency (forward ~11 us, backward ~35 us for any small problem) and, for
// tensors TMA cannot describe (W % 4 != 0), a single-warp cp.async producer.  Measured with tools/microbench.py
// (variant 5 = force literal, 6 = force tiled): the literal forward wins up to ~1000 output pixels (1x192x6x10:
// 13 vs 277 us, 1x96x24x40: 12 vs 25 us; 16x32x12x16: 14 vs 12 us), the literal backward only for tiny non-TMA
// shapes (1x192x6x10: 98 vs 342 us; 1x128x12x20: 99 vs 34 us).
inline bool is_md4(const CorrGeom& g) {
    return g.ks == 1 && g.s1 == 1 && g.s2 == 1 && g.md == kMD && g.pad == kMD;
}
inline bool is_fast(const CorrGeom& g, bool bwd) {
    if (!is_md4(g) || g_variant == 5) return false;
    if (g_variant == 6) return true;
    const long long px = (long long)g.B * g.H * g.W;
    if (bwd) return px > 128;       // below: corr_bwd_small_md4 (1x192x6x10: 10 us against 98 literal / 342 tiled)
    return px > 1024;               // below: literal, or corr_fwd_small_md4 up to 128 pixels (1x192x6x10: 4 vs 13 us)
}

}  // namespace

#if ARF_TEST_HOOKS
extern thread_local int g_warp_variant;   // warp.cu
extern thread_local int g_trisolve_variant;   // stencil.cu
extern thread_local int g_census_variant;     // census.cu
#endif

extern "C" int arf_debug_set(int key, int value) {
#if !ARF_TEST_HOOKS
    (void)key; (void)value;
    return ARF_EUNSUPPORTED;
#else
    if (key == 0) { g_force_no_tma = value; return ARF_OK; }
    if (key == 1) { g_variant = value; return ARF_OK; }
    if (key == 2) { g_probe = value; return ARF_OK; }
    if (key == 3) { g_warp_variant = value; return ARF_OK; }
    if (key == 4) { g_trisolve_variant = value; return ARF_OK; }
    if (key == 5) { g_census_variant = value; return ARF_OK; }
    return ARF_EINVAL;
#endif
}

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
    if (is_fast(g, false)) {
        const int tiles_x = arf_cdiv(W, kTW);
        const float inv_c = 1.0f / (float)C;
        const bool tma = !g_force_no_tma && arf::tma_ok_nchw(f1, W) && arf::tma_ok_nchw(f2, W);
        const long long sms = ARF_NUM_SMS;
        // Three tiled launches.  "p2" (pair-shared 64-bit operands) needs TMA and runs with 32x12 tiles (one 16-warp CTA per
        // SM) or, for problems that would leave SMs idle that way (a coarse pyramid level, batch 1), with 32x4 tiles (one
        // row group, 6 warps, two CTAs per SM); the column-thread kernel has 32x8 tiles (two CTAs per SM) and serves the
        // tensors TMA cannot describe.  A launch lasts (tiles on the busiest SM) x (tile rows) x (time per row and
        // channel); measured on B200 (tools/microbench.py corr_fwd --variant 30 | 31 | 32): 0.0235 us (p2), 0.025 (p2
        // small), 0.0529 per CTA with two column-thread CTAs sharing an SM.  E.g. 64x32x96x128: 126 (p2) / 149 us (column);
        // 16x96x24x32: 25 / 25 / 15 (small tiles); 1x64x48x80: 18 / 19 / 11; 8x64x56x128 (160 p2 tiles = 2 rounds): 35 / 28.
        const long long n_p2 = (long long)tiles_x * arf_cdiv(H, 12) * B, n_ct = (long long)tiles_x * arf_cdiv(H, kTH) * B;
        const long long n_p2s = (long long)tiles_x * arf_cdiv(H, 4) * B;
        if (n_p2s > 0x7fffffffLL) return ARF_EINVAL;
        const double t_p2 = (double)((n_p2 + sms - 1) / sms) * 12 * 0.0235;
        const double t_p2s = (double)((n_p2s + sms - 1) / sms) * 4 * 0.025;
        const double t_ct = (double)((n_ct + 2 * sms - 1) / (2 * sms)) * kTH * 0.0529;
        bool use_p2s = tma && t_p2s < 0.97 * t_p2 && t_p2s < 0.97 * t_ct;
        bool use_p2 = tma && !use_p2s && t_p2 < 0.97 * t_ct;
        if (g_variant == 30) { use_p2 = tma; use_p2s = false; }          // tuning hooks: force one of the three
        if (g_variant == 31) { use_p2 = false; use_p2s = false; }
        if (g_variant == 32) { use_p2s = tma; use_p2 = false; }
#define ARF_LAUNCH_P2(R, RG, STG, MINB, NT)                                                                          \
    do {                                                                                                             \
        CUtensorMap m1, m2;                                                                                          \
        if (arf::make_map_nchw(&m1, f1, B, C, H, W, kTW, R * RG, 8) &&                                               \
            arf::make_map_nchw(&m2, f2, B, C, H, W, kHW, R * RG + 2 * kMD, 8)) {                                     \
            auto kern = corr_fwd_md4_p2<R, RG, STG, MINB>;                                                           \
            ARF_ENSURE_SMEM(kern, (fwd_p2_smem<R * RG, 8>(STG)));                                                    \
            const int tiles_y = arf_cdiv(H, R * RG);                                                                 \
            const int grid = (int)((NT) < (MINB) * sms ? (NT) : (MINB) * sms);                                       \
            kern<<<grid, 32 * (kPW * RG + 1), fwd_p2_smem<R * RG, 8>(STG), st>>>(m1, m2, out, B, C, H, W, tiles_x,   \
                                                                                 tiles_y, inv_c, g_probe);           \
            ARF_CHECK_LAUNCH();                                                                                      \
            return ARF_OK;                                                                                           \
        }                                                                                                            \
    } while (0)
        if (use_p2s) ARF_LAUNCH_P2(4, 1, 4, 2, n_p2s);
        if (use_p2) ARF_LAUNCH_P2(4, 3, 4, 1, n_p2);
#undef ARF_LAUNCH_P2
        const int tiles_y = arf_cdiv(H, kTH);
        const long long ntiles = n_ct;
        CUtensorMap m1, m2;
        bool tma_ct = tma && arf::make_map_nchw(&m1, f1, B, C, H, W, kTW, kTH, kCcDefault) &&
                      arf::make_map_nchw(&m2, f2, B, C, H, W, kHW, kTH + 2 * kMD, kCcDefault);
        if (!tma_ct) {
            memset(&m1, 0, sizeof(m1));
            memset(&m2, 0, sizeof(m2));
        }
#define ARF_LAUNCH_FWD(TMA, RG, STG, MINB, UNR, F2, CC)                                                          \
    do {                                                                                                         \
        auto kern = corr_fwd_md4<TMA, RG, STG, MINB, UNR, F2, CC>;                                               \
        ARF_ENSURE_SMEM(kern, (fwd_smem<RG, CC>(STG)));                                                          \
        const int grid = (int)(ntiles < (MINB) * ARF_NUM_SMS ? ntiles : (MINB) * ARF_NUM_SMS);                   \
        kern<<<grid, 32 * (kD * RG + 1), fwd_smem<RG, CC>(STG), st>>>(m1, m2, f1, f2, out, B, C, H, W, tiles_x,  \
                                                                      tiles_y, inv_c, g_probe);                  \
    } while (0)
        if (!tma_ct) ARF_LAUNCH_FWD(false, 1, 3, 2, 1, false, 8);
        else if (g_variant == 7) ARF_LAUNCH_FWD(true, 1, 3, 2, 1, false, 8);   // scalar FFMA loop (same bits)
        else ARF_LAUNCH_FWD(true, 1, 3, 2, 2, true, 8);                        // FFMA2 loop, two channels in flight
#undef ARF_LAUNCH_FWD
        ARF_CHECK_LAUNCH();
        return ARF_OK;
    } else if (is_md4(g) && g_variant != 5 && (long long)B * H * W <= 128) {
        // a few dozen pixels, hundreds of channels: a warp per output element (at 240+ pixels its strided channel
        // loads lose to the literal kernel: 1x128x12x20 14 vs 12 us)
        const long long nout = (long long)B * 81 * H * W;
        corr_fwd_small_md4<<<arf_grid_1d(nout * 32, 256, 16), 256, 0, st>>>(f1, f2, out, B, C, H, W);
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
    if (is_fast(g, true)) {
        if (!g1 && !g2) return ARF_OK;
        // Six item shapes; a launch takes ceil(items / resident CTAs) rounds of roughly constant duration, so the shape is
        // chosen by rounds x measured round time (B200, tools/microbench.py corr_bwd --variant 10..15, both gradients):
        //   0: <2,4> 32x16 px x 32 ch, 8 rows per warp, 1 CTA/SM, ~16 us     1: <2,2> 32x16 px x 16 ch, 1 CTA/SM, ~11 us
        //   2: <1,4> 32x8 px x 32 ch, 2 CTAs/SM, ~22 us for a full pair
        // and with 4 rows per warp (more LDS per FMA, but short items for the levels that cannot fill the machine):
        //   3: <1,4> 32x4 px x 32 ch, 2 CTAs/SM, ~10 us    4: <3,4> 32x12 px x 32 ch, 1 CTA/SM, ~13.5 us    5: <2,4> 32x8, ~10.3 us
        // e.g. 16x32x48x64: 34 / 41 / 44 / 30 / 29 / 30 us, 16x32x12x16: 13 / 13 / 13 / 8.7, 16x96x24x32: 33 / 33 / 26 / 21 /
        // 28 / 21, 16x32x96x128: 95 (0) vs 92.5 (4), 64x32x96x128: 321 (0) vs 357 (4).  (8-channel items re-stream the gO slabs
        // four times as often and lose everywhere: 65 / 68 us at 16x32x48x64.)
        const int ngrad = (g1 ? 1 : 0) + (g2 ? 1 : 0);
        const long long tx = arf_cdiv(W, kTW), sms = ARF_NUM_SMS;
        const long long c32 = arf_cdiv(C, 32), c16 = arf_cdiv(C, 16);
        const long long n[6] = {tx * arf_cdiv(H, 16) * B * c32 * ngrad, tx * arf_cdiv(H, 16) * B * c16 * ngrad,
                                tx * arf_cdiv(H, 8) * B * c32 * ngrad,  tx * arf_cdiv(H, 4) * B * c32 * ngrad,
                                tx * arf_cdiv(H, 12) * B * c32 * ngrad, tx * arf_cdiv(H, 8) * B * c32 * ngrad};
        const double t[6] = {(double)((n[0] + sms - 1) / sms) * 16.0,          (double)((n[1] + sms - 1) / sms) * 11.0,
                             (double)((n[2] + 2 * sms - 1) / (2 * sms)) * 22.0, (double)((n[3] + 2 * sms - 1) / (2 * sms)) * 10.0,
                             (double)((n[4] + sms - 1) / sms) * 13.5,          (double)((n[5] + sms - 1) / sms) * 10.3};
        int cfg = 0;
        for (int k = 1; k < 6; ++k)
            if (t[k] < t[cfg]) cfg = k;
        if (g_variant >= 10 && g_variant <= 15) cfg = g_variant - 10;   // tuning hook
        const bool tma = !g_force_no_tma;
        switch (cfg) {
            case 0: return launch_bwd_md4<2, 4, 3>(f1, f2, gout, g1, g2, B, C, H, W, tma, 1, st);
            case 1: return launch_bwd_md4<2, 2, 2>(f1, f2, gout, g1, g2, B, C, H, W, tma, 1, st);
            case 2: return launch_bwd_md4<1, 4, 3>(f1, f2, gout, g1, g2, B, C, H, W, tma, 2, st);
            case 3: return launch_bwd_md4<1, 4, 2, 4, 2>(f1, f2, gout, g1, g2, B, C, H, W, tma, 2, st);
            case 4: return launch_bwd_md4<3, 4, 3, 4, 1>(f1, f2, gout, g1, g2, B, C, H, W, tma, 1, st);
            default: return launch_bwd_md4<2, 4, 2, 4, 1>(f1, f2, gout, g1, g2, B, C, H, W, tma, 1, st);
        }
    }
    long long total = (long long)B * C * H * W;
    if (is_md4(g) && g_variant != 5) {
        if (!g1 && !g2) return ARF_OK;
        dim3 grid(arf_grid_1d(total * 32, 256, 16), (g1 ? 1 : 0) + (g2 ? 1 : 0));
        corr_bwd_small_md4<<<grid, 256, 0, st>>>(f1, f2, gout, g1, g2, g1 ? 0 : 1, B, C, H, W);
        ARF_CHECK_LAUNCH();
        return ARF_OK;
    }
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

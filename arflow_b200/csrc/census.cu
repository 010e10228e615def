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

// out[0] = num, out[1] = den, out[2] = num / (den + 1e-6); fixed summation order (deterministic)
__global__ void census_finalize_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
    __shared__ double sn[256], sd[256];
    double a = 0.0, d = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a += (double)partials[2 * (size_t)i];
        d += (double)partials[2 * (size_t)i + 1];
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
        out[0] = num;
        out[1] = den;
        out[2] = num / (den + 1e-6f);
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
                  int W, int tiles_x, int tiles_y, float scale, float eps, float q) {
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
        if (!ghamming) {
            gl = __ldg(gloss);
            idn = 1.f / (__ldg(sums + 1) + 1e-6f);
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

}  // namespace

extern "C" int arf_census_num_partials(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return ARF_EINVAL;
    long long n = (long long)arf_cdiv(W, kCTW) * arf_cdiv(H, kCTH) * B;
    return n > 0x7fffffffLL ? ARF_EINVAL : (int)n;
}

extern "C" int arf_census_fwd(const float* im_a, const float* im_b, const float* mask, float* hamming,
                              float* partials, float* sums, int B, int H, int W, int patch, float scale,
                              float eps, float q, void* stream) {
    ARF_REQUIRE(im_a && im_b && hamming);
    ARF_REQUIRE(B > 0 && H > 0 && W > 0 && patch >= 1 && (patch & 1));
    const int want = sums != nullptr;
    if (want) ARF_REQUIRE(partials != nullptr);
    const int tiles_x = arf_cdiv(W, kCTW), tiles_y = arf_cdiv(H, kCTH);
    const int n = arf_census_num_partials(B, H, W);
    if (n < 0) return n;
    cudaStream_t st = (cudaStream_t)stream;
    switch (patch / 2) {
        case 1: census_fwd_kernel<1><<<n, kCThreads, 0, st>>>(im_a, im_b, mask, hamming, partials, B, H, W, tiles_x, tiles_y, scale, want, eps, q); break;
        case 2: census_fwd_kernel<2><<<n, kCThreads, 0, st>>>(im_a, im_b, mask, hamming, partials, B, H, W, tiles_x, tiles_y, scale, want, eps, q); break;
        case 3: census_fwd_kernel<3><<<n, kCThreads, 0, st>>>(im_a, im_b, mask, hamming, partials, B, H, W, tiles_x, tiles_y, scale, want, eps, q); break;
        default: return ARF_EUNSUPPORTED;
    }
    ARF_CHECK_LAUNCH();
    if (want) {
        census_finalize_kernel<<<1, 256, 0, st>>>(partials, n, sums);
        ARF_CHECK_LAUNCH();
    }
    return ARF_OK;
}

extern "C" int arf_census_bwd(const float* im_a, const float* im_b, const float* ghamming, const float* hamming,
                              const float* mask, const float* sums, const float* gloss, float* g_a, float* g_b,
                              int B, int H, int W, int patch, float scale, float eps, float q, void* stream) {
    ARF_REQUIRE(im_a && im_b);
    ARF_REQUIRE(ghamming || (hamming && sums && gloss));
    ARF_REQUIRE(B > 0 && H > 0 && W > 0 && patch >= 1 && (patch & 1));
    if (!g_a && !g_b) return ARF_OK;
    const int tiles_x = arf_cdiv(W, kCTW), tiles_y = arf_cdiv(H, kCTH);
    const int n = arf_census_num_partials(B, H, W);
    if (n < 0) return n;
    cudaStream_t st = (cudaStream_t)stream;
    switch (patch / 2) {
        case 1:
            if (g_a && g_b) census_bwd_kernel<1, true, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            else if (g_b) census_bwd_kernel<1, false, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            else census_bwd_kernel<1, true, false><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            break;
        case 2:
            if (g_a && g_b) census_bwd_kernel<2, true, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            else if (g_b) census_bwd_kernel<2, false, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            else census_bwd_kernel<2, true, false><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            break;
        case 3:
            if (g_a && g_b) census_bwd_kernel<3, true, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            else if (g_b) census_bwd_kernel<3, false, true><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            else census_bwd_kernel<3, true, false><<<n, kCThreads, 0, st>>>(im_a, im_b, ghamming, hamming, mask, sums, gloss, g_a, g_b, B, H, W, tiles_x, tiles_y, scale, eps, q);
            break;
        default: return ARF_EUNSUPPORTED;
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

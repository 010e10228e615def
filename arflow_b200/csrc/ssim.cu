// Fused SSIM blocks.
//   ssim_loss  utils/uflow_utils.py:309-334   five AvgPool2d(P,1,P/2) (zeros in, divisor P*P), two maps
//              d1 = clamp(1-S1,0,1), d2 = clamp(1-S2,0,1), S1 = (2 mx my + C1)/(mx^2+my^2+C1), S2 = (2 sxy + C2)/(sx+sy+C2)
//   SSIM       losses/loss_blocks.py:65-84    five VALID AvgPool2d(P,1,0), one map clamp((1 - S1*S2)/2, 0, 1)
// Forward: the five box filters are evaluated together from one shared-memory tile pair (x, y) with a P/2
// halo; nothing but the final maps is written.  Backward: kernel 1 turns the upstream gradients into the
// five per-window coefficients (d/d mean_x, mean_y, E[xx], E[yy], E[xy], already divided by P*P), kernel 2
// gathers them over the windows that contain each input pixel (deterministic, no atomics).
#include "common.cuh"

namespace {

constexpr int kSTW = 32, kSTH = 8, kSThreads = 256;
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

struct SsimGeom {
    int planes, H, W, Ho, Wo, r, valid, mode;   // mode 0: two maps (uflow), 1: one map (loss_blocks)
};

template <int R>
__device__ __forceinline__ void load_tile(float (*t)[kSTW + 2 * R], const float* __restrict__ img, int x0, int y0,
                                          int H, int W) {
    constexpr int TW = kSTW + 2 * R, THh = kSTH + 2 * R;
    for (int e = threadIdx.x; e < TW * THh; e += kSThreads) {
        int xx = e % TW, yy = e / TW;
        int gx = x0 + xx, gy = y0 + yy;
        t[yy][xx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? __ldg(img + (size_t)gy * W + gx) : 0.f;
    }
}

struct Stats { float mx, my, sx, sy, sxy, exx, eyy, exy; };

template <int R>
__device__ __forceinline__ Stats window_stats(float (*tx)[kSTW + 2 * R], float (*ty)[kSTW + 2 * R], int ly, int lx) {
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int dy = 0; dy <= 2 * R; ++dy)
#pragma unroll
        for (int dx = 0; dx <= 2 * R; ++dx) {
            float u = tx[ly + dy][lx + dx], v = ty[ly + dy][lx + dx];
            a += u; b += v;
            aa = fmaf(u, u, aa); bb = fmaf(v, v, bb); ab = fmaf(u, v, ab);
        }
    const float inv = 1.f / (float)((2 * R + 1) * (2 * R + 1));
    Stats s;
    s.mx = a * inv; s.my = b * inv;
    s.exx = aa * inv; s.eyy = bb * inv; s.exy = ab * inv;
    s.sx = s.exx - s.mx * s.mx;
    s.sy = s.eyy - s.my * s.my;
    s.sxy = s.exy - s.mx * s.my;
    return s;
}

// tile origin in input coordinates of output pixel (ox, oy): same -> (ox - r, oy - r), valid -> (ox, oy)
template <int R>
__global__ void __launch_bounds__(kSThreads)
ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out1,
                float* __restrict__ out2, SsimGeom g, int tiles_x, int tiles_y) {
    __shared__ float tx[kSTH + 2 * R][kSTW + 2 * R];
    __shared__ float ty[kSTH + 2 * R][kSTW + 2 * R];
    const int tile = blockIdx.x;
    const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y, p = tile / (tiles_x * tiles_y);
    const int ox0 = txi * kSTW, oy0 = tyi * kSTH;
    const int off = g.valid ? 0 : -R;
    const float* xp = x + (size_t)p * g.H * g.W;
    const float* yp = y + (size_t)p * g.H * g.W;
    load_tile<R>(tx, xp, ox0 + off, oy0 + off, g.H, g.W);
    load_tile<R>(ty, yp, ox0 + off, oy0 + off, g.H, g.W);
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int ox = ox0 + lx, oy = oy0 + ly;
    if (ox >= g.Wo || oy >= g.Ho) return;
    Stats s = window_stats<R>(tx, ty, ly, lx);
    float S1 = (2.f * s.mx * s.my + kC1) / (s.mx * s.mx + s.my * s.my + kC1);
    float S2 = (2.f * s.sxy + kC2) / (s.sx + s.sy + kC2);
    size_t o = (size_t)p * g.Ho * g.Wo + (size_t)oy * g.Wo + ox;
    if (g.mode == 0) {
        out1[o] = fminf(fmaxf(1.f - S1, 0.f), 1.f);
        out2[o] = fminf(fmaxf(1.f - S2, 0.f), 1.f);
    } else {
        out1[o] = fminf(fmaxf((1.f - S1 * S2) * 0.5f, 0.f), 1.f);
    }
}

// coefficient planes (5 x planes x Ho x Wo): d loss / d (mx, my, exx, eyy, exy) of each window, / P^2
template <int R>
__global__ void __launch_bounds__(kSThreads)
ssim_coeff_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ g1,
                  const float* __restrict__ g2, float* __restrict__ coef, SsimGeom g, int tiles_x, int tiles_y) {
    __shared__ float tx[kSTH + 2 * R][kSTW + 2 * R];
    __shared__ float ty[kSTH + 2 * R][kSTW + 2 * R];
    const int tile = blockIdx.x;
    const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y, p = tile / (tiles_x * tiles_y);
    const int ox0 = txi * kSTW, oy0 = tyi * kSTH;
    const int off = g.valid ? 0 : -R;
    load_tile<R>(tx, x + (size_t)p * g.H * g.W, ox0 + off, oy0 + off, g.H, g.W);
    load_tile<R>(ty, y + (size_t)p * g.H * g.W, ox0 + off, oy0 + off, g.H, g.W);
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int ox = ox0 + lx, oy = oy0 + ly;
    if (ox >= g.Wo || oy >= g.Ho) return;
    Stats s = window_stats<R>(tx, ty, ly, lx);
    const float n1 = 2.f * s.mx * s.my + kC1, d1 = s.mx * s.mx + s.my * s.my + kC1;
    const float n2 = 2.f * s.sxy + kC2, d2 = s.sx + s.sy + kC2;
    const float S1 = n1 / d1, S2 = n2 / d2;
    const size_t plane = (size_t)g.planes * g.Ho * g.Wo;
    const size_t o = (size_t)p * g.Ho * g.Wo + (size_t)oy * g.Wo + ox;
    float gS1, gS2;   // d loss / d S1, d S2
    if (g.mode == 0) {
        float v1 = 1.f - S1, v2 = 1.f - S2;
        gS1 = (v1 >= 0.f && v1 <= 1.f) ? -__ldg(g1 + o) : 0.f;
        gS2 = (v2 >= 0.f && v2 <= 1.f) ? -__ldg(g2 + o) : 0.f;
    } else {
        float v = (1.f - S1 * S2) * 0.5f;
        float gu = (v >= 0.f && v <= 1.f) ? -0.5f * __ldg(g1 + o) : 0.f;
        gS1 = gu * S2;
        gS2 = gu * S1;
    }
    // S1(mx,my); S2(sxy,sx,sy) with sx = exx - mx^2, sy = eyy - my^2, sxy = exy - mx my
    const float dS1_dmx = (2.f * s.my * d1 - n1 * 2.f * s.mx) / (d1 * d1);
    const float dS1_dmy = (2.f * s.mx * d1 - n1 * 2.f * s.my) / (d1 * d1);
    const float dS2_dsxy = 2.f / d2, dS2_ds = -n2 / (d2 * d2);
    const float inv = 1.f / (float)((2 * R + 1) * (2 * R + 1));
    float c_mx = gS1 * dS1_dmx + gS2 * (dS2_dsxy * (-s.my) + dS2_ds * (-2.f * s.mx));
    float c_my = gS1 * dS1_dmy + gS2 * (dS2_dsxy * (-s.mx) + dS2_ds * (-2.f * s.my));
    coef[o] = c_mx * inv;
    coef[plane + o] = c_my * inv;
    coef[2 * plane + o] = gS2 * dS2_ds * inv;      // exx and eyy (the same coefficient)
    coef[3 * plane + o] = gS2 * dS2_dsxy * inv;    // exy
}

// grid = (column blocks, rows, planes strided); the (2R+1)^2 window walk is unrolled, four coefficient planes
// (d / d exx == d / d eyy), 32-bit offsets inside a plane
template <int R>
__global__ void __launch_bounds__(256)
ssim_gather_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ coef,
                   float* __restrict__ gx, float* __restrict__ gy, SsimGeom g) {
    const size_t plane = (size_t)g.planes * g.Ho * g.Wo;
    const int qx = blockIdx.x * blockDim.x + threadIdx.x, qy = blockIdx.y;
    if (qx >= g.W) return;
    // windows (output pixels) containing input pixel q: centres q-R..q+R (same) or top-left corners q-2R..q (valid)
    const int oy0 = g.valid ? qy - 2 * R : qy - R, ox0 = g.valid ? qx - 2 * R : qx - R;
    for (int p = blockIdx.z; p < g.planes; p += gridDim.z) {
        const float* cp = coef + (size_t)p * g.Ho * g.Wo;
        float a_mx = 0.f, a_my = 0.f, a_ss = 0.f, a_xy = 0.f;
#pragma unroll
        for (int dy = 0; dy <= 2 * R; ++dy) {
            const int oy = oy0 + dy;
            if (oy < 0 || oy >= g.Ho) continue;
#pragma unroll
            for (int dx = 0; dx <= 2 * R; ++dx) {
                const int ox = ox0 + dx;
                if (ox < 0 || ox >= g.Wo) continue;
                const int o = oy * g.Wo + ox;
                a_mx += __ldg(cp + o);
                a_my += __ldg(cp + plane + o);
                a_ss += __ldg(cp + 2 * plane + o);
                a_xy += __ldg(cp + 3 * plane + o);
            }
        }
        const size_t idx = ((size_t)p * g.H + qy) * g.W + qx;
        const float xv = __ldg(x + idx), yv = __ldg(y + idx);
        if (gx) gx[idx] = a_mx + 2.f * xv * a_ss + yv * a_xy;
        if (gy) gy[idx] = a_my + 2.f * yv * a_ss + xv * a_xy;
    }
}

int make_geom(SsimGeom& g, long long planes, int H, int W, int patch, int valid, int mode) {
    if (planes <= 0 || planes > 0x7fffffffLL || H <= 0 || W <= 0 || patch < 1 || !(patch & 1)) return ARF_EINVAL;
    if (mode < 0 || mode > 1) return ARF_EINVAL;
    g.planes = (int)planes; g.H = H; g.W = W; g.r = patch / 2; g.valid = valid ? 1 : 0; g.mode = mode;
    g.Ho = valid ? H - 2 * g.r : H;
    g.Wo = valid ? W - 2 * g.r : W;
    if (g.Ho <= 0 || g.Wo <= 0) return ARF_EINVAL;
    return ARF_OK;
}

}  // namespace

extern "C" int arf_ssim_fwd(const float* x, const float* y, float* out1, float* out2, long long planes, int H, int W,
                            int patch, int valid, int mode, void* stream) {
    ARF_REQUIRE(x && y && out1 && (mode == 1 || out2));
    SsimGeom g;
    int rc = make_geom(g, planes, H, W, patch, valid, mode);
    if (rc) return rc;
    const int tiles_x = arf_cdiv(g.Wo, kSTW), tiles_y = arf_cdiv(g.Ho, kSTH);
    const long long n = (long long)tiles_x * tiles_y * planes;
    if (n > 0x7fffffffLL) return ARF_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    switch (g.r) {
        case 1: ssim_fwd_kernel<1><<<(int)n, kSThreads, 0, st>>>(x, y, out1, out2, g, tiles_x, tiles_y); break;
        case 2: ssim_fwd_kernel<2><<<(int)n, kSThreads, 0, st>>>(x, y, out1, out2, g, tiles_x, tiles_y); break;
        case 3: ssim_fwd_kernel<3><<<(int)n, kSThreads, 0, st>>>(x, y, out1, out2, g, tiles_x, tiles_y); break;
        default: return ARF_EUNSUPPORTED;
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

/* coef: workspace of 5 * planes * Ho * Wo floats */
extern "C" int arf_ssim_bwd(const float* x, const float* y, const float* g1, const float* g2, float* coef, float* gx,
                            float* gy, long long planes, int H, int W, int patch, int valid, int mode, void* stream) {
    ARF_REQUIRE(x && y && g1 && coef && (mode == 1 || g2));
    if (!gx && !gy) return ARF_OK;
    SsimGeom g;
    int rc = make_geom(g, planes, H, W, patch, valid, mode);
    if (rc) return rc;
    const int tiles_x = arf_cdiv(g.Wo, kSTW), tiles_y = arf_cdiv(g.Ho, kSTH);
    const long long n = (long long)tiles_x * tiles_y * planes;
    if (n > 0x7fffffffLL) return ARF_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    switch (g.r) {
        case 1: ssim_coeff_kernel<1><<<(int)n, kSThreads, 0, st>>>(x, y, g1, g2, coef, g, tiles_x, tiles_y); break;
        case 2: ssim_coeff_kernel<2><<<(int)n, kSThreads, 0, st>>>(x, y, g1, g2, coef, g, tiles_x, tiles_y); break;
        case 3: ssim_coeff_kernel<3><<<(int)n, kSThreads, 0, st>>>(x, y, g1, g2, coef, g, tiles_x, tiles_y); break;
        default: return ARF_EUNSUPPORTED;
    }
    ARF_CHECK_LAUNCH();
    {
        const int bx = arf_cdiv(W, 256);
        long long bz = (16LL * ARF_NUM_SMS + (long long)bx * H - 1) / ((long long)bx * H);
        if (bz > planes) bz = planes;
        if (bz > 65535) bz = 65535;
        if (bz < 1) bz = 1;
        if (H > 65535) return ARF_EINVAL;
        dim3 grid((unsigned)bx, (unsigned)H, (unsigned)bz);
        switch (g.r) {
            case 1: ssim_gather_kernel<1><<<grid, 256, 0, st>>>(x, y, coef, gx, gy, g); break;
            case 2: ssim_gather_kernel<2><<<grid, 256, 0, st>>>(x, y, coef, gx, gy, g); break;
            case 3: ssim_gather_kernel<3><<<grid, 256, 0, st>>>(x, y, coef, gx, gy, g); break;
            default: return ARF_EUNSUPPORTED;
        }
    }
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// Channel-concatenation into a packed NHWC tensor, and its inverse (the "decoder concat" of SURVEY §8f item 4).
//
// cuDNN's Blackwell convolution kernels are NHWC kernels: handed NCHW activations it transposes every input and
// output itself (6.6 ms of a 24 ms chairs_uflow step, ncu launch list), and NHWC activations whose channel count is
// not a multiple of 4 go through a padding copy.  The PWC decoder therefore keeps its dense-block activations as
// packed NHWC tensors with the channel count rounded up to a multiple of 8; this file builds them:
//   pack:   dst[n, p, c_off + c] = src part (either NHWC [n, p, c] or NCHW [n, c, p]);  pad channels are zeroed
//   unpack: gsrc part            = gdst[n, p, c_off + c]                  (backward of pack, same two formats)
// replacing torch.cat (models/uflow_model.py:189-205) and CatBackward.  NCHW parts (the cost volume and the flow
// come from the NCHW hot-path kernels) go through a 32-pixel x 32-channel shared-memory transpose so that both the
// reads (along pixels) and the writes (along channels) are coalesced.
#include "common.cuh"

namespace {

constexpr int kLThreads = 256;

// NHWC part <-> NHWC destination slice: rows = N*HW pixels, part width Cs, destination width Cd, offset c_off.
template <bool kPack, bool kAcc = false>
__global__ void __launch_bounds__(kLThreads)
nhwc_part_kernel(float* __restrict__ dst, const float* __restrict__ src, long long rows, int Cs, int Cd, int c_off, int vec) {
    // kPack:  dst[row*Cd + c_off + c] = src[row*Cs + c]     (src may be null: zero fill)
    // !kPack: dst is the PART (width Cs), src the packed tensor (width Cd)
    if (vec) {
        const int q = Cs >> 2;
        const long long total = rows * q;
        for (long long e = blockIdx.x * (long long)kLThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kLThreads) {
            const long long row = e / q;
            const int c = (int)(e - row * q) << 2;
            if (kPack) {
                float4 v = src ? *reinterpret_cast<const float4*>(src + row * Cs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(dst + row * Cd + c_off + c) = v;
            } else {
                float4 v = *reinterpret_cast<const float4*>(src + row * Cd + c_off + c);
                if (kAcc) {     // unpack-accumulate: part += packed slice (gradient of a tensor that is also a conv input)
                    const float4 o = *reinterpret_cast<const float4*>(dst + row * Cs + c);
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                *reinterpret_cast<float4*>(dst + row * Cs + c) = v;
            }
        }
    } else {
        const long long total = rows * Cs;
        for (long long e = blockIdx.x * (long long)kLThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kLThreads) {
            const long long row = e / Cs;
            const int c = (int)(e - row * Cs);
            if (kPack) dst[row * Cd + c_off + c] = src ? src[row * Cs + c] : 0.f;
            else dst[row * Cs + c] = (kAcc ? dst[row * Cs + c] : 0.f) + src[row * Cd + c_off + c];
        }
    }
}

// NCHW part [n][c][p] <-> packed NHWC [n][p][Cd] slice, through a 32 x 32 (+1) shared-memory tile.
// grid = (pixel tiles, channel tiles, n), block = (32, 8).
// batch_shift: the PART's batch index is (n + batch_shift) % N — the two flow directions of a stacked batch read
// each other's features, so the half-batch swap rides on this copy instead of a torch.cat (and, in the backward, a
// zero-fill + slice copy + add).
// slope != 1: a leaky ReLU rides on the copy (the cost volume's activation, models/uflow_model.py:185-186): pack writes
// leaky(src); unpack returns grad * (fwd > 0 ? 1 : slope) with fwd = the packed forward tensor (same addressing).
template <bool kPack>
__global__ void __launch_bounds__(256)
nchw_part_kernel(float* __restrict__ dst, const float* __restrict__ src, long long HW, int Cs, int Cd, int c_off,
                 int batch_shift, float slope = 1.f, const float* __restrict__ fwd = nullptr) {
    __shared__ float tile[32][33];
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const long long n = blockIdx.z;
    const long long n_part = (n + batch_shift) % gridDim.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (kPack) {
        // read: lanes along pixels of one channel
        const float* s = src + n_part * Cs * HW;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + ty + 8 * k;
            const long long p = p0 + tx;
            float v = (c < Cs && p < HW) ? __ldg(s + (long long)c * HW + p) : 0.f;
            tile[ty + 8 * k][tx] = v > 0.f ? v : v * slope;
        }
        __syncthreads();
        // write: lanes along channels of one pixel
        float* d = dst + n * HW * Cd + c_off;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long p = p0 + ty + 8 * k;
            const int c = c0 + tx;
            if (c < Cs && p < HW) d[p * Cd + c] = tile[tx][ty + 8 * k];
        }
    } else {
        // dst = NCHW part, src = packed NHWC
        const float* s = src + n * HW * Cd + c_off;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long p = p0 + ty + 8 * k;
            const int c = c0 + tx;
            float v = (c < Cs && p < HW) ? __ldg(s + p * Cd + c) : 0.f;
            if (fwd && c < Cs && p < HW && !(__ldg(fwd + n * HW * Cd + c_off + p * Cd + c) > 0.f)) v *= slope;
            tile[ty + 8 * k][tx] = v;
        }
        __syncthreads();
        float* d = dst + n_part * Cs * HW;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + ty + 8 * k;
            const long long p = p0 + tx;
            if (c < Cs && p < HW) d[(long long)c * HW + p] = tile[tx][ty + 8 * k];
        }
    }
}

// NCHW parts with a handful of channels (the 3-channel image, the 2-channel flow): thread <-> pixel, the reads of one
// channel are coalesced across the warp, the Cs values of a pixel are written next to each other.  (The 32 x 32
// transpose above would leave 29 of 32 channel lanes idle: 143 us for the 16 x 3 x 384 x 512 image pack.)
template <bool kPack, int kMaxC>
__global__ void __launch_bounds__(kLThreads)
nchw_smallc_kernel(float* __restrict__ dst, const float* __restrict__ src, long long N, long long HW, int Cs, int Cd, int c_off) {
    const long long total = N * HW;
    for (long long e = blockIdx.x * (long long)kLThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kLThreads) {
        const long long n = e / HW, p = e - n * HW;
        if (kPack) {
            const float* s = src + n * Cs * HW + p;
            float* d = dst + e * Cd + c_off;
            float v[kMaxC];
#pragma unroll
            for (int c = 0; c < kMaxC; ++c) v[c] = c < Cs ? __ldg(s + (long long)c * HW) : 0.f;
#pragma unroll
            for (int c = 0; c < kMaxC; ++c)
                if (c < Cs) d[c] = v[c];
        } else {
            const float* s = src + e * Cd + c_off;
            float* d = dst + n * Cs * HW + p;
            float v[kMaxC];
#pragma unroll
            for (int c = 0; c < kMaxC; ++c) v[c] = c < Cs ? __ldg(s + c) : 0.f;
#pragma unroll
            for (int c = 0; c < kMaxC; ++c)
                if (c < Cs) d[(long long)c * HW] = v[c];
        }
    }
}

// The networks' input stage in one pass: (B, 2C, H, W) image pairs -> the stacked batch [first images; second images]
// as packed NHWC with 8 channels, value * scale + shift on the C real channels and zeros behind them
// (torch.cat of the two slices, x * 2 - 1, the 3 -> 8 channel pack and its zero tail were five launches, 160 us at
// 8 x 6 x 384 x 512; one read of 75 MB and one write of 201 MB is 45 us).
__global__ void __launch_bounds__(kLThreads)
image_pair_pack_kernel(float* __restrict__ dst, const float* __restrict__ src, long long B, long long HW, int C, float scale,
                       float shift) {
    const long long total = 2 * B * HW;
    for (long long e = blockIdx.x * (long long)kLThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kLThreads) {
        const long long img = e / HW, p = e - img * HW;
        const long long d = img / B, b = img - d * B;             // d: first / second image of the pair
        const float* s = src + (b * 2 * C + d * C) * HW + p;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = c < C ? fmaf(__ldg(s + (long long)c * HW), scale, shift) : 0.f;
        float4* o = reinterpret_cast<float4*>(dst + e * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

template <bool kPack>
int launch_part(float* dst, const float* src, long long N, long long HW, int Cs, int Cd, int c_off, int src_nhwc,
                cudaStream_t st) {
    if (N <= 0 || HW <= 0 || Cs <= 0 || Cd <= 0 || c_off < 0 || c_off + Cs > Cd) return ARF_EINVAL;
    const float* part = kPack ? src : dst;       // the tensor that has the part's own layout
    if (src_nhwc || (kPack && !src)) {
        const float* packed = kPack ? dst : src;
        const int vec = (Cs % 4 == 0) && (Cd % 4 == 0) && (c_off % 4 == 0) && ((uintptr_t)packed % 16 == 0) &&
                        (!part || (uintptr_t)part % 16 == 0);
        const long long work = N * HW * (vec ? Cs / 4 : Cs);
        nhwc_part_kernel<kPack><<<arf_grid_1d(work, kLThreads, 16), kLThreads, 0, st>>>(dst, src, N * HW, Cs, Cd, c_off, vec);
    } else if (Cs <= 4) {
        nchw_smallc_kernel<kPack, 4><<<arf_grid_1d(N * HW, kLThreads, 16), kLThreads, 0, st>>>(dst, src, N, HW, Cs, Cd, c_off);
    } else {
        if (N > 65535 || (Cs + 31) / 32 > 65535) return ARF_EINVAL;
        dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cs + 31) / 32), (unsigned)N);
        nchw_part_kernel<kPack><<<grid, dim3(32, 8), 0, st>>>(dst, src, HW, Cs, Cd, c_off, 0);
    }
    return ARF_OK;
}

}  // namespace

// Weight re-layout for the channels-last conv stacks: dst (Co_d, Ci_d, KH, KW) in channels-last memory order
// [co][kh][kw][ci] <- src (Co_s, Ci_s, KH, KW) with arbitrary strides, zero input channels inserted at up to four
// positions and zero output channels appended (to_padded = 1); or the inverse gather (to_padded = 0: the weight
// gradient back in the parameter's own layout, inserted rows / columns dropped).  One launch per layer and
// direction instead of zeros + cat + contiguous (+ their autograd counterparts).
struct PadMap {
    int n;           // number of insertions
    int at[4];       // position in the ORIGINAL input-channel order (increasing)
    int cnt[4];      // zero channels inserted there
};

__global__ void __launch_bounds__(256)
pad_weight_kernel(float* __restrict__ dst, const float* __restrict__ src, int Co_s, int Ci_s, int KH, int KW, int Co_d,
                  int Ci_d, long long s_co, long long s_ci, long long s_kh, long long s_kw, PadMap pm, int to_padded) {
    // iterate over the PADDED index space; (co, kh, kw, cip) with cip fastest (the padded tensor's memory order)
    const long long total = (long long)Co_d * KH * KW * Ci_d;
    for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int cip = (int)(e % Ci_d);
        long long t = e / Ci_d;
        const int kw = (int)(t % KW); t /= KW;
        const int kh = (int)(t % KH);
        const int co = (int)(t / KH);
        // padded input channel -> original input channel (or -1 inside an inserted block)
        int ci = cip, shift = 0;
        bool zero = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < pm.n) {
                const int start = pm.at[k] + shift;          // first padded index of the inserted block
                if (cip >= start + pm.cnt[k]) { shift += pm.cnt[k]; }
                else if (cip >= start) { zero = true; }
            }
        }
        ci = cip - shift;
        zero = zero || co >= Co_s || ci >= Ci_s;
        const long long so = (long long)co * s_co + (long long)ci * s_ci + (long long)kh * s_kh + (long long)kw * s_kw;
        if (to_padded) dst[e] = zero ? 0.f : __ldg(src + so);
        else if (!zero) dst[so] = __ldg(src + e);             // dst = original layout, src = padded
    }
}

extern "C" int arf_pad_weight(float* dst, const float* src, int Co, int Ci, int KH, int KW, int Co_pad, int Ci_pad,
                              long long s_co, long long s_ci, long long s_kh, long long s_kw, int n_pads,
                              const int* pad_at, const int* pad_cnt, int to_padded, void* stream) {
    ARF_REQUIRE(dst && src);
    ARF_REQUIRE(Co > 0 && Ci > 0 && KH > 0 && KW > 0 && Co_pad >= Co && Ci_pad >= Ci && n_pads >= 0 && n_pads <= 4);
    PadMap pm;
    pm.n = n_pads;
    int extra = 0;
    for (int k = 0; k < 4; ++k) {
        pm.at[k] = k < n_pads ? pad_at[k] : 0;
        pm.cnt[k] = k < n_pads ? pad_cnt[k] : 0;
        if (k < n_pads) {
            if (pad_at[k] < 0 || pad_at[k] > Ci || pad_cnt[k] < 0 || (k > 0 && pad_at[k] < pad_at[k - 1])) return ARF_EINVAL;
            extra += pad_cnt[k];
        }
    }
    if (Ci + extra != Ci_pad) return ARF_EINVAL;
    const long long total = (long long)Co_pad * KH * KW * Ci_pad;
    pad_weight_kernel<<<arf_grid_1d(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(dst, src, Co, Ci, KH, KW, Co_pad, Ci_pad, s_co,
                                                                                   s_ci, s_kh, s_kw, pm, to_padded);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_nhwc_pack(float* dst, const float* src, long long N, long long HW, int Cs, int Cd, int c_off,
                             int src_nhwc, void* stream) {
    ARF_REQUIRE(dst);
    int rc = launch_part<true>(dst, src, N, HW, Cs, Cd, c_off, src_nhwc, (cudaStream_t)stream);
    if (rc) return rc;
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_nhwc_unpack(float* part, const float* packed, long long N, long long HW, int Cs, int Cd, int c_off,
                               int part_nhwc, void* stream) {
    ARF_REQUIRE(part && packed);
    int rc = launch_part<false>(part, packed, N, HW, Cs, Cd, c_off, part_nhwc, (cudaStream_t)stream);
    if (rc) return rc;
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// NCHW tensor (N,C,HW) <-> channels-last tensor (N,HW,C) with the NCHW side's batch rotated by batch_shift:
// to_nchw = 1: nchw[(n + shift) % N] = nhwc[n];   to_nchw = 0: nhwc[n] = nchw[(n + shift) % N].
extern "C" int arf_nhwc_transpose(float* dst, const float* src, long long N, long long HW, int C, int to_nchw,
                                  int batch_shift, void* stream) {
    ARF_REQUIRE(dst && src);
    if (N <= 0 || N > 65535 || HW <= 0 || C <= 0 || (C + 31) / 32 > 65535 || batch_shift < 0) return ARF_EINVAL;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
    if (to_nchw) nchw_part_kernel<false><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(dst, src, HW, C, C, 0, batch_shift);
    else nchw_part_kernel<true><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(dst, src, HW, C, C, 0, batch_shift);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// NCHW part with a leaky ReLU fused into the copy (see nchw_part_kernel)
extern "C" int arf_nhwc_pack_act(float* dst, const float* src, long long N, long long HW, int Cs, int Cd, int c_off,
                                 float slope, void* stream) {
    ARF_REQUIRE(dst && src);
    if (N <= 0 || N > 65535 || HW <= 0 || Cs <= 0 || Cd <= 0 || c_off < 0 || c_off + Cs > Cd || (Cs + 31) / 32 > 65535) return ARF_EINVAL;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cs + 31) / 32), (unsigned)N);
    nchw_part_kernel<true><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(dst, src, HW, Cs, Cd, c_off, 0, slope, nullptr);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_nhwc_unpack_act(float* part, const float* packed_grad, const float* packed_fwd, long long N, long long HW,
                                   int Cs, int Cd, int c_off, float slope, void* stream) {
    ARF_REQUIRE(part && packed_grad && packed_fwd);
    if (N <= 0 || N > 65535 || HW <= 0 || Cs <= 0 || Cd <= 0 || c_off < 0 || c_off + Cs > Cd || (Cs + 31) / 32 > 65535) return ARF_EINVAL;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cs + 31) / 32), (unsigned)N);
    nchw_part_kernel<false><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(part, packed_grad, HW, Cs, Cd, c_off, 0, slope,
                                                                            packed_fwd);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// part (NHWC, width Cs) += packed[..., c_off : c_off + Cs]
extern "C" int arf_nhwc_unpack_add(float* part, const float* packed, long long N, long long HW, int Cs, int Cd, int c_off,
                                   void* stream) {
    ARF_REQUIRE(part && packed);
    if (N <= 0 || HW <= 0 || Cs <= 0 || Cd <= 0 || c_off < 0 || c_off + Cs > Cd) return ARF_EINVAL;
    const int vec = (Cs % 4 == 0) && (Cd % 4 == 0) && (c_off % 4 == 0) && ((uintptr_t)packed % 16 == 0) && ((uintptr_t)part % 16 == 0);
    const long long work = N * HW * (vec ? Cs / 4 : Cs);
    nhwc_part_kernel<false, true><<<arf_grid_1d(work, kLThreads, 16), kLThreads, 0, (cudaStream_t)stream>>>(
        part, packed, N * HW, Cs, Cd, c_off, vec);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

extern "C" int arf_image_pair_pack(float* dst, const float* src, long long B, long long HW, int C, int Cd, float scale,
                                   float shift, void* stream) {
    ARF_REQUIRE(dst && src && B > 0 && HW > 0 && C > 0);
    if (Cd != 8 || C > 8 || (uintptr_t)dst % 16 != 0) return ARF_EUNSUPPORTED;
    image_pair_pack_kernel<<<arf_grid_1d(2 * B * HW, kLThreads, 16), kLThreads, 0, (cudaStream_t)stream>>>(dst, src, B, HW, C,
                                                                                                        scale, shift);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// Shared device/host helpers for the arflow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/arflow_b200.h"

#define ARF_NUM_SMS 148

// number of kernels this library has launched in this process (bench.py reports it as gpu_launches)
extern long long g_arf_launches;

#define ARF_CHECK_LAUNCH()                                   \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return (int)e__;             \
        ++g_arf_launches;                                    \
    } while (0)

#define ARF_REQUIRE(cond)                                    \
    do {                                                     \
        if (!(cond)) return ARF_EINVAL;                      \
    } while (0)

static inline int arf_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Grid size for grid-stride kernels: enough CTAs to fill every SM a few times, never more
// than the work needs.
static inline int arf_grid_1d(long long work_items, int block, int ctas_per_sm = 8) {
    long long need = (work_items + block - 1) / block;
    long long cap = (long long)ARF_NUM_SMS * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float arf_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024). Result valid in thread 0.
__device__ __forceinline__ float arf_block_sum(float v, float* smem32) {
    v = arf_warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    float r = 0.f;
    if (w == 0) {
        r = lane < nw ? smem32[lane] : 0.f;
        r = arf_warp_sum(r);
    }
    __syncthreads();
    return r;
}

__device__ __forceinline__ float arf_ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// idx -> (x, y, b) for a (B, H, W) index space.  Indices below 2^32 (every shape the networks use) take 32-bit
// divisions; a 64-bit div/mod pair per element costs more than the rest of an elementwise kernel.
__device__ __forceinline__ void arf_split3(long long idx, int W, int H, int& x, int& y, int& b) {
    if (idx <= 0xffffffffLL) {
        const unsigned u = (unsigned)idx, t = u / (unsigned)W, bb = t / (unsigned)H;
        x = (int)(u - t * (unsigned)W);
        y = (int)(t - bb * (unsigned)H);
        b = (int)bb;
    } else {
        const long long t = idx / W;
        x = (int)(idx - t * W);
        y = (int)(t % H);
        b = (int)(t / H);
    }
}

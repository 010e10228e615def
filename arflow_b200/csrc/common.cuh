// Shared device/host helpers for the arflow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/arflow_b200.h"

#include <atomic>

// Multiprocessor count of the CURRENT device (148 on a B200), queried once per device: grids are sized from it,
// and a process may drive several devices (the reference's DataParallel does).
#define ARF_NUM_SMS arf_num_sms()
int arf_num_sms();

// number of kernels this library has launched in this process (bench.py reports it as gpu_launches)
extern std::atomic<long long> g_arf_launches;

// Peek, do not get: an error left behind by earlier, unrelated work (cuDNN, NCCL) is reported but stays latched
// for its owner instead of being cleared here.
#define ARF_CHECK_LAUNCH()                                   \
    do {                                                     \
        cudaError_t e__ = cudaPeekAtLastError();             \
        if (e__ != cudaSuccess) return (int)e__;             \
        g_arf_launches.fetch_add(1, std::memory_order_relaxed); \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to (function, device), so the
// "already done" flag is one bit per device ordinal (devices >= 64 set it on every launch); thread-safe.
template <typename K>
static inline cudaError_t arf_ensure_smem(K kern, size_t bytes, std::atomic<unsigned long long>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
    return e;
}
#define ARF_ENSURE_SMEM(kern, bytes)                                        \
    do {                                                                    \
        static std::atomic<unsigned long long> done__{0};                   \
        cudaError_t e__ = arf_ensure_smem(kern, bytes, done__);             \
        if (e__ != cudaSuccess) return (int)e__;                            \
    } while (0)

// Kernel-selection / probe hooks of arf_debug_set (the parity tests force each kernel variant through them).  The in-tree
// build keeps them (ARF_TEST_HOOKS=1); a release build compiles them out: `ARF_RELEASE=1 python -m arflow_b200.build`
// passes -DARF_TEST_HOOKS=0, every hook becomes the constant 0, the variant branches fold away and arf_debug_set returns
// ARF_EUNSUPPORTED.
#ifndef ARF_TEST_HOOKS
#define ARF_TEST_HOOKS 1
#endif
#if ARF_TEST_HOOKS
#define ARF_HOOK thread_local int
#else
#define ARF_HOOK static constexpr int
#endif

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

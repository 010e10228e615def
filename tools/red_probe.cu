// Global fp32 reduction throughput probe (design input for the warp source-gradient flush): scalar red.global.add.f32
// against the vector form red.global.add.v4.f32 (sm_90+), both fully coalesced over a buffer larger than a pass needs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_probe tools/red_probe.cu && ./tools/red_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void red1(float* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + i), "f"(1.0f) : "memory");
}
__global__ void red4(float* p, size_t n) {
    for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4; i < n; i += (size_t)gridDim.x * blockDim.x * 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p + i), "f"(1.0f) : "memory");
}
__global__ void red2(float* p, size_t n) {
    for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2; i < n; i += (size_t)gridDim.x * blockDim.x * 2)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(p + i), "f"(1.0f) : "memory");
}
__global__ void st1(float* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1.0f;
}
// the pattern of the direct kernel: every lane adds to element i and i + 1 (two reds, the second shifted by one float)
__global__ void red1x2(float* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i + 1 < n; i += (size_t)gridDim.x * blockDim.x) {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + i), "f"(1.0f) : "memory");
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + i + 1), "f"(1.0f) : "memory");
    }
}

// TMA bulk reduction: every CTA adds a shared-memory tile of `tile_bytes` into its own global regions, chunk by chunk
// (cp.reduce.async.bulk.global.shared::cta.add.f32) - the flush a shared-memory accumulation window would use.
__global__ void red_bulk(float* p, size_t n, int tile_floats) {
    extern __shared__ __align__(128) float tile[];
    for (int i = threadIdx.x; i < tile_floats; i += blockDim.x) tile[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(tile);
        int k = 0;
        for (size_t i = (size_t)blockIdx.x * tile_floats; i + tile_floats <= n; i += (size_t)gridDim.x * tile_floats) {
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                         ::"l"(p + i), "r"(src), "r"(tile_floats * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++k % 8 == 0) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(s); f(); cudaEventRecord(e); cudaEventSynchronize(e);
        float ms; cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t n = 64ull << 20;   // 64 M floats = 256 MB (> L2)
    float* p;
    if (cudaMalloc(&p, n * 4) != cudaSuccess) return 1;
    cudaMemset(p, 0, n * 4);
    const int blocks = 148 * 8, threads = 256;
    float a = time_ms([&] { red1<<<blocks, threads>>>(p, n); });
    float b = time_ms([&] { red4<<<blocks, threads>>>(p, n); });
    float b2 = time_ms([&] { red2<<<blocks, threads>>>(p, n); });
    float c = time_ms([&] { st1<<<blocks, threads>>>(p, n); });
    float d = time_ms([&] { red1x2<<<blocks, threads>>>(p, n); });
    const size_t m = 6ull << 20;    // 24 MB: L2-resident target (the warp gradient of one level)
    float a2 = time_ms([&] { red1<<<blocks, threads>>>(p, m); });
    float b3 = time_ms([&] { red4<<<blocks, threads>>>(p, m); });
    float d2 = time_ms([&] { red1x2<<<blocks, threads>>>(p, m); });
    float e1 = time_ms([&] { red_bulk<<<blocks, 128, 4096>>>(p, n, 1024); });
    float e2 = time_ms([&] { red_bulk<<<blocks, 128, 512>>>(p, n, 128); });
    float e3 = time_ms([&] { red_bulk<<<blocks, 128, 4096>>>(p, m, 1024); });
    printf("{\"bulk_reduce_4KB_Gelem_s\": %.1f, \"bulk_reduce_512B_Gelem_s\": %.1f, \"l2_bulk_reduce_4KB_Gelem_s\": %.1f}\n",
           n / e1 / 1e6, n / e2 / 1e6, m / e3 / 1e6);
    printf("{\"red_f32_Gelem_s\": %.1f, \"red_v2_Gelem_s\": %.1f, \"red_v4_Gelem_s\": %.1f, \"st_f32_Gelem_s\": %.1f, "
           "\"red_f32_pair_Gadds_s\": %.1f, \"l2_red_f32_Gelem_s\": %.1f, \"l2_red_v4_Gelem_s\": %.1f, "
           "\"l2_red_f32_pair_Gadds_s\": %.1f}\n",
           n / a / 1e6, n / b2 / 1e6, n / b / 1e6, n / c / 1e6, 2.0 * n / d / 1e6, m / a2 / 1e6, m / b3 / 1e6, 2.0 * m / d2 / 1e6);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}

// Shared-memory wavefront probe: how many crossbar passes does one warp-wide LDS.{32,64,128} take when lanes share
// addresses?  (Design input for the correlation kernels' operand layout; B300_MICROARCH.md only states the
// 128/N B/cyc/SM rule for 32-bit accesses.)  One CTA of 8 warps on one SM; every warp issues independent loads in a
// loop; cycles / (warp-level loads) = crossbar passes per instruction once the pipe is saturated.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lds_probe tools/lds_probe.cu && ./tools/lds_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kBytes>
__device__ __forceinline__ void lds(uint32_t addr, float (&v)[4]) {
    if (kBytes == 16)
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr) : "memory");
    else if (kBytes == 8)
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr) : "memory");
    else
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(addr) : "memory");
}

template <int kBytes>
__global__ void probe(const int* __restrict__ chunk_of_lane, int iters, float* sink, long long* cycles, int zero_mask) {
    __shared__ __align__(16) float buf[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(buf) + chunk_of_lane[lane] * kBytes;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float v[8][4] = {};
#pragma unroll
        for (int u = 0; u < 8; ++u) lds<kBytes>(base + u * 2000, v[u]);   // 8 independent loads, same bank pattern (2000 = 16 * 125 keeps the alignment, moves the banks)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u][0] + v[u][1] + v[u][2] + v[u][3];
        base += __float_as_int(v[7][0]) & zero_mask;   // run-time zero: the next iteration's addresses depend on this one's data
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    if (acc == 1234.5f) *sink = acc;
}

template <int kBytes>
double run(const int* h_pat, int warps) {
    int* d_pat; float* sink; long long* cyc;
    cudaMalloc(&d_pat, 32 * sizeof(int)); cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
    cudaMemcpy(d_pat, h_pat, 32 * sizeof(int), cudaMemcpyHostToDevice);
    const int iters = 4096;
    probe<kBytes><<<1, 32 * warps>>>(d_pat, 64, sink, cyc, 0);
    probe<kBytes><<<1, 32 * warps>>>(d_pat, iters, sink, cyc, 0);
    long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_pat); cudaFree(sink); cudaFree(cyc);
    return (double)c / ((double)iters * 8 * warps);
}

int main() {
    struct Pat { const char* name; int bytes; int (*f)(int); };
    Pat pats[] = {
        {"b128 distinct (lane)", 16, [](int l) { return l; }},
        {"b128 all same", 16, [](int) { return 0; }},
        {"b128 lane&7 (quarters identical)", 16, [](int l) { return l & 7; }},
        {"b128 lane>>2 (8 distinct, 2 per quarter)", 16, [](int l) { return l >> 2; }},
        {"b128 lane>>1 (16 distinct)", 16, [](int l) { return l >> 1; }},
        {"b128 lane&15 (halves identical)", 16, [](int l) { return l & 15; }},
        {"b128 lane&3 (4 distinct)", 16, [](int l) { return l & 3; }},
        {"b128 (lane&3)+4*(lane>>4) (8 distinct)", 16, [](int l) { return (l & 3) + 4 * (l >> 4); }},
        {"b128 lane>>3 (4 distinct, 1 per quarter)", 16, [](int l) { return l >> 3; }},
        {"b128 stride 2 chunks (lane*2): 2-way conflict", 16, [](int l) { return l * 2; }},
        {"b64 distinct (lane)", 8, [](int l) { return l; }},
        {"b64 all same", 8, [](int) { return 0; }},
        {"b64 lane&15 (halves identical)", 8, [](int l) { return l & 15; }},
        {"b64 lane>>1 (16 distinct)", 8, [](int l) { return l >> 1; }},
        {"b64 lane>>2 (8 distinct)", 8, [](int l) { return l >> 2; }},
        {"b32 distinct (lane)", 4, [](int l) { return l; }},
        {"b32 all same", 4, [](int) { return 0; }},
        {"b32 stride 2 words: 2-way conflict", 4, [](int l) { return l * 2; }},
    };
    printf("[\n");
    const int n = sizeof(pats) / sizeof(pats[0]);
    for (int i = 0; i < n; ++i) {
        int h[32];
        for (int l = 0; l < 32; ++l) h[l] = pats[i].f(l);
        double c8 = pats[i].bytes == 16 ? run<16>(h, 8) : pats[i].bytes == 8 ? run<8>(h, 8) : run<4>(h, 8);
        double c16 = pats[i].bytes == 16 ? run<16>(h, 16) : pats[i].bytes == 8 ? run<8>(h, 16) : run<4>(h, 16);
        printf(" {\"pattern\": \"%s\", \"cycles_per_warp_load_8w\": %.2f, \"cycles_per_warp_load_16w\": %.2f}%s\n", pats[i].name,
               c8, c16, i + 1 < n ? "," : "");
    }
    printf("]\n");
    return cudaDeviceSynchronize() != cudaSuccess;
}

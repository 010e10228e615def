// FP32 FMA and MUFU micro-peaks (roofline denominators for the FMA-bound correlation and the
// MUFU-bound census kernels).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void fma_kernel(float* out, float a, float b, int iters) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], b, a);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// a*b+c with three distinct, changing register operands (like the correlation inner loop)
template <int ILP>
__global__ void fma3_kernel(float* out, const float* in, int iters) {
    float acc[ILP], x[8], y[8];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 8 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(x[i & 7], y[(i >> 3) & 7], acc[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += 1.0f;   // keep operands live and changing
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// Packed fma.rn.f32x2 (FFMA2, sm_100+): two FMAs per lane per instruction.  Same structure as fma3_kernel:
// 64 accumulators = 32 register pairs, distinct changing operands.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
template <int PAIRS>
__global__ void fma2_kernel(float* out, const float* in, int iters) {
    unsigned long long acc[PAIRS];
    float x[8], y[8];
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) acc[i] = 0ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 8 + i]; }
    for (int it = 0; it < iters; ++it) {
        unsigned long long xp[4], yp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { xp[i] = pack2(x[2 * i], x[2 * i + 1]); yp[i] = pack2(y[2 * i], y[2 * i + 1]); }
#pragma unroll
        for (int i = 0; i < PAIRS; ++i) ffma2(acc[i], xp[i & 3], yp[(i >> 2) & 3]);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += 1.0f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void mufu_kernel(float* out, float a, int iters) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 1.0f + threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = rsqrtf(acc[i]) + a;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    f(); f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(s); f(); cudaEventRecord(e); cudaEventSynchronize(e);
        float ms; cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float *out, *in;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&in, 4096 * sizeof(float));
    cudaMemset(in, 0, 4096 * sizeof(float));
    double n = (double)blocks * threads * iters;
    float ms = time_ms([&] { fma_kernel<8><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
    printf("{\"kernel\":\"ffma_2reg_ilp8\",\"tflops\":%.2f}\n", n * 16 * 2 / ms / 1e9);
    ms = time_ms([&] { fma_kernel<16><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
    printf("{\"kernel\":\"ffma_2reg_ilp16\",\"tflops\":%.2f}\n", n * 32 * 2 / ms / 1e9);
    ms = time_ms([&] { fma3_kernel<64><<<blocks, threads>>>(out, in, iters); });
    printf("{\"kernel\":\"ffma_3reg_ilp64\",\"tflops\":%.2f}\n", n * 64 * 2 / ms / 1e9);
    ms = time_ms([&] { fma2_kernel<32><<<blocks, threads>>>(out, in, iters); });
    printf("{\"kernel\":\"ffma2_packed_32pairs\",\"tflops\":%.2f}\n", n * 64 * 2 / ms / 1e9);
    ms = time_ms([&] { mufu_kernel<<<blocks, threads>>>(out, 0.5f, iters); });
    printf("{\"kernel\":\"mufu_rsq\",\"gops\":%.1f}\n", n * 8 / ms / 1e6);
    return 0;
}

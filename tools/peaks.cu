// Micro-peaks of one B200 that MEASURED_PEAKS.json does not carry: the roofline denominators of the FMA-bound
// correlation kernels (FP32 FMA, scalar and packed), the MUFU-bound census kernels (MUFU.RSQ), the shared-memory
// crossbar that bounds the FP32 correlation loop (LDS.32 wavefronts) and the legacy tensor path (mma.sync TF32),
// which settles whether a split-TF32 correlation variant can pay.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/libarf_peaks.so tools/peaks.cu
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DARF_PEAKS_MAIN -o tools/peaks tools/peaks.cu
//
// Tool / measurement infrastructure: bench.py and tools/microbench.py load the .so through ctypes and call
// arf_peaks_measure() once per run, so every fraction they print is against a number measured in the same process
// on the same GPU.  Not part of libarflow_b200.so.
#include <cstdio>
#include <cuda_runtime.h>

namespace {

// a*b+c with three distinct, changing register operands (the shape of the correlation inner loop)
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
        for (int i = 0; i < 8; ++i) x[i] += 1.0f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

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
        for (int i = 0; i < 8; ++i) {
            float r;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i]));
            acc[i] = r + a;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// a*acc+b with two loop-invariant operands (the friendliest register pattern: the upper bound of the FMA pipe)
template <int ILP>
__global__ void fma2op_kernel(float* out, float a, float b, int iters) {
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

// conflict-free LDS.32: every warp-wide load is one 128-byte wavefront (volatile: the loads stay in the loop)
__global__ void lds_kernel(float* out, int iters) {
    __shared__ volatile float buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = (float)i;
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    int o = threadIdx.x & 1023;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += buf[(o + 32 * i) & 2047];
        o = (o + 256) & 1023;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// legacy tensor path: mma.sync.m16n8k8 TF32, NT independent accumulator tiles per warp
template <int NT>
__global__ void mma_tf32_kernel(float* out, int iters) {
    float d[NT][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[t][i] = 0.f;
    unsigned a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    b[0] = __float_as_uint(0.5f);
    b[1] = __float_as_uint(0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[t][0]), "+f"(d[t][1]), "+f"(d[t][2]), "+f"(d[t][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += d[t][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// fp32 atomicAdd on shared memory, spread addresses (one bank per lane)
__global__ void atoms_kernel(float* out, int iters) {
    __shared__ float buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = 0.f;
    __syncthreads();
    int o = threadIdx.x & 1023;
    for (int it = 0; it < iters; ++it) {
        atomicAdd(&buf[o], 1.0f);
        o = (o + 33) & 1023;
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = buf[threadIdx.x & 1023];
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t s, e;
    cudaEventCreate(&s);
    cudaEventCreate(&e);
    f();
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(s);
        f();
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms;
        cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(s);
    cudaEventDestroy(e);
    return best;
}

}  // namespace

// out[0] FP32 FFMA TFLOP/s (best of the 3-register-operand and the 2-invariant-operand loops), out[1] packed FFMA2
// TFLOP/s, out[2] MUFU.RSQ Gop/s,
// out[3] LDS.32 GB/s (all SMs), out[4] mma.sync TF32 m16n8k8 dense TFLOP/s, out[5] shared fp32 atomicAdd G lane-adds/s,
// out[6] SM count.  Returns 0 or a cudaError_t.
extern "C" int arf_peaks_measure(double* out) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 2048;
    float *obuf, *in;
    if (cudaMalloc(&obuf, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return 2;
    cudaMalloc(&in, 4096 * sizeof(float));
    cudaMemset(in, 0, 4096 * sizeof(float));
    const double n = (double)blocks * threads * iters;
    float ms = time_ms([&] { fma3_kernel<64><<<blocks, threads>>>(obuf, in, iters); });
    out[0] = n * 64 * 2 / ms / 1e9;
    ms = time_ms([&] { fma2op_kernel<16><<<blocks, threads>>>(obuf, 1.0001f, 0.5f, iters); });
    if (n * 32 * 2 / ms / 1e9 > out[0]) out[0] = n * 32 * 2 / ms / 1e9;
    ms = time_ms([&] { fma2_kernel<32><<<blocks, threads>>>(obuf, in, iters); });
    out[1] = n * 64 * 2 / ms / 1e9;
    ms = time_ms([&] { mufu_kernel<<<blocks, threads>>>(obuf, 0.5f, iters); });
    out[2] = n * 8 / ms / 1e6;
    ms = time_ms([&] { lds_kernel<<<blocks, threads>>>(obuf, iters); });
    out[3] = n * 8 * 4 / ms / 1e6;
    ms = time_ms([&] { mma_tf32_kernel<8><<<blocks, threads>>>(obuf, iters); });
    out[4] = (double)blocks * (threads / 32) * iters * 8 * (16.0 * 8 * 8 * 2) / ms / 1e9;
    ms = time_ms([&] { atoms_kernel<<<blocks, threads>>>(obuf, 256); });
    out[5] = (double)blocks * threads * 256 / ms / 1e6;
    out[6] = sms;
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(obuf);
    cudaFree(in);
    return (int)e;
}

#ifdef ARF_PEAKS_MAIN
int main() {
    double v[7];
    int rc = arf_peaks_measure(v);
    printf("{\"rc\": %d, \"fp32_fma_tflops\": %.2f, \"fp32_ffma2_tflops\": %.2f, \"mufu_gops\": %.1f, \"lds_gbs\": %.1f, "
           "\"mma_sync_tf32_tflops\": %.1f, \"atoms_f32_glanes\": %.2f, \"sms\": %d}\n",
           rc, v[0], v[1], v[2], v[3], v[4], v[5], (int)v[6]);
    return rc;
}
#endif

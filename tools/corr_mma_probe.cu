// Tensor-pipe variant of the cost-volume contraction, measured (north_star: "a tensor-pipe correlation variant, kept or
// dropped on evidence").  The contraction out[x, dx] = sum_c f1[c, x] * f2[c, x + dx] is a BANDED GEMM: with the
// smallest tensor tile (mma.sync m16n8k8, TF32) a block of 16 pixels needs the 24 f2 columns x0-4 .. x0+19, i.e. three
// n8 tiles of which 9 of 24 columns are used.  FP32 parity (1e-5) needs the 3xTF32 split (hi*hi + lo*hi + hi*lo).
// This probe runs exactly that micro-tile from shared memory - one warp <-> (16 pixels, 1 row, 9 dy x 3 n-tiles = 27
// accumulator tiles), operands staged as the product kernels stage them ([c][row][x]) - and reports
//   * useful TFLOP/s (2 * 81 flops per pixel and channel, as for the FFMA kernels),
//   * the error of 1xTF32 and 3xTF32 against a float64 reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/corr_mma_probe tools/corr_mma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int C = 16, ROWS = 8, TW = 16, HW = TW + 8, HR = ROWS + 8;   // tile 16 x 8 px, halo 24 x 16

__device__ __forceinline__ unsigned tf32_hi(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// one CTA = 8 warps = the 8 rows of one tile; out[row][dy][x][24 columns] (the band is extracted on the host)
template <int kSplit>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ f1, const float* __restrict__ f2,
                                             float* __restrict__ out, int iters) {
    __shared__ float s1[C][ROWS][TW];
    __shared__ float s2[C][HR][HW];
    for (int i = threadIdx.x; i < C * ROWS * TW; i += 256) (&s1[0][0][0])[i] = f1[i];
    for (int i = threadIdx.x; i < C * HR * HW; i += 256) (&s2[0][0][0])[i] = f2[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    float acc[9][3][4];
#pragma unroll
    for (int dy = 0; dy < 9; ++dy)
#pragma unroll
        for (int n = 0; n < 3; ++n)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[dy][n][k] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 8) {
            float av[4] = {s1[c0 + t][row][g], s1[c0 + t][row][g + 8], s1[c0 + t + 4][row][g], s1[c0 + t + 4][row][g + 8]};
            unsigned ah[4], al[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ah[k] = tf32_hi(av[k]);
                al[k] = tf32_hi(av[k] - __uint_as_float(ah[k]));
            }
#pragma unroll
            for (int dy = 0; dy < 9; ++dy)
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    float bv[2] = {s2[c0 + t][row + dy][n * 8 + g], s2[c0 + t + 4][row + dy][n * 8 + g]};
                    unsigned bh[2] = {tf32_hi(bv[0]), tf32_hi(bv[1])};
                    mma_tf32(acc[dy][n], ah, bh);
                    if (kSplit == 3) {
                        unsigned bl[2] = {tf32_hi(bv[0] - __uint_as_float(bh[0])), tf32_hi(bv[1] - __uint_as_float(bh[1]))};
                        mma_tf32(acc[dy][n], al, bh);
                        mma_tf32(acc[dy][n], ah, bl);
                    }
                }
        }
    }
    // D fragment: (row g, cols 2t, 2t+1), (row g+8, cols 2t, 2t+1)
    float* o = out + (size_t)blockIdx.x * ROWS * 9 * TW * HW + (size_t)row * 9 * TW * HW;
#pragma unroll
    for (int dy = 0; dy < 9; ++dy)
#pragma unroll
        for (int n = 0; n < 3; ++n) {
            o[(dy * TW + g) * HW + n * 8 + 2 * t] = acc[dy][n][0];
            o[(dy * TW + g) * HW + n * 8 + 2 * t + 1] = acc[dy][n][1];
            o[(dy * TW + g + 8) * HW + n * 8 + 2 * t] = acc[dy][n][2];
            o[(dy * TW + g + 8) * HW + n * 8 + 2 * t + 1] = acc[dy][n][3];
        }
}

int main() {
    std::vector<float> h1(C * ROWS * TW), h2(C * HR * HW);
    srand(1);
    for (auto& v : h1) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : h2) v = (float)rand() / RAND_MAX * 2 - 1;
    float *d1, *d2, *dout;
    const int blocks = 148 * 8;
    cudaMalloc(&d1, h1.size() * 4);
    cudaMalloc(&d2, h2.size() * 4);
    cudaMalloc(&dout, (size_t)blocks * ROWS * 9 * TW * HW * 4);
    cudaMemcpy(d1, h1.data(), h1.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d2, h2.data(), h2.size() * 4, cudaMemcpyHostToDevice);
    std::vector<float> ho((size_t)ROWS * 9 * TW * HW);
    double err[2] = {0, 0}, mx = 0;
    for (int split = 0; split < 2; ++split) {
        if (split == 0) probe<1><<<blocks, 256>>>(d1, d2, dout, 1);
        else probe<3><<<blocks, 256>>>(d1, d2, dout, 1);
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        for (int r = 0; r < ROWS; ++r)
            for (int dy = 0; dy < 9; ++dy)
                for (int x = 0; x < TW; ++x)
                    for (int dx = 0; dx < 9; ++dx) {
                        double ref = 0;
                        for (int c = 0; c < C; ++c) ref += (double)h1[(c * ROWS + r) * TW + x] * h2[(c * HR + r + dy) * HW + x + dx];
                        double got = ho[((size_t)(r * 9 + dy) * TW + x) * HW + x + dx];
                        err[split] = fmax(err[split], fabs(got - ref));
                        mx = fmax(mx, fabs(ref));
                    }
    }
    cudaEvent_t s, e;
    cudaEventCreate(&s); cudaEventCreate(&e);
    float ms[2];
    const int iters = 200;
    for (int split = 0; split < 2; ++split) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(s);
            if (split == 0) probe<1><<<blocks, 256>>>(d1, d2, dout, iters);
            else probe<3><<<blocks, 256>>>(d1, d2, dout, iters);
            cudaEventRecord(e);
            cudaEventSynchronize(e);
            float t;
            cudaEventElapsedTime(&t, s, e);
            if (rep > 0 && t < best) best = t;
        }
        ms[split] = best;
    }
    const double useful = (double)blocks * ROWS * TW * C * 162.0 * iters;   // flops the FFMA kernels are credited with
    printf("{\"tile\": \"16 px x 8 rows x 81 displacements x 16 ch, smem-resident, no HBM traffic\", "
           "\"tf32x1_useful_tflops\": %.2f, \"tf32x3_useful_tflops\": %.2f, \"tf32x1_rel_err\": %.2e, \"tf32x3_rel_err\": %.2e, "
           "\"mma_per_useful_fma\": %.2f}\n",
           useful / ms[0] / 1e9, useful / ms[1] / 1e9, err[0] / mx, err[1] / mx, 81.0 * 3 * 1024 / (16 * 81 * 8.0));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

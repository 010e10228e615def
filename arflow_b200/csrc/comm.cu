// Gradient all-reduce over NVLink peer memory (the one real exchange of the path; the reference's own
// multi-GPU code is nn.DataParallel's gather of gradients onto GPU 0, trainer/base_trainer.py:75,131-147).
//
// One process per GPU.  Every rank allocates its gradient buffer and a small flag area with arf_comm_alloc, exports both
// as CUDA IPC handles and opens its peers' (arflow_b200/comm.py moves the 64-byte handles through torch.distributed).
// arf_allreduce_f32 is then ONE kernel launch per rank, in place, with no host involvement - it can be captured in a CUDA
// graph and forked onto a side stream while backward is still running (NCCL calls inside a captured step hung on this
// pool):
//   barrier (every peer's gradients are complete: a kernel starts after its stream's earlier work)
//   reduce-scatter: rank r sums shard r of all N buffers (16-byte loads over NVLink, fixed rank order -> every rank
//                   later holds bit-identical results) and writes mean or sum into its own shard
//   barrier
//   all-gather:     rank r copies the N-1 other reduced shards from their owners
//   barrier         (nobody still reads this rank's buffer when the next backward starts to overwrite it)
// A barrier is per CTA index: CTA b of rank r stores an epoch into slot [b][r] of every peer's flag area
// (st.release.sys after a system-scope fence) and spins on its own slots (ld.acquire.sys).  CTA b of every rank works on
// the same sub-range of every shard, so CTA-level pairing is enough.  Epochs count up per CTA in local memory, so graph
// replays need no host-side state.  Spins are bounded (~2 s): a missing peer sets an error word instead of hanging the
// GPU.  Peer loads use ld.global.cg: peer lines may sit in this SM's L1 from the previous step, L2 is bypassed by the
// hardware for peer addresses.
#include "common.cuh"
#include <string.h>

namespace {

constexpr int kMaxRanks = 8;
constexpr int kMaxCtas = 64;
constexpr int kCommThreads = 512;
// flag area (uint32 words): [kMaxCtas][kMaxRanks] arrival slots, then [kMaxCtas] local epoch counters, then 1 error word
constexpr int kFlagWords = kMaxCtas * kMaxRanks + kMaxCtas + 1;

struct CommPtrs {
    float* data[kMaxRanks];
    unsigned* flags[kMaxRanks];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all ranks' CTAs with this blockIdx meet; returns after every peer has arrived at the same barrier
__device__ __forceinline__ void peer_barrier(const CommPtrs& P, int rank, int nranks, unsigned& epoch, unsigned* err) {
    // The CTA barrier orders every thread's data writes before the signalling threads' release stores, and a release is
    // cumulative: no per-thread system fence is needed (512 MEMBAR.SYS per barrier cost ~10 us each time).
    __syncthreads();
    epoch += 1;
    if ((int)threadIdx.x < nranks) {
        const int p = threadIdx.x;
        st_release_sys(P.flags[p] + blockIdx.x * kMaxRanks + rank, epoch);
        const unsigned* mine = P.flags[rank] + blockIdx.x * kMaxRanks + p;
        const long long t0 = clock64();
        // epochs are compared with wrap-around arithmetic
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > 4000000000LL) { *err = 1u; break; }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kCommThreads)
allreduce_f32_kernel(CommPtrs P, int rank, int nranks, size_t offset, size_t count, float scale) {
    unsigned* local = P.flags[rank];
    unsigned* err = local + kMaxCtas * kMaxRanks + kMaxCtas;
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) s_epoch = local[kMaxCtas * kMaxRanks + blockIdx.x];
    __syncthreads();
    unsigned epoch = s_epoch;

    const size_t nv = count / 4;                                   // float4 units (count % 4 == 0, offset % 4 == 0)
    const size_t per = (nv + nranks - 1) / nranks;                 // shard length
    const size_t chunk = (per + gridDim.x - 1) / gridDim.x;        // this CTA's part of every shard
    const size_t c_lo = (size_t)blockIdx.x * chunk;

    peer_barrier(P, rank, nranks, epoch, err);
    {   // reduce-scatter: my shard.  Two float4 per peer and thread in flight (2 x nranks independent 16-byte loads).
        const size_t s_lo = (size_t)rank * per, s_hi = min(nv, s_lo + per);
        const size_t lo = min(s_hi, s_lo + c_lo), hi = min(s_hi, lo + chunk);
        float4* mine = reinterpret_cast<float4*>(P.data[rank] + offset);
        for (size_t i0 = lo + threadIdx.x; i0 < hi; i0 += 2 * kCommThreads) {
            const size_t i1 = i0 + kCommThreads;
            const bool two = i1 < hi;
            float4 v0[kMaxRanks], v1[kMaxRanks];
#pragma unroll
            for (int p = 0; p < kMaxRanks; ++p) {
                if (p < nranks) {
                    const float4* src = reinterpret_cast<const float4*>(P.data[p] + offset);
                    v0[p] = __ldcg(src + i0);
                    v1[p] = two ? __ldcg(src + i1) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
            for (int p = 0; p < kMaxRanks; ++p) {
                if (p < nranks) {
                    a0.x += v0[p].x; a0.y += v0[p].y; a0.z += v0[p].z; a0.w += v0[p].w;
                    a1.x += v1[p].x; a1.y += v1[p].y; a1.z += v1[p].z; a1.w += v1[p].w;
                }
            }
            a0.x *= scale; a0.y *= scale; a0.z *= scale; a0.w *= scale;
            mine[i0] = a0;
            if (two) {
                a1.x *= scale; a1.y *= scale; a1.z *= scale; a1.w *= scale;
                mine[i1] = a1;
            }
        }
    }
    peer_barrier(P, rank, nranks, epoch, err);
    {   // all-gather: everybody else's reduced shard, one float4 from every peer in flight per thread
        float4* mine = reinterpret_cast<float4*>(P.data[rank] + offset);
        for (size_t j = threadIdx.x; j < chunk; j += kCommThreads) {
            float4 v[kMaxRanks];
            bool ok[kMaxRanks];
#pragma unroll
            for (int q = 1; q < kMaxRanks; ++q) {
                ok[q] = false;
                if (q < nranks) {
                    const int p = (rank + q) % nranks;                    // spread the traffic over the peers
                    const size_t s_lo = (size_t)p * per, s_hi = min(nv, s_lo + per);
                    const size_t i = s_lo + c_lo + j;
                    ok[q] = i < s_hi;
                    if (ok[q]) v[q] = __ldcg(reinterpret_cast<const float4*>(P.data[p] + offset) + i);
                }
            }
#pragma unroll
            for (int q = 1; q < kMaxRanks; ++q) {
                if (q < nranks && ok[q]) {
                    const int p = (rank + q) % nranks;
                    mine[(size_t)p * per + c_lo + j] = v[q];
                }
            }
        }
    }
    peer_barrier(P, rank, nranks, epoch, err);
    if (threadIdx.x == 0) local[kMaxCtas * kMaxRanks + blockIdx.x] = epoch;
}

}  // namespace

extern "C" int arf_comm_flag_bytes(void) { return kFlagWords * (int)sizeof(unsigned); }

extern "C" int arf_comm_alloc(void** ptr, size_t bytes) {
    ARF_REQUIRE(ptr && bytes > 0);
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    return e == cudaSuccess ? ARF_OK : (int)e;
}

extern "C" int arf_comm_free(void* ptr) {
    if (!ptr) return ARF_OK;
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? ARF_OK : (int)e;
}

extern "C" int arf_comm_ipc_get(void* ptr, void* handle64) {
    ARF_REQUIRE(ptr && handle64);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr);
    return e == cudaSuccess ? ARF_OK : (int)e;
}

extern "C" int arf_comm_ipc_open(const void* handle64, void** ptr) {
    ARF_REQUIRE(ptr && handle64);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? ARF_OK : (int)e;
}

extern "C" int arf_comm_ipc_close(void* ptr) {
    if (!ptr) return ARF_OK;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? ARF_OK : (int)e;
}

extern "C" int arf_allreduce_f32(float* const* data, unsigned* const* flags, int rank, int nranks, size_t offset,
                                 size_t count, float scale, int ctas, void* stream) {
    ARF_REQUIRE(data && flags && nranks >= 1 && nranks <= kMaxRanks && rank >= 0 && rank < nranks);
    ARF_REQUIRE(count % 4 == 0 && offset % 4 == 0);
    if (count == 0) return ARF_OK;
    if (ctas < 1) ctas = 1;
    if (ctas > kMaxCtas) ctas = kMaxCtas;
    CommPtrs P;
    for (int p = 0; p < kMaxRanks; ++p) {
        P.data[p] = p < nranks ? data[p] : nullptr;
        P.flags[p] = p < nranks ? flags[p] : nullptr;
        if (p < nranks) ARF_REQUIRE(P.data[p] && P.flags[p]);
    }
    allreduce_f32_kernel<<<ctas, kCommThreads, 0, (cudaStream_t)stream>>>(P, rank, nranks, offset, count, scale);
    ARF_CHECK_LAUNCH();
    return ARF_OK;
}

// error word of a flag area (non-zero after a barrier timed out); synchronises the stream
extern "C" int arf_comm_error(const unsigned* flags, void* stream) {
    ARF_REQUIRE(flags);
    unsigned v = 0;
    cudaError_t e = cudaMemcpyAsync(&v, flags + kMaxCtas * kMaxRanks + kMaxCtas, sizeof(v), cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    return v ? ARF_ETIMEOUT : ARF_OK;
}

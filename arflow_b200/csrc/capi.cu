// Library-level entry points of the C-ABI (version, error strings).
#include "common.cuh"

std::atomic<long long> g_arf_launches{0};

extern "C" long long arf_launch_count(void) { return g_arf_launches.load(std::memory_order_relaxed); }

int arf_num_sms() {
    static std::atomic<int> cache[64];   // zero-initialised; 0 = not queried yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64) {
        int v = cache[dev].load(std::memory_order_relaxed);
        if (v > 0) return v;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

extern "C" int arf_version(void) { return 100; }  // 0.1.0

extern "C" const char* arf_error_string(int code) {
    switch (code) {
        case ARF_OK: return "ok";
        case ARF_EINVAL: return "invalid argument (shape, parameter or null pointer)";
        case ARF_EUNSUPPORTED: return "combination valid in the reference but not implemented";
        case ARF_EWORKSPACE: return "workspace too small";
        case ARF_ETIMEOUT: return "a peer did not reach an all-reduce barrier in time";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown arflow_b200 error";
}

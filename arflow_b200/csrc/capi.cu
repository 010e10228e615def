// Library-level entry points of the C-ABI (version, error strings).
#include "common.cuh"

long long g_arf_launches = 0;

extern "C" long long arf_launch_count(void) { return g_arf_launches; }

extern "C" int arf_version(void) { return 100; }  // 0.1.0

extern "C" const char* arf_error_string(int code) {
    switch (code) {
        case ARF_OK: return "ok";
        case ARF_EINVAL: return "invalid argument (shape, parameter or null pointer)";
        case ARF_EUNSUPPORTED: return "combination valid in the reference but not implemented";
        case ARF_EWORKSPACE: return "workspace too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown arflow_b200 error";
}

// Library-level entry points of the C ABI: version, error strings, device check.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace mk {

static thread_local char g_cuda_error[256] = {0};

void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s (%s)", where, cudaGetErrorName(e),
             cudaGetErrorString(e));
}

}  // namespace mk

extern "C" int mk_version(void) { return MK_VERSION; }

extern "C" const char* mk_error_string(int code) {
    switch (code) {
        case MK_OK: return "ok";
        case MK_EINVAL: return "invalid argument";
        case MK_EUNSUPPORTED: return "unsupported size combination";
        case MK_ECUDA: return "CUDA runtime error";
        case MK_ENODEVICE: return "no sm_100 CUDA device";
        default: return "unknown error";
    }
}

extern "C" const char* mk_last_cuda_error(void) { return mk::g_cuda_error; }

extern "C" int mk_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return MK_ENODEVICE;
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return MK_ENODEVICE;
    }
    return major == 10 ? MK_OK : MK_ENODEVICE;
}

// Process-level pieces of the C ABI: version, per-thread error string, device probe.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace dpft {
namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}
}  // namespace dpft

extern "C" int dpft_abi_version(void) { return DPFT_ABI_VERSION; }

extern "C" const char* dpft_last_error(void) { return dpft::g_err; }

extern "C" int dpft_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    cudaDeviceProp p;
    int st = dpft::cuda_status(cudaGetDeviceProperties(&p, device), "cudaGetDeviceProperties");
    if (st) return st;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return DPFT_OK;
}

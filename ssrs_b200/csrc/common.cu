// Error plumbing + device queries behind the C-ABI (include/ssrs_b200.h).
#include "common.cuh"

#include <mutex>
#include <string.h>

namespace ssrs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return SSRS_ERR_CUDA;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace ssrs

extern "C" int ssrs_abi_version(void) { return SSRS_ABI_VERSION; }

extern "C" const char* ssrs_last_error(void) { return ssrs::g_err; }

extern "C" int ssrs_device_info(int* sm, int* major, int* minor) {
    int dev = 0;
    SSRS_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SSRS_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    if (sm) *sm = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return SSRS_OK;
}

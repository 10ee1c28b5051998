// Error plumbing and device queries of the C ABI (include/ia_b200.h).
#include <stdarg.h>
#include <string.h>

#include "ia_common.cuh"

static thread_local char g_err[512] = "";

void ia_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ia_sm_count()
{
    static thread_local int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

extern "C" const char *ia_last_error_string(void) { return g_err; }

extern "C" int32_t ia_abi_version(void) { return IA_ABI_VERSION; }

extern "C" int32_t ia_device_arch(void)
{
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        ia_set_error("no CUDA device available");
        return IA_ERR_NO_DEVICE;
    }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}

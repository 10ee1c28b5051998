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

// Persisting L2 window over the hash tables (include/ia_b200.h "L2 residency of the hash tables").
extern "C" int32_t ia_l2_persist(const void *base, int64_t bytes, float hit_ratio, int64_t *info_host, void *stream)
{
    int dev = 0, l2 = 0, max_persist = 0, max_window = 0;
    IA_CUDA_OK(cudaGetDevice(&dev));
    IA_CUDA_OK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    IA_CUDA_OK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    IA_CUDA_OK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    IA_REQUIRE(bytes >= 0 && (bytes == 0 || base != nullptr), "l2_persist: bad window (%p, %lld)", base, (long long)bytes);
    IA_REQUIRE(hit_ratio >= 0.f && hit_ratio <= 1.f, "l2_persist: hit_ratio %f outside [0,1]", hit_ratio);
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    size_t set_aside = 0, window = 0;
    if (bytes > 0) {
        window = (size_t)(bytes < (int64_t)max_window ? bytes : (int64_t)max_window);
        set_aside = (size_t)(window < (size_t)max_persist ? window : (size_t)max_persist);
        IA_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside));
        attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
        attr.accessPolicyWindow.num_bytes = window;
        attr.accessPolicyWindow.hitRatio = hit_ratio;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        IA_CUDA_OK(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        IA_CUDA_OK(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        IA_CUDA_OK(cudaCtxResetPersistingL2Cache());
        IA_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
    }
    if (info_host) {
        info_host[0] = l2;
        info_host[1] = (int64_t)set_aside;
        info_host[2] = (int64_t)window;
    }
    return IA_OK;
}

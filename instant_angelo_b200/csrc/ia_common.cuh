// Shared helpers for the sm_100a kernels behind include/ia_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ia_b200.h"

// thread-local error text (capi.cu)
void ia_set_error(const char *fmt, ...);

#define IA_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            ia_set_error(__VA_ARGS__);        \
            return IA_ERR_INVALID_ARG;        \
        }                                     \
    } while (0)

#define IA_CUDA_OK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            ia_set_error("%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return IA_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

#define IA_LAUNCH_OK(name)                                                            \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            ia_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));    \
            return IA_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

static inline int64_t ia_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (cached per thread).
int ia_sm_count();

#ifdef __CUDACC__
// [N, L*F]-sized activation streams are written once and read once, a whole L2 later: evict-first hints keep them from
// pushing the hash tables out of the L2 (IA_NO_STREAM_HINTS: plain accesses, for A/B builds)
#ifndef IA_NO_STREAM_HINTS
#define IA_LD_STREAM(p) __ldcs(p)
#define IA_ST_STREAM(p, v) __stcs((p), (v))
#else
#define IA_LD_STREAM(p) __ldg(p)
#define IA_ST_STREAM(p, v) (*(p) = (v))
#endif

__device__ __forceinline__ float ia_warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

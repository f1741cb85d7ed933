// common.cuh -- shared helpers for libmvedit_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmvedit_b200 is written for sm_100a (B200) only"
#endif

void mve_set_error(const char* fmt, ...);

#define MVE_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            mve_set_error("%s: %s", name, cudaGetErrorString(e__));                   \
            return (int)e__;                                                          \
        }                                                                             \
    } while (0)

#define MVE_CUDA(call)                                                                \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            mve_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                          \
        }                                                                             \
    } while (0)

#define MVE_ARG(cond, msg)                                                            \
    do {                                                                              \
        if (!(cond)) { mve_set_error("bad argument: %s", msg); return -1; }           \
    } while (0)

static inline unsigned int cdiv(unsigned long long a, unsigned int b) { return (unsigned int)((a + b - 1) / b); }

// SM count of the current device, queried once per device (148 on a B200; grids are sized from it, nothing is hard-wired)
static inline int mve_num_sms() {
    static int cached[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& c = cached[dev & 15];
    if (c == 0 && (cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || c <= 0)) c = 148;
    return c;
}
#define kNumSM (mve_num_sms())

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// cuda_host_shim.h -- TEST INFRASTRUCTURE ONLY.  Lets an element-wise CUDA source of the product (one thread per output element, float
// atomics, warp_sum + "lane 0 adds" reductions; no shared memory, no other warp intrinsics) compile UNCHANGED as plain C++:
// tests/host_harness.py rewrites `k<<<grid, block, 0, s>>>(args)` into SHIM_LAUNCH(k, grid, block, args) and the include of common.cuh
// into this file.  Every host "thread" is lane 0 of its own warp (warp_sum is the identity), so a reduction adds every term once.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __logf logf
#define __expf expf

struct ShimLane {
    uint32_t v;
    operator uint32_t() const { return v; }
};
static inline uint32_t operator&(ShimLane, int) { return 0; }
struct ShimThreadIdx { ShimLane x; };
struct ShimBlockIdx { uint32_t x; };
static thread_local ShimThreadIdx threadIdx;
static thread_local ShimBlockIdx blockIdx;

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }

static inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
static inline float __fdividef(float a, float b) { return a / b; }

static inline float warp_sum(float v) { return v; }
static inline float atomicAdd(float* p, float v) { const float o = *p; *p = o + v; return o; }

typedef void* cudaStream_t;
#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), 0)
#define MVE_CUDA(call) (void)(call)
#define MVE_CHECK_LAUNCH(name)
#define MVE_ARG(cond, msg) do { if (!(cond)) { fprintf(stderr, "bad argument: %s\n", msg); return -1; } } while (0)
static inline unsigned int cdiv(unsigned long long a, unsigned int b) { return (unsigned int)((a + b - 1) / b); }

#define SHIM_LAUNCH(kernel, grid, block, ...)                                   \
    for (uint32_t b__ = 0; b__ < (uint32_t)(grid); b__++)                       \
        for (uint32_t t__ = 0; t__ < (uint32_t)(block); t__++) {                \
            blockIdx.x = b__; threadIdx.x.v = t__;                              \
            kernel(__VA_ARGS__);                                                \
        }

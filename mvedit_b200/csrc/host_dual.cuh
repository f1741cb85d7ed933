// host_dual.cuh -- prelude of the kernels that also compile as plain C++ (-DMVE_HOST_HARNESS, tests/host_harness.py): per-element
// functions are written once as MVE_HD; MVE_ELEMENT_KERNEL turns one into a __global__ kernel (CUDA) or a serial loop (host), MVE_LAUNCH
// launches / runs it, atomics and memsets are abstracted the same way.  The host build is test infrastructure: the CPU suite runs the very
// same arithmetic (bit for bit where FMA contraction is off) and the Python autograd mirrors without a GPU.  The product library is the
// CUDA build only.
#pragma once
#ifdef MVE_HOST_HARNESS
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#define MVE_HD static inline
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static char g_harness_err[512];
static void mve_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_harness_err, sizeof g_harness_err, fmt, ap); va_end(ap); }
#define MVE_ARG(cond, msg) do { if (!(cond)) { mve_set_error("bad argument: %s", msg); return -1; } } while (0)
#define MVE_CHECK_LAUNCH(name) do { } while (0)
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline void atomic_add_f(float* p, float v) { *p += v; }
static inline void atomic_min_u64(unsigned long long* p, unsigned long long v) { if (v < *p) *p = v; }
static inline uint32_t atomic_inc_u32(uint32_t* p) { return (*p)++; }
static inline void atomic_min_u32(uint32_t* p, uint32_t v) { if (v < *p) *p = v; }
static inline unsigned long long atomic_cas_u64(unsigned long long* p, unsigned long long cmp, unsigned long long v) { unsigned long long o = *p; if (o == cmp) *p = v; return o; }
#define MVE_ELEMENT_KERNEL(kname, P, fn) static void kname(const P& p, uint32_t n) { for (uint32_t i = 0; i < n; ++i) fn(p, i); }
// a per-element function that also accumulates K partial sums into p.field[0..K): fn(p, i, acc)
#define MVE_REDUCE_KERNEL(kname, P, fn, K, field)                                          \
    static void kname(const P& p, uint32_t n) {                                           \
        double tot[K] = {};                                                               \
        for (uint32_t i = 0; i < n; ++i) { float acc[K] = {}; fn(p, i, acc); for (int k = 0; k < K; ++k) tot[k] += acc[k]; }  \
        for (int k = 0; k < K; ++k) p.field[k] += (float)tot[k];                          \
    }
#define MVE_LAUNCH(kname, p, n, st) kname(p, n)
#define MVE_MEMSET(ptr, byte, bytes, st) memset(ptr, byte, bytes)
#define MVE_MEMCPY(dst, src, bytes, st) memcpy(dst, src, bytes)
#define MVE_EXPORT extern "C" __attribute__((visibility("default")))
#else
#include "common.cuh"
#define MVE_HD __device__ __forceinline__
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ void atomic_add_f(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_min_u64(unsigned long long* p, unsigned long long v) { atomicMin(p, v); }
__device__ __forceinline__ uint32_t atomic_inc_u32(uint32_t* p) { return atomicAdd(p, 1u); }
__device__ __forceinline__ void atomic_min_u32(uint32_t* p, uint32_t v) { atomicMin(p, v); }
__device__ __forceinline__ unsigned long long atomic_cas_u64(unsigned long long* p, unsigned long long cmp, unsigned long long v) { return atomicCAS(p, cmp, v); }
#define MVE_ELEMENT_KERNEL(kname, P, fn)                                                  \
    __global__ void __launch_bounds__(256) kname(const P p, uint32_t n) {                 \
        uint32_t i = blockIdx.x * 256u + threadIdx.x;                                     \
        if (i < n) fn(p, i);                                                              \
    }
// per-element function with K partial sums: warp shuffle reduction, one atomic per warp and sum
#define MVE_REDUCE_KERNEL(kname, P, fn, K, field)                                         \
    __global__ void __launch_bounds__(256) kname(const P p, uint32_t n) {                 \
        uint32_t i = blockIdx.x * 256u + threadIdx.x;                                     \
        float acc[K] = {};                                                                \
        if (i < n) fn(p, i, acc);                                                         \
        _Pragma("unroll")                                                                 \
        for (int k = 0; k < K; ++k) {                                                     \
            float v = warp_sum(acc[k]);                                                   \
            if ((threadIdx.x & 31u) == 0u && v != 0.f) atomicAdd(p.field + k, v);         \
        }                                                                                 \
    }
#define MVE_LAUNCH(kname, p, n, st)                                                       \
    do { if ((n) > 0) { kname<<<cdiv((n), 256), 256, 0, (cudaStream_t)(st)>>>(p, (uint32_t)(n)); MVE_CHECK_LAUNCH(#kname); } } while (0)
#define MVE_MEMSET(ptr, byte, bytes, st) MVE_CUDA(cudaMemsetAsync(ptr, byte, bytes, (cudaStream_t)(st)))
#define MVE_MEMCPY(dst, src, bytes, st) MVE_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)(st)))
#define MVE_EXPORT extern "C"
#endif


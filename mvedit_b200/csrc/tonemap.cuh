// tonemap.cuh -- the 16-knot piecewise-linear tone curve of the reference (lib/models/decoders/tonemapping.py:5-52) as device code.
//   lut(x)          log2-exposure -> display value: bucketize(x, lut_x, right=True) clamped to [1, n-1], linear interpolation (the end
//                   segments extrapolate)
//   inverse_lut(y)  display value -> log2-exposure, the same over (lut_y, lut_x)
// Shading in tone-mapped space (mvedit_3d_pipeline.py:564-570, :1377-1384): out = lut(inverse_lut(rgb / alpha) + log2(shading)).
// The knots travel by value in the kernel parameters (n <= 32): nothing to keep alive for a captured graph.
#pragma once
#include <stdint.h>

#define MVE_TONEMAP_MAX_KNOTS 32

struct ToneLut {
    int n;                                  // 0 = tone mapping off
    float x[MVE_TONEMAP_MAX_KNOTS], y[MVE_TONEMAP_MAX_KNOTS];
};

// host: knots = [lut_x (n) | lut_y (n)] or NULL
static inline int fill_tone_lut(ToneLut& t, const float* knots, uint32_t n) {
    t.n = 0;
    if (!knots || n == 0) return 0;
    if (n < 2 || n > MVE_TONEMAP_MAX_KNOTS) return -1;
    t.n = (int)n;
    for (uint32_t i = 0; i < n; i++) { t.x[i] = knots[i]; t.y[i] = knots[n + i]; }
    return 0;
}

// piecewise-linear map through knots (a -> b); *slope receives d out / d v
__device__ __forceinline__ float tone_interp(const float* __restrict__ a, const float* __restrict__ b, int n, float v, float* slope) {
    int i = 1;                              // torch.bucketize(v, a, right=True): number of knots <= v, then clamp to [1, n-1]
    while (i < n - 1 && a[i] <= v) i++;
    const float k = (b[i] - b[i - 1]) / (a[i] - a[i - 1]);
    if (slope) *slope = k;
    return b[i - 1] + (v - a[i - 1]) * k;
}
__device__ __forceinline__ float tone_lut(const ToneLut& t, float x, float* slope = nullptr) { return tone_interp(t.x, t.y, t.n, x, slope); }
__device__ __forceinline__ float tone_inverse_lut(const ToneLut& t, float y, float* slope = nullptr) { return tone_interp(t.y, t.x, t.n, y, slope); }

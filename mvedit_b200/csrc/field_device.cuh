// field_device.cuh -- device code of the Instant-NGP field shared by field.cu (training fwd/bwd) and render.cu
// (fused inference renderer).  See field.cu for the algorithm notes and reference citations.
#pragma once
#include "common.cuh"

namespace field {

constexpr int HID = 64;
constexpr int MAXL = 16;

struct Levels {
    float scale[MAXL];
    uint32_t res[MAXL], size[MAXL], off[MAXL];
    uint32_t hashed;  // bit l set: level l uses the coherent prime hash
};

struct FieldCfg {
    float bound, inv2b, blob_density, blob_k, sat_scale, sat_shift;
};

// grid.h grid_index: dense levels index linearly, hashed levels use the coherent prime hash; both end in `% size`.
// A generic integer modulo costs ~20 instructions and there are 96 of them per sample, so it is specialised (same result):
//   hashed: size is a power of two for every config the reference builds (2^19) -> mask;  otherwise the generic modulo;
//   dense : idx <= res^3 + res^2 + res < 2 * size, so one conditional subtraction is the exact modulo.
__device__ __forceinline__ uint32_t grid_index(const bool hashed, const uint32_t res, const uint32_t size, const uint32_t cx, const uint32_t cy,
                                               const uint32_t cz) {
    if (hashed) {
        const uint32_t idx = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
        return ((size & (size - 1)) == 0) ? (idx & (size - 1)) : (idx % size);
    }
    const uint32_t idx = cx + cy * res + cz * res * res;
    return idx >= size ? idx - size : idx;
}

// World coordinate -> the hash grid's [0, 1] input (ingp_decoder.py:112), clamped to the box.  Ray-marched samples never leave the
// aabb, but the mesh stage queries arbitrary points (init_tet's rescaled tet grid, lib/pipelines/utils.py:177-183; surface points of a
// drifting mesh): outside [0, 1] the dense-level index of grid_index() would leave the table.  tcnn wraps such indices (`% size` on a
// value that no longer fits the one-subtraction shortcut below); here the query is evaluated at the nearest point of the box instead --
// the reference discards those values anyway (init_tet overwrites the SDF outside [-1, 1]).
__device__ __forceinline__ float unit_coord(const FieldCfg& c, const float x) { return fminf(fmaxf((x + c.bound) * c.inv2b, 0.f), 1.f); }

// positional part of one level: base cell, smoothstep weights and their derivatives
struct Cell {
    uint32_t g[3];
    float w[3], dw[3];
};
__device__ __forceinline__ Cell locate(const float x0, const float x1, const float x2, const float scale) {
    Cell c;
    const float x[3] = {x0, x1, x2};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float pos = fmaf(scale, x[d], 0.5f);
        const float fl = floorf(pos);
        c.g[d] = (uint32_t)(int)fl;
        const float f = pos - fl;
        c.w[d] = f * f * (3.0f - 2.0f * f);
        c.dw[d] = 6.0f * f * (1.0f - f);
    }
    return c;
}

template <int L>
__device__ __forceinline__ void encode(const Levels& lv, const float2* __restrict__ table, const float x0, const float x1, const float x2,
                                       float (&enc)[2 * L]) {
#pragma unroll
    for (int l = 0; l < L; l++) {
        const Cell c = locate(x0, x1, x2, lv.scale[l]);
        const bool hashed = (lv.hashed >> l) & 1u;
        const uint32_t res = lv.res[l], size = lv.size[l];
        const float2* __restrict__ t = table + lv.off[l];
        float2 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
            v[k] = __ldg(t + grid_index(hashed, res, size, c.g[0] + (k & 1), c.g[1] + ((k >> 1) & 1), c.g[2] + (k >> 2)));
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float wgt = ((k & 1) ? c.w[0] : 1.f - c.w[0]) * ((k & 2) ? c.w[1] : 1.f - c.w[1]) * ((k & 4) ? c.w[2] : 1.f - c.w[2]);
            a0 = fmaf(wgt, v[k].x, a0);
            a1 = fmaf(wgt, v[k].y, a1);
        }
        enc[2 * l] = a0;
        enc[2 * l + 1] = a1;
    }
}

// shared-memory record of hidden unit j: [W1[j][0..IN) | W2[0..3][j] | b1[j] | pad]  (float4 aligned)
template <int L>
struct Rec {
    static constexpr int IN = 2 * L;
    static constexpr int W2O = IN, B1O = IN + 4, STRIDE = ((IN + 5 + 3) / 4) * 4;
};

template <int L>
__device__ __forceinline__ void stage_mlp(float* __restrict__ rec, const float* __restrict__ w1, const float* __restrict__ b1,
                                          const float* __restrict__ w2) {
    using R = Rec<L>;
    for (int i = threadIdx.x; i < HID * R::STRIDE; i += blockDim.x) {
        const int j = i / R::STRIDE, o = i % R::STRIDE;
        float v = 0.f;
        if (o < R::IN) v = w1[j * R::IN + o];
        else if (o < R::IN + 4) v = w2[(o - R::IN) * HID + j];
        else if (o == R::B1O) v = b1[j];
        rec[i] = v;
    }
}

__device__ __forceinline__ float blob_of(const FieldCfg& c, const float x, const float y, const float z) {
    const float d = fmaxf(x * x + y * y + z * z, 0.2f);
    return c.blob_density * __expf(-d * c.blob_k);
}


inline int fill_levels(Levels& lv, uint32_t n_levels, const float* scale, const uint32_t* res, const uint32_t* size, const uint32_t* off) {
    if (n_levels > (uint32_t)MAXL) return -1;
    lv.hashed = 0;
    for (uint32_t l = 0; l < n_levels; l++) {
        lv.scale[l] = scale[l]; lv.res[l] = res[l]; lv.size[l] = size[l]; lv.off[l] = off[l];
        if ((uint64_t)res[l] * res[l] * res[l] > size[l]) lv.hashed |= 1u << l;
    }
    return 0;
}

inline FieldCfg make_cfg(float bound, float blob_density, float blob_radius, float sigmoid_saturation) {
    FieldCfg c;
    c.bound = bound; c.inv2b = 1.0f / (2 * bound);
    c.blob_density = blob_density; c.blob_k = 1.0f / (2 * blob_radius * blob_radius);
    c.sat_scale = 1 + 2 * sigmoid_saturation; c.sat_shift = -sigmoid_saturation;
    return c;
}

// MLP forward for one sample from the staged records: 4 raw outputs
template <int L, bool DENSITY_ONLY>
__device__ __forceinline__ void mlp_forward(const float* __restrict__ rec, const float (&enc)[2 * L], float& o0, float& o1, float& o2, float& o3) {
    using R = Rec<L>;
#pragma unroll 4
    for (int j = 0; j < HID; j++) {
        const float4* r4 = reinterpret_cast<const float4*>(rec + j * R::STRIDE);
        float a = rec[j * R::STRIDE + R::B1O];
#pragma unroll
        for (int q = 0; q < R::IN / 4; q++) {
            const float4 w = r4[q];
            a = fmaf(w.x, enc[4 * q], a); a = fmaf(w.y, enc[4 * q + 1], a); a = fmaf(w.z, enc[4 * q + 2], a); a = fmaf(w.w, enc[4 * q + 3], a);
        }
        a = fmaxf(a, 0.f);
        const float4 v = r4[R::W2O / 4];
        o0 = fmaf(v.x, a, o0);
        if (!DENSITY_ONLY) { o1 = fmaf(v.y, a, o1); o2 = fmaf(v.z, a, o2); o3 = fmaf(v.w, a, o3); }
    }
}

}  // namespace field

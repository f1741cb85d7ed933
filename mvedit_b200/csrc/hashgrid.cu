// hashgrid.cu -- the multi-resolution hash-grid encoding as a stand-alone differentiable op: the `tcnn.Encoding(x)` seam (B4, SURVEY.md
// §8b: /root/reference/lib/models/decoders/ingp_decoder.py:62-74,112; triplane_ingp_decoder.py:102-114,150).
//
// tiny-cuda-nn is an un-vendored dependency (requirements.txt:5); its grid.h algorithm (grid_scale / grid_resolution, Smoothstep
// interpolation, dense indexing below 2^19 entries and the coherent prime hash above) is restated in oracle/field_oracle.py, which
// these kernels are checked against.  The iNGP decoder itself uses the FUSED field kernels (field.cu: encoding + MLP + activations in
// one launch); this file serves callers that need the bare encoding -- TriPlaneiNGPDecoder adds it to tri-plane features before its
// own layers -- with gradients w.r.t. the table (atomic red.v2 scatter) and w.r.t. the input positions.
//
// One thread per (sample, level): 8 gathers of 8 bytes; consecutive threads of a warp walk the levels of the same sample, so the
// [M, 2L] output row is written as 2L consecutive floats.  L2-gather bound (the table is 28.7 MB at L = 12).
#include "field_device.cuh"
#include "../../include/mvedit_b200.h"

namespace {

using namespace field;

__global__ void __launch_bounds__(256) k_hashgrid_fwd(const float* __restrict__ x01, const uint32_t M, const float2* __restrict__ table,
                                                      const Levels lv, const uint32_t L, float* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= (size_t)M * L) return;
    const uint32_t i = (uint32_t)(t / L), l = (uint32_t)(t % L);
    const Cell c = locate(x01[(size_t)i * 3], x01[(size_t)i * 3 + 1], x01[(size_t)i * 3 + 2], lv.scale[l]);
    const bool hashed = (lv.hashed >> l) & 1u;
    const float2* __restrict__ tb = table + lv.off[l];
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float2 v = __ldg(tb + grid_index(hashed, lv.res[l], lv.size[l], c.g[0] + (k & 1), c.g[1] + ((k >> 1) & 1), c.g[2] + (k >> 2)));
        const float wgt = ((k & 1) ? c.w[0] : 1.f - c.w[0]) * ((k & 2) ? c.w[1] : 1.f - c.w[1]) * ((k & 4) ? c.w[2] : 1.f - c.w[2]);
        a0 = fmaf(wgt, v.x, a0);
        a1 = fmaf(wgt, v.y, a1);
    }
    out[t * 2] = a0;
    out[t * 2 + 1] = a1;
}

__global__ void __launch_bounds__(256) k_hashgrid_bwd(const float* __restrict__ x01, const uint32_t M, const float2* __restrict__ table,
                                                      const Levels lv, const uint32_t L, const float* __restrict__ g_out,
                                                      float2* __restrict__ g_table, float* __restrict__ g_x) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= (size_t)M * L) return;
    const uint32_t i = (uint32_t)(t / L), l = (uint32_t)(t % L);
    const float ga = g_out[t * 2], gb = g_out[t * 2 + 1];
    if (ga == 0.f && gb == 0.f) return;
    const Cell c = locate(x01[(size_t)i * 3], x01[(size_t)i * 3 + 1], x01[(size_t)i * 3 + 2], lv.scale[l]);
    const bool hashed = (lv.hashed >> l) & 1u;
    float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f;
    const float sc = lv.scale[l];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float wx = (k & 1) ? c.w[0] : 1.f - c.w[0], wy = (k & 2) ? c.w[1] : 1.f - c.w[1], wz = (k & 4) ? c.w[2] : 1.f - c.w[2];
        const uint32_t idx = grid_index(hashed, lv.res[l], lv.size[l], c.g[0] + (k & 1), c.g[1] + ((k >> 1) & 1), c.g[2] + (k >> 2));
        if (g_table) {
            const float wgt = wx * wy * wz;
            float2* addr = g_table + lv.off[l] + idx;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(wgt * ga), "f"(wgt * gb) : "memory");
        }
        if (g_x) {
            const float2 v = __ldg(table + lv.off[l] + idx);
            const float gv = v.x * ga + v.y * gb;
            gx0 = fmaf(((k & 1) ? 1.f : -1.f) * c.dw[0] * sc * wy * wz, gv, gx0);
            gx1 = fmaf(((k & 2) ? 1.f : -1.f) * c.dw[1] * sc * wx * wz, gv, gx1);
            gx2 = fmaf(((k & 4) ? 1.f : -1.f) * c.dw[2] * sc * wx * wy, gv, gx2);
        }
    }
    if (g_x) { atomicAdd(&g_x[(size_t)i * 3], gx0); atomicAdd(&g_x[(size_t)i * 3 + 1], gx1); atomicAdd(&g_x[(size_t)i * 3 + 2], gx2); }
}

}  // namespace

extern "C" {

int mve_hashgrid_forward(const float* x01, uint32_t M, const float* table, uint32_t n_levels, const float* level_scale, const uint32_t* level_res,
                         const uint32_t* level_size, const uint32_t* level_offset, float* out, void* stream) {
    if (M == 0) return 0;
    Levels lv;
    MVE_ARG(fill_levels(lv, n_levels, level_scale, level_res, level_size, level_offset) == 0, "hashgrid: at most 16 levels");
    k_hashgrid_fwd<<<cdiv((unsigned long long)M * n_levels, 256), 256, 0, (cudaStream_t)stream>>>(x01, M, (const float2*)table, lv, n_levels, out);
    MVE_CHECK_LAUNCH("mve_hashgrid_forward");
    return 0;
}

int mve_hashgrid_backward(const float* x01, uint32_t M, const float* table, uint32_t n_levels, const float* level_scale, const uint32_t* level_res,
                          const uint32_t* level_size, const uint32_t* level_offset, const float* grad_out, float* grad_table, float* grad_x,
                          void* stream) {
    if (M == 0) return 0;
    Levels lv;
    MVE_ARG(fill_levels(lv, n_levels, level_scale, level_res, level_size, level_offset) == 0, "hashgrid: at most 16 levels");
    MVE_ARG(grad_table || grad_x, "hashgrid_backward: nothing to compute");
    k_hashgrid_bwd<<<cdiv((unsigned long long)M * n_levels, 256), 256, 0, (cudaStream_t)stream>>>(x01, M, (const float2*)table, lv, n_levels, grad_out,
                                                                                               (float2*)grad_table, grad_x);
    MVE_CHECK_LAUNCH("mve_hashgrid_backward");
    return 0;
}

}  // extern "C"

// field.cu -- fused Instant-NGP field for the 3D adapter:
//   multi-resolution hash grid (Smoothstep) -> Linear(2L,64)+ReLU -> Linear(64,4) -> trunc_exp(+blob) / sigmoid
// in ONE kernel per direction, one thread per sample.  Replaces, on the reference path,
//   tinycudann.Encoding fwd/bwd + torch nn.Linear x2 (cuBLAS) + activations
//   (/root/reference/lib/models/decoders/ingp_decoder.py:106-120, lib/ops/activation.py:8-23).
//
// Forward: the 64 hidden units are never stored -- unit j is computed from the 2L encoded features and folded
// straight into the 4 outputs (weights are smem records of [W1[j][:], W2[:][j], b1[j]], read as broadcast float4).
// Backward: re-gathers and re-computes the forward (no activations are saved: xyz is the only saved tensor),
// then  d(enc) -> vectorised red.global.add.v2.f32 scatter into the table gradient (8 corners x L levels),
// and MLP weight gradients reduced warp-cooperatively through shared memory in chunks of 8 hidden units, kept
// in registers across the CTA's grid-stride loop and written once per CTA to a [CTA][n_param] workspace that a
// second tiny kernel sums in a fixed order (deterministic MLP gradients; the table scatter is atomic like tcnn's).
// Optionally d/d xyz (needed by the DMTet stage, base_mesh_renderer.py:277-283).
#include "field_device.cuh"
#include "mlp_mma.cuh"
#include "field_bwd_mma.cuh"
#include "../../include/mvedit_b200.h"

using namespace field;

namespace {

template <int L, bool DENSITY_ONLY>
__global__ void __launch_bounds__(256) k_field_fwd(const float* __restrict__ xyz, uint32_t M, const int* __restrict__ M_dev,
                                                   const float2* __restrict__ table, const float* __restrict__ w1,
                                                   const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                   const Levels lv, const FieldCfg cfg, float* __restrict__ sigma, float* __restrict__ rgb) {
    using R = Rec<L>;
    __shared__ __align__(16) float rec[HID * R::STRIDE];
    stage_mlp<L>(rec, w1, b1, w2);
    __syncthreads();
    if (M_dev) M = min(M, (uint32_t)*M_dev);
    const float ob0 = b2[0], ob1 = b2[1], ob2 = b2[2], ob3 = b2[3];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float x = xyz[(size_t)i * 3], y = xyz[(size_t)i * 3 + 1], z = xyz[(size_t)i * 3 + 2];
        float enc[R::IN];
        encode<L>(lv, table, unit_coord(cfg, x), unit_coord(cfg, y), unit_coord(cfg, z), enc);
        float o0 = ob0, o1 = ob1, o2 = ob2, o3 = ob3;
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
        sigma[i] = __expf(o0 + blob_of(cfg, x, y, z));
        if (!DENSITY_ONLY) {
            rgb[(size_t)i * 3] = fmaf(1.f / (1.f + __expf(-o1)), cfg.sat_scale, cfg.sat_shift);
            rgb[(size_t)i * 3 + 1] = fmaf(1.f / (1.f + __expf(-o2)), cfg.sat_scale, cfg.sat_shift);
            rgb[(size_t)i * 3 + 2] = fmaf(1.f / (1.f + __expf(-o3)), cfg.sat_scale, cfg.sat_shift);
        }
    }
}

// Tensor-core forward: the MLP runs as per-warp TF32 mma.sync GEMMs (mlp_mma.cuh; TF32 is the reference's own matmul precision,
// allow_tf32) and the hash-grid encoder is a rolled per-level loop that writes straight into the MMA staging tile.
// DENSITY_ONLY: the culling pre-pass (point_density_decode before weight culling, base_volume_renderer.py:222-227).
template <int L, bool DENSITY_ONLY>
__global__ void __launch_bounds__(128) k_field_fwd_mma(const float* __restrict__ xyz, uint32_t M, const int* __restrict__ M_dev,
                                                       const float2* __restrict__ table, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, const Levels lv, const FieldCfg cfg,
                                                       float* __restrict__ sigma, float* __restrict__ rgb) {
    using MC = mlpmma::Cfg<L>;
    __shared__ __align__(16) float frags[MC::FRAG_FLOATS];
    __shared__ __align__(16) float stage[4][MC::STAGE_FLOATS];
    mlpmma::stage_frags<L>(frags, w1, b1, w2);
    float* const stg = stage[threadIdx.x >> 5];
    mlpmma::zero_stage_pad<L>(stg);
    __syncthreads();
    if (M_dev) M = min(M, (uint32_t)*M_dev);
    const float ob0 = b2[0], ob1 = b2[1], ob2 = b2[2], ob3 = b2[3];
    const uint32_t lane = threadIdx.x & 31, warp_g = blockIdx.x * 4 + (threadIdx.x >> 5), n_warps = gridDim.x * 4;
    for (uint32_t base = warp_g * 32; base < M; base += n_warps * 32) {      // warp-uniform trip count
        const uint32_t i = base + lane;
        const bool live = i < M;
        float x = 0.f, y = 0.f, z = 0.f;
        if (live) { x = xyz[(size_t)i * 3]; y = xyz[(size_t)i * 3 + 1]; z = xyz[(size_t)i * 3 + 2]; }
        mlpmma::encode_staged<L, 1>(lv, table, unit_coord(cfg, x), unit_coord(cfg, y), unit_coord(cfg, z), live, stg);
        float o[4];
        mlpmma::mlp_forward_staged<L>(stg, frags, o);
        if (live) {
            sigma[i] = __expf(o[0] + ob0 + blob_of(cfg, x, y, z));
            if (!DENSITY_ONLY) {
                rgb[(size_t)i * 3] = fmaf(1.f / (1.f + __expf(-(o[1] + ob1))), cfg.sat_scale, cfg.sat_shift);
                rgb[(size_t)i * 3 + 1] = fmaf(1.f / (1.f + __expf(-(o[2] + ob2))), cfg.sat_scale, cfg.sat_shift);
                rgb[(size_t)i * 3 + 2] = fmaf(1.f / (1.f + __expf(-(o[3] + ob3))), cfg.sat_scale, cfg.sat_shift);
            }
        }
    }
}

__device__ __forceinline__ void red_add_v2(float2* addr, const float a, const float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// number of MLP parameters / workspace floats per CTA
template <int L>
constexpr int n_mlp() { return HID * 2 * L + HID + 4 * HID + 4; }

constexpr int BW_T = 128;  // backward CTA: 4 warps
constexpr int JC = 8;      // hidden units per weight-gradient chunk

template <int L, bool WITH_DX>
__global__ void __launch_bounds__(BW_T) k_field_bwd(const float* __restrict__ xyz, uint32_t M, const int* __restrict__ M_dev,
                                                    const float2* __restrict__ table, const float* __restrict__ w1,
                                                    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                    const Levels lv, const FieldCfg cfg, const float* __restrict__ g_sigma,
                                                    const float* __restrict__ g_rgb, float2* __restrict__ g_table,
                                                    float* __restrict__ workspace, float* __restrict__ g_xyz) {
    using R = Rec<L>;
    constexpr int IN = R::IN;
    constexpr int ENC_LD = IN + 1;  // odd stride: conflict-free column reads
    constexpr int IB = IN / 4;      // dW1 entries per lane per chunk: lane -> (j = lane/4, i-block = lane%4 of IB entries)
    __shared__ __align__(16) float rec[HID * R::STRIDE];
    __shared__ float s_enc[BW_T / 32][32 * ENC_LD];
    __shared__ float s_dh[BW_T / 32][32 * (JC + 1)];
    __shared__ float s_hr[BW_T / 32][32 * (JC + 1)];
    __shared__ float s_do[BW_T / 32][32 * 5];
    stage_mlp<L>(rec, w1, b1, w2);
    __syncthreads();
    if (M_dev) M = min(M, (uint32_t)*M_dev);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* enc_s = s_enc[warp];
    float* dh_s = s_dh[warp];
    float* hr_s = s_hr[warp];
    float* do_s = s_do[warp];
    const float ob0 = b2[0], ob1 = b2[1], ob2 = b2[2], ob3 = b2[3];

    // per-lane gradient accumulators (kept across the grid-stride loop)
    float aW1[HID / JC][IB];   // chunk c, entry (j = c*JC + lane/4, i = (lane%4)*IB + e)
    float aW2[HID / JC];       // chunk c, entry (k = lane/8, j = c*JC + lane%8)
    float aB1[HID / JC];       // chunk c, lanes 0..7: j = c*JC + lane
    float aB2 = 0.f;           // lanes 0..3: k = lane
#pragma unroll
    for (int c = 0; c < HID / JC; c++) {
        aW2[c] = 0.f; aB1[c] = 0.f;
#pragma unroll
        for (int e = 0; e < IB; e++) aW1[c][e] = 0.f;
    }

    // tiles are WARP-sized (32 samples): after culling a reconstruction iteration keeps only a few thousand samples, and the
    // weight-gradient reduction is warp-cooperative anyway, so spreading 32-sample tiles over all resident warps (instead of
    // 128-sample tiles over CTAs) turns a ~50-CTA latency-bound launch into one tile per warp.
    const uint32_t n_tiles = (M + 31) / 32, total_warps = gridDim.x * (BW_T / 32);
    for (uint32_t tile = blockIdx.x * (BW_T / 32) + warp; tile < n_tiles; tile += total_warps) {
        const uint32_t i = tile * 32 + lane;
        const bool live = i < M;
        float x = 0.f, y = 0.f, z = 0.f, gs = 0.f, gr = 0.f, gg = 0.f, gb = 0.f;
        if (live) {
            x = xyz[(size_t)i * 3]; y = xyz[(size_t)i * 3 + 1]; z = xyz[(size_t)i * 3 + 2];
            gs = g_sigma[i];
            if (g_rgb) { gr = g_rgb[(size_t)i * 3]; gg = g_rgb[(size_t)i * 3 + 1]; gb = g_rgb[(size_t)i * 3 + 2]; }
        }
        const float x0 = unit_coord(cfg, x), x1 = unit_coord(cfg, y), x2 = unit_coord(cfg, z);
        float enc[IN];
        encode<L>(lv, table, x0, x1, x2, enc);
        // ---- forward recompute of the 4 outputs
        float o0 = ob0, o1 = ob1, o2 = ob2, o3 = ob3;
#pragma unroll 4
        for (int j = 0; j < HID; j++) {
            const float4* r4 = reinterpret_cast<const float4*>(rec + j * R::STRIDE);
            float a = rec[j * R::STRIDE + R::B1O];
#pragma unroll
            for (int q = 0; q < IN / 4; q++) {
                const float4 w = r4[q];
                a = fmaf(w.x, enc[4 * q], a); a = fmaf(w.y, enc[4 * q + 1], a); a = fmaf(w.z, enc[4 * q + 2], a); a = fmaf(w.w, enc[4 * q + 3], a);
            }
            a = fmaxf(a, 0.f);
            const float4 v = r4[R::W2O / 4];
            o0 = fmaf(v.x, a, o0); o1 = fmaf(v.y, a, o1); o2 = fmaf(v.z, a, o2); o3 = fmaf(v.w, a, o3);
        }
        // ---- output activations backward (activation.py:18-22: clamp exp to [1e-6, 1e6] in the backward)
        const float blob = blob_of(cfg, x, y, z);
        const float e = __expf(o0 + blob);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        if (live) {
            d0 = gs * fminf(fmaxf(e, 1e-6f), 1e6f);
            const float s1 = 1.f / (1.f + __expf(-o1)), s2 = 1.f / (1.f + __expf(-o2)), s3 = 1.f / (1.f + __expf(-o3));
            d1 = gr * cfg.sat_scale * s1 * (1.f - s1);
            d2 = gg * cfg.sat_scale * s2 * (1.f - s2);
            d3 = gb * cfg.sat_scale * s3 * (1.f - s3);
        }
        // stage enc and dout for the warp-cooperative weight gradients
#pragma unroll
        for (int q = 0; q < IN; q++) enc_s[lane * ENC_LD + q] = enc[q];
        do_s[lane * 5 + 0] = d0; do_s[lane * 5 + 1] = d1; do_s[lane * 5 + 2] = d2; do_s[lane * 5 + 3] = d3;
        float denc[IN];
#pragma unroll
        for (int q = 0; q < IN; q++) denc[q] = 0.f;
        __syncwarp();
        if (lane < 4) {
            float sacc = 0.f;
            for (int s = 0; s < 32; s++) sacc += do_s[s * 5 + lane];
            aB2 += sacc;
        }
        // ---- hidden layer backward, JC units at a time
#pragma unroll
        for (int c = 0; c < HID / JC; c++) {
#pragma unroll
            for (int jj = 0; jj < JC; jj++) {
                const int j = c * JC + jj;
                const float4* r4 = reinterpret_cast<const float4*>(rec + j * R::STRIDE);
                float a = rec[j * R::STRIDE + R::B1O];
#pragma unroll
                for (int q = 0; q < IN / 4; q++) {
                    const float4 w = r4[q];
                    a = fmaf(w.x, enc[4 * q], a); a = fmaf(w.y, enc[4 * q + 1], a); a = fmaf(w.z, enc[4 * q + 2], a); a = fmaf(w.w, enc[4 * q + 3], a);
                }
                const float4 v = r4[R::W2O / 4];
                const float dh = (a > 0.f) ? (v.x * d0 + v.y * d1 + v.z * d2 + v.w * d3) : 0.f;
#pragma unroll
                for (int q = 0; q < IN / 4; q++) {
                    const float4 w = r4[q];
                    denc[4 * q] = fmaf(w.x, dh, denc[4 * q]); denc[4 * q + 1] = fmaf(w.y, dh, denc[4 * q + 1]);
                    denc[4 * q + 2] = fmaf(w.z, dh, denc[4 * q + 2]); denc[4 * q + 3] = fmaf(w.w, dh, denc[4 * q + 3]);
                }
                dh_s[lane * (JC + 1) + jj] = dh;
                hr_s[lane * (JC + 1) + jj] = fmaxf(a, 0.f);
            }
            __syncwarp();
            {   // weight gradients of this chunk: every lane sums over the warp's 32 samples
                const int jl = lane >> 2, ib = (lane & 3) * IB;
                float t1[IB];
#pragma unroll
                for (int e2 = 0; e2 < IB; e2++) t1[e2] = 0.f;
                float t2 = 0.f, t3 = 0.f;
                const int k2 = lane >> 3, j2 = lane & 7;
#pragma unroll 4
                for (int s = 0; s < 32; s++) {
                    const float dhv = dh_s[s * (JC + 1) + jl];
#pragma unroll
                    for (int e2 = 0; e2 < IB; e2++) t1[e2] = fmaf(dhv, enc_s[s * ENC_LD + ib + e2], t1[e2]);
                    t2 = fmaf(do_s[s * 5 + k2], hr_s[s * (JC + 1) + j2], t2);
                    if (lane < JC) t3 += dh_s[s * (JC + 1) + lane];
                }
#pragma unroll
                for (int e2 = 0; e2 < IB; e2++) aW1[c][e2] += t1[e2];
                aW2[c] += t2;
                aB1[c] += t3;
            }
            __syncwarp();
        }
        // ---- scatter d(enc) into the table gradient (+ optional d/dx)
        if (live) {
            float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f;
#pragma unroll
            for (int l = 0; l < L; l++) {
                const Cell cl = locate(x0, x1, x2, lv.scale[l]);
                const bool hashed = (lv.hashed >> l) & 1u;
                const uint32_t res = lv.res[l], size = lv.size[l];
                float2* __restrict__ gt = g_table + lv.off[l];
                const float ga = denc[2 * l], gb2 = denc[2 * l + 1];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float wx = (k & 1) ? cl.w[0] : 1.f - cl.w[0], wy = (k & 2) ? cl.w[1] : 1.f - cl.w[1], wz = (k & 4) ? cl.w[2] : 1.f - cl.w[2];
                    const uint32_t idx = grid_index(hashed, res, size, cl.g[0] + (k & 1), cl.g[1] + ((k >> 1) & 1), cl.g[2] + (k >> 2));
                    const float wgt = wx * wy * wz;
                    if (ga != 0.f || gb2 != 0.f) red_add_v2(gt + idx, wgt * ga, wgt * gb2);
                    if (WITH_DX) {
                        const float2 v = __ldg(table + lv.off[l] + idx);
                        const float gv = v.x * ga + v.y * gb2;
                        const float sc = lv.scale[l];
                        gx0 = fmaf(((k & 1) ? 1.f : -1.f) * cl.dw[0] * sc * wy * wz, gv, gx0);
                        gx1 = fmaf(((k & 2) ? 1.f : -1.f) * cl.dw[1] * sc * wx * wz, gv, gx1);
                        gx2 = fmaf(((k & 4) ? 1.f : -1.f) * cl.dw[2] * sc * wx * wy, gv, gx2);
                    }
                }
            }
            if (WITH_DX) {
                // x01 = (x + bound) / (2 bound); blob(x) also depends on x where |x|^2 > 0.2 (ingp_decoder.py:101-104)
                const float r2 = x * x + y * y + z * z;
                const float gblob = (r2 > 0.2f) ? d0 * blob * (-2.f * cfg.blob_k) : 0.f;
                g_xyz[(size_t)i * 3] = gx0 * cfg.inv2b + gblob * x;
                g_xyz[(size_t)i * 3 + 1] = gx1 * cfg.inv2b + gblob * y;
                g_xyz[(size_t)i * 3 + 2] = gx2 * cfg.inv2b + gblob * z;
            }
        }
        __syncwarp();
    }

    // ---- CTA reduction of the MLP gradients through shared memory, one workspace row per CTA
    // layout of a row: [dW1 (HID x IN) | db1 (HID) | dW2 (4 x HID) | db2 (4)]
    __syncthreads();
    float* row = workspace + (size_t)blockIdx.x * n_mlp<L>();
    __shared__ float s_red[BW_T / 32][JC * 2 * L + 32 + JC];  // per warp: chunk's dW1 (JC x IN) | dW2 (4 x JC) | db1 (JC)
    for (int c = 0; c < HID / JC; c++) {
        const int jl = lane >> 2, ib = (lane & 3) * IB;
        // select chunk c without dynamic register indexing
        float v1[IB], v2 = 0.f, v3 = 0.f;
#pragma unroll
        for (int cc = 0; cc < HID / JC; cc++)
            if (cc == c) {
#pragma unroll
                for (int e2 = 0; e2 < IB; e2++) v1[e2] = aW1[cc][e2];
                v2 = aW2[cc]; v3 = aB1[cc];
            }
#pragma unroll
        for (int e2 = 0; e2 < IB; e2++) s_red[warp][jl * IN + ib + e2] = v1[e2];
        s_red[warp][JC * IN + lane] = v2;                 // (k = lane/8, j = lane%8)
        if (lane < JC) s_red[warp][JC * IN + 32 + lane] = v3;
        __syncthreads();
        for (int t = threadIdx.x; t < JC * IN + 32 + JC; t += BW_T) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < BW_T / 32; w++) s += s_red[w][t];
            if (t < JC * IN) row[(c * JC + t / IN) * IN + t % IN] = s;                       // dW1[j][i]
            else if (t < JC * IN + 32) { const int q = t - JC * IN; row[HID * IN + HID + (q >> 3) * HID + c * JC + (q & 7)] = s; }  // dW2[k][j]
            else row[HID * IN + c * JC + (t - JC * IN - 32)] = s;                            // db1[j]
        }
        __syncthreads();
    }
    {
        __shared__ float s_b2[BW_T / 32][4];
        if (lane < 4) s_b2[warp][lane] = aB2;
        __syncthreads();
        if (threadIdx.x < 4) {
            float s = 0.f;
            for (int w = 0; w < BW_T / 32; w++) s += s_b2[w][threadIdx.x];
            row[HID * IN + HID + 4 * HID + threadIdx.x] = s;
        }
    }
}

// Tensor-core backward (field_bwd_mma.cuh): rolled encoder -> per-warp TF32 MMA backward of the MLP -> rolled scatter of d(enc)
// into the table gradient.  Same interface, workspace layout and two-stage deterministic MLP reduction as k_field_bwd.
template <int L, bool WITH_DX>
__global__ void __launch_bounds__(BW_T, 3) k_field_bwd_mma(const float* __restrict__ xyz, uint32_t M, const int* __restrict__ M_dev,
                                                           const float2* __restrict__ table, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, const Levels lv, const FieldCfg cfg,
                                                           const float* __restrict__ g_sigma, const float* __restrict__ g_rgb,
                                                           float2* __restrict__ g_table, float* __restrict__ workspace,
                                                           float* __restrict__ g_xyz) {
    using B = mlpmma::BCfg<L>;
    constexpr int IN = B::IN, LD = B::LD, WARP_FLOATS = B::STAGE_FLOATS + B::XCH_FLOATS;
    extern __shared__ __align__(16) float smem[];
    float* const fr = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* const stage = smem + B::FRAG_FLOATS + warp * WARP_FLOATS;
    float* const xch = stage + B::STAGE_FLOATS;
    mlpmma::stage_bwd_frags<L>(fr, w1, b1, w2);
    mlpmma::init_bwd_stage<L>(stage);
    __syncthreads();
    if (M_dev) M = min(M, (uint32_t)*M_dev);
    const float ob0 = b2[0], ob1 = b2[1], ob2 = b2[2], ob3 = b2[3];
    mlpmma::MlpGradAcc<L> acc;
    acc.clear();

    const uint32_t n_tiles = (M + 31) / 32, total_warps = gridDim.x * (BW_T / 32);
    for (uint32_t tile = blockIdx.x * (BW_T / 32) + warp; tile < n_tiles; tile += total_warps) {
        const uint32_t i = tile * 32 + lane;
        const bool live = i < M;
        float x = 0.f, y = 0.f, z = 0.f, gs = 0.f, gr = 0.f, gg = 0.f, gb = 0.f;
        if (live) {
            x = xyz[(size_t)i * 3]; y = xyz[(size_t)i * 3 + 1]; z = xyz[(size_t)i * 3 + 2];
            gs = g_sigma[i];
            if (g_rgb) { gr = g_rgb[(size_t)i * 3]; gg = g_rgb[(size_t)i * 3 + 1]; gb = g_rgb[(size_t)i * 3 + 2]; }
        }
        const float x0 = unit_coord(cfg, x), x1 = unit_coord(cfg, y), x2 = unit_coord(cfg, z);
        mlpmma::encode_staged<L, 1>(lv, table, x0, x1, x2, live, stage);
        const float blob = blob_of(cfg, x, y, z);
        float d0 = 0.f;
        mlpmma::mlp_backward_tile<L>(stage, xch, fr, acc, [&](const float (&o)[4], float (&d)[4]) {
            // output activations backward (activation.py:18-22: clamp exp to [1e-6, 1e6] in the backward)
            d[0] = d[1] = d[2] = d[3] = 0.f;
            if (live) {
                const float e = __expf(o[0] + ob0 + blob);
                d[0] = gs * fminf(fmaxf(e, 1e-6f), 1e6f);
                const float s1 = 1.f / (1.f + __expf(-(o[1] + ob1))), s2 = 1.f / (1.f + __expf(-(o[2] + ob2))),
                            s3 = 1.f / (1.f + __expf(-(o[3] + ob3)));
                d[1] = gr * cfg.sat_scale * s1 * (1.f - s1);
                d[2] = gg * cfg.sat_scale * s2 * (1.f - s2);
                d[3] = gb * cfg.sat_scale * s3 * (1.f - s3);
            }
            d0 = d[0];
        });
        // ---- scatter d(enc) into the table gradient (+ optional d/dx), one level per trip
        float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f;
        // warps start at different levels: the coarse levels have few entries (17^3 at level 0) and the L2 atomic unit serialises
        // per address, so warps that walk the levels in lockstep queue up on the same lines
        const int rot = (int)(tile % (uint32_t)L);
#pragma unroll 1
        for (int lq = 0; lq < L; lq++) {
            const int l = lq + rot < L ? lq + rot : lq + rot - L;
            const float ga = stage[(2 * l) * LD + lane], gb2 = stage[(2 * l + 1) * LD + lane];
            if (live && (ga != 0.f || gb2 != 0.f)) {
                const Cell cl = locate(x0, x1, x2, lv.scale[l]);
                const bool hashed = (lv.hashed >> l) & 1u;
                const uint32_t res = lv.res[l], size = lv.size[l];
                float2* __restrict__ gt = g_table + lv.off[l];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float wx = (k & 1) ? cl.w[0] : 1.f - cl.w[0], wy = (k & 2) ? cl.w[1] : 1.f - cl.w[1], wz = (k & 4) ? cl.w[2] : 1.f - cl.w[2];
                    const uint32_t idx = grid_index(hashed, res, size, cl.g[0] + (k & 1), cl.g[1] + ((k >> 1) & 1), cl.g[2] + (k >> 2));
                    const float wgt = wx * wy * wz;
                    red_add_v2(gt + idx, wgt * ga, wgt * gb2);
                    if (WITH_DX) {
                        const float2 v = __ldg(table + lv.off[l] + idx);
                        const float gv = v.x * ga + v.y * gb2;
                        const float sc = lv.scale[l];
                        gx0 = fmaf(((k & 1) ? 1.f : -1.f) * cl.dw[0] * sc * wy * wz, gv, gx0);
                        gx1 = fmaf(((k & 2) ? 1.f : -1.f) * cl.dw[1] * sc * wx * wz, gv, gx1);
                        gx2 = fmaf(((k & 4) ? 1.f : -1.f) * cl.dw[2] * sc * wx * wy, gv, gx2);
                    }
                }
            }
        }
        if (WITH_DX && live) {
            // x01 = (x + bound) / (2 bound); blob(x) also depends on x where |x|^2 > 0.2 (ingp_decoder.py:101-104)
            const float r2 = x * x + y * y + z * z;
            const float gblob = (r2 > 0.2f) ? d0 * blob * (-2.f * cfg.blob_k) : 0.f;
            g_xyz[(size_t)i * 3] = gx0 * cfg.inv2b + gblob * x;
            g_xyz[(size_t)i * 3 + 1] = gx1 * cfg.inv2b + gblob * y;
            g_xyz[(size_t)i * 3 + 2] = gx2 * cfg.inv2b + gblob * z;
        }
        __syncwarp();
    }

    // ---- CTA reduction of the MLP gradients in a fixed warp order, one workspace row per CTA
    // layout of a row: [dW1 (HID x IN) | db1 (HID) | dW2 (4 x HID) | db2 (4)]
    __syncthreads();
    constexpr int NP = n_mlp<L>();
    float* const s_acc = smem + B::FRAG_FLOATS;           // the staging tiles are free now
    static_assert(NP <= (BW_T / 32) * WARP_FLOATS, "CTA accumulator does not fit in the staging area");
    for (int t = threadIdx.x; t < NP; t += BW_T) s_acc[t] = 0.f;
    __syncthreads();
    const int g = lane >> 2, t4 = lane & 3;
    for (int w = 0; w < BW_T / 32; w++) {
        if (warp == w) {
#pragma unroll
            for (int mt = 0; mt < B::MT; mt++) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int hid = 16 * mt + g + 8 * (c >> 1);
#pragma unroll
                    for (int ft = 0; ft < B::WT; ft++) {
                        const int f = 8 * ft + 2 * t4 + (c & 1);
                        if (f < IN) s_acc[hid * IN + f] += acc.w1[mt][ft][c];
                        else if (f == IN) s_acc[HID * IN + hid] += acc.w1[mt][ft][c];
                    }
                    const int o = 2 * t4 + (c & 1);
                    if (o < 4) s_acc[HID * IN + HID + o * HID + hid] += acc.w2[mt][c];
                }
            }
            if (lane < 4) s_acc[HID * IN + HID + 4 * HID + lane] += acc.b2;
        }
        __syncthreads();
    }
    float* row = workspace + (size_t)blockIdx.x * NP;
    for (int t = threadIdx.x; t < NP; t += BW_T) row[t] = s_acc[t];
}

// sum the per-CTA rows in a fixed order -> g_w1, g_b1, g_w2, g_b2 (accumulate = add to existing .grad)
template <int L>
__global__ void k_field_reduce_mlp(const float* __restrict__ workspace, const uint32_t n_rows, float* __restrict__ g_w1,
                                   float* __restrict__ g_b1, float* __restrict__ g_w2, float* __restrict__ g_b2, const bool accumulate) {
    constexpr int IN = 2 * L, NP = n_mlp<L>();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NP) return;
    float s = 0.f;
    for (uint32_t r = 0; r < n_rows; r++) s += workspace[(size_t)r * NP + t];
    float* dst;
    if (t < HID * IN) dst = g_w1 + t;
    else if (t < HID * IN + HID) dst = g_b1 + (t - HID * IN);
    else if (t < HID * IN + HID + 4 * HID) dst = g_w2 + (t - HID * IN - HID);
    else dst = g_b2 + (t - HID * IN - HID - 4 * HID);
    *dst = accumulate ? *dst + s : s;
}

}  // namespace

extern "C" {

uint32_t mve_field_backward_workspace_floats(uint32_t n_levels) {
    return (uint32_t)(3 * kNumSM) * (uint32_t)(HID * 2 * n_levels + HID + 4 * HID + 4);
}

int mve_field_forward(const float* xyz, uint32_t M, const int32_t* M_dev, const float* table, const float* w1, const float* b1,
                      const float* w2, const float* b2, uint32_t n_levels, const float* level_scale, const uint32_t* level_res,
                      const uint32_t* level_size, const uint32_t* level_offset, float bound, float blob_density, float blob_radius,
                      float sigmoid_saturation, int density_only, float* sigma, float* rgb, void* stream) {
    if (M == 0) return 0;
    Levels lv;
    MVE_ARG(fill_levels(lv, n_levels, level_scale, level_res, level_size, level_offset) == 0, "field: n_levels > 16");
    MVE_ARG(n_levels == 12 || n_levels == 14 || n_levels == 16, "field: n_levels must be 12, 14 or 16");
    const FieldCfg cfg = make_cfg(bound, blob_density, blob_radius, sigmoid_saturation);
    uint32_t grid = cdiv(M, 256);
    if (grid > (uint32_t)(8 * kNumSM)) grid = 8 * kNumSM;
    uint32_t grid2 = cdiv(M, 128);
    if (grid2 > (uint32_t)(8 * kNumSM)) grid2 = 8 * kNumSM;
    const float2* t2 = reinterpret_cast<const float2*>(table);
    cudaStream_t s = (cudaStream_t)stream;
#define FWD(LL)                                                                                                                          \
    if (density_only == 2) k_field_fwd_mma<LL, true><<<grid2, 128, 0, s>>>(xyz, M, M_dev, t2, w1, b1, w2, b2, lv, cfg, sigma, rgb);      \
    else if (density_only == 3) k_field_fwd_mma<LL, false><<<grid2, 128, 0, s>>>(xyz, M, M_dev, t2, w1, b1, w2, b2, lv, cfg, sigma, rgb); \
    else if (density_only) k_field_fwd<LL, true><<<grid, 256, 0, s>>>(xyz, M, M_dev, t2, w1, b1, w2, b2, lv, cfg, sigma, rgb);           \
    else k_field_fwd<LL, false><<<grid, 256, 0, s>>>(xyz, M, M_dev, t2, w1, b1, w2, b2, lv, cfg, sigma, rgb);
    if (n_levels == 12) { FWD(12) } else if (n_levels == 14) { FWD(14) } else { FWD(16) }
#undef FWD
    MVE_CHECK_LAUNCH("mve_field_forward");
    return 0;
}

int mve_field_backward(const float* xyz, uint32_t M, const int32_t* M_dev, const float* table, const float* w1, const float* b1,
                       const float* w2, const float* b2, uint32_t n_levels, const float* level_scale, const uint32_t* level_res,
                       const uint32_t* level_size, const uint32_t* level_offset, float bound, float blob_density, float blob_radius,
                       float sigmoid_saturation, const float* grad_sigma, const float* grad_rgb, float* grad_table, float* grad_w1,
                       float* grad_b1, float* grad_w2, float* grad_b2, int accumulate_mlp, int mlp_tf32, float* workspace,
                       float* grad_xyz, void* stream) {
    Levels lv;
    MVE_ARG(fill_levels(lv, n_levels, level_scale, level_res, level_size, level_offset) == 0, "field: n_levels > 16");
    MVE_ARG(n_levels == 12 || n_levels == 14 || n_levels == 16, "field: n_levels must be 12, 14 or 16");
    MVE_ARG(workspace != nullptr, "field backward: workspace required (mve_field_backward_workspace_floats)");
    const FieldCfg cfg = make_cfg(bound, blob_density, blob_radius, sigmoid_saturation);
    const int ctas_per_sm = mlp_tf32 ? 3 : 2;
    uint32_t grid = cdiv(M > 0 ? M : 1, BW_T);
    if (grid > (uint32_t)(ctas_per_sm * kNumSM)) grid = ctas_per_sm * kNumSM;   // M is the buffer capacity when M_dev is given: all CTAs launch
    const float2* t2 = reinterpret_cast<const float2*>(table);
    float2* gt2 = reinterpret_cast<float2*>(grad_table);
    cudaStream_t s = (cudaStream_t)stream;
#define BWD_MMA(LL, DX)                                                                                                                  \
    {                                                                                                                                    \
        using BC = mlpmma::BCfg<LL>;                                                                                                     \
        const size_t smem = sizeof(float) * (BC::FRAG_FLOATS + (BW_T / 32) * (BC::STAGE_FLOATS + BC::XCH_FLOATS));                       \
        static bool attr_done = false;                                                                                                   \
        if (!attr_done) {                                                                                                                \
            MVE_CUDA(cudaFuncSetAttribute(k_field_bwd_mma<LL, DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
            attr_done = true;                                                                                                            \
        }                                                                                                                                \
        k_field_bwd_mma<LL, DX><<<grid, BW_T, smem, s>>>(xyz, M, M_dev, t2, w1, b1, w2, b2, lv, cfg, grad_sigma, grad_rgb, gt2,          \
                                                         workspace, grad_xyz);                                                           \
    }
#define BWD(LL)                                                                                                                          \
    {                                                                                                                                    \
        if (mlp_tf32) {                                                                                                                  \
            if (grad_xyz) BWD_MMA(LL, true) else BWD_MMA(LL, false)                                                                      \
        } else if (grad_xyz)                                                                                                             \
            k_field_bwd<LL, true><<<grid, BW_T, 0, s>>>(xyz, M, M_dev, t2, w1, b1, w2, b2, lv, cfg, grad_sigma, grad_rgb, gt2,           \
                                                        workspace, grad_xyz);                                                            \
        else                                                                                                                             \
            k_field_bwd<LL, false><<<grid, BW_T, 0, s>>>(xyz, M, M_dev, t2, w1, b1, w2, b2, lv, cfg, grad_sigma, grad_rgb, gt2,          \
                                                         workspace, grad_xyz);                                                           \
        k_field_reduce_mlp<LL><<<cdiv(n_mlp<LL>(), 256), 256, 0, s>>>(workspace, grid, grad_w1, grad_b1, grad_w2, grad_b2,               \
                                                                       accumulate_mlp != 0);                                             \
    }
    if (n_levels == 12) BWD(12) else if (n_levels == 14) BWD(14) else BWD(16)
#undef BWD
#undef BWD_MMA
    MVE_CHECK_LAUNCH("mve_field_backward");
    return 0;
}

}  // extern "C"

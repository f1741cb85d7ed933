// mlp_mma.cuh -- the field's 2-layer MLP (2L -> 64 -> 4) on tensor cores, per warp, for 32 samples at a time.
//
// Why: ncu on the FFMA version of the fused renderer shows an issue-bound kernel (IPC 2.1 of 4, 5 700 warp-instructions per
// 32 samples) in which the MLP is ~1 800 FFMA + ~450 broadcast LDS of those (profiles/r01_ncu_k_render_rays.txt).  As a per-warp
// GEMM  H[32 x 64] = Enc[32 x K] W1^T,  Out[32 x 8] = relu(H) W2^T  it is 64 mma.sync.m16n8k8 (TF32 inputs, FP32 accumulate) plus
// ~170 staging instructions.  TF32 is the reference's own precision for these matmuls (it runs with
// torch.backends.cuda.matmul.allow_tf32 = True, /root/reference/lib/apis/adapter3d.py:51-61).
// A 32-sample x 64 tile per warp is far below tcgen05's 128-row CTA-wide tiles (and its TMEM round trip), so the warp-level
// mma.sync path is the right tool here; the tcgen05 kernels are the UNet's (gemm_tc.cu, attention.cu).
//
// Layouts (m16n8k8, PTX ISA): g = lane >> 2, t = lane & 3
//   A: a0 (row g, k t)  a1 (row g+8, k t)  a2 (row g, k t+4)  a3 (row g+8, k t+4)
//   B: b0 (k t, n g)    b1 (k t+4, n g)
//   C: c0 (row g, n 2t) c1 (row g, n 2t+1) c2 (row g+8, n 2t) c3 (row g+8, n 2t+1)
// Layer 2 consumes layer 1's C fragments directly as A fragments (a0=c0, a1=c2, a2=c1, a3=c3): that maps A's logical k = t to
// hidden unit 8*nt + 2t and k = t+4 to unit 8*nt + 2t + 1, and W2's B fragments are laid out with the same permutation.
#pragma once
#include "field_device.cuh"

namespace mlpmma {

using field::HID;

template <int L>
struct Cfg {
    static constexpr int IN = 2 * L;
    static constexpr int KS = (IN + 7) / 8;          // k-steps of layer 1 (K padded to a multiple of 8 with zeros)
    static constexpr int NT = HID / 8;               // 8 n-tiles of layer 1 == 8 k-steps of layer 2
    static constexpr int ENC_LD = 40;                // [k][40]: conflict-free for both the per-lane stores and the fragment loads
    // shared-memory floats: B1 fragments | B2 fragments | b1
    static constexpr int B1_FLOATS = NT * KS * 32 * 2, B2_FLOATS = NT * 32 * 2, FRAG_FLOATS = B1_FLOATS + B2_FLOATS + HID;
    static constexpr int STAGE_FLOATS = KS * 8 * ENC_LD;   // per warp
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// once per CTA: weights -> fragment-ordered shared memory (TF32-rounded)
template <int L>
__device__ __forceinline__ void stage_frags(float* __restrict__ fr, const float* __restrict__ w1, const float* __restrict__ b1,
                                            const float* __restrict__ w2) {
    using C = Cfg<L>;
    uint32_t* fb = reinterpret_cast<uint32_t*>(fr);
    for (int i = threadIdx.x; i < C::NT * C::KS * 32; i += blockDim.x) {
        const int lane = i & 31, s = (i >> 5) % C::KS, nt = (i >> 5) / C::KS, g = lane >> 2, t = lane & 3;
        const int n = 8 * nt + g, k0 = 8 * s + t, k1 = k0 + 4;
        fb[2 * i] = to_tf32(k0 < C::IN ? w1[n * C::IN + k0] : 0.f);
        fb[2 * i + 1] = to_tf32(k1 < C::IN ? w1[n * C::IN + k1] : 0.f);
    }
    for (int i = threadIdx.x; i < C::NT * 32; i += blockDim.x) {
        const int lane = i & 31, nt = i >> 5, g = lane >> 2, t = lane & 3;
        fb[C::B1_FLOATS + 2 * i] = to_tf32(g < 4 ? w2[g * HID + 8 * nt + 2 * t] : 0.f);
        fb[C::B1_FLOATS + 2 * i + 1] = to_tf32(g < 4 ? w2[g * HID + 8 * nt + 2 * t + 1] : 0.f);
    }
    for (int i = threadIdx.x; i < HID; i += blockDim.x) fr[C::B1_FLOATS + C::B2_FLOATS + i] = b1[i];
}

// Hash-grid encoding of one sample per lane written straight into the warp's [k][sample] staging tile (TF32), one level per
// loop trip (UNROLL = 1; two levels per trip measured 10 % slower in the renderer and 30 % slower in the field forward).
// The level loop is deliberately NOT unrolled: the fully unrolled encoder is ~4 000 straight-line instructions
// (64 KB) that every warp streams through once per round, and with the warps of an SM at different points of the
// march/shade loop the fused renderer became instruction-fetch bound (ncu: stall_no_instruction 11 of 16 cycles per issue,
// profiles/r01_ncu_k_render_rays_bench_state.txt).  A ~130-instruction loop body stays resident in the instruction caches.
template <int L, int UNROLL>
__device__ __forceinline__ void encode_staged(const field::Levels& lv, const float2* __restrict__ table, const float x0, const float x1,
                                              const float x2, const bool live, float* __restrict__ stage) {
    using C = Cfg<L>;
    const int lane = threadIdx.x & 31;
    uint32_t* st = reinterpret_cast<uint32_t*>(stage);
#pragma unroll UNROLL
    for (int l = 0; l < L; l++) {
        float a0 = 0.f, a1 = 0.f;
        if (live) {
            const field::Cell c = field::locate(x0, x1, x2, lv.scale[l]);
            const bool hashed = (lv.hashed >> l) & 1u;
            const uint32_t res = lv.res[l], size = lv.size[l];
            const float2* __restrict__ t = table + lv.off[l];
            float2 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
                v[k] = __ldg(t + field::grid_index(hashed, res, size, c.g[0] + (k & 1), c.g[1] + ((k >> 1) & 1), c.g[2] + (k >> 2)));
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float wgt = ((k & 1) ? c.w[0] : 1.f - c.w[0]) * ((k & 2) ? c.w[1] : 1.f - c.w[1]) * ((k & 4) ? c.w[2] : 1.f - c.w[2]);
                a0 = fmaf(wgt, v[k].x, a0);
                a1 = fmaf(wgt, v[k].y, a1);
            }
        }
        st[(2 * l) * C::ENC_LD + lane] = to_tf32(a0);
        st[(2 * l + 1) * C::ENC_LD + lane] = to_tf32(a1);
    }
}

// zero the K-padding rows of a warp's staging tile (never written afterwards: the output exchange only touches floats 0..127)
template <int L>
__device__ __forceinline__ void zero_stage_pad(float* __restrict__ stage) {
    using C = Cfg<L>;
    const int lane = threadIdx.x & 31;
    static_assert(C::IN * C::ENC_LD >= 128, "output exchange would overlap the padding rows");
#pragma unroll
    for (int k = C::IN; k < C::KS * 8; k++) stage[k * C::ENC_LD + lane] = 0.f;
    __syncwarp();
}

template <int L>
__device__ __forceinline__ void mlp_forward_staged(float* __restrict__ stage, const float* __restrict__ fr, float (&out)[4]);

// All 32 lanes must call (mma.sync); lanes without a live sample pass zeros in enc.
// out[0..3] = W2 relu(W1 enc + b1) for THIS lane's sample (the output bias b2 is added by the caller).
template <int L>
__device__ __forceinline__ void mlp_forward(const float (&enc)[2 * L], float* __restrict__ stage, const float* __restrict__ fr, float (&out)[4]) {
    using C = Cfg<L>;
    const int lane = threadIdx.x & 31;
    uint32_t* st = reinterpret_cast<uint32_t*>(stage);
    // sample-major registers -> [k][sample] in shared memory
#pragma unroll
    for (int k = 0; k < C::KS * 8; k++) st[k * C::ENC_LD + lane] = (k < C::IN) ? to_tf32(enc[k < C::IN ? k : 0]) : 0u;
    mlp_forward_staged<L>(stage, fr, out);
}

// the MLP on an already staged [k][sample] tile (encode_staged); all 32 lanes must call
template <int L>
__device__ __forceinline__ void mlp_forward_staged(float* __restrict__ stage, const float* __restrict__ fr, float (&out)[4]) {
    using C = Cfg<L>;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint32_t* st = reinterpret_cast<uint32_t*>(stage);
    __syncwarp();
    uint32_t a[2][C::KS][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int s = 0; s < C::KS; s++) {
            a[mt][s][0] = st[(8 * s + t) * C::ENC_LD + 16 * mt + g];
            a[mt][s][1] = st[(8 * s + t) * C::ENC_LD + 16 * mt + g + 8];
            a[mt][s][2] = st[(8 * s + t + 4) * C::ENC_LD + 16 * mt + g];
            a[mt][s][3] = st[(8 * s + t + 4) * C::ENC_LD + 16 * mt + g + 8];
        }
    float d2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const float2* b1f = reinterpret_cast<const float2*>(fr);
    const float2* b2f = reinterpret_cast<const float2*>(fr + C::B1_FLOATS);
    const float* bias = fr + C::B1_FLOATS + C::B2_FLOATS;
#pragma unroll
    for (int nt = 0; nt < C::NT; nt++) {
        const float2 bi = *reinterpret_cast<const float2*>(bias + 8 * nt + 2 * t);
        float h[2][4] = {{bi.x, bi.y, bi.x, bi.y}, {bi.x, bi.y, bi.x, bi.y}};
#pragma unroll
        for (int s = 0; s < C::KS; s++) {
            const float2 b = b1f[(nt * C::KS + s) * 32 + lane];
            mma_tf32(h[0], a[0][s], __float_as_uint(b.x), __float_as_uint(b.y));
            mma_tf32(h[1], a[1][s], __float_as_uint(b.x), __float_as_uint(b.y));
        }
        const float2 b2 = b2f[nt * 32 + lane];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const uint32_t a2[4] = {to_tf32(fmaxf(h[mt][0], 0.f)), to_tf32(fmaxf(h[mt][2], 0.f)), to_tf32(fmaxf(h[mt][1], 0.f)),
                                    to_tf32(fmaxf(h[mt][3], 0.f))};
            mma_tf32(d2[mt], a2, __float_as_uint(b2.x), __float_as_uint(b2.y));
        }
    }
    // C fragments of the [32 x 8] result -> each lane's own 4 outputs, through the (now free) staging buffer
    __syncwarp();
    if (t < 2) {
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            stage[(16 * mt + g) * 4 + 2 * t] = d2[mt][0];
            stage[(16 * mt + g) * 4 + 2 * t + 1] = d2[mt][1];
            stage[(16 * mt + g + 8) * 4 + 2 * t] = d2[mt][2];
            stage[(16 * mt + g + 8) * 4 + 2 * t + 1] = d2[mt][3];
        }
    }
    __syncwarp();
    const float4 o = *reinterpret_cast<const float4*>(stage + lane * 4);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
    __syncwarp();
}

}  // namespace mlpmma

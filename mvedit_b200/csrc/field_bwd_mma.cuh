// field_bwd_mma.cuh -- backward of the field's MLP (2L -> 64 -> 4) on tensor cores, per warp, 32 samples at a time.
//
// Why: the FFMA backward (k_field_bwd) is 19 000 straight-line instructions (305 KB of code, 255 registers, 8 warps per SM) and
// runs at ~0.5 G samples/s -- 9x slower than the forward (profiles/r01_ncu_k_field.txt) and 70 ms of every pipeline step.  As
// per-warp mma.sync.m16n8k8 TF32 GEMMs (the reference's own matmul precision for these layers, lib/apis/adapter3d.py:51-61) the
// whole backward is ~270 MMAs per 32 samples in a ~2 000-instruction kernel.
//
// Two orientations of the same tiny GEMMs are used so that every product finds its operands already in fragment layout
// (a C fragment can be fed back as an A fragment with a fixed permutation of k, never as a B fragment):
//   sample-major (rows = samples):   H = Enc W1^T, Out = relu(H) W2^T            (forward recompute, relu mask bits)
//                                    dH = dOut W2 (masked), dEnc = dH W1          (C -> A reuse of dH)
//   hidden-major (rows = hidden):    H^T = W1 Enc^T, dH^T = W2^T dOut^T (masked)  (B fragments of Enc^T are pass A's A fragments)
//                                    dW1 += dH^T Enc, dW2^T += relu(H^T) dOut     (C -> A reuse of dH^T / relu(H^T), k = samples)
// db1 falls out of the dW1 product through a constant-one feature row; db2 is a 32-term column sum.
#pragma once
#include "mlp_mma.cuh"

namespace mlpmma {

template <int L>
struct BCfg {
    using C = Cfg<L>;
    static constexpr int IN = C::IN, KS = C::KS, NT = C::NT, LD = C::ENC_LD;
    static constexpr int FT = KS;                                    // feature n-tiles of dEnc (IN padded to 8)
    static constexpr int WT = (IN + 1 + 7) / 8;                      // feature n-tiles of dW1 incl. the ones column (db1)
    static constexpr int ROWS = (WT > KS ? WT : KS) * 8;             // rows of the [feature][sample] staging tile
    static constexpr int MT = HID / 16;                              // hidden m-tiles (hidden-major orientation)
    // fragment-ordered TF32 weights in shared memory (float offsets)
    static constexpr int O_B1 = 0;                                   // [NT][KS][32][2]  layer 1, sample-major
    static constexpr int O_B2 = O_B1 + NT * KS * 64;                 // [NT][32][2]      layer 2, sample-major
    static constexpr int O_BIAS = O_B2 + NT * 64;                    // [HID] fp32
    static constexpr int O_B4 = O_BIAS + HID;                        // [NT][32]         dH = dOut W2
    static constexpr int O_B3 = O_B4 + NT * 32;                      // [NT][FT][32][2]  dEnc = dH W1
    static constexpr int O_A1 = O_B3 + NT * FT * 64;                 // [MT][KS][32][4]  H^T = W1 Enc^T
    static constexpr int O_A4 = O_A1 + MT * KS * 128;                // [MT][32][2]      dH^T = W2^T dOut^T
    static constexpr int FRAG_FLOATS = O_A4 + MT * 64;
    static constexpr int STAGE_FLOATS = ROWS * LD;                   // per warp: [feature][sample] tile
    static constexpr int XCH_FLOATS = 128;                           // per warp: outputs / dOut [32][4]
    static_assert(O_A1 % 4 == 0, "float4 fragment loads need 16-byte alignment");
};

template <int L>
__device__ __forceinline__ void stage_bwd_frags(float* __restrict__ fr, const float* __restrict__ w1, const float* __restrict__ b1,
                                                const float* __restrict__ w2) {
    using B = BCfg<L>;
    constexpr int IN = B::IN, KS = B::KS, NT = B::NT, FT = B::FT, MT = B::MT;
    uint32_t* fb = reinterpret_cast<uint32_t*>(fr);
    auto W1 = [&](int h, int f) { return to_tf32(f < IN ? w1[h * IN + f] : 0.f); };
    auto W2 = [&](int o, int h) { return to_tf32(o < 4 ? w2[o * HID + h] : 0.f); };
    for (int i = threadIdx.x; i < NT * KS * 32; i += blockDim.x) {
        const int lane = i & 31, s = (i >> 5) % KS, nt = (i >> 5) / KS, g = lane >> 2, t = lane & 3;
        fb[B::O_B1 + 2 * i] = W1(8 * nt + g, 8 * s + t);
        fb[B::O_B1 + 2 * i + 1] = W1(8 * nt + g, 8 * s + t + 4);
    }
    for (int i = threadIdx.x; i < NT * 32; i += blockDim.x) {
        const int lane = i & 31, nt = i >> 5, g = lane >> 2, t = lane & 3;
        fb[B::O_B2 + 2 * i] = W2(g, 8 * nt + 2 * t);
        fb[B::O_B2 + 2 * i + 1] = W2(g, 8 * nt + 2 * t + 1);
        fb[B::O_B4 + i] = W2(t, 8 * nt + g);
    }
    for (int i = threadIdx.x; i < HID; i += blockDim.x) fr[B::O_BIAS + i] = b1[i];
    for (int i = threadIdx.x; i < NT * FT * 32; i += blockDim.x) {
        const int lane = i & 31, ft = (i >> 5) % FT, nt = (i >> 5) / FT, g = lane >> 2, t = lane & 3;
        fb[B::O_B3 + 2 * i] = W1(8 * nt + 2 * t, 8 * ft + g);
        fb[B::O_B3 + 2 * i + 1] = W1(8 * nt + 2 * t + 1, 8 * ft + g);
    }
    for (int i = threadIdx.x; i < MT * KS * 32; i += blockDim.x) {
        const int lane = i & 31, ks = (i >> 5) % KS, mt = (i >> 5) / KS, g = lane >> 2, t = lane & 3;
        fb[B::O_A1 + 4 * i] = W1(16 * mt + g, 8 * ks + t);
        fb[B::O_A1 + 4 * i + 1] = W1(16 * mt + g + 8, 8 * ks + t);
        fb[B::O_A1 + 4 * i + 2] = W1(16 * mt + g, 8 * ks + t + 4);
        fb[B::O_A1 + 4 * i + 3] = W1(16 * mt + g + 8, 8 * ks + t + 4);
    }
    for (int i = threadIdx.x; i < MT * 32; i += blockDim.x) {
        const int lane = i & 31, mt = i >> 5, g = lane >> 2, t = lane & 3;
        fb[B::O_A4 + 2 * i] = W2(t, 16 * mt + g);
        fb[B::O_A4 + 2 * i + 1] = W2(t, 16 * mt + g + 8);
    }
}

// C fragment -> A fragment of the next product (logical k = t <-> column 2t, k = t+4 <-> column 2t+1)
__device__ __forceinline__ void c_to_a(const float (&c)[4], uint32_t (&a)[4]) {
    a[0] = to_tf32(c[0]); a[1] = to_tf32(c[2]); a[2] = to_tf32(c[1]); a[3] = to_tf32(c[3]);
}

// Per-warp accumulators of the MLP gradients (C fragments, kept in registers across the warp's tiles)
template <int L>
struct MlpGradAcc {
    using B = BCfg<L>;
    float w1[B::MT][B::WT][4];   // (hidden 16mt+g [+8], feature 8ft+2t [+1]); feature == IN is db1
    float w2[B::MT][4];          // (hidden 16mt+g [+8], out 2t [+1])  (t < 2)
    float b2;                    // lanes 0..3: out = lane
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int mt = 0; mt < B::MT; mt++) {
#pragma unroll
            for (int c = 0; c < 4; c++) w2[mt][c] = 0.f;
#pragma unroll
            for (int ft = 0; ft < B::WT; ft++)
#pragma unroll
                for (int c = 0; c < 4; c++) w1[mt][ft][c] = 0.f;
        }
        b2 = 0.f;
    }
};

// The MLP part of the backward for one 32-sample tile.  In: `stage` rows 0..IN-1 hold the TF32 encodings [feature][sample], row
// IN holds ones, other rows zeros; `xch` receives the raw outputs, `dout_fn(o, d)` maps this lane's 4 raw outputs to its 4
// output gradients.  Out: the gradient w.r.t. the encoding overwrites `stage` rows 0..IN-1 (fp32, [feature][sample]); the MLP
// gradients are accumulated into `acc`.  All 32 lanes must call.
template <int L, typename DoutFn>
__device__ __forceinline__ void mlp_backward_tile(float* __restrict__ stage, float* __restrict__ xch, const float* __restrict__ fr,
                                                  MlpGradAcc<L>& acc, DoutFn dout_fn) {
    using B = BCfg<L>;
    constexpr int KS = B::KS, NT = B::NT, FT = B::FT, WT = B::WT, MT = B::MT, LD = B::LD, IN = B::IN;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const uint32_t* st = reinterpret_cast<const uint32_t*>(stage);
    __syncwarp();
    // ------------------------------------------------ pass A: sample-major forward, relu mask bits
    uint32_t a[2][KS][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int s = 0; s < KS; s++) {
            a[mt][s][0] = st[(8 * s + t) * LD + 16 * mt + g];
            a[mt][s][1] = st[(8 * s + t) * LD + 16 * mt + g + 8];
            a[mt][s][2] = st[(8 * s + t + 4) * LD + 16 * mt + g];
            a[mt][s][3] = st[(8 * s + t + 4) * LD + 16 * mt + g + 8];
        }
    uint32_t mask[2] = {0u, 0u};
    {
        float d2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        const float2* b1f = reinterpret_cast<const float2*>(fr + B::O_B1);
        const float2* b2f = reinterpret_cast<const float2*>(fr + B::O_B2);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const float2 bi = *reinterpret_cast<const float2*>(fr + B::O_BIAS + 8 * nt + 2 * t);
            float h[2][4] = {{bi.x, bi.y, bi.x, bi.y}, {bi.x, bi.y, bi.x, bi.y}};
#pragma unroll
            for (int s = 0; s < KS; s++) {
                const float2 b = b1f[(nt * KS + s) * 32 + lane];
                mma_tf32(h[0], a[0][s], __float_as_uint(b.x), __float_as_uint(b.y));
                mma_tf32(h[1], a[1][s], __float_as_uint(b.x), __float_as_uint(b.y));
            }
            const float2 b2 = b2f[nt * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (h[mt][c] > 0.f) mask[mt] |= 1u << (nt * 4 + c);
                    h[mt][c] = fmaxf(h[mt][c], 0.f);
                }
                uint32_t a2[4];
                c_to_a(h[mt], a2);
                mma_tf32(d2[mt], a2, __float_as_uint(b2.x), __float_as_uint(b2.y));
            }
        }
        if (t < 2) {
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                *reinterpret_cast<float2*>(xch + (16 * mt + g) * 4 + 2 * t) = make_float2(d2[mt][0], d2[mt][1]);
                *reinterpret_cast<float2*>(xch + (16 * mt + g + 8) * 4 + 2 * t) = make_float2(d2[mt][2], d2[mt][3]);
            }
        }
    }
    __syncwarp();
    {
        const float4 o4 = *reinterpret_cast<const float4*>(xch + lane * 4);
        const float o[4] = {o4.x, o4.y, o4.z, o4.w};
        float d[4];
        dout_fn(o, d);
        __syncwarp();
        *reinterpret_cast<float4*>(xch + lane * 4) = make_float4(d[0], d[1], d[2], d[3]);
    }
    __syncwarp();
    if (lane < 4) {
        float sacc = 0.f;
#pragma unroll 8
        for (int s = 0; s < 32; s++) sacc += xch[s * 4 + lane];
        acc.b2 += sacc;
    }
    // ------------------------------------------------ pass C: hidden-major, weight gradients (k = samples)
    {
        uint32_t bdo[4], bd2[4][2];
#pragma unroll
        for (int ns = 0; ns < 4; ns++) {
            bdo[ns] = to_tf32(xch[(8 * ns + g) * 4 + t]);
            bd2[ns][0] = g < 4 ? to_tf32(xch[(8 * ns + 2 * t) * 4 + g]) : 0u;
            bd2[ns][1] = g < 4 ? to_tf32(xch[(8 * ns + 2 * t + 1) * 4 + g]) : 0u;
        }
        const float4* a1f = reinterpret_cast<const float4*>(fr + B::O_A1);
        const float2* a4f = reinterpret_cast<const float2*>(fr + B::O_A4);
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const float bg = fr[B::O_BIAS + 16 * mt + g], bg8 = fr[B::O_BIAS + 16 * mt + g + 8];
            float hT[4][4], dT[4][4];
#pragma unroll
            for (int ns = 0; ns < 4; ns++) {
                hT[ns][0] = bg; hT[ns][1] = bg; hT[ns][2] = bg8; hT[ns][3] = bg8;
                dT[ns][0] = 0.f; dT[ns][1] = 0.f; dT[ns][2] = 0.f; dT[ns][3] = 0.f;
            }
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                const float4 w = a1f[(mt * KS + ks) * 32 + lane];
                const uint32_t aw[4] = {__float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z), __float_as_uint(w.w)};
#pragma unroll
                for (int ns = 0; ns < 4; ns++) mma_tf32(hT[ns], aw, a[ns >> 1][ks][ns & 1], a[ns >> 1][ks][2 + (ns & 1)]);
            }
            const float2 w4 = a4f[mt * 32 + lane];
            const uint32_t a4[4] = {__float_as_uint(w4.x), __float_as_uint(w4.y), 0u, 0u};
#pragma unroll
            for (int ns = 0; ns < 4; ns++) {
                mma_tf32(dT[ns], a4, bdo[ns], 0u);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (!(hT[ns][c] > 0.f)) dT[ns][c] = 0.f;
                    hT[ns][c] = fmaxf(hT[ns][c], 0.f);
                }
                uint32_t ad[4], ah[4];
                c_to_a(dT[ns], ad);
                c_to_a(hT[ns], ah);
#pragma unroll
                for (int ft = 0; ft < WT; ft++) {
                    const uint2 be = *reinterpret_cast<const uint2*>(st + (8 * ft + g) * LD + 8 * ns + 2 * t);
                    mma_tf32(acc.w1[mt][ft], ad, be.x, be.y);
                }
                mma_tf32(acc.w2[mt], ah, bd2[ns][0], bd2[ns][1]);
            }
        }
    }
    // ------------------------------------------------ pass B: sample-major dH -> dEnc
    float de[2][FT][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int ft = 0; ft < FT; ft++) { de[mt][ft][0] = 0.f; de[mt][ft][1] = 0.f; de[mt][ft][2] = 0.f; de[mt][ft][3] = 0.f; }
    {
        uint32_t ado[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            ado[mt][0] = to_tf32(xch[(16 * mt + g) * 4 + t]);
            ado[mt][1] = to_tf32(xch[(16 * mt + g + 8) * 4 + t]);
            ado[mt][2] = 0u; ado[mt][3] = 0u;
        }
        const uint32_t* b4f = reinterpret_cast<const uint32_t*>(fr + B::O_B4);
        const uint2* b3f = reinterpret_cast<const uint2*>(fr + B::O_B3);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const uint32_t b4 = b4f[nt * 32 + lane];
            uint2 b3[FT];
#pragma unroll
            for (int ft = 0; ft < FT; ft++) b3[ft] = b3f[(nt * FT + ft) * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                float dh[4] = {0.f, 0.f, 0.f, 0.f};
                mma_tf32(dh, ado[mt], b4, 0u);
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (!((mask[mt] >> (nt * 4 + c)) & 1u)) dh[c] = 0.f;
                uint32_t ad[4];
                c_to_a(dh, ad);
#pragma unroll
                for (int ft = 0; ft < FT; ft++) mma_tf32(de[mt][ft], ad, b3[ft].x, b3[ft].y);
            }
        }
    }
    // dEnc C fragments -> [feature][sample] over the (no longer needed) encodings; the ones / padding rows are left alone
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int ft = 0; ft < FT; ft++) {
            const int f = 8 * ft + 2 * t;
            if (f < IN) {
                stage[f * LD + 16 * mt + g] = de[mt][ft][0];
                stage[(f + 1) * LD + 16 * mt + g] = de[mt][ft][1];
                stage[f * LD + 16 * mt + g + 8] = de[mt][ft][2];
                stage[(f + 1) * LD + 16 * mt + g + 8] = de[mt][ft][3];
            }
        }
    __syncwarp();
}

// ones row (db1) and zero padding rows of a warp's staging tile; written once, never overwritten
template <int L>
__device__ __forceinline__ void init_bwd_stage(float* __restrict__ stage) {
    using B = BCfg<L>;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = B::IN; k < B::ROWS; k++) stage[k * B::LD + lane] = (k == B::IN) ? 1.0f : 0.f;
    __syncwarp();
}

}  // namespace mlpmma

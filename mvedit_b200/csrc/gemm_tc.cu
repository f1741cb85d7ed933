// gemm_tc.cu -- bf16 GEMM and implicit-GEMM 3x3 convolution on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel:
//   warp 0     TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1     MMA issuer     (one elected lane: tcgen05.mma kind::f16, M=128, N=BN, K=16; fp32 accumulators in TMEM,
//                              two accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1)
//   warps 2-9  epilogue       (tcgen05.ld 32x32b -> bias / per-image bias / activation / scale / residual -> bf16 -> global;
//                              two warps per TMEM lane quadrant take alternate 32-column chunks)
//
//   C[M,N] = epi( A[M,K] . B[N,K]^T )       A, B bf16 K-major (row-major with K contiguous), C bf16 row-major.
//
// MT = 2 ("tall" tiles): one CTA owns a 256 x BN output tile as TWO 128-row MMAs per K step that share the B stage in shared memory
// (A stage = 32 KB, accumulators at TMEM columns [0,BN) and [BN,2BN)).  A 128 x BN tile needs 2*128*BN / (2*(128+BN)) FLOP per byte of
// operands staged through shared memory -- 64 (BN=128) ... 71 (BN=160) -- and the ~100 B/clk an SM can pull from L2 then caps the
// tensor pipe at 60-70 % (ncu: profiles/r02_ncu_kernels_summary.txt); 256 rows per B tile raise that to 85-98 FLOP/B.
// ACC = number of accumulator stages in TMEM: 2 overlaps the epilogue of tile i with the main loop of tile i+1 (needs 2*MT*BN <= 512
// columns); 1 is used for 256 x 160 tiles (320 columns), whose long K loops (3x3 convolutions) hide the un-overlapped epilogue.
//
// MODE 1 (conv3x3, stride 1, pad 1, NHWC): A is never materialised.  The producer walks K as 9 taps x (Cin/64) chunks and
// fetches, for tap (dy,dx), the box {64 ch, BW, BH, BB} of the NHWC activation at (w0+dx-1, h0+dy-1, b0) with a 4-D tensor map;
// TMA's out-of-bounds zero fill supplies the padding.  The 128 rows of the box are 128 consecutive output pixels.
// Weights are [Cout][ky][kx][Cin] so that B's K index = tap*Cin + c.
//
// Replaces, on the reference's path, the cuDNN/cuBLAS calls under diffusers' UNet2DConditionModel / ControlNetModel
// (SURVEY.md §8 a-1/a-2, Appendix A).
#include <cstdlib>
#include "tc_common.cuh"
#include <mutex>
#include <string.h>
#include "../../include/mvedit_b200.h"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two warps per TMEM lane quadrant, interleaved column chunks)
constexpr int A_BYTES = BM * BK * 2;

struct GemmParams {
    __nv_bfloat16* C;
    uint32_t M, N, ldc;
    uint32_t num_kb;        // K / 64 (conv: 9 * Cin/64)
    uint32_t m_tiles, n_tiles;
    // epilogue
    const float* bias;              // [N] or null
    const float* row_bias;          // [M / rows_per_group, N] or null
    uint32_t rows_per_group, ldrb;
    const __nv_bfloat16* residual;  // [M, ldr] or null
    uint32_t ldr;
    int act;                        // 0 none, 1 SiLU, 2 GELU(erf), 3 GEGLU (tile = [BN/2 value | BN/2 gate] columns -> BN/2 outputs),
                                    // 4 ReLU, 5 ReLU-backward gate: out = residual > 0 ? (acc + bias) * alpha : 0 (residual = the forward
                                    // activation; nothing is added)
                                    // 6 PReLU with per-column slopes act_param[N]
    float alpha;                    // out = act(acc + bias + row_bias) * alpha + residual
    const float* act_param;         // [N] for act 6, else unused
    // conv geometry (MODE 1)
    uint32_t H, W, cin_chunks;
    // split-K (few-tile problems with a long K: the 8^2 / 16^2 levels of the LPIPS VGG): `splits` CTAs share one output tile, each
    // reduces a slice of K and adds its fp32 partial into ws [M, N] (zero on entry); k_splitk_finish applies the epilogue and re-zeroes
    uint32_t splits;
    float* ws;
};

constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int BN, int MT = 1, int ACC = 2>
struct Cfg {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
    static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : (MT == 1 ? ((BN >= 256) ? 4 : ((BN >= 128) ? 6 : 8)) : STAGES_RAW);
    // column offset between accumulator stages; the MT sub-tiles of one stage are BN columns apart
    static constexpr int ACC_STRIDE = (MT == 1) ? ((BN > 128) ? 256 : BN) : MT * BN;
    static constexpr int TMEM_COLS = pow2_cols(ACC == 1 ? MT * BN : ACC_STRIDE + MT * BN);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(TMEM_COLS <= 512 && (ACC == 1 ? MT * BN : ACC_STRIDE + MT * BN) <= 512, "accumulators do not fit TMEM");
};

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == 1) return v / (1.0f + __expf(-v));
    if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    if (act == 4) return fmaxf(v, 0.0f);
    return v;
}

// RARE = true: the instantiation that also carries the split-K reduction and the PReLU epilogue (64-wide conv tiles only: the LPIPS VGG
// and the SRVGG enhancer).  The hot instantiations (RARE = false) do not contain that code: with it inlined the generic epilogue -- the
// bottleneck of the short-K GEMMs and of the 256-row tiles -- ran 1.4-2.4x slower (profiles/r02_bench_*: 74 -> 143 ms of GEMM per step).
template <int BN, int MODE, int MT = 1, int ACC = 2, bool RARE = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                            const GemmParams p) {
    using C_ = Cfg<BN, MT, ACC>;
    constexpr int BMT = BM * MT;            // rows of one CTA tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_::STAGES * C_::STAGE_BYTES);
    uint64_t* full = bars;                        // [STAGES]
    uint64_t* empty = bars + C_::STAGES;          // [STAGES]
    uint64_t* tfull = bars + 2 * C_::STAGES;      // [2]
    uint64_t* tempty = bars + 2 * C_::STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C_::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        for (int s = 0; s < C_::STAGES; s++) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; s++) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], 8); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, C_::TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t num_tiles = p.m_tiles * p.n_tiles * (RARE ? p.splits : 1u);
    // tile -> (m tile, n tile, first and one-past-last k block of this CTA's K slice)
    const uint32_t d_splits = RARE ? p.splits : 1u, d_n_tiles = p.n_tiles, d_num_kb = p.num_kb;
    auto decode = [d_splits, d_n_tiles, d_num_kb](const uint32_t tile, uint32_t& mt, uint32_t& nt, uint32_t& kb0, uint32_t& kb1) {
        if (!RARE || d_splits == 1) {
            mt = tile / d_n_tiles; nt = tile % d_n_tiles; kb0 = 0; kb1 = d_num_kb;
            return;
        }
        const uint32_t sp = tile % d_splits, t2 = tile / d_splits;
        mt = t2 / d_n_tiles; nt = t2 % d_n_tiles;
        kb0 = sp * d_num_kb / d_splits; kb1 = (sp + 1) * d_num_kb / d_splits;
    };

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                uint32_t mt, nt, kb0, kb1;
                decode(tile, mt, nt, kb0, kb1);
                const int m0 = mt * BMT, n0 = nt * BN;
                int b0 = 0, h0 = 0, w0 = 0;
                if (MODE == 1) {
                    const uint32_t hw = p.H * p.W;
                    b0 = m0 / hw;
                    const uint32_t rem = m0 % hw;
                    h0 = rem / p.W;
                    w0 = rem % p.W;
                }
                for (uint32_t kb = kb0; kb < kb1; kb++) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * C_::STAGE_BYTES;
                    uint8_t* sb = sa + MT * A_BYTES;
                    tc::mbar_arrive_expect_tx(&full[stage], C_::STAGE_BYTES);
                    if (MODE == 0) {
                        tc::tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
                    } else {
                        const uint32_t tap = kb / p.cin_chunks, cc = kb % p.cin_chunks;
                        const int dy = (int)(tap / 3) - 1, dx = (int)(tap % 3) - 1;
                        tc::tma_load_4d(sa, &tmA, &full[stage], cc * BK, w0 + dx, h0 + dy, b0);
                    }
                    tc::tma_load_2d(sb, &tmB, &full[stage], kb * BK, n0);
                    if (++stage == C_::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                uint32_t mt_, nt_, kb0, kb1;
                decode(tile, mt_, nt_, kb0, kb1);
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C_::ACC_STRIDE;
                for (uint32_t kb = kb0; kb < kb1; kb++) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t sa = tc::smem_u32(smem + stage * C_::STAGE_BYTES);
                    const uint64_t da = tc::make_desc_k_sw128(sa), db = tc::make_desc_k_sw128(sa + MT * A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++) {
                        // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr >> 4) field;
                        // the second 128-row sub-tile of A sits A_BYTES further on and accumulates BN columns further on
#pragma unroll
                        for (int s = 0; s < MT; s++)
                            tc::umma_f16(d_tmem + s * BN, da + (uint64_t)(k * 2 + s * (A_BYTES >> 4)), db + (uint64_t)(k * 2), idesc, ((kb - kb0) | k) ? 1u : 0u);
                    }
                    tc::umma_commit(&empty[stage]);
                    if (++stage == C_::STAGES) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(&tfull[acc]);
                if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------ epilogue warps (TMEM lane quadrant = warp % 4)
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;     // MT == 1: 0 = even column chunks, 1 = odd column chunks; MT == 2: the 128-row sub-tile
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            uint32_t mt, nt, kb0_, kb1_;
            decode(tile, mt, nt, kb0_, kb1_);
            const uint32_t row = mt * BMT + (MT == 2 ? half * BM : 0) + q * 32 + lane;
            const uint32_t n0 = nt * BN;
            tc::mbar_wait(&tfull[acc], acc_phase);
            tc::tc_fence_after();
            const uint32_t t_row = tmem_base + acc * C_::ACC_STRIDE + (MT == 2 ? half * BN : 0) + ((uint32_t)(q * 32) << 16);
            const bool row_ok = row < p.M;
            const float* rb = (p.row_bias && row_ok) ? p.row_bias + (size_t)(row / p.rows_per_group) * p.ldrb : nullptr;
            constexpr int CH = (BN >= 32) ? 32 : 16;
            if (MT == 1 && BN >= 64 && p.act == 3) {
                // GEGLU fused into the feed-forward's first projection (diffusers GEGLU: hidden, gate = proj(x).chunk(2); hidden *
                // gelu(gate)): the weight rows are pre-ordered so that this tile's first BN/2 columns are values and the last BN/2
                // the matching gates; only the BN/2 products are written (the [M, 2F] intermediate never exists in HBM).
                constexpr int HB = BN / 2;
#pragma unroll 1
                for (int c = half * 32; c < HB; c += 64) {
                    uint32_t v[32], g[32];
                    tc::tmem_ld32(t_row + c, v);
                    tc::tmem_ld32(t_row + HB + c, g);
                    tc::tmem_ld_wait();
                    if (row_ok) {
                        __nv_bfloat16* out = p.C + (size_t)row * p.ldc + nt * HB + c;
#pragma unroll
                        for (int gq = 0; gq < 4; gq++) {
                            float f[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                float a = __uint_as_float(v[gq * 8 + i]), b = __uint_as_float(g[gq * 8 + i]);
                                if (p.bias) { a += p.bias[n0 + c + gq * 8 + i]; b += p.bias[n0 + HB + c + gq * 8 + i]; }
                                f[i] = a * (0.5f * b * (1.0f + erff(b * 0.70710678118654752f))) * p.alpha;
                            }
                            uint4 o;
                            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                            for (int i = 0; i < 4; i++) o2[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                            *reinterpret_cast<uint4*>(out + gq * 8) = o;
                        }
                    }
                }
            } else
#pragma unroll 1
            for (int c = (MT == 2 ? 0 : half * CH); c < BN; c += (MT == 2 ? CH : 2 * CH)) {
                uint32_t v[32];
                if (CH == 32) {
                    tc::tmem_ld32(t_row + c, v);
                } else {
                    uint32_t v16[16];
                    tc::tmem_ld16(t_row + c, v16);
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = v16[i];
                }
                tc::tmem_ld_wait();
                const uint32_t col0 = n0 + c;
                if (RARE && p.splits > 1) {
                    if (row_ok) {
                        float* wrow = p.ws + (size_t)row * p.N + col0;
#pragma unroll
                        for (int i = 0; i < CH; i++)
                            if (col0 + i < p.N) atomicAdd(wrow + i, __uint_as_float(v[i]));
                    }
                } else if (row_ok && col0 < p.N) {
                    __nv_bfloat16* out = p.C + (size_t)row * p.ldc + col0;
                    const __nv_bfloat16* res = p.residual ? p.residual + (size_t)row * p.ldr + col0 : nullptr;
                    const bool full_chunk = (col0 + CH <= p.N) && ((p.ldc & 7) == 0) && (!res || (p.ldr & 7) == 0);
                    if (full_chunk) {
#pragma unroll
                        for (int g = 0; g < CH / 8; g++) {
                            float f[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                float x = __uint_as_float(v[g * 8 + i]);
                                if (p.bias) x += p.bias[col0 + g * 8 + i];
                                if (rb) x += rb[col0 + g * 8 + i];
                                if (RARE && p.act == 6) x = x > 0.f ? x : x * p.act_param[col0 + g * 8 + i];
                                f[i] = act_apply(x, p.act) * p.alpha;
                            }
                            if (res) {
                                const uint4 r4 = *reinterpret_cast<const uint4*>(res + g * 8);
                                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    const float2 t = __bfloat1622float2(r2[i]);
                                    if (p.act == 5) { f[2 * i] = t.x > 0.f ? f[2 * i] : 0.f; f[2 * i + 1] = t.y > 0.f ? f[2 * i + 1] : 0.f; }
                                    else { f[2 * i] += t.x; f[2 * i + 1] += t.y; }
                                }
                            }
                            uint4 o;
                            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                            for (int i = 0; i < 4; i++) o2[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                            *reinterpret_cast<uint4*>(out + g * 8) = o;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < CH; i++) {
                            if (col0 + i >= p.N) break;
                            float x = __uint_as_float(v[i]);
                            if (p.bias) x += p.bias[col0 + i];
                            if (rb) x += rb[col0 + i];
                            if (RARE && p.act == 6) x = x > 0.f ? x : x * p.act_param[col0 + i];
                            x = act_apply(x, p.act) * p.alpha;
                            if (res) x = p.act == 5 ? (__bfloat162float(res[i]) > 0.f ? x : 0.f) : x + __bfloat162float(res[i]);
                            out[i] = __float2bfloat16(x);
                        }
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, C_::TMEM_COLS);
    }
}

// epilogue of a split-K launch: ws [M,N] fp32 sums -> act(sum + bias) * alpha (+ residual / gated by it) -> bf16, and ws back to zero
__global__ void __launch_bounds__(256) k_splitk_finish(const GemmParams p) {
    const uint32_t n8 = p.N / 8;
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= (size_t)p.M * n8) return;
    const uint32_t row = t / n8, col0 = (t % n8) * 8;
    float4* w4 = reinterpret_cast<float4*>(p.ws + (size_t)row * p.N + col0);
    const float4 a = w4[0], b = w4[1];
    w4[0] = make_float4(0.f, 0.f, 0.f, 0.f); w4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const float* rb = p.row_bias ? p.row_bias + (size_t)(row / p.rows_per_group) * p.ldrb : nullptr;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float x = f[i];
        if (p.bias) x += p.bias[col0 + i];
        if (rb) x += rb[col0 + i];
        if (p.act == 6) x = x > 0.f ? x : x * p.act_param[col0 + i];
        x = act_apply(x, p.act) * p.alpha;
        if (p.residual) {
            const float r = __bfloat162float(p.residual[(size_t)row * p.ldr + col0 + i]);
            x = p.act == 5 ? (r > 0.f ? x : 0.f) : x + r;
        }
        p.C[(size_t)row * p.ldc + col0 + i] = __float2bfloat16(x);
    }
}

// per-device fp32 workspace of the split-K path (zero between launches by construction).  Never allocated under stream capture.
constexpr size_t SPLITK_WS_FLOATS = 2u << 20;      // 8 MB: M * N <= 2 Mi elements
float* splitk_workspace(cudaStream_t s) {
    static float* ws[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    float*& w = ws[dev & 15];
    if (!w) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
        if (cudaMalloc(&w, SPLITK_WS_FLOATS * sizeof(float)) != cudaSuccess) { w = nullptr; return nullptr; }
        if (cudaMemset(w, 0, SPLITK_WS_FLOATS * sizeof(float)) != cudaSuccess) return nullptr;
    }
    return w;
}

int pick_bn(uint32_t N) {
    if (N % 256 == 0) return 256;
    if (N % 160 == 0) return 160;
    if (N % 128 == 0) return 128;
    if (N >= 192) return 128;
    if (N > 32) return 64;
    if (N > 16) return 32;
    return 16;
}

// Tile shape for a problem: (BN, MT, ACC).  Tall (256-row) tiles when M is large and the shape has a tall configuration:
//   N % 128 == 0 but not % 256 (128, 384, 640 ...) -> 256 x 128, two accumulator stages;
//   N % 160 == 0 (320, 640 with long K)            -> 256 x 160, one accumulator stage (K >= 1024: the main loop hides the epilogue).
// (A 256 x 256 single-stage tile was measured too: 3 smem stages and no epilogue overlap cost more than the extra reuse gains --
//  1012 vs 1350 TF/s on the VAE's 256-channel convolutions -- so N % 256 == 0 stays on 128 x 256 tiles with two accumulator stages.)
// MVE_GEMM_TALL=0 in the environment disables them (A/B timing).
struct TileCfg { int bn, mt, acc; };
TileCfg pick_tile(uint32_t M, uint32_t N, uint32_t K, int act) {
    static int tall = -1;
    if (tall < 0) { const char* e = getenv("MVE_GEMM_TALL"); tall = (e && e[0] == '0') ? 0 : 1; }
    int bn = pick_bn(N);
    if (act != 3) {
        // few-tile problems (the 8^2 .. 32^2 levels of the LPIPS VGG, a patch at a time): narrower tiles put more SMs to work
        const uint32_t mt_ = (M + BM - 1) / BM;
        if (bn == 256 && mt_ * ((N + 255) / 256) * 4 <= (uint32_t)kNumSM) bn = 128;
        if (bn == 128 && N % 64 == 0 && mt_ * ((N + 127) / 128) * 4 <= (uint32_t)kNumSM) bn = 64;
    }
    if (!tall || act == 3 || M < 148u * 256u) return {bn, 1, 2};
    if (bn == 128 && N % 128 == 0) return {128, 2, 2};
    if (bn == 160 && K >= 1024) return {160, 2, 1};
    return {bn, 1, 2};
}

template <int BN, int MODE, int MT = 1, int ACC = 2, bool RARE = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
    using C_ = Cfg<BN, MT, ACC>;
    static bool configured[16] = {};
    int dev = 0;
    MVE_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 15]) {
        MVE_CUDA(cudaFuncSetAttribute(k_gemm_tc<BN, MODE, MT, ACC, RARE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
        configured[dev & 15] = true;
    }
    const uint32_t tiles = p.m_tiles * p.n_tiles * p.splits;
    const uint32_t grid = tiles < (uint32_t)kNumSM ? tiles : (uint32_t)kNumSM;
    k_gemm_tc<BN, MODE, MT, ACC, RARE><<<grid, NUM_THREADS, C_::SMEM_BYTES, stream>>>(tmA, tmB, p);
    MVE_CHECK_LAUNCH("k_gemm_tc");
    if (p.splits > 1) {
        k_splitk_finish<<<cdiv((size_t)p.M * (p.N / 8), 256), 256, 0, stream>>>(p);
        MVE_CHECK_LAUNCH("k_splitk_finish");
    }
    return 0;
}

template <int MODE>
int dispatch(TileCfg t, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t s) {
    if (p.splits > 1 || p.act == 6) {
        if (MODE != 1 || t.bn != 64 || t.mt != 1) { mve_set_error("split-K / PReLU epilogues exist for 64-wide convolution tiles only"); return -1; }
        return launch<64, 1, 1, 2, true>(tmA, tmB, p, s);
    }
    if (t.mt == 2)
        return t.bn == 128 ? launch<128, MODE, 2, 2>(tmA, tmB, p, s)
                           : (t.bn == 160 ? launch<160, MODE, 2, 1>(tmA, tmB, p, s) : launch<256, MODE, 2, 1>(tmA, tmB, p, s));
    switch (t.bn) {
        case 256: return launch<256, MODE>(tmA, tmB, p, s);
        case 160: return launch<160, MODE>(tmA, tmB, p, s);
        case 128: return launch<128, MODE>(tmA, tmB, p, s);
        case 64: return launch<64, MODE>(tmA, tmB, p, s);
        case 32: return launch<32, MODE>(tmA, tmB, p, s);
        default: return launch<16, MODE>(tmA, tmB, p, s);
    }
}

}  // namespace

// ---------------------------------------------------------------- host helpers (declared in tc_common.cuh)
PFN_tmapEncodeTiled mve_get_tmap_encode() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (PFN_tmapEncodeTiled)f;
    }
    return fn;
}

int mve_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const char* what) {
    // Descriptors are pure functions of (base, dims, strides, box): a step issues ~650 GEMM / conv / attention launches over a few dozen
    // distinct operand layouts at recurring addresses (weights; activations at the caching allocator's recycled pointers), so a small
    // direct-mapped cache removes the driver call from the launch path.
    struct Entry { uint64_t key[14]; CUtensorMap map; bool valid; };
    static Entry cache[512];
    static std::mutex mu;
    uint64_t key[14] = {(uint64_t)(uintptr_t)base, (uint64_t)rank};
    for (int i = 0; i < 4; i++) {
        key[2 + i] = i < rank ? dims[i] : 0;
        key[6 + i] = i + 1 < rank ? strides_bytes[i] : 0;
        key[10 + i] = i < rank ? box[i] : 0;
    }
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 14; i++) { h ^= key[i]; h *= 1099511628211ull; }
    Entry& e = cache[(h >> 20) & 511];
    static int use_cache = -1;
    if (use_cache < 0) { const char* ev = getenv("MVE_TMAP_CACHE"); use_cache = (ev && ev[0] == '0') ? 0 : 1; }
    if (use_cache) {
        std::lock_guard<std::mutex> lock(mu);
        if (e.valid && memcmp(e.key, key, sizeof(key)) == 0) { *out = e.map; return 0; }
    }
    PFN_tmapEncodeTiled enc = mve_get_tmap_encode();
    if (!enc) { mve_set_error("%s: cuTensorMapEncodeTiled entry point not available (no CUDA driver?)", what); return -2; }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        mve_set_error("%s: cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", what, (int)r, rank,
                      (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                      (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return -3;
    }
    {
        std::lock_guard<std::mutex> lock(mu);
        memcpy(e.key, key, sizeof(key)); e.map = *out; e.valid = true;
    }
    return 0;
}

extern "C" {

int mve_gemm_bf16(const void* A, const void* B, void* C, uint32_t M, uint32_t N, uint32_t K, uint32_t lda, uint32_t ldb, uint32_t ldc,
                  const float* bias, const float* row_bias, uint32_t rows_per_group, uint32_t ldrb, const void* residual, uint32_t ldr,
                  int act, float alpha, const float* act_param, void* stream) {
    if (M == 0 || N == 0) return 0;
    MVE_ARG(act != 6, "gemm: the PReLU epilogue (act 6) exists for mve_conv3x3_bf16 only");
    MVE_ARG(K % BK == 0 && K > 0, "gemm: K must be a positive multiple of 64");
    MVE_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements (16-byte TMA strides)");
    MVE_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0, "gemm: pointers must be 16-byte aligned");
    MVE_ARG(!row_bias || rows_per_group > 0, "gemm: rows_per_group must be > 0 with row_bias");
    if (act == 3)
        MVE_ARG(N % 256 == 0 && !row_bias && !residual, "gemm: GEGLU epilogue needs N % 256 == 0 (value/gate interleaved per 256 columns), no row_bias, no residual");
    const TileCfg tc_ = pick_tile(M, N, K, act);
    const int bn = tc_.bn;
    const uint32_t bm = BM * tc_.mt;
    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[2] = {K, M}, str[1] = {(uint64_t)lda * 2};
        const uint32_t box[2] = {BK, bm};
        int r = mve_make_tmap_bf16(&tmA, A, 2, dims, str, box, "gemm A");
        if (r) return r;
    }
    {
        const uint64_t dims[2] = {K, N}, str[1] = {(uint64_t)ldb * 2};
        const uint32_t box[2] = {BK, (uint32_t)bn};
        int r = mve_make_tmap_bf16(&tmB, B, 2, dims, str, box, "gemm B");
        if (r) return r;
    }
    GemmParams p{};
    p.C = (__nv_bfloat16*)C; p.M = M; p.N = N; p.ldc = ldc; p.num_kb = K / BK;
    p.m_tiles = (M + bm - 1) / bm; p.n_tiles = (N + bn - 1) / bn;
    p.bias = bias; p.row_bias = row_bias; p.rows_per_group = rows_per_group; p.ldrb = ldrb ? ldrb : N;
    p.residual = (const __nv_bfloat16*)residual; p.ldr = ldr; p.act = act; p.alpha = alpha; p.act_param = act_param;
    p.splits = 1;
    return dispatch<0>(tc_, tmA, tmB, p, (cudaStream_t)stream);
}

int mve_conv3x3_bf16(const void* X, const void* Wt, void* Y, uint32_t Bn, uint32_t H, uint32_t W, uint32_t Cin, uint32_t Cout,
                     uint32_t ldy, const float* bias, const float* row_bias, uint32_t ldrb, const void* residual, uint32_t ldr, int act,
                     float alpha, const float* act_param, void* stream) {
    if (Bn == 0) return 0;
    const bool allow_split_k = (act & 0x100) != 0;      // opt-in: fp32 partial sums meet in atomic order -> not bit-reproducible
    act &= 0xff;
    MVE_ARG(act != 6 || act_param != nullptr, "conv3x3: act 6 (PReLU) needs act_param [Cout]");
    MVE_ARG(Cin % BK == 0, "conv3x3: Cin must be a multiple of 64 (pad channels)");
    MVE_ARG((W <= 128 && 128 % W == 0) || W % 128 == 0, "conv3x3: W must divide 128 or be a multiple of 128");
    const uint32_t M = Bn * H * W;
    TileCfg tc_ = pick_tile(M, Cout, 9 * Cin, act);
    if (act == 6) tc_ = {64, 1, 2};            // the PReLU epilogue lives in the 64-wide instantiation (SRVGG: Cout = 64 / 48)
    // the tile's pixels as a TMA box {64 ch, BW, BH, BB}: whole image rows / whole images (a tall tile falls back if it cannot)
    uint32_t PIX, BW, BH, BB;
    for (;;) {
        PIX = BM * tc_.mt;
        BW = W < PIX ? W : PIX;
        BH = PIX / BW;
        if (BH > H) BH = H;
        BB = PIX / (BW * BH);
        const bool ok = BW * BH * BB == PIX && (W % BW) == 0 && (H % BH) == 0 && BB <= 256 && BW <= 256 && BH <= 256;
        if (ok || tc_.mt == 1) break;
        tc_ = {pick_bn(Cout), 1, 2};
    }
    MVE_ARG(BW * BH * BB == PIX && (H % BH) == 0 && BB <= 256, "conv3x3: the pixel tile must cover whole rows / images");
    const int bn = tc_.bn;
    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[4] = {Cin, W, H, Bn};
        const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t box[4] = {BK, BW, BH, BB};
        int r = mve_make_tmap_bf16(&tmA, X, 4, dims, str, box, "conv3x3 X");
        if (r) return r;
    }
    {
        const uint64_t dims[2] = {(uint64_t)9 * Cin, Cout}, str[1] = {(uint64_t)9 * Cin * 2};
        const uint32_t box[2] = {BK, (uint32_t)bn};
        int r = mve_make_tmap_bf16(&tmB, Wt, 2, dims, str, box, "conv3x3 W");
        if (r) return r;
    }
    GemmParams p{};
    p.C = (__nv_bfloat16*)Y; p.M = M; p.N = Cout; p.ldc = ldy; p.num_kb = 9 * (Cin / BK);
    p.m_tiles = (M + PIX - 1) / PIX; p.n_tiles = (Cout + bn - 1) / bn;   // a last partial tile reads zero-filled images (TMA OOB)
    p.bias = bias; p.row_bias = row_bias; p.rows_per_group = H * W; p.ldrb = ldrb ? ldrb : Cout;
    p.residual = (const __nv_bfloat16*)residual; p.ldr = ldr; p.act = act; p.alpha = alpha; p.act_param = act_param;
    p.H = H; p.W = W; p.cin_chunks = Cin / BK;
    // split-K: a handful of tiles each looping over >= 16 k blocks is latency-bound (28 us for the 512 -> 512 convolutions of an 8^2 or
    // 16^2 feature map whatever the tile count) -- spread K over idle SMs, >= 8 k blocks per CTA
    p.splits = 1;
    {
        static int enabled = -1;
        if (enabled < 0) { const char* e = getenv("MVE_CONV_SPLITK"); enabled = (e && e[0] == '0') ? 0 : 1; }
        const uint32_t tiles = p.m_tiles * p.n_tiles;
        if (enabled && allow_split_k && tc_.mt == 1 && tc_.bn == 64 && act != 3 && Cout % 8 == 0 && (ldy % 8) == 0 && tiles * 4 <= (uint32_t)kNumSM && p.num_kb >= 16 &&
            (size_t)M * Cout <= SPLITK_WS_FLOATS) {
            uint32_t sp = p.num_kb / 8;
            if (sp > (uint32_t)kNumSM / tiles) sp = (uint32_t)kNumSM / tiles;
            if (sp >= 2 && (p.ws = splitk_workspace((cudaStream_t)stream)) != nullptr) p.splits = sp;
        }
    }
    return dispatch<1>(tc_, tmA, tmB, p, (cudaStream_t)stream);
}

}  // extern "C"

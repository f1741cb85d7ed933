// attention.cu -- flash attention on tcgen05 tensor cores for the SD-1.5 UNet / ControlNet transformer blocks (sm_100a).
//
// Replaces torch.nn.functional.scaled_dot_product_attention under diffusers' AttnProcessor2_0 /
// CrossImageAttnProcWrapper (/root/reference/lib/models/architecture/joint_attn.py:11-37):
// self-attention with S in {4096,1024,256,64} (x2 with the reference image concatenated), head dim 40/80/160, and
// text cross-attention (77 keys).  bf16 in/out, fp32 softmax statistics and accumulators.
//
// One CTA = 128 query rows of one (batch, head).  6 warps:
//   warps 0-3  softmax: thread r owns query row r = TMEM lane r.  S is read from TMEM (tcgen05.ld), the row max / exp2 /
//              row sum are per-thread (no shuffles), P is written as bf16 into a 128B-swizzled K-major smem tile,
//              the O accumulator in TMEM is rescaled (tcgen05.ld/st) only when some row max of the warp moved.
//   warp 4     MMA issuer: S_j = Q K_j^T (A,B K-major), O += P_j V_j (A = P K-major, B = V MN-major straight from the
//              [keys, d] layout -- no transpose pass), S double-buffered in TMEM so QK^T(j+1) overlaps softmax(j).
//   warp 5     TMA producer: Q once, K_j / V_j tiles through a 1-2 stage ring; head-dim tails and sequence tails are
//              zero-filled by TMA (4-D maps: d, head, token, batch), padded keys are masked in the softmax.
#include "tc_common.cuh"
#include "../../include/mvedit_b200.h"

namespace {

constexpr int BQ = 128, BKV = 128;
constexpr int ATOM_BYTES = 128 * 128;  // [128 rows][64 bf16]
constexpr int NUM_THREADS_A = 192;

struct AttnParams {
    __nv_bfloat16* O;
    uint32_t ldo;          // elements
    uint32_t q_len, kv_len, heads, d;
    float scale_log2;      // softmax scale * log2(e)
    uint32_t n_kv_blocks;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// DPAD = head dim rounded up to 16 (48 / 80 / 160); DA = number of 64-wide atoms; ST = K/V ring stages
template <int DPAD, int ST>
__global__ void __launch_bounds__(NUM_THREADS_A, 1) k_attention(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    constexpr int DA = (DPAD + 63) / 64;
    constexpr int Q_BYTES = DA * ATOM_BYTES, KV_BYTES = DA * ATOM_BYTES, P_BYTES = 2 * ATOM_BYTES;
    constexpr uint32_t S_COL0 = 0, S_COL1 = 128, O_COL = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Q_BYTES;                 // [ST][KV_BYTES]
    uint8_t* sV = sK + ST * KV_BYTES;           // [ST][KV_BYTES]
    uint8_t* sP = sV + ST * KV_BYTES;           // [P_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
    uint64_t* bar_q = bars;            // 1
    uint64_t* kv_full = bars + 1;      // [ST]
    uint64_t* kv_empty = bars + 1 + ST;  // [ST]
    uint64_t* s_full = bars + 1 + 2 * ST;  // [2]
    uint64_t* p_full = bars + 3 + 2 * ST;  // 1 (128 arrivals)
    uint64_t* pv_done = bars + 4 + 2 * ST; // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * ST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t qb = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const uint32_t nblk = p.n_kv_blocks;

    if (warp == 5 && lane == 0) {
        tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV);
        tc::mbar_init(bar_q, 1);
        for (int s = 0; s < ST; s++) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
        tc::mbar_init(&s_full[0], 1); tc::mbar_init(&s_full[1], 1);
        tc::mbar_init(p_full, 128);
        tc::mbar_init(pv_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 4) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 5) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(bar_q, Q_BYTES);
#pragma unroll
            for (int a = 0; a < DA; a++) tc::tma_load_4d(sQ + a * ATOM_BYTES, &tmQ, bar_q, a * 64, head, qb * BQ, batch);
            uint32_t stage = 0, phase = 0;
            for (uint32_t j = 0; j < nblk; j++) {
                tc::mbar_wait(&kv_empty[stage], phase ^ 1);
                tc::mbar_arrive_expect_tx(&kv_full[stage], 2 * KV_BYTES);
#pragma unroll
                for (int a = 0; a < DA; a++) {
                    tc::tma_load_4d(sK + stage * KV_BYTES + a * ATOM_BYTES, &tmK, &kv_full[stage], a * 64, head, j * BKV, batch);
                    tc::tma_load_4d(sV + stage * KV_BYTES + a * ATOM_BYTES, &tmV, &kv_full[stage], a * 64, head, j * BKV, batch);
                }
                if (++stage == ST) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_qk = tc::make_idesc_bf16(128, BKV, false, false);
            constexpr uint32_t idesc_pv = tc::make_idesc_bf16(128, DPAD, false, true);   // B = V is MN-major
            const uint32_t aQ = tc::smem_u32(sQ), aP = tc::smem_u32(sP);
            uint32_t kv_phase[ST];
#pragma unroll
            for (int s = 0; s < ST; s++) kv_phase[s] = 0;
            auto issue_qk = [&](uint32_t j) {
                const uint32_t st = j % ST;
                tc::mbar_wait(&kv_full[st], kv_phase[st]);
                kv_phase[st] ^= 1;
                tc::tc_fence_after();
                const uint32_t aK = tc::smem_u32(sK + st * KV_BYTES);
                const uint32_t d_tmem = tmem_base + ((j & 1) ? S_COL1 : S_COL0);
#pragma unroll
                for (int kk = 0; kk < DPAD / 16; kk++) {
                    const uint32_t off = (kk / 4) * ATOM_BYTES + (kk % 4) * 32;
                    tc::umma_f16(d_tmem, tc::make_desc_k_sw128(aQ + off), tc::make_desc_k_sw128(aK + off), idesc_qk, kk ? 1u : 0u);
                }
                tc::umma_commit(&s_full[j & 1]);
            };
            tc::mbar_wait(bar_q, 0);
            issue_qk(0);
            for (uint32_t j = 0; j < nblk; j++) {
                if (ST > 1 && j + 1 < nblk) issue_qk(j + 1);
                tc::mbar_wait(p_full, j & 1);
                tc::tc_fence_after();
                const uint32_t st = j % ST;
                const uint32_t aV = tc::smem_u32(sV + st * KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < BKV / 16; kk++) {
                    // A = P: K-major, 64-key atoms;  B = V: MN-major, 16 keys = 2048 B down the tile, d-atoms 16 KB apart
                    const uint64_t da = tc::make_desc_k_sw128(aP + (kk / 4) * ATOM_BYTES + (kk % 4) * 32);
                    const uint64_t db = tc::make_desc_mn_sw128(aV + kk * 2048, ATOM_BYTES, 1024);
                    tc::umma_f16(tmem_base + O_COL, da, db, idesc_pv, (j | kk) ? 1u : 0u);
                }
                tc::umma_commit(&kv_empty[st]);
                tc::umma_commit(pv_done);
                if (ST == 1 && j + 1 < nblk) issue_qk(j + 1);
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warps: thread <-> query row <-> TMEM lane
        const uint32_t row = warp * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float m = -INFINITY, l = 0.f;
        uint8_t* prow = sP + row * 128;
        const uint32_t sw = row & 7;
        for (uint32_t j = 0; j < nblk; j++) {
            tc::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t s_addr = lane_addr + ((j & 1) ? S_COL1 : S_COL0);
            const uint32_t kv_valid = min((uint32_t)BKV, p.kv_len - j * BKV);
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < BKV; c += 32) {
                uint32_t v[32];
                tc::tmem_ld32(s_addr + c, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i++)
                    if ((uint32_t)(c + i) < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float m_new = fmaxf(m, mx * p.scale_log2);
            const float alpha = exp2f(m - m_new);   // first block: exp2(-inf) = 0
            // previous PV must be complete before P is overwritten / O is rescaled
            if (j > 0) {
                tc::mbar_wait(pv_done, (j - 1) & 1);
                tc::tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                    for (int c = 0; c < DPAD; c += 16) {
                        uint32_t v[16];
                        tc::tmem_ld16(lane_addr + O_COL + c, v);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st16(lane_addr + O_COL + c, v);
                    }
                    tmem_st_wait();
                }
            }
            // pass 2: p = exp2(s*scale - m_new) -> bf16 -> swizzled smem; row sum
            float lsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < BKV; c += 32) {
                uint32_t v[32];
                tc::tmem_ld32(s_addr + c, v);
                tc::tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = ((uint32_t)(c + i) < kv_valid) ? exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new)) : 0.f;
                    float p1 = ((uint32_t)(c + i + 1) < kv_valid) ? exp2f(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_new)) : 0.f;
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
                    const float2 back = __bfloat1622float2(b2);   // sum what the tensor core will actually see
                    lsum += back.x + back.y;
                    pk[i / 2] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                // 32 keys = 4 chunks of 16 B inside atom (c / 64)
                uint8_t* base = prow + (c / 64) * ATOM_BYTES;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t chunk = ((c % 64) / 8 + q) ^ sw;
                    *reinterpret_cast<uint4*>(base + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            l = l * alpha + lsum;
            m = m_new;
            tc::fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc::tc_fence_before();
            tc::mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> global
        tc::mbar_wait(pv_done, (nblk - 1) & 1);
        tc::tc_fence_after();
        const uint32_t qrow = qb * BQ + row;
        const float inv_l = 1.0f / l;
        __nv_bfloat16* out = p.O + ((size_t)batch * p.q_len + qrow) * p.ldo + head * p.d;
#pragma unroll 1
        for (int c = 0; c < DPAD; c += 16) {
            uint32_t v[16];
            tc::tmem_ld16(lane_addr + O_COL + c, v);
            tc::tmem_ld_wait();
            if (qrow < p.q_len) {
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    if ((uint32_t)(c + g * 8) < p.d) {   // d is a multiple of 8
                        uint4 o;
                        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            o2[i] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * i]) * inv_l, __uint_as_float(v[g * 8 + 2 * i + 1]) * inv_l);
                        *reinterpret_cast<uint4*>(out + c + g * 8) = o;
                    }
                }
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Ping-pong variant for head dims <= 64 (one 64-wide atom): one CTA owns TWO 128-row query tiles (A, B) of the same (batch, head).
// 18 warps: 0-7 softmax A, 8-15 softmax B (two warps per TMEM lane quadrant, 64 key columns each), 16 MMA issuer, 17 TMA producer.  While the softmax warps of tile A work on S_A(j), the tensor
// core runs Q_B K_j^T / P_B V_j and vice versa, so softmax (MUFU / FMA pipes) and MMA overlap inside one SM, and every K/V tile
// fetched by TMA is used by both query tiles.  TMEM: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// explicit shared-memory accesses (the exchange buffer is carved out of the dynamic allocation: plain dereferences compile to
// generic LD/ST, which showed up as the second-largest stall of the softmax loop)
__device__ __forceinline__ void sts_f32(float* p, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(tc::smem_u32(p)), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(const float* p) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(tc::smem_u32(p)) : "memory");
    return v;
}

template <int DPAD>
__global__ void __launch_bounds__(576, 1) k_attention_pp(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                         const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    static_assert(DPAD <= 64, "ping-pong kernel: one 64-wide head-dim atom");
    constexpr int ST = 2;
    constexpr uint32_t O_COL0 = 256, O_COL1 = 320;
    constexpr int NCH = DPAD / 16;                 // 16-column chunks of O
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                          // [2][ATOM]
    uint8_t* sK = sQ + 2 * ATOM_BYTES;           // [ST][ATOM]
    uint8_t* sV = sK + ST * ATOM_BYTES;          // [ST][ATOM]
    uint8_t* sP = sV + ST * ATOM_BYTES;          // [2][2*ATOM]
    float* xch = reinterpret_cast<float*>(sP + 4 * ATOM_BYTES);   // [2 parity][2 tiles][2 halves][128] row-max exchange (+ final row sums)
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * 2 * 2 * 128);
    uint64_t* bar_q = bars;               // 1
    uint64_t* k_full = bars + 1;          // [2]
    uint64_t* k_empty = bars + 3;         // [2]
    uint64_t* v_full = bars + 5;          // [2]
    uint64_t* v_empty = bars + 7;         // [2]
    uint64_t* s_full = bars + 9;          // [2 tiles]
    uint64_t* s_free = bars + 11;         // [2 tiles], 256 arrivals: S(j) has been read into registers
    uint64_t* p_full = bars + 13;         // [2 tiles], 256 arrivals
    uint64_t* pv_done = bars + 15;        // [2 tiles]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t qb = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const uint32_t nblk = p.n_kv_blocks;

    if (warp == 17 && lane == 0) {
        tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV);
        tc::mbar_init(bar_q, 1);
        for (int s = 0; s < 2; s++) {
            tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1); tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], 1);
            tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_free[s], 256); tc::mbar_init(&p_full[s], 256); tc::mbar_init(&pv_done[s], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 16) tc::tmem_alloc(tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 17) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(bar_q, 2 * ATOM_BYTES);
            tc::tma_load_4d(sQ, &tmQ, bar_q, 0, head, qb * 256, batch);
            tc::tma_load_4d(sQ + ATOM_BYTES, &tmQ, bar_q, 0, head, qb * 256 + 128, batch);
            uint32_t stage = 0, phase = 0;
            for (uint32_t j = 0; j < nblk; j++) {       // K and V have their own rings: K(j+2) only waits for QK(j), V(j+2) for PV(j)
                tc::mbar_wait(&k_empty[stage], phase ^ 1);
                tc::mbar_arrive_expect_tx(&k_full[stage], ATOM_BYTES);
                tc::tma_load_4d(sK + stage * ATOM_BYTES, &tmK, &k_full[stage], 0, head, j * BKV, batch);
                tc::mbar_wait(&v_empty[stage], phase ^ 1);
                tc::mbar_arrive_expect_tx(&v_full[stage], ATOM_BYTES);
                tc::tma_load_4d(sV + stage * ATOM_BYTES, &tmV, &v_full[stage], 0, head, j * BKV, batch);
                if (++stage == ST) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 16) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = tc::make_idesc_bf16(128, BKV, false, false);
            constexpr uint32_t idesc_pv = tc::make_idesc_bf16(128, DPAD, false, true);
            const uint32_t aQ = tc::smem_u32(sQ), aP = tc::smem_u32(sP);
            auto issue_qk = [&](uint32_t tile, uint32_t j) {
                const uint32_t aK = tc::smem_u32(sK + (j % ST) * ATOM_BYTES);
#pragma unroll
                for (int kk = 0; kk < DPAD / 16; kk++)
                    tc::umma_f16(tmem_base + tile * 128, tc::make_desc_k_sw128(aQ + tile * ATOM_BYTES + kk * 32), tc::make_desc_k_sw128(aK + kk * 32),
                                 idesc_qk, kk ? 1u : 0u);
                tc::umma_commit(&s_full[tile]);
            };
            // Software pipeline (FlashAttention-3 style): the softmax warps pull S(j) into registers and release the S columns at
            // once (s_free), so QK(j+1) runs on the tensor pipe while the exponentials of block j are still being computed; PV(j)
            // follows when P(j) is in shared memory.  The exp phase (the MUFU-bound part) never waits for an MMA round trip.
            tc::mbar_wait(bar_q, 0);
            tc::mbar_wait(&k_full[0], 0);
            tc::tc_fence_after();
            issue_qk(0, 0);
            issue_qk(1, 0);
            tc::umma_commit(&k_empty[0]);
            for (uint32_t j = 0; j < nblk; j++) {
                const uint32_t st = j % ST;
                if (j + 1 < nblk) {
                    const uint32_t sn = (j + 1) % ST;
                    tc::mbar_wait(&k_full[sn], ((j + 1) / ST) & 1);
#pragma unroll
                    for (uint32_t tile = 0; tile < 2; tile++) {
                        tc::mbar_wait(&s_free[tile], j & 1);
                        tc::tc_fence_after();
                        issue_qk(tile, j + 1);
                    }
                    tc::umma_commit(&k_empty[sn]);
                }
                tc::mbar_wait(&v_full[st], (j / ST) & 1);
                const uint32_t aV = tc::smem_u32(sV + st * ATOM_BYTES);
#pragma unroll
                for (uint32_t tile = 0; tile < 2; tile++) {
                    tc::mbar_wait(&p_full[tile], j & 1);
                    tc::tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < BKV / 16; kk++) {
                        const uint64_t da = tc::make_desc_k_sw128(aP + tile * 2 * ATOM_BYTES + (kk / 4) * ATOM_BYTES + (kk % 4) * 32);
                        const uint64_t db = tc::make_desc_mn_sw128(aV + kk * 2048, ATOM_BYTES, 1024);
                        tc::umma_f16(tmem_base + (tile ? O_COL1 : O_COL0), da, db, idesc_pv, (j | kk) ? 1u : 0u);
                    }
                    if (tile == 1) tc::umma_commit(&v_empty[st]);
                    tc::umma_commit(&pv_done[tile]);
                }
            }
        }
    } else {
        // 16 softmax warps: tile = warp / 8; inside a tile two warps share each TMEM lane quadrant and split the 128 key columns
        const uint32_t tile = warp >> 3, w8 = warp & 7, quad = w8 & 3, half = w8 >> 2;
        const uint32_t row = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t s_addr = lane_addr + tile * 128 + half * 64;
        const uint32_t o_addr = lane_addr + (tile ? O_COL1 : O_COL0);
        constexpr int C0 = 0, C1 = (NCH + 1) / 2;                  // O chunks [C0,C1) -> half 0, [C1,NCH) -> half 1
        const int oc_lo = half ? C1 : C0, oc_hi = half ? NCH : C1;
        float m = -INFINITY, l = 0.f;
        uint8_t* prow = sP + tile * 2 * ATOM_BYTES + half * ATOM_BYTES + row * 128;
        const uint32_t sw = row & 7;
        for (uint32_t j = 0; j < nblk; j++) {
            tc::mbar_wait(&s_full[tile], j & 1);
            tc::tc_fence_after();
            const uint32_t kv_valid = min((uint32_t)BKV, p.kv_len - j * BKV);
            const bool full_blk = kv_valid == BKV;
            const uint32_t col0 = half * 64;
            // S(j): this thread's 64 scores stay in registers for both passes; the TMEM columns are released immediately
            uint32_t v0[32], v1[32];
            tc::tmem_ld32(s_addr, v0);
            tc::tmem_ld32(s_addr + 32, v1);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&s_free[tile]);
            float mx = -INFINITY;
            if (full_blk) {
#pragma unroll
                for (int i = 0; i < 32; i++) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    if (col0 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v0[i]));
                    if (col0 + 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v1[i]));
                }
            }
            float* xm = xch + (((j & 1) * 2 + tile) * 2) * 128;
            sts_f32(xm + half * 128 + row, mx);
            asm volatile("bar.sync %0, 256;" ::"r"(1 + (int)tile) : "memory");
            mx = fmaxf(mx, lds_f32(xm + (half ^ 1) * 128 + row));
            // Lazy rescaling (FlashAttention-4): the running reference m only moves when the row maximum grows by more than 2^8;
            // until then P = 2^(s - m) may exceed 1 (<= 256, exact in fp32 / fine in bf16) and O, l keep their scale -- the final
            // O / l is unchanged, but the TMEM round trip that rescales O disappears from almost every block.
            const float mx_s = mx * p.scale_log2;
            const float m_new = (mx_s > m + 8.0f) ? mx_s : m;
            const float alpha = ex2_approx(m - m_new);
            // PV(j-1) must have finished before O is rescaled or P is overwritten: with lazy rescaling the rescale is rare, so
            // the wait normally moves behind the exponentials (the MUFU-bound part) where PV(j-1) has long completed.
            const bool resc = j > 0 && __any_sync(0xffffffffu, alpha != 1.0f);
            if (resc) {
                tc::mbar_wait(&pv_done[tile], (j - 1) & 1);
                tc::tc_fence_after();
#pragma unroll 1
                for (int c = oc_lo; c < oc_hi; c++) {
                    uint32_t v[16];
                    tc::tmem_ld16(o_addr + c * 16, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    tmem_st16(o_addr + c * 16, v);
                }
                tmem_st_wait();
            }
            float lsum0 = 0.f, lsum1 = 0.f;
            uint32_t pk2[2][16];
#pragma unroll
            for (int cb = 0; cb < 2; cb++) {
                const uint32_t* v = cb ? v1 : v0;
                const int c = cb * 32;
                uint32_t (&pk)[16] = pk2[cb];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
                    float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -m_new));
                    if (!full_blk) {
                        if (col0 + c + i >= kv_valid) p0 = 0.f;
                        if (col0 + c + i + 1 >= kv_valid) p1 = 0.f;
                    }
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
                    lsum0 += p0; lsum1 += p1;
                    pk[i / 2] = *reinterpret_cast<const uint32_t*>(&b2);
                }
            }
            if (j > 0 && !resc) {
                tc::mbar_wait(&pv_done[tile], (j - 1) & 1);
                tc::tc_fence_after();
            }
#pragma unroll
            for (int cb = 0; cb < 2; cb++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t chunk = (cb * 4 + q) ^ sw;
                    *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk2[cb][4 * q], pk2[cb][4 * q + 1], pk2[cb][4 * q + 2], pk2[cb][4 * q + 3]);
                }
            const float lsum = lsum0 + lsum1;
            l = l * alpha + lsum;
            m = m_new;
            tc::fence_proxy_async_smem();
            tc::tc_fence_before();
            tc::mbar_arrive(&p_full[tile]);
        }
        // combine the two halves' row sums, then O / l -> bf16 -> global (each half stores its O chunks)
        float* xs = xch + (((nblk & 1) * 2 + tile) * 2) * 128;
        xs[half * 128 + row] = l;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + (int)tile) : "memory");
        l += xs[(half ^ 1) * 128 + row];
        tc::mbar_wait(&pv_done[tile], (nblk - 1) & 1);
        tc::tc_fence_after();
        const uint32_t qrow = qb * 256 + tile * 128 + row;
        const float inv_l = 1.0f / l;
        __nv_bfloat16* out = p.O + ((size_t)batch * p.q_len + qrow) * p.ldo + head * p.d;
#pragma unroll 1
        for (int cc = oc_lo; cc < oc_hi; cc++) {
            const int c = cc * 16;
            uint32_t v[16];
            tc::tmem_ld16(o_addr + c, v);
            tc::tmem_ld_wait();
            if (qrow < p.q_len) {
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    if ((uint32_t)(c + g * 8) < p.d) {
                        uint4 o;
                        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            o2[i] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * i]) * inv_l, __uint_as_float(v[g * 8 + 2 * i + 1]) * inv_l);
                        *reinterpret_cast<uint4*>(out + c + g * 8) = o;
                    }
                }
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

template <int DPAD>
int launch_attn_pp(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, uint32_t heads, uint32_t batch,
                   cudaStream_t s) {
    constexpr int SMEM = ATOM_BYTES * (2 + 2 + 2 + 4) + 2 * 2 * 2 * 128 * 4 + 1024 + 256;   // 256 B of barriers: 17 x 8 + 4
    static bool configured = false;
    if (!configured) {
        MVE_CUDA(cudaFuncSetAttribute(k_attention_pp<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    const dim3 grid((p.q_len + 255) / 256, heads, batch);
    k_attention_pp<DPAD><<<grid, 576, SMEM, s>>>(tq, tk, tv, p);
    MVE_CHECK_LAUNCH("k_attention_pp");
    return 0;
}

template <int DPAD, int ST>
int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid, cudaStream_t s) {
    constexpr int DA = (DPAD + 63) / 64;
    constexpr int SMEM = DA * ATOM_BYTES * (1 + 2 * ST) + 2 * ATOM_BYTES + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        MVE_CUDA(cudaFuncSetAttribute(k_attention<DPAD, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    k_attention<DPAD, ST><<<grid, NUM_THREADS_A, SMEM, s>>>(tq, tk, tv, p);
    MVE_CHECK_LAUNCH("k_attention");
    return 0;
}

int make_qkv_map(CUtensorMap* m, const void* base, uint32_t d, uint32_t heads, uint32_t len, uint32_t batch, uint32_t ld, const char* what) {
    const uint64_t dims[4] = {d, heads, len, batch};
    const uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)len * ld * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    return mve_make_tmap_bf16(m, base, 4, dims, str, box, what);
}

}  // namespace

extern "C" {

int mve_attention_bf16(const void* Q, const void* K, const void* V, void* O, uint32_t batch, uint32_t heads, uint32_t q_len,
                       uint32_t kv_len, uint32_t d, uint32_t ldq, uint32_t ldk, uint32_t ldv, uint32_t ldo, float scale, void* stream) {
    if (batch == 0 || q_len == 0) return 0;
    MVE_ARG(d == 40 || d == 80 || d == 160 || d == 64 || d == 128, "attention: head dim must be 40, 64, 80, 128 or 160");
    MVE_ARG(kv_len > 0, "attention: kv_len must be > 0");
    MVE_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention: leading dims must be multiples of 8 elements");
    MVE_ARG(((uintptr_t)Q & 15) == 0 && ((uintptr_t)K & 15) == 0 && ((uintptr_t)V & 15) == 0 && ((uintptr_t)O & 15) == 0,
            "attention: pointers must be 16-byte aligned");
    CUtensorMap tq, tk, tv;
    int r;
    if ((r = make_qkv_map(&tq, Q, d, heads, q_len, batch, ldq, "attention Q"))) return r;
    if ((r = make_qkv_map(&tk, K, d, heads, kv_len, batch, ldk, "attention K"))) return r;
    if ((r = make_qkv_map(&tv, V, d, heads, kv_len, batch, ldv, "attention V"))) return r;
    AttnParams p{};
    p.O = (__nv_bfloat16*)O; p.ldo = ldo; p.q_len = q_len; p.kv_len = kv_len; p.heads = heads; p.d = d;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.n_kv_blocks = (kv_len + BKV - 1) / BKV;
    const dim3 grid((q_len + BQ - 1) / BQ, heads, batch);
    cudaStream_t s = (cudaStream_t)stream;
    if (d <= 48) return launch_attn_pp<48>(tq, tk, tv, p, heads, batch, s);
    if (d <= 64) return launch_attn_pp<64>(tq, tk, tv, p, heads, batch, s);
    if (d <= 80) return launch_attn<80, 2>(tq, tk, tv, p, grid, s);
    if (d <= 128) return launch_attn<128, 2>(tq, tk, tv, p, grid, s);
    return launch_attn<160, 1>(tq, tk, tv, p, grid, s);
}

}  // extern "C"

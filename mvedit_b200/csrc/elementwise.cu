// elementwise.cu -- HBM-bound glue of the UNet / ControlNet path (bf16, NHWC / row-major, 16-byte vector accesses):
//   GroupNorm(32)+SiLU, LayerNorm, GEGLU gate, nearest x2 upsample, stride-2 3x3 im2col, timestep sinusoid.
// These replace torch.nn.GroupNorm / LayerNorm / F.gelu / F.interpolate calls inside diffusers' ResnetBlock2D,
// Transformer2DModel, Upsample2D, Downsample2D; each is one read + one write of its tensor.
#include "common.cuh"
#include "../../include/mvedit_b200.h"

namespace {

using bf16 = __nv_bfloat16;
using bf162 = __nv_bfloat162;

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
    const bf162* h = reinterpret_cast<const bf162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; i++) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    bf162* h = reinterpret_cast<bf162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return o;
}
__device__ __forceinline__ float silu(float v) { return v / (1.0f + __expf(-v)); }

// ---------------------------------------------------------------- GroupNorm
// Block = k pixel lanes x (C/8) channel vectors, so a thread keeps ONE 8-channel vector for its whole pixel range: statistics
// accumulate per channel in registers (folded into their groups once at the end) and the per-channel scale / shift of
// the apply pass live in registers too.  (The first version re-derived the vector per visit and folded every 16-byte load into
// its group with shared-memory atomics: 256 threads contending for 32 counters, 110 us for a 168 MB tensor.)
// stats: grid (chunks, B); stats[b][g] = (sum, sumsq) fp32, zeroed by the launcher; one global atomic per (group, CTA).
__global__ void __launch_bounds__(512) k_gn_stats(const bf16* __restrict__ x, const uint32_t HW, const uint32_t C, const uint32_t G,
                                                  const uint32_t px_per_cta, float* __restrict__ stats) {
    __shared__ float s_sum[64], s_sq[64];
    const uint32_t b = blockIdx.y;
    const uint32_t vpp = C / 8, lanes = blockDim.x / vpp;
    if (threadIdx.x < 64) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
    __syncthreads();
    const uint32_t cv = threadIdx.x % vpp, pl = threadIdx.x / vpp;
    const uint32_t cpg = C / G, c0 = cv * 8;
    const uint32_t p0 = blockIdx.x * px_per_cta, p1 = min(HW, p0 + px_per_cta);
    const bf16* xb = x + (size_t)b * HW * C + c0;
    float sa[8], qa[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { sa[k] = 0.f; qa[k] = 0.f; }
    // four independent 16-byte loads in flight per thread (a 2 GB VAE activation is pure HBM streaming: with one load per
    // iteration the kernel ran at ~37 % of the copy bandwidth)
    uint32_t px = p0 + pl;
    for (; px + 3 * lanes < p1; px += 4 * lanes) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; u++) r[u] = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)(px + u * lanes) * C));
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float f[8];
            unpack8(r[u], f);
#pragma unroll
            for (int k = 0; k < 8; k++) { sa[k] += f[k]; qa[k] = fmaf(f[k], f[k], qa[k]); }
        }
    }
    for (; px < p1; px += lanes) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(xb + (size_t)px * C), f);
#pragma unroll
        for (int k = 0; k < 8; k++) { sa[k] += f[k]; qa[k] = fmaf(f[k], f[k], qa[k]); }
    }
    // fold the 8 channels into their groups (runs of equal group first, so C/G >= 8 costs at most two atomics per sum)
    {
        uint32_t g = c0 / cpg;
        float s = 0.f, q = 0.f;
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) {
            const uint32_t gk = (c0 + k) / cpg;
            if (gk != g) { atomicAdd(&s_sum[g], s); atomicAdd(&s_sq[g], q); g = gk; s = 0.f; q = 0.f; }
            s += sa[k]; q += qa[k];
        }
        atomicAdd(&s_sum[g], s); atomicAdd(&s_sq[g], q);
    }
    __syncthreads();
    if (threadIdx.x < G) {
        atomicAdd(&stats[((size_t)b * G + threadIdx.x) * 2], s_sum[threadIdx.x]);
        atomicAdd(&stats[((size_t)b * G + threadIdx.x) * 2 + 1], s_sq[threadIdx.x]);
    }
}

// apply: y = (x - mean) * rstd * gamma + beta, optional SiLU.  grid (chunks, B); same thread <-> channel-vector mapping.
__global__ void __launch_bounds__(512) k_gn_apply(const bf16* __restrict__ x, bf16* __restrict__ y, const uint32_t HW, const uint32_t C,
                                                  const uint32_t G, const uint32_t px_per_cta, const float* __restrict__ stats,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta, const float eps,
                                                  const int act) {
    const uint32_t b = blockIdx.y;
    const uint32_t vpp = C / 8, lanes = blockDim.x / vpp;
    const uint32_t cv = threadIdx.x % vpp, pl = threadIdx.x / vpp;
    const uint32_t cpg = C / G, c0 = cv * 8;
    const float inv_n = 1.0f / ((float)HW * (float)cpg);
    float sc[8], sh[8];
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
        const uint32_t c = c0 + k, g = c / cpg;
        const float mean = stats[((size_t)b * G + g) * 2] * inv_n;
        const float var = fmaxf(stats[((size_t)b * G + g) * 2 + 1] * inv_n - mean * mean, 0.f);
        sc[k] = rsqrtf(var + eps) * gamma[c];
        sh[k] = beta[c] - mean * sc[k];
    }
    const uint32_t p0 = blockIdx.x * px_per_cta, p1 = min(HW, p0 + px_per_cta);
    const size_t base = (size_t)b * HW * C + c0;
    uint32_t px = p0 + pl;
    for (; px + 3 * lanes < p1; px += 4 * lanes) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; u++) r[u] = __ldg(reinterpret_cast<const uint4*>(x + base + (size_t)(px + u * lanes) * C));
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float f[8];
            unpack8(r[u], f);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float v = fmaf(f[k], sc[k], sh[k]);
                f[k] = act ? silu(v) : v;
            }
            *reinterpret_cast<uint4*>(y + base + (size_t)(px + u * lanes) * C) = pack8(f);
        }
    }
    for (; px < p1; px += lanes) {
        const size_t off = base + (size_t)px * C;
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + off), f);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float v = fmaf(f[k], sc[k], sh[k]);
            f[k] = act ? silu(v) : v;
        }
        *reinterpret_cast<uint4*>(y + off) = pack8(f);
    }
}

// ---------------------------------------------------------------- LayerNorm over the last dim
// Every width on the SD1.5 path is 320 * {1, 2, 4} = 40 * LPR vectors of 8: LPR = 8 / 16 / 32 lanes share a row with exactly 5
// vectors each (no idle lanes, the row stays in registers between the statistics and the normalisation: one read, one write).
template <int LPR>
__global__ void __launch_bounds__(256) k_layernorm5(const bf16* __restrict__ x, bf16* __restrict__ y, const uint32_t rows, const uint32_t C,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta, const float eps) {
    constexpr int RPW = 32 / LPR;                      // rows per warp
    const uint32_t lane = threadIdx.x & 31, sub = lane % LPR;
    const uint32_t row = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / LPR;
    const bool ok = row < rows;
    const bf16* xr = x + (size_t)(ok ? row : 0) * C;
    float f[5][8];
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        unpack8(*reinterpret_cast<const uint4*>(xr + (sub + LPR * c) * 8), f[c]);
#pragma unroll
        for (int k = 0; k < 8; k++) { s += f[c][k]; q += f[c][k] * f[c][k]; }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    const float mean = s / C;
    const float rstd = rsqrtf(fmaxf(q / C - mean * mean, 0.f) + eps);
    if (!ok) return;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const uint32_t v = sub + LPR * c;
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = (f[c][k] - mean) * rstd * g[k] + b[k];
        *reinterpret_cast<uint4*>(y + (size_t)row * C + v * 8) = pack8(o);
    }
}

// generic width (C % 8 == 0, C <= 1280): one warp per row, up to 5 vectors per lane
__global__ void __launch_bounds__(256) k_layernorm(const bf16* __restrict__ x, bf16* __restrict__ y, const uint32_t rows, const uint32_t C,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta, const float eps) {
    const uint32_t row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const bf16* xr = x + (size_t)row * C;
    const uint32_t nvec = C / 8;
    float s = 0.f, q = 0.f;
    float f[5][8];
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const uint32_t v = lane + 32 * c;
        if (v < nvec) {
            unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f[c]);
#pragma unroll
            for (int k = 0; k < 8; k++) { s += f[c][k]; q += f[c][k] * f[c][k]; }
        }
    }
    s = warp_sum(s); q = warp_sum(q);
    const float mean = s / C;
    const float rstd = rsqrtf(fmaxf(q / C - mean * mean, 0.f) + eps);
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const uint32_t v = lane + 32 * c;
        if (v < nvec) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] = (f[c][k] - mean) * rstd * gamma[v * 8 + k] + beta[v * 8 + k];
            *reinterpret_cast<uint4*>(y + (size_t)row * C + v * 8) = pack8(o);
        }
    }
}

// ---------------------------------------------------------------- GEGLU: y[m, j] = h[m, j] * gelu(h[m, F + j]),  h [M, 2F]
__global__ void __launch_bounds__(256) k_geglu(const bf16* __restrict__ h, bf16* __restrict__ y, const size_t M, const uint32_t F) {
    const size_t nvec = M * (F / 8);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
        const size_t m = i / (F / 8);
        const uint32_t j = (uint32_t)(i % (F / 8)) * 8;
        float a[8], g[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(h + m * 2 * F + j), a);
        unpack8(*reinterpret_cast<const uint4*>(h + m * 2 * F + F + j), g);
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = a[k] * (0.5f * g[k] * (1.0f + erff(g[k] * 0.70710678118654752f)));
        *reinterpret_cast<uint4*>(y + m * F + j) = pack8(o);
    }
}


// ---------------------------------------------------------------- row softmax (the AutoencoderKL mid-block attention: one head, d = 512,
// scores materialised by the GEMM).  One CTA of 128 threads per row, the row stays in registers: one read, one write.
__global__ void __launch_bounds__(128) k_softmax_rows(const bf16* __restrict__ x, bf16* __restrict__ y, const uint32_t cols, const uint32_t ldx,
                                                      const uint32_t ldy, const float scale_log2e) {
    constexpr int MAXV = 8;                     // cols <= 128 * 8 * 8 = 8192
    __shared__ float s_red[4];
    const uint32_t row = blockIdx.x, nvec = cols / 8;
    const bf16* xr = x + (size_t)row * ldx;
    float f[MAXV][8];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXV; c++) {
        const uint32_t v = threadIdx.x + 128 * c;
        if (v < nvec) {
            unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f[c]);
#pragma unroll
            for (int k = 0; k < 8; k++) m = fmaxf(m, f[c][k]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXV; c++) {
        const uint32_t v = threadIdx.x + 128 * c;
        if (v < nvec) {
#pragma unroll
            for (int k = 0; k < 8; k++) { f[c][k] = exp2f((f[c][k] - m) * scale_log2e); s += f[c][k]; }
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    const float inv = 1.0f / (s_red[0] + s_red[1] + s_red[2] + s_red[3]);
    bf16* yr = y + (size_t)row * ldy;
#pragma unroll
    for (int c = 0; c < MAXV; c++) {
        const uint32_t v = threadIdx.x + 128 * c;
        if (v < nvec) {
#pragma unroll
            for (int k = 0; k < 8; k++) f[c][k] *= inv;
            *reinterpret_cast<uint4*>(yr + v * 8) = pack8(f[c]);
        }
    }
}

// ---------------------------------------------------------------- nearest x2 upsample NHWC
__global__ void __launch_bounds__(256) k_upsample2x(const bf16* __restrict__ x, bf16* __restrict__ y, const uint32_t B, const uint32_t H,
                                                    const uint32_t W, const uint32_t C) {
    const uint32_t vpp = C / 8;
    const size_t total = (size_t)B * 2 * H * 2 * W * vpp;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const uint32_t cv = (uint32_t)(i % vpp);
        size_t px = i / vpp;
        const uint32_t ow = (uint32_t)(px % (2 * W)); px /= 2 * W;
        const uint32_t oh = (uint32_t)(px % (2 * H));
        const uint32_t b = (uint32_t)(px / (2 * H));
        const uint4 v = *reinterpret_cast<const uint4*>(x + (((size_t)b * H + oh / 2) * W + ow / 2) * C + cv * 8);
        *reinterpret_cast<uint4*>(y + i * 8) = v;
    }
}

// ---------------------------------------------------------------- im2col for 3x3 stride-2 pad-1 conv: out [B*Ho*Wo, 9*C], K index = tap*C + c
__global__ void __launch_bounds__(256) k_im2col_s2(const bf16* __restrict__ x, bf16* __restrict__ y, const uint32_t B, const uint32_t H,
                                                   const uint32_t W, const uint32_t C, const int pad_lo) {
    const uint32_t Ho = H / 2, Wo = W / 2, vpp = C / 8;
    const size_t total = (size_t)B * Ho * Wo * 9 * vpp;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const uint32_t cv = (uint32_t)(i % vpp);
        size_t r = i / vpp;
        const uint32_t tap = (uint32_t)(r % 9); r /= 9;
        const uint32_t wo = (uint32_t)(r % Wo); r /= Wo;
        const uint32_t ho = (uint32_t)(r % Ho);
        const uint32_t b = (uint32_t)(r / Ho);
        const int ih = (int)(2 * ho) + (int)(tap / 3) - pad_lo, iw = (int)(2 * wo) + (int)(tap % 3) - pad_lo;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (ih >= 0 && ih < (int)H && iw >= 0 && iw < (int)W) v = *reinterpret_cast<const uint4*>(x + (((size_t)b * H + ih) * W + iw) * C + cv * 8);
        *reinterpret_cast<uint4*>(y + i * 8) = v;
    }
}

// ---------------------------------------------------------------- NCHW fp32/bf16 <-> NHWC bf16 with channel padding (latents 4ch, control images 3ch)
template <typename T>
__global__ void k_nchw_to_nhwc_pad(const T* __restrict__ x, bf16* __restrict__ y, const uint32_t B, const uint32_t C, const uint32_t HW,
                                   const uint32_t Cpad) {
    const size_t total = (size_t)B * HW * Cpad;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const uint32_t c = (uint32_t)(i % Cpad);
        const size_t px = i / Cpad;
        const uint32_t b = (uint32_t)(px / HW), p = (uint32_t)(px % HW);
        y[i] = (c < C) ? __float2bfloat16((float)x[((size_t)b * C + c) * HW + p]) : __float2bfloat16(0.f);
    }
}

// ---------------------------------------------------------------- L2 gather ceiling (diagnostic: the roof of the fused renderer)
// Every thread issues `per_thread` random 8-byte gathers from a table (indices from an in-register LCG, 8 independent loads in
// flight), sums them and writes one float: nothing but the gathers touches memory.
__global__ void __launch_bounds__(256) k_gather_ceiling(const float2* __restrict__ table, const uint32_t n_entries, const uint32_t per_thread,
                                                        float* __restrict__ out) {
    uint32_t s = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (uint32_t i = 0; i < per_thread; i += 8) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            s = s * 1664525u + 1013904223u;
            v[u] = __ldg(table + (uint32_t)(((uint64_t)s * n_entries) >> 32));
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u].x + v[u].y;
    }
    out[blockIdx.x * 256u + threadIdx.x] = acc;
}

}  // namespace

extern "C" {

int mve_gather_ceiling(const float* table, uint32_t n_entries, uint32_t blocks, uint32_t per_thread, float* out, void* stream) {
    MVE_ARG(n_entries > 0 && per_thread % 8 == 0, "gather_ceiling: per_thread must be a multiple of 8");
    k_gather_ceiling<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float2*)table, n_entries, per_thread, out);
    MVE_CHECK_LAUNCH("mve_gather_ceiling");
    return 0;
}

int mve_groupnorm_bf16(const void* x, void* y, uint32_t B, uint32_t HW, uint32_t C, uint32_t G, const float* gamma, const float* beta,
                       float eps, int silu_act, float* stats_scratch, void* stream) {
    if (B == 0) return 0;
    MVE_ARG(C % 8 == 0 && C % G == 0 && G <= 64, "groupnorm: C % 8 == 0, C % G == 0, G <= 64 required");
    MVE_ARG(C / 8 <= 512, "groupnorm: C <= 4096 required");
    cudaStream_t s = (cudaStream_t)stream;
    MVE_CUDA(cudaMemsetAsync(stats_scratch, 0, (size_t)B * G * 2 * sizeof(float), s));
    const uint32_t vpp = C / 8;
    uint32_t lanes = 384 / vpp;                      // pixel lanes per CTA: block = lanes * vpp threads (240 .. 512)
    if (lanes < 1) lanes = 1;
    const uint32_t threads = lanes * vpp;
    // aim for ~6 CTAs per SM in total, a whole number of pixel-lane rounds per CTA
    uint32_t chunks = (6 * kNumSM + B - 1) / B;
    if (chunks < 1) chunks = 1;
    uint32_t px_per_cta = (HW + chunks - 1) / chunks;
    px_per_cta = ((px_per_cta + lanes - 1) / lanes) * lanes;
    chunks = (HW + px_per_cta - 1) / px_per_cta;
    const dim3 grid(chunks, B);
    k_gn_stats<<<grid, threads, 0, s>>>((const bf16*)x, HW, C, G, px_per_cta, stats_scratch);
    k_gn_apply<<<grid, threads, 0, s>>>((const bf16*)x, (bf16*)y, HW, C, G, px_per_cta, stats_scratch, gamma, beta, eps, silu_act);
    MVE_CHECK_LAUNCH("mve_groupnorm_bf16");
    return 0;
}

int mve_layernorm_bf16(const void* x, void* y, uint32_t rows, uint32_t C, const float* gamma, const float* beta, float eps, void* stream) {
    if (rows == 0) return 0;
    MVE_ARG(C % 8 == 0 && C <= 1280, "layernorm: C % 8 == 0 and C <= 1280 required");
    const bool al = (((uintptr_t)gamma | (uintptr_t)beta) & 15) == 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (al && C == 320) k_layernorm5<8><<<cdiv(rows, 32), 256, 0, s>>>((const bf16*)x, (bf16*)y, rows, C, gamma, beta, eps);
    else if (al && C == 640) k_layernorm5<16><<<cdiv(rows, 16), 256, 0, s>>>((const bf16*)x, (bf16*)y, rows, C, gamma, beta, eps);
    else if (al && C == 1280) k_layernorm5<32><<<cdiv(rows, 8), 256, 0, s>>>((const bf16*)x, (bf16*)y, rows, C, gamma, beta, eps);
    else k_layernorm<<<cdiv(rows, 8), 256, 0, s>>>((const bf16*)x, (bf16*)y, rows, C, gamma, beta, eps);
    MVE_CHECK_LAUNCH("mve_layernorm_bf16");
    return 0;
}

int mve_geglu_bf16(const void* h, void* y, uint64_t M, uint32_t F, void* stream) {
    if (M == 0) return 0;
    MVE_ARG(F % 8 == 0, "geglu: F % 8 == 0 required");
    const size_t nvec = (size_t)M * (F / 8);
    uint32_t grid = (uint32_t)((nvec + 255) / 256 < (size_t)(16 * kNumSM) ? (nvec + 255) / 256 : (size_t)(16 * kNumSM));
    k_geglu<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)h, (bf16*)y, (size_t)M, F);
    MVE_CHECK_LAUNCH("mve_geglu_bf16");
    return 0;
}

int mve_upsample2x_bf16(const void* x, void* y, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* stream) {
    if (B == 0) return 0;
    MVE_ARG(C % 8 == 0, "upsample2x: C % 8 == 0 required");
    const size_t total = (size_t)B * 4 * H * W * (C / 8);
    uint32_t grid = (uint32_t)((total + 255) / 256 < (size_t)(16 * kNumSM) ? (total + 255) / 256 : (size_t)(16 * kNumSM));
    k_upsample2x<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, H, W, C);
    MVE_CHECK_LAUNCH("mve_upsample2x_bf16");
    return 0;
}

int mve_im2col3x3s2_bf16(const void* x, void* y, uint32_t B, uint32_t H, uint32_t W, uint32_t C, int pad_lo, void* stream) {
    if (B == 0) return 0;
    MVE_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "im2col3x3s2: C % 8 == 0, even H and W required");
    MVE_ARG(pad_lo == 0 || pad_lo == 1, "im2col3x3s2: pad_lo must be 1 (pad 1) or 0 (diffusers Downsample2D padding=0: F.pad (0,1,0,1))");
    const size_t total = (size_t)B * (H / 2) * (W / 2) * 9 * (C / 8);
    uint32_t grid = (uint32_t)((total + 255) / 256 < (size_t)(16 * kNumSM) ? (total + 255) / 256 : (size_t)(16 * kNumSM));
    k_im2col_s2<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, H, W, C, pad_lo);
    MVE_CHECK_LAUNCH("mve_im2col3x3s2_bf16");
    return 0;
}

int mve_softmax_rows_bf16(const void* x, void* y, uint32_t rows, uint32_t cols, uint32_t ldx, uint32_t ldy, float scale, void* stream) {
    if (rows == 0) return 0;
    MVE_ARG(cols % 8 == 0 && cols <= 8192 && ldx % 8 == 0 && ldy % 8 == 0, "softmax_rows: cols % 8 == 0, cols <= 8192, 16-byte aligned rows required");
    k_softmax_rows<<<rows, 128, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, cols, ldx, ldy, scale * 1.4426950408889634f);
    MVE_CHECK_LAUNCH("mve_softmax_rows_bf16");
    return 0;
}

int mve_nchw_to_nhwc_pad_bf16(const void* x, int x_is_f32, void* y, uint32_t B, uint32_t C, uint32_t HW, uint32_t Cpad, void* stream) {
    if (B == 0) return 0;
    MVE_ARG(Cpad >= C, "nchw_to_nhwc_pad: Cpad >= C required");
    const size_t total = (size_t)B * HW * Cpad;
    uint32_t grid = (uint32_t)((total + 255) / 256 < (size_t)(16 * kNumSM) ? (total + 255) / 256 : (size_t)(16 * kNumSM));
    if (x_is_f32) k_nchw_to_nhwc_pad<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (bf16*)y, B, C, HW, Cpad);
    else k_nchw_to_nhwc_pad<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, C, HW, Cpad);
    MVE_CHECK_LAUNCH("mve_nchw_to_nhwc_pad_bf16");
    return 0;
}

}  // extern "C"


// ---- tail of SRVGGNetCompact.forward (lib/models/decoders/image_space_ss.py:66-75): PixelShuffle(r) of the last conv's output plus the
// nearest-upsampled input.  y [B,H,W,ldy >= C r^2] bf16 NHWC (channel c r^2 + i r + j), x [B,H,W,ldx >= C] bf16 NHWC ->
// out [B,C,rH,rW] bf16 NCHW: out[b,c,h r + i,w r + j] = y[b,h,w,c r^2 + i r + j] + x[b,h,w,c].  One thread per output pixel pair row.
namespace {
__global__ void __launch_bounds__(256) k_pixel_shuffle_add(const bf16* __restrict__ y, const bf16* __restrict__ x, uint32_t B, uint32_t H,
                                                           uint32_t W, uint32_t C, uint32_t r, uint32_t ldy, uint32_t ldx,
                                                           bf16* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;                 // over (b, c, oh, ow), ow fastest: coalesced NCHW writes
    const uint32_t OW = W * r, OH = H * r;
    if (t >= (size_t)B * C * OH * OW) return;
    const uint32_t ow = t % OW, oh = (t / OW) % OH, c = (t / ((size_t)OW * OH)) % C, b = t / ((size_t)OW * OH * C);
    const uint32_t w = ow / r, j = ow % r, h = oh / r, i = oh % r;
    const size_t pix = ((size_t)b * H + h) * W + w;
    const float v = __bfloat162float(y[pix * ldy + c * r * r + i * r + j]) + __bfloat162float(x[pix * ldx + c]);
    out[t] = __float2bfloat16(v);
}
}  // namespace

extern "C" int mve_pixel_shuffle_add_bf16(const void* y, const void* x, uint32_t B, uint32_t H, uint32_t W, uint32_t C, uint32_t r,
                                          uint32_t ldy, uint32_t ldx, void* out, void* stream) {
    const size_t n = (size_t)B * C * H * r * W * r;
    if (n == 0) return 0;
    MVE_ARG(ldy >= C * r * r && ldx >= C, "pixel_shuffle_add: ldy >= C r^2 and ldx >= C");
    k_pixel_shuffle_add<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)y, (const bf16*)x, B, H, W, C, r, ldy, ldx, (bf16*)out);
    MVE_CHECK_LAUNCH("mve_pixel_shuffle_add_bf16");
    return 0;
}

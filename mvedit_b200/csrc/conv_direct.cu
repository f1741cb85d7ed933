// conv_direct.cu -- 3x3 convolutions with FEW channels (Cin <= 32) on the CUDA cores: the front of the ControlNet hint embedding.
//
// diffusers' ControlNetConditioningEmbedding (called under Adapter3DMixin.get_noise_pred*, /root/reference/
// lib/pipelines/adapter3d_mixin.py:101-109) runs Conv3x3 3->16, 16->16, 16->32 (stride 2), 32->32, 32->96 (stride 2) on the 512^2
// condition images before anything is wide enough for a tensor-core tile.  Through the implicit-GEMM kernel those layers needed their
// channels zero-padded to 64 (K = 9*64 of which 27 ... 288 are real) and an im2col pass for the strided ones: ~15 ms per step of
// padding traffic.  Here one thread owns one output pixel, keeps all Cout accumulators in registers, reads its 3x3 x Cin neighbourhood
// with 16-byte loads (NHWC bf16; the first layer reads the NCHW image directly) and the weights from shared memory (broadcast):
// HBM-bound on the true tensor sizes (16-96 channels), no padding, no im2col.
#include "common.cuh"
#include "../../include/mvedit_b200.h"

namespace {

using bf16 = __nv_bfloat16;

template <int CIN, int COUT, int STRIDE, int IN_FMT>     // IN_FMT 0: NHWC bf16, 1: NCHW f32, 2: NCHW bf16
__global__ void __launch_bounds__(128) k_conv3x3_direct(const void* __restrict__ xin, const float* __restrict__ wpk, const float* __restrict__ bias,
                                                        bf16* __restrict__ y, const uint32_t B, const uint32_t H, const uint32_t W,
                                                        const uint32_t ldy, const int act) {
    extern __shared__ float s_w[];                       // [9][CIN][COUT]
    for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) s_w[i] = wpk[i];
    __syncthreads();
    const uint32_t Ho = H / STRIDE, Wo = W / STRIDE;
    const size_t n = (size_t)B * Ho * Wo;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t ox = (uint32_t)(i % Wo), oy = (uint32_t)((i / Wo) % Ho), b = (uint32_t)(i / ((size_t)Wo * Ho));
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; co++) acc[co] = bias ? bias[co] : 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; tap++) {
        const int iy = (int)(oy * STRIDE) + tap / 3 - 1, ix = (int)(ox * STRIDE) + tap % 3 - 1;
        if (iy < 0 || iy >= (int)H || ix < 0 || ix >= (int)W) continue;
        float xv[CIN];
        if (IN_FMT == 0) {
            const bf16* px = reinterpret_cast<const bf16*>(xin) + (((size_t)b * H + iy) * W + ix) * CIN;
            if (CIN % 8 == 0) {
#pragma unroll
                for (int v = 0; v < CIN / 8; v++) {
                    const uint4 r = *reinterpret_cast<const uint4*>(px + v * 8);
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
                    for (int k = 0; k < 4; k++) { const float2 f = __bfloat1622float2(h[k]); xv[v * 8 + 2 * k] = f.x; xv[v * 8 + 2 * k + 1] = f.y; }
                }
            } else {
#pragma unroll
                for (int c = 0; c < CIN; c++) xv[c] = __bfloat162float(px[c]);
            }
        } else {
            const size_t plane = (size_t)H * W, o = (size_t)b * CIN * plane + (size_t)iy * W + ix;
#pragma unroll
            for (int c = 0; c < CIN; c++)
                xv[c] = (IN_FMT == 1) ? reinterpret_cast<const float*>(xin)[o + c * plane]
                                      : __bfloat162float(reinterpret_cast<const bf16*>(xin)[o + c * plane]);
        }
        const float* wt = s_w + tap * CIN * COUT;
#pragma unroll
        for (int c = 0; c < CIN; c++) {
            const float4* w4 = reinterpret_cast<const float4*>(wt + c * COUT);
#pragma unroll
            for (int q = 0; q < COUT / 4; q++) {
                const float4 w = w4[q];
                acc[4 * q] = fmaf(xv[c], w.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv[c], w.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(xv[c], w.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv[c], w.w, acc[4 * q + 3]);
            }
        }
    }
    bf16* out = y + i * ldy;
#pragma unroll
    for (int v = 0; v < COUT / 8; v++) {
        uint4 o;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float a0 = acc[v * 8 + 2 * k], a1 = acc[v * 8 + 2 * k + 1];
            if (act == 1) { a0 = a0 / (1.0f + __expf(-a0)); a1 = a1 / (1.0f + __expf(-a1)); }
            h[k] = __floats2bfloat162_rn(a0, a1);
        }
        *reinterpret_cast<uint4*>(out + v * 8) = o;
    }
}

template <int CIN, int COUT, int STRIDE>
int launch_direct(const void* x, int fmt, const float* w, const float* bias, void* y, uint32_t B, uint32_t H, uint32_t W, uint32_t ldy, int act,
                  cudaStream_t s) {
    const size_t n = (size_t)B * (H / STRIDE) * (W / STRIDE);
    const int smem = 9 * CIN * COUT * (int)sizeof(float);
    const unsigned grid = cdiv(n, 128);
#define MVE_DIRECT(FMT)                                                                                                              \
    {                                                                                                                                \
        if (smem > 48 * 1024) MVE_CUDA(cudaFuncSetAttribute(k_conv3x3_direct<CIN, COUT, STRIDE, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        k_conv3x3_direct<CIN, COUT, STRIDE, FMT><<<grid, 128, smem, s>>>(x, w, bias, (bf16*)y, B, H, W, ldy, act);                   \
    }
    if (fmt == 0) MVE_DIRECT(0) else if (fmt == 1) MVE_DIRECT(1) else MVE_DIRECT(2)
#undef MVE_DIRECT
    MVE_CHECK_LAUNCH("mve_conv3x3_direct_bf16");
    return 0;
}

}  // namespace

extern "C" int mve_conv3x3_direct_bf16(const void* x, int x_format, const float* w_packed, const float* bias, void* y, uint32_t B, uint32_t H,
                                       uint32_t W, uint32_t Cin, uint32_t Cout, uint32_t stride, uint32_t ldy, int act, void* stream) {
    if (B == 0) return 0;
    MVE_ARG(x_format >= 0 && x_format <= 2, "conv3x3_direct: x_format 0 (NHWC bf16), 1 (NCHW f32) or 2 (NCHW bf16)");
    MVE_ARG(stride == 1 || (stride == 2 && H % 2 == 0 && W % 2 == 0), "conv3x3_direct: stride 1, or 2 with even H, W");
    MVE_ARG(ldy >= Cout && ldy % 8 == 0 && (((uintptr_t)y) & 15) == 0, "conv3x3_direct: ldy >= Cout, 16-byte aligned output rows");
    MVE_ARG(x_format != 0 || (((uintptr_t)x) & 15) == 0, "conv3x3_direct: 16-byte aligned NHWC input");
    cudaStream_t s = (cudaStream_t)stream;
#define MVE_CASE(CI, CO, ST) \
    if (Cin == CI && Cout == CO && stride == ST) return launch_direct<CI, CO, ST>(x, x_format, w_packed, bias, y, B, H, W, ldy, act, s);
    MVE_CASE(3, 16, 1) MVE_CASE(16, 16, 1) MVE_CASE(16, 32, 2) MVE_CASE(32, 32, 1) MVE_CASE(32, 96, 2)
    MVE_CASE(3, 8, 1) MVE_CASE(8, 8, 1) MVE_CASE(8, 16, 2) MVE_CASE(16, 32, 1)
#undef MVE_CASE
    mve_set_error("conv3x3_direct: (Cin, Cout, stride) = (%u, %u, %u) is not an instantiated configuration", Cin, Cout, stride);
    return -1;
}

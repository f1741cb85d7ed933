// lpips.cu -- the elementwise half of the perceptual patch loss of the reconstruction objective
// (LPIPSLoss(net='vgg'), /root/reference/lib/models/losses/lpips_loss.py:14-43, called at lib/pipelines/mvedit_3d_pipeline.py:611-617):
// the VGG16 convolutions run on the tcgen05 implicit-GEMM kernel (gemm_tc.cu, ReLU and ReLU-gate epilogues); here are
//   k_lpips_prep        fp32 patches in [0,1] -> normalised (x*2-1, ScalingLayer shift/scale) bf16 NHWC, zero-padded to 64 channels
//   k_maxpool2x2 / _bwd 2x2/2 max pool, and its backward fused with "+= into the gradient of the pre-pool features" and their ReLU gate
//   k_lpips_layer       per feature level: channel-unit-normalise pred and target, squared difference, 1x1 'lin' weights, spatial mean
//                       -> loss, AND its gradient w.r.t. the pred features (ReLU-gated), in one pass, one warp per pixel
//   k_lpips_input_grad  gradient of the padded first-layer input -> fp32 gradient of the [0,1] patch
// All HBM-bound streaming kernels over <= 4 MB feature maps (a 128^2 patch); they exist to keep the iteration at a few dozen
// launches inside the CUDA graph, not for bandwidth.
#include "common.cuh"
#include "../../include/mvedit_b200.h"

namespace {

__constant__ float kShift[3] = {-0.030f, -0.088f, -0.188f};
__constant__ float kScale[3] = {0.458f, 0.448f, 0.450f};

// one thread per pixel: 64 bf16 = 8 x uint4, only the first carries data
__global__ void k_lpips_prep(const float* __restrict__ pred, const float* __restrict__ target, uint32_t n_pix_each,
                             __nv_bfloat16* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n_pix_each) return;
    const float* src = i < n_pix_each ? pred + (size_t)i * 3 : target + (size_t)(i - n_pix_each) * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) v[c] = ((src[c] * 2.f - 1.f) - kShift[c]) / kScale[c];
    uint4 first = make_uint4(0, 0, 0, 0);
    __nv_bfloat16* f = reinterpret_cast<__nv_bfloat16*>(&first);
    f[0] = __float2bfloat16(v[0]); f[1] = __float2bfloat16(v[1]); f[2] = __float2bfloat16(v[2]);
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 64);
    o[0] = first;
#pragma unroll
    for (int k = 1; k < 8; k++) o[k] = make_uint4(0, 0, 0, 0);
}

__global__ void k_lpips_input_grad(const __nv_bfloat16* __restrict__ g64, uint32_t n_pix, float* __restrict__ g_pred) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
#pragma unroll
    for (int c = 0; c < 3; c++) g_pred[(size_t)i * 3 + c] = __bfloat162float(g64[(size_t)i * 64 + c]) * (2.f / kScale[c]);
}

// thread per (output pixel, 8 channels)
__global__ void k_maxpool2x2(const __nv_bfloat16* __restrict__ x, uint32_t B, uint32_t H, uint32_t W, uint32_t C,
                             __nv_bfloat16* __restrict__ y) {
    const uint32_t c8 = C / 8, Ho = H / 2, Wo = W / 2;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * Ho * Wo * c8) return;
    const uint32_t cg = t % c8;
    size_t pix = t / c8;
    const uint32_t xo = pix % Wo, yo = (pix / Wo) % Ho, b = pix / ((size_t)Wo * Ho);
    const __nv_bfloat16* base = x + (((size_t)b * H + 2 * yo) * W + 2 * xo) * C + cg * 8;
    uint4 v[4] = {*reinterpret_cast<const uint4*>(base), *reinterpret_cast<const uint4*>(base + C),
                  *reinterpret_cast<const uint4*>(base + (size_t)W * C), *reinterpret_cast<const uint4*>(base + (size_t)W * C + C)};
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        __nv_bfloat162 m = reinterpret_cast<__nv_bfloat162*>(&v[0])[i];
#pragma unroll
        for (int k = 1; k < 4; k++) m = __hmax2(m, reinterpret_cast<__nv_bfloat162*>(&v[k])[i]);
        o2[i] = m;
    }
    *reinterpret_cast<uint4*>(y + (((size_t)b * Ho + yo) * Wo + xo) * C + cg * 8) = o;
}

// g_x += route(g_y) at the window's first maximum, then gated by x > 0 (x is a ReLU output).  thread per (output pixel, 8 channels);
// every input pixel belongs to exactly one window, so the read-modify-write of g_x is race free.
__global__ void k_maxpool2x2_bwd(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g_y, uint32_t B, uint32_t H,
                                 uint32_t W, uint32_t C, __nv_bfloat16* __restrict__ g_x) {
    const uint32_t c8 = C / 8, Ho = H / 2, Wo = W / 2;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * Ho * Wo * c8) return;
    const uint32_t cg = t % c8;
    size_t pix = t / c8;
    const uint32_t xo = pix % Wo, yo = (pix / Wo) % Ho, b = pix / ((size_t)Wo * Ho);
    const size_t off[4] = {0, C, (size_t)W * C, (size_t)W * C + C};
    const size_t base = (((size_t)b * H + 2 * yo) * W + 2 * xo) * C + cg * 8;
    uint4 xv[4], gv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        xv[k] = *reinterpret_cast<const uint4*>(x + base + off[k]);
        gv[k] = *reinterpret_cast<const uint4*>(g_x + base + off[k]);
    }
    const uint4 gy4 = *reinterpret_cast<const uint4*>(g_y + (((size_t)b * Ho + yo) * Wo + xo) * C + cg * 8);
    const __nv_bfloat16* gy = reinterpret_cast<const __nv_bfloat16*>(&gy4);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float xs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) xs[k] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&xv[k])[i]);
        int am = 0;
#pragma unroll
        for (int k = 1; k < 4; k++) if (xs[k] > xs[am]) am = k;          // first maximum, torch's max_pool2d rule
#pragma unroll
        for (int k = 0; k < 4; k++) {
            __nv_bfloat16* g = reinterpret_cast<__nv_bfloat16*>(&gv[k]) + i;
            const float tot = __bfloat162float(*g) + (k == am ? __bfloat162float(gy[i]) : 0.f);
            *g = __float2bfloat16(xs[k] > 0.f ? tot : 0.f);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) *reinterpret_cast<uint4*>(g_x + base + off[k]) = gv[k];
}

// One warp per pixel of a pred image; lane owns CPL = C / 32 contiguous channels.  feat = [pred images (P) | target images (P)].
//   u = f / (|f| + 1e-10),  val = sum_c w_c (u_p - u_t)_c^2,  loss[img] += val / HW
//   d loss_total / d f_p = gscale[img] / HW * d val / d f_p,   d val / d f_k = q_k / (n + eps) - f_k (q . f) / (n (n + eps)^2),  q = 2 w (u_p - u_t)
// The gradient is written ReLU-gated (f_p > 0) in bf16: it is the gradient w.r.t. the pre-activation of the layer's last conv.
template <int CPL>
__global__ void __launch_bounds__(256) k_lpips_layer(const __nv_bfloat16* __restrict__ feat, uint32_t P, uint32_t HW,
                                                     const float* __restrict__ lin_w, const float* __restrict__ gscale,
                                                     float* __restrict__ loss, __nv_bfloat16* __restrict__ g_feat,
                                                     const uint32_t PIX_PER_WARP) {
    constexpr int C = CPL * 32;
    // a CTA covers 8 * PIX_PER_WARP consecutive pixels of ONE image and issues one loss atomic -- 16 384 same-address atomics, one
    // per pixel, serialise in L2 and cost 27 us on the 128^2 level; the small levels keep one pixel per warp (latency, not atomics)
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, per_block = 8 * PIX_PER_WARP;
    const uint32_t blocks_per_img = (HW + per_block - 1) / per_block, img = blockIdx.x / blocks_per_img;
    const uint32_t pix0 = (blockIdx.x % blocks_per_img) * per_block + wib * PIX_PER_WARP;
    float w[CPL];
#pragma unroll
    for (int j = 0; j < CPL; j++) w[j] = lin_w[lane * CPL + j];
    const float gs = gscale[img] / (float)HW;
    float acc = 0.f;
    for (uint32_t k = 0; k < PIX_PER_WARP; k++) {
        const uint32_t pix = pix0 + k;
        if (pix >= HW) break;
        const size_t row = (size_t)img * HW + pix;
        const __nv_bfloat16* fp = feat + row * C + lane * CPL;
        const __nv_bfloat16* ft = feat + ((size_t)P * HW + row) * C + lane * CPL;
        float p[CPL], t[CPL];
        float sp = 0.f, st = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; j += 2) {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(fp + j));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ft + j));
            p[j] = a.x; p[j + 1] = a.y; t[j] = b.x; t[j + 1] = b.y;
            sp += a.x * a.x + a.y * a.y; st += b.x * b.x + b.y * b.y;
        }
        sp = warp_sum(sp); st = warp_sum(st);
        const float np_ = sqrtf(sp), nt_ = sqrtf(st), ip = 1.f / (np_ + 1e-10f), it = 1.f / (nt_ + 1e-10f);
        float val = 0.f, dot = 0.f, q[CPL];
#pragma unroll
        for (int j = 0; j < CPL; j++) {
            const float d = p[j] * ip - t[j] * it;
            val += w[j] * d * d;
            q[j] = 2.f * w[j] * d;
            dot += q[j] * p[j];
        }
        val = warp_sum(val); dot = warp_sum(dot);
        acc += val;
        const float k2 = np_ > 0.f ? dot * ip * ip / np_ : 0.f;
        __nv_bfloat16* g = g_feat + row * C + lane * CPL;
#pragma unroll
        for (int j = 0; j < CPL; j += 2) {
            const float g0 = p[j] > 0.f ? gs * (q[j] * ip - p[j] * k2) : 0.f;
            const float g1 = p[j + 1] > 0.f ? gs * (q[j + 1] * ip - p[j + 1] * k2) : 0.f;
            *reinterpret_cast<__nv_bfloat162*>(g + j) = __floats2bfloat162_rn(g0, g1);
        }
    }
    __shared__ float s_acc[8];
    if (lane == 0) s_acc[wib] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) tot += s_acc[i];
        atomicAdd(loss + img, tot / (float)HW);
    }
}

}  // namespace

extern "C" {

int mve_lpips_prep(const float* pred, const float* target, uint32_t n_pix_each, void* out, void* stream) {
    if (n_pix_each == 0) return 0;
    k_lpips_prep<<<cdiv(2ull * n_pix_each, 256), 256, 0, (cudaStream_t)stream>>>(pred, target, n_pix_each, (__nv_bfloat16*)out);
    MVE_CHECK_LAUNCH("mve_lpips_prep");
    return 0;
}

int mve_lpips_input_grad(const void* g64, uint32_t n_pix, float* g_pred, void* stream) {
    if (n_pix == 0) return 0;
    k_lpips_input_grad<<<cdiv(n_pix, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)g64, n_pix, g_pred);
    MVE_CHECK_LAUNCH("mve_lpips_input_grad");
    return 0;
}

int mve_maxpool2x2_bf16(const void* x, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* y, void* stream) {
    MVE_ARG(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool2x2: even H, W and C % 8 == 0");
    const size_t n = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    if (n == 0) return 0;
    k_maxpool2x2<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, B, H, W, C, (__nv_bfloat16*)y);
    MVE_CHECK_LAUNCH("mve_maxpool2x2_bf16");
    return 0;
}

int mve_maxpool2x2_relu_backward_bf16(const void* x, const void* g_y, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* g_x,
                                      void* stream) {
    MVE_ARG(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool2x2 backward: even H, W and C % 8 == 0");
    const size_t n = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    if (n == 0) return 0;
    k_maxpool2x2_bwd<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)g_y, B, H, W, C,
                                                                      (__nv_bfloat16*)g_x);
    MVE_CHECK_LAUNCH("mve_maxpool2x2_relu_backward_bf16");
    return 0;
}

int mve_lpips_layer(const void* feat, uint32_t P, uint32_t HW, uint32_t C, const float* lin_w, const float* gscale, float* loss,
                    void* g_feat, void* stream) {
    MVE_ARG(C == 64 || C == 128 || C == 256 || C == 512, "lpips layer: C must be 64, 128, 256 or 512 (VGG16)");
    if (P * HW == 0) return 0;
    uint32_t ppw = (uint32_t)((size_t)P * HW / (8u * (uint32_t)kNumSM));       // >= one CTA per SM before warps take several pixels
    ppw = ppw < 1 ? 1 : (ppw > 8 ? 8 : ppw);
    const uint32_t blocks = P * cdiv(HW, 8 * ppw);
    cudaStream_t s = (cudaStream_t)stream;
    const __nv_bfloat16* f = (const __nv_bfloat16*)feat;
    __nv_bfloat16* g = (__nv_bfloat16*)g_feat;
    if (C == 64) k_lpips_layer<2><<<blocks, 256, 0, s>>>(f, P, HW, lin_w, gscale, loss, g, ppw);
    else if (C == 128) k_lpips_layer<4><<<blocks, 256, 0, s>>>(f, P, HW, lin_w, gscale, loss, g, ppw);
    else if (C == 256) k_lpips_layer<8><<<blocks, 256, 0, s>>>(f, P, HW, lin_w, gscale, loss, g, ppw);
    else k_lpips_layer<16><<<blocks, 256, 0, s>>>(f, P, HW, lin_w, gscale, loss, g, ppw);
    MVE_CHECK_LAUNCH("mve_lpips_layer");
    return 0;
}

}  // extern "C"

// shade.cu -- everything between the fused renderer's raw outputs and the ControlNet inputs of denoise P2, in two launches.
//
// Reference path (per step, on 32 x 512^2 pixels, ~40 elementwise torch kernels and as many full-size temporaries):
//   BaseNeRF.render tail      /root/reference/lib/models/autoencoders/base_nerf.py:536-556  (inverse-z depth, depth / alpha,
//                              depth_to_normal, normal compositing)
//   depth_to_normal            lib/core/utils/geometry_utils.py:119-148 (4 finite-difference crosses with replicate padding)
//   Lambert shading + compose  lib/pipelines/mvedit_3d_pipeline.py:1352-1380
//   normalize_depth            lib/core/utils/geometry_utils.py:151-168 (per-view max / masked min, then affine map)
// Here: k_shade_reduce (per-view depth max and foreground-min, float-as-int atomics: all values are >= 0) and k_shade_apply
// (one thread per pixel: recomputes the 5-point stencil of camera-space points from the raw depth, writes the shaded image
// and the normalised depth as bf16 NCHW -- the layout the denoiser's hint path takes).
#include <cstring>
#include "common.cuh"
#include "tonemap.cuh"
#include "../../include/mvedit_b200.h"

namespace {

struct ShadeParams {
    const float* ws;        // [V,h,w]   accumulated weights (alpha)
    const float* depth;     // [V,h,w]   sum w / t
    const float* image;     // [V,h,w,3] premultiplied rgb
    const float* intr;      // [V,4] fx fy cx cy at the render size
    const float* lights;    // [V,3] camera-space light directions (OpenCV)
    uint32_t V, h, w;
    float ambient, bg;
    float far_depth, alpha_clip, eps;
    int* red;               // [V,2] float bits: max inverse-z depth, min foreground depth
    __nv_bfloat16* out_img; // [V,3,h,w]
    __nv_bfloat16* out_dep; // [V,3,h,w]
    float* out_nrm;         // [V,h,w,3] optional: normal_fg in the opengl [0,1] encoding
    ToneLut tone;           // n == 0: linear shading; else shading in tone-mapped space (mvedit_3d_pipeline.py:1377-1384)
};

__device__ __forceinline__ float dir_norm(const float* K, const uint32_t x, const uint32_t y, float& dx, float& dy) {
    dx = ((float)x + 0.5f - K[2]) / K[0];
    dy = ((float)y + 0.5f - K[3]) / K[1];
    return sqrtf(dx * dx + dy * dy + 1.0f);
}

__global__ void __launch_bounds__(256) k_shade_reduce(const ShadeParams p) {
    const uint32_t v = blockIdx.y;
    const float* K = p.intr + v * 4;
    const size_t base = (size_t)v * p.h * p.w;
    float mx = 0.f, mn = 1.0f / p.eps;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.h * p.w; i += gridDim.x * blockDim.x) {
        float dx, dy;
        const float d = p.depth[base + i] * dir_norm(K, i % p.w, i / p.w, dx, dy);
        const float a = p.ws[base + i];
        mx = fmaxf(mx, d);
        if (!(a < p.alpha_clip)) mn = fminf(mn, d / fmaxf(a, p.eps));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(p.red + v * 2, __float_as_int(mx));
        atomicMin(p.red + v * 2 + 1, __float_as_int(mn));
    }
}

__global__ void k_shade_init(int* red, uint32_t V, int min_bits) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) { red[v * 2] = 0; red[v * 2 + 1] = min_bits; }
}

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 sub(const V3 a, const V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 ncross(const V3 a, const V3 b) {      // F.normalize(torch.cross(a, b)): x / max(|x|, 1e-12)
    const V3 c = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    const float inv = 1.0f / fmaxf(sqrtf(c.x * c.x + c.y * c.y + c.z * c.z), 1e-12f);
    return {c.x * inv, c.y * inv, c.z * inv};
}

__global__ void __launch_bounds__(256) k_shade_apply(const ShadeParams p) {
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), v = blockIdx.z;
    if (x >= p.w || y >= p.h) return;
    const float* K = p.intr + v * 4;
    const size_t base = (size_t)v * p.h * p.w;
    // camera-space point of a pixel: directions / clamp(inverse-z depth of the foreground, 1e-6)
    auto point = [&](const uint32_t px, const uint32_t py) {
        float dx, dy;
        const float nrm = dir_norm(K, px, py, dx, dy);
        const size_t i = base + (size_t)py * p.w + px;
        const float dfg = p.depth[i] * nrm / fmaxf(p.ws[i], 1e-6f);
        const float inv = 1.0f / fmaxf(dfg, 1e-6f);
        return V3{dx * inv, dy * inv, inv};
    };
    const V3 c = point(x, y);
    // replicate-padded finite differences (geometry_utils.py:124-131)
    const uint32_t xr = x + 1 < p.w ? x : x - 1, xl = x > 0 ? x : 1, yd = y + 1 < p.h ? y : y - 1, yu = y > 0 ? y : 1;
    const V3 right = sub(point(xr + 1, y), point(xr, y));
    const V3 left = sub(point(xl - 1, y), point(xl, y));
    const V3 down = sub(point(x, yd + 1), point(x, yd));
    const V3 up = sub(point(x, yu - 1), point(x, yu));
    const V3 n0 = ncross(right, up), n1 = ncross(up, left), n2 = ncross(left, down), n3 = ncross(down, right);
    V3 n = {n0.x + n1.x + n2.x + n3.x, n0.y + n1.y + n2.y + n3.y, n0.z + n1.z + n2.z + n3.z};
    const float inv = 1.0f / fmaxf(sqrtf(n.x * n.x + n.y * n.y + n.z * n.z), 1e-12f);
    n = {n.x * inv, n.y * inv, n.z * inv};
    // opengl [0,1] encoding and back to OpenCV, as the reference round-trips it (base_nerf.py:552, mvedit_3d_pipeline.py:1357)
    const float fx = n.x / 2 + 0.5f, fy = -n.y / 2 + 0.5f, fz = -n.z / 2 + 0.5f;
    const size_t i = base + (size_t)y * p.w + x;
    if (p.out_nrm) { p.out_nrm[i * 3] = fx; p.out_nrm[i * 3 + 1] = fy; p.out_nrm[i * 3 + 2] = fz; }
    if (!p.out_img) return;
    const float ox = fx * 2 - 1, oy = -fy * 2 + 1, oz = -fz * 2 + 1;
    const float* Lg = p.lights + v * 3;
    const float shading = fmaxf(Lg[0] * ox + Lg[1] * oy + Lg[2] * oz, 0.f) * (1 - p.ambient) + p.ambient;
    const float a = p.ws[i];
    const size_t plane = (size_t)p.h * p.w, o = (size_t)v * 3 * plane + (size_t)y * p.w + x;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float val;
        if (p.tone.n)
            val = tone_lut(p.tone, tone_inverse_lut(p.tone, p.image[i * 3 + ch] / fmaxf(a, 1e-6f)) + log2f(fmaxf(shading, 1e-6f))) * a +
                  p.bg * (1 - a);
        else
            val = p.image[i * 3 + ch] * shading + p.bg * (1 - a);
        p.out_img[o + ch * plane] = __float2bfloat16(fminf(fmaxf(__bfloat162float(__float2bfloat16(val)), 0.f), 1.f));
    }
    // normalize_depth
    float dx, dy;
    const float d = p.depth[i] * dir_norm(K, x, y, dx, dy);
    const float dmax = __int_as_float(p.red[v * 2]), dmin = __int_as_float(p.red[v * 2 + 1]);
    float dn = (d / fmaxf(a, p.eps) - dmin) / fmaxf(dmax - dmin, p.eps);
    dn = dn * (1 - p.far_depth) + p.far_depth;
    const __nv_bfloat16 db = __float2bfloat16(fminf(fmaxf(dn * a, 0.f), 1.f));
    p.out_dep[o] = db; p.out_dep[o + plane] = db; p.out_dep[o + 2 * plane] = db;
}

}  // namespace

extern "C" int mve_shade_views(const float* weights_sum, const float* depth, const float* image, const float* intrinsics, const float* lights,
                               uint32_t V, uint32_t h, uint32_t w, float ambient, float bg_color, float far_depth, float alpha_clip, float eps,
                               int32_t* reduce_scratch, void* out_images, void* out_depths, float* out_normals_fg,
                               const float* tonemap_knots, uint32_t tonemap_n, void* stream) {
    if (V == 0) return 0;
    MVE_ARG(h >= 2 && w >= 2, "shade_views: h, w >= 2 required");
    MVE_ARG((out_images == nullptr) == (out_depths == nullptr), "shade_views: out_images and out_depths go together");
    MVE_ARG(out_images != nullptr || out_normals_fg != nullptr, "shade_views: nothing to produce");
    MVE_ARG(out_images == nullptr || (reduce_scratch != nullptr && lights != nullptr), "shade_views: reduce_scratch [V,2] i32 and lights required");
    cudaStream_t s = (cudaStream_t)stream;
    ShadeParams p{weights_sum, depth, image, intrinsics, lights, V, h, w, ambient, bg_color, far_depth, alpha_clip, eps, reduce_scratch,
                  (__nv_bfloat16*)out_images, (__nv_bfloat16*)out_depths, out_normals_fg, {}};
    MVE_ARG(fill_tone_lut(p.tone, tonemap_knots, tonemap_n) == 0, "shade_views: tone curve needs 2..32 knots");
    if (out_images == nullptr) {
        k_shade_apply<<<dim3(cdiv(w, 32), cdiv(h, 8), V), 256, 0, s>>>(p);
        MVE_CHECK_LAUNCH("mve_shade_views");
        return 0;
    }
    const float init_min = 1.0f / eps;             // max starts at 0, the masked min at 1 / eps (geometry_utils.py:158)
    int min_bits;
    memcpy(&min_bits, &init_min, 4);
    k_shade_init<<<cdiv(V, 128), 128, 0, s>>>(reduce_scratch, V, min_bits);
    uint32_t gx = cdiv(h * w, 256 * 8);
    if (gx < 1) gx = 1;
    k_shade_reduce<<<dim3(gx, V), 256, 0, s>>>(p);
    k_shade_apply<<<dim3(cdiv(w, 32), cdiv(h, 8), V), 256, 0, s>>>(p);
    MVE_CHECK_LAUNCH("mve_shade_views");
    return 0;
}


// nerf_loss.cu -- the per-iteration objective of MVEdit3DPipeline.nerf_optim, fused.
//
// The reference builds the loss of one reconstruction iteration out of ~100 eager torch ops on a 128x128 patch
// (/root/reference/lib/pipelines/mvedit_3d_pipeline.py:541-603: depth -> normals (geometry_utils.depth_to_normal :119-148), Lambert
// shading, L1 rgb, L1 alpha, TV-1.5 normal regulariser (lib/models/losses/tv_loss.py:7-42), background entropy) and autograd replays
// ~200 more in the backward.  At 16 384 rays that is ~400 kernel launches of a few microseconds each per iteration -- pure launch
// latency.  Here the same arithmetic is four kernels that produce the loss terms AND the gradients w.r.t. the renderer outputs
// (image, weights_sum, depth) in one go (the loss is a scalar, so "backward" is a scale by the upstream gradient):
//   k_normals    depth -> camera-space points -> 4-neighbour normals (opengl, [0,1]) + 3x3 min-pooled foreground weight
//   k_terms      shading, L1 rgb / alpha, background entropy, TV-1.5 on the normals; writes d(image), part of d(alpha), d(normal)
//   k_normal_bwd d(normal) -> d(points) of the 5-point stencil (atomics into a [N,3] buffer)
//   k_finish     d(points) -> d(depth), d(alpha)
// Optional terms (mve_nerf_patch_loss_targets; every pointer NULL in the text-to-3D recipe): target normals inside the TV term
// (tv_loss.py:27: diff(pred) - diff(target)), the L1 term on 1/z against target depths (:586-592) and the gradient of the high-passed
// normal patch term (:619-626) w.r.t. the alpha-composited normals, chained here to d(normal_fg) and d(alpha).
#include "cuda_host_shim.h"
#include "tonemap.cuh"
#include "mvedit_b200.h"

namespace {

struct LossParams {
    // renderer outputs for P patches of ps x ps rays (ray index = (patch*ps + y)*ps + x)
    const float* image;   // [N,3] premultiplied rgb
    const float* alpha;   // [N]   weights_sum
    const float* depth;   // [N]   sum w / t  (inverse distance along the ray)
    // targets / per-patch data
    const float* tgt_rgb;    // [N,3]
    const float* tgt_mask;   // [N]
    const float* dirs;       // [N,3] un-normalised camera-space directions (z = 1)
    const float* patch_w;    // [P]  cam_weight / mean(cam_weights)
    const float* lights;     // [P,3]
    uint32_t P, ps;
    int shaded;              // apply Lambert shading (not is_init or init_shaded)
    float ambient, bg_color, bg_width;
    // term weights: device scalars (schedule dependent, must stay valid inside a captured graph)
    const float* w_alpha_mul;   // 5.0 on the first call else 1.0
    const float* w_normal_reg;  // normal_reg_weight * 10
    const float* w_entropy;     // entropy_weight
    float pixel_loss_weight;    // L1LossMod.loss_weight (1.2)
    // scratch / outputs
    float* normals;      // [N,3]
    float* fgw;          // [N]
    float* d_normals;    // [N,3]  (zeroed by the launcher)
    float* d_xyz;        // [N,3]  (zeroed by the launcher)
    float* g_image;      // [N,3]
    float* g_alpha;      // [N]
    float* g_depth;      // [N]
    float* loss;         // [5] total, pixel_rgb, alpha, normal_reg, entropy(background part)   (zeroed by the launcher)
    const float* g_out_extra;   // [N,3] or NULL: gradient of further terms w.r.t. the composited / shaded rgb (the LPIPS patch loss)
    float* out_rgb;             // [N,3] (k_out_rgb only): image * shading + bg_color (1 - alpha), what pixel and patch losses see
    ToneLut tone;               // n > 0 (and shaded): out = lut(inverse_lut(image / alpha) + log2(shading)) * alpha + bg (1 - alpha) (:564-570)
    // optional targets (NULL = term absent)
    const float* tgt_normal;       // [N,3] opengl normals in [0,1]: the TV term sees diff(normal_fg) - diff(tgt_normal)
    const float* g_normal_extra;   // [N,3] gradient of further terms w.r.t. out_normal = normal_fg * alpha + normal_bg (1 - alpha)
    float normal_bg[3];
    float* out_normal;             // [N,3] (k_out_normal only)
    const float* tgt_depth;        // [N]   target 1/z
    const float* w_depth;          // device scalar: depth_weight
    float* loss_depth;             // [1] (zeroed by the launcher)
};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
// F.normalize: v / max(|v|, 1e-12)
__device__ __forceinline__ V3 normalize(V3 v, float& len) { len = fmaxf(sqrtf(dot(v, v)), 1e-12f); return v * (1.0f / len); }
// backward of u = v / |v|: dv = (g - u (u.g)) / |v|
__device__ __forceinline__ V3 normalize_bwd(V3 u, float len, V3 g) { return (g - u * dot(u, g)) * (1.0f / len); }

// camera-space point of ray i: dir / max(depth_fg, 1e-6), depth_fg = depth*|dir| / max(alpha, 1e-6)   (mvedit_3d_pipeline.py:544-548)
__device__ __forceinline__ V3 point_of(const LossParams& p, uint32_t i) {
    const V3 d = {p.dirs[i * 3], p.dirs[i * 3 + 1], p.dirs[i * 3 + 2]};
    const float z = p.depth[i] * sqrtf(dot(d, d));
    const float dfg = z / fmaxf(p.alpha[i], 1e-6f);
    return d * (1.0f / fmaxf(dfg, 1e-6f));
}

// the four stencil differences with replicate padding (geometry_utils.py:128-133)
struct Stencil { uint32_t r0, r1, u0, u1, l0, l1, d0, d1; };   // vec = point[*1] - point[*0]
__device__ __forceinline__ Stencil stencil_of(uint32_t base, uint32_t x, uint32_t y, uint32_t ps) {
    Stencil s;
    const uint32_t xr = (x + 1 < ps) ? x : x - 1;            // right: dx[x] (last column replicates dx[W-2])
    s.r0 = base + y * ps + xr; s.r1 = s.r0 + 1;
    const uint32_t xl = (x > 0) ? x - 1 : 0;                 // left = -dx[x-1] (first column replicates -dx[0])
    s.l0 = base + y * ps + xl + 1; s.l1 = base + y * ps + xl;
    const uint32_t yu = (y > 0) ? y - 1 : 0;                 // up = -dy[y-1]
    s.u0 = base + (yu + 1) * ps + x; s.u1 = base + yu * ps + x;
    const uint32_t yd = (y + 1 < ps) ? y : y - 1;            // down = dy[y]
    s.d0 = base + yd * ps + x; s.d1 = s.d0 + ps;
    return s;
}

__global__ void __launch_bounds__(256) k_normals(const LossParams p) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x, N = p.P * p.ps * p.ps;
    if (i >= N) return;
    const uint32_t ps = p.ps, pp = ps * ps, base = (i / pp) * pp, rem = i % pp, y = rem / ps, x = rem % ps;
    const Stencil s = stencil_of(base, x, y, ps);
    const V3 right = point_of(p, s.r1) - point_of(p, s.r0), up = point_of(p, s.u1) - point_of(p, s.u0);
    const V3 left = point_of(p, s.l1) - point_of(p, s.l0), down = point_of(p, s.d1) - point_of(p, s.d0);
    float l;
    V3 n = normalize(cross(right, up), l) + normalize(cross(up, left), l) + normalize(cross(left, down), l) + normalize(cross(down, right), l);
    n = normalize(n, l);
    // opengl flip of y,z then /2 + 0.5
    p.normals[i * 3] = n.x * 0.5f + 0.5f;
    p.normals[i * 3 + 1] = -n.y * 0.5f + 0.5f;
    p.normals[i * 3 + 2] = -n.z * 0.5f + 0.5f;
    // 3x3 min pool of alpha, implicit +inf padding (== -max_pool2d(-a, 3, 1, 1))
    float m = 3.4e38f;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            const int yy = (int)y + dy, xx = (int)x + dx;
            if (yy >= 0 && yy < (int)ps && xx >= 0 && xx < (int)ps) m = fminf(m, p.alpha[base + yy * ps + xx]);
        }
    p.fgw[i] = m;
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(256) k_terms(const LossParams p) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x, N = p.P * p.ps * p.ps;
    float l_pix = 0.f, l_alpha = 0.f, l_tv = 0.f, l_ent = 0.f;
    if (i < N) {
        const uint32_t ps = p.ps, pp = ps * ps, patch = i / pp, base = patch * pp, rem = i % pp, y = rem / ps, x = rem % ps;
        const float w = p.patch_w[patch];
        const float A = p.alpha[i];
        const float n0 = p.normals[i * 3], n1 = p.normals[i * 3 + 1], n2 = p.normals[i * 3 + 2];
        float dn0 = 0.f, dn1 = 0.f, dn2 = 0.f, gA = 0.f;
        // ---- rgb (mvedit_3d_pipeline.py:558-576): mean over N*3 elements, x 1.2 x 4.5
        float s = 1.0f, lcv = 0.f;
        const float lx = p.lights[patch * 3], ly = p.lights[patch * 3 + 1], lz = p.lights[patch * 3 + 2];
        if (p.shaded) {
            lcv = lx * (n0 * 2 - 1) + ly * (-n1 * 2 + 1) + lz * (-n2 * 2 + 1);
            s = fmaxf(lcv, 0.f) * (1 - p.ambient) + p.ambient;
        }
        const float c_pix = p.pixel_loss_weight * 4.5f / (float)(N * 3);
        float ds = 0.f, dsum = 0.f;
        const bool tone = p.shaded && p.tone.n > 0;
        const float Ac = fmaxf(A, 1e-6f), sc = fmaxf(s, 1e-6f), l2s = tone ? log2f(sc) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float C = p.image[i * 3 + c];
            float out, k_inv = 0.f, k_lut = 0.f, t = 0.f;
            if (tone) {
                t = tone_lut(p.tone, tone_inverse_lut(p.tone, C / Ac, &k_inv) + l2s, &k_lut);
                out = t * A + p.bg_color * (1 - A);
            } else {
                out = C * s + p.bg_color * (1 - A);
            }
            const float diff = out - p.tgt_rgb[i * 3 + c];
            l_pix += fabsf(diff) * w * c_pix;
            const float g = sgn(diff) * w * c_pix + (p.g_out_extra ? p.g_out_extra[i * 3 + c] : 0.f);
            if (tone) {
                // out = lut(inv(C / Ac) + log2(sc)) A + bg (1 - A)
                const float gz = g * A * k_lut;                      // d / d (inverse_lut(.) + log2 s)
                p.g_image[i * 3 + c] = gz * k_inv / Ac;
                if (s > 1e-6f) ds += gz / (sc * 0.6931471805599453f);
                gA += g * t;
                if (A > 1e-6f) gA += -gz * k_inv * C / (Ac * Ac);
            } else {
                p.g_image[i * 3 + c] = g * s;
                ds += g * C;
            }
            dsum += g;
        }
        gA += -p.bg_color * dsum;
        if (p.shaded && lcv > 0.f) {
            const float k = ds * (1 - p.ambient);
            dn0 += 2 * lx * k; dn1 += -2 * ly * k; dn2 += -2 * lz * k;
        }
        // ---- alpha (:578-580): mean over N
        const float c_a = p.pixel_loss_weight * p.w_alpha_mul[0] / (float)N;
        const float da = A - p.tgt_mask[i];
        l_alpha = fabsf(da) * w * c_a;
        gA += sgn(da) * w * c_a;
        // ---- background entropy (:597-603): -(1-A)(log max(1-A,1e-6) - log bg_width) * entropy_weight / N
        const float c_e = p.w_entropy[0] / (float)N;
        const float bgw = 1 - A, lb = __logf(p.bg_width);
        const float lg = __logf(fmaxf(bgw, 1e-6f));
        l_ent = -bgw * (lg - lb) * c_e;
        gA += c_e * ((lg - lb) + ((bgw > 1e-6f) ? 1.f : 0.f));
        if (p.g_normal_extra) {   // out_normal = normal_fg A + normal_bg (1 - A)   (:553-554)
            const float e0 = p.g_normal_extra[i * 3], e1 = p.g_normal_extra[i * 3 + 1], e2 = p.g_normal_extra[i * 3 + 2];
            dn0 += e0 * A; dn1 += e1 * A; dn2 += e2 * A;
            gA += e0 * (n0 - p.normal_bg[0]) + e1 * (n1 - p.normal_bg[1]) + e2 * (n2 - p.normal_bg[2]);
        }
        p.g_alpha[i] = gA;
        // ---- TV^1.5 on the normals (tv_loss.py:7-42; NCHW view, dims = (H, W)): mean over P*3*ps*ps
        const float c_tv = p.w_normal_reg[0] / (float)(N * 3);
        const bool has_h = y + 1 < ps, has_w = x + 1 < ps;
        const uint32_t ih = i + ps, iw = i + 1;
        const float wh = has_h ? fminf(p.fgw[i], p.fgw[ih]) : 0.f, ww = has_w ? fminf(p.fgw[i], p.fgw[iw]) : 0.f;
        const float nn[3] = {n0, n1, n2};
        float dnn[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float dh = has_h ? (p.normals[ih * 3 + c] - nn[c]) : 0.f;
            float dw = has_w ? (p.normals[iw * 3 + c] - nn[c]) : 0.f;
            if (p.tgt_normal) {
                const float tn = p.tgt_normal[i * 3 + c];
                if (has_h) dh -= p.tgt_normal[ih * 3 + c] - tn;
                if (has_w) dw -= p.tgt_normal[iw * 3 + c] - tn;
            }
            dh *= wh; dw *= ww;
            const float v = sqrtf(dh * dh + dw * dw);
            l_tv += v * sqrtf(v) * c_tv;
            if (v > 0.f) {
                const float k = 1.5f * c_tv / sqrtf(v);      // d(v^1.5)/d(dh) = 1.5 sqrt(v) dh / v
                const float gh = k * dh * wh, gw_ = k * dw * ww;
                dnn[c] -= gh + gw_;
                if (has_h) atomicAdd(&p.d_normals[ih * 3 + c], gh);
                if (has_w) atomicAdd(&p.d_normals[iw * 3 + c], gw_);
            }
        }
        atomicAdd(&p.d_normals[i * 3], dn0 + dnn[0]);
        atomicAdd(&p.d_normals[i * 3 + 1], dn1 + dnn[1]);
        atomicAdd(&p.d_normals[i * 3 + 2], dn2 + dnn[2]);
    }
    // block reduction of the loss terms
    l_pix = warp_sum(l_pix); l_alpha = warp_sum(l_alpha); l_tv = warp_sum(l_tv); l_ent = warp_sum(l_ent);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&p.loss[1], l_pix); atomicAdd(&p.loss[2], l_alpha); atomicAdd(&p.loss[3], l_tv); atomicAdd(&p.loss[4], l_ent);
        atomicAdd(&p.loss[0], l_pix + l_alpha + l_tv + l_ent);
    }
}

// the rendered rgb the losses compare with the target (mvedit_3d_pipeline.py:558-571): the input of the LPIPS patch term
__global__ void __launch_bounds__(256) k_out_rgb(const LossParams p) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x, N = p.P * p.ps * p.ps;
    if (i >= N) return;
    const uint32_t patch = i / (p.ps * p.ps);
    float s = 1.0f;
    if (p.shaded) {
        const float lcv = p.lights[patch * 3] * (p.normals[i * 3] * 2 - 1) + p.lights[patch * 3 + 1] * (-p.normals[i * 3 + 1] * 2 + 1) +
                          p.lights[patch * 3 + 2] * (-p.normals[i * 3 + 2] * 2 + 1);
        s = fmaxf(lcv, 0.f) * (1 - p.ambient) + p.ambient;
    }
    const float A = p.alpha[i], bg = p.bg_color * (1 - A);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float C = p.image[i * 3 + c];
        p.out_rgb[i * 3 + c] = (p.shaded && p.tone.n > 0)
            ? tone_lut(p.tone, tone_inverse_lut(p.tone, C / fmaxf(A, 1e-6f)) + log2f(fmaxf(s, 1e-6f))) * A + bg
            : C * s + bg;
    }
}

// the alpha-composited normal map the high-passed normal patch term looks at (:553-554)
__global__ void __launch_bounds__(256) k_out_normal(const LossParams p) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x, N = p.P * p.ps * p.ps;
    if (i >= N) return;
    const float A = p.alpha[i];
#pragma unroll
    for (int c = 0; c < 3; c++) p.out_normal[i * 3 + c] = p.normals[i * 3 + c] * A + p.normal_bg[c] * (1 - A);
}

__device__ __forceinline__ void add3(float* dst, uint32_t i, V3 g) {
    atomicAdd(&dst[i * 3], g.x); atomicAdd(&dst[i * 3 + 1], g.y); atomicAdd(&dst[i * 3 + 2], g.z);
}

__global__ void __launch_bounds__(256) k_normal_bwd(const LossParams p) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x, N = p.P * p.ps * p.ps;
    if (i >= N) return;
    // d(normal in [0,1], opengl) -> d(n)
    V3 g = {p.d_normals[i * 3] * 0.5f, -p.d_normals[i * 3 + 1] * 0.5f, -p.d_normals[i * 3 + 2] * 0.5f};
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f) return;
    const uint32_t ps = p.ps, pp = ps * ps, base = (i / pp) * pp, rem = i % pp, y = rem / ps, x = rem % ps;
    const Stencil s = stencil_of(base, x, y, ps);
    const V3 a[4] = {point_of(p, s.r1) - point_of(p, s.r0), point_of(p, s.u1) - point_of(p, s.u0), point_of(p, s.l1) - point_of(p, s.l0),
                     point_of(p, s.d1) - point_of(p, s.d0)};   // right, up, left, down
    V3 u[4];
    float len[4];
    V3 sum = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; k++) { u[k] = normalize(cross(a[k], a[(k + 1) & 3]), len[k]); sum = sum + u[k]; }
    float ls;
    const V3 n = normalize(sum, ls);
    const V3 gsum = normalize_bwd(n, ls, g);
    V3 da[4] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const V3 gc = normalize_bwd(u[k], len[k], gsum);          // d cross(a_k, a_{k+1})
        da[k] = da[k] + cross(a[(k + 1) & 3], gc);                // c = a x b: da = b x g
        da[(k + 1) & 3] = da[(k + 1) & 3] + cross(gc, a[k]);      //            db = g x a
    }
    const uint32_t i1[4] = {s.r1, s.u1, s.l1, s.d1}, i0[4] = {s.r0, s.u0, s.l0, s.d0};
#pragma unroll
    for (int k = 0; k < 4; k++) { add3(p.d_xyz, i1[k], da[k]); add3(p.d_xyz, i0[k], neg(da[k])); }
}

__global__ void __launch_bounds__(256) k_finish(const LossParams p) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x, N = p.P * p.ps * p.ps;
    if (i >= N) return;
    const V3 d = {p.dirs[i * 3], p.dirs[i * 3 + 1], p.dirs[i * 3 + 2]};
    const V3 gx = {p.d_xyz[i * 3], p.d_xyz[i * 3 + 1], p.d_xyz[i * 3 + 2]};
    const float nd = sqrtf(dot(d, d));
    const float A = p.alpha[i], Ac = fmaxf(A, 1e-6f);
    const float z = p.depth[i] * nd;
    const float dfg = z / Ac;
    float g_depth = 0.f, gA = 0.f;
    if (dfg > 1e-6f) {
        const float inv = 1.0f / dfg;
        const float d_dfg = -dot(gx, d) * inv * inv;       // xyz = dir / dfg
        g_depth = d_dfg / Ac * nd;
        if (A > 1e-6f) gA = -d_dfg * z / (Ac * Ac);
    }
    if (p.tgt_depth) {            // L1 on 1/z (:586-592): mean over N, x loss_weight x depth_weight
        const float c_d = p.pixel_loss_weight * p.w_depth[0] / (float)N * p.patch_w[i / (p.ps * p.ps)];
        const float dd = z - p.tgt_depth[i];
        const float l = fabsf(dd) * c_d;
        g_depth += sgn(dd) * c_d * nd;
        atomicAdd(&p.loss_depth[0], l);
        atomicAdd(&p.loss[0], l);
    }
    p.g_depth[i] = g_depth;
    p.g_alpha[i] += gA;
}

}  // namespace

extern "C" {

uint32_t mve_nerf_patch_loss_scratch_floats(uint32_t n_rays) { return n_rays * 10u; }

int mve_nerf_patch_loss_targets(const float* image, const float* alpha, const float* depth, const float* tgt_rgb, const float* tgt_mask,
                                const float* dirs, const float* patch_w, const float* lights, uint32_t n_patches, uint32_t patch_size,
                                int shaded, float ambient, float bg_color, float bg_width, float pixel_loss_weight,
                                const float* w_alpha_mul, const float* w_normal_reg, const float* w_entropy, float* scratch, float* g_image,
                                float* g_alpha, float* g_depth, float* loss5, const float* g_out_extra, const float* tonemap_knots,
                                uint32_t tonemap_n, const float* tgt_normal, const float* g_normal_extra, float normal_bg_x,
                                float normal_bg_y, float normal_bg_z, const float* tgt_depth, const float* w_depth, float* loss_depth,
                                void* stream) {
    const uint32_t N = n_patches * patch_size * patch_size;
    if (N == 0) return 0;
    MVE_ARG(patch_size >= 2, "nerf_patch_loss: patch_size must be >= 2");
    MVE_ARG(tgt_depth == nullptr || (w_depth != nullptr && loss_depth != nullptr), "nerf_patch_loss: tgt_depth needs w_depth and loss_depth");
    cudaStream_t s = (cudaStream_t)stream;
    LossParams p{};
    p.image = image; p.alpha = alpha; p.depth = depth; p.tgt_rgb = tgt_rgb; p.tgt_mask = tgt_mask; p.dirs = dirs; p.patch_w = patch_w;
    p.lights = lights; p.P = n_patches; p.ps = patch_size; p.shaded = shaded; p.ambient = ambient; p.bg_color = bg_color; p.bg_width = bg_width;
    p.w_alpha_mul = w_alpha_mul; p.w_normal_reg = w_normal_reg; p.w_entropy = w_entropy; p.pixel_loss_weight = pixel_loss_weight;
    p.normals = scratch; p.fgw = scratch + (size_t)N * 3; p.d_normals = scratch + (size_t)N * 4; p.d_xyz = scratch + (size_t)N * 7;
    p.g_image = g_image; p.g_alpha = g_alpha; p.g_depth = g_depth; p.loss = loss5; p.g_out_extra = g_out_extra;
    p.tgt_normal = tgt_normal; p.g_normal_extra = g_normal_extra; p.normal_bg[0] = normal_bg_x; p.normal_bg[1] = normal_bg_y;
    p.normal_bg[2] = normal_bg_z; p.tgt_depth = tgt_depth; p.w_depth = w_depth; p.loss_depth = loss_depth;
    MVE_ARG(fill_tone_lut(p.tone, tonemap_knots, tonemap_n) == 0, "nerf_patch_loss: tone curve needs 2..32 knots");
    MVE_CUDA(cudaMemsetAsync(p.d_normals, 0, (size_t)N * 6 * sizeof(float), s));
    MVE_CUDA(cudaMemsetAsync(loss5, 0, 5 * sizeof(float), s));
    if (loss_depth) MVE_CUDA(cudaMemsetAsync(loss_depth, 0, sizeof(float), s));
    const uint32_t grid = cdiv(N, 256);
    SHIM_LAUNCH(k_normals, grid, 256, p);
    SHIM_LAUNCH(k_terms, grid, 256, p);
    SHIM_LAUNCH(k_normal_bwd, grid, 256, p);
    SHIM_LAUNCH(k_finish, grid, 256, p);
    MVE_CHECK_LAUNCH("mve_nerf_patch_loss");
    return 0;
}

int mve_nerf_patch_loss(const float* image, const float* alpha, const float* depth, const float* tgt_rgb, const float* tgt_mask,
                        const float* dirs, const float* patch_w, const float* lights, uint32_t n_patches, uint32_t patch_size, int shaded,
                        float ambient, float bg_color, float bg_width, float pixel_loss_weight, const float* w_alpha_mul,
                        const float* w_normal_reg, const float* w_entropy, float* scratch, float* g_image, float* g_alpha, float* g_depth,
                        float* loss5, const float* g_out_extra, const float* tonemap_knots, uint32_t tonemap_n, void* stream) {
    return mve_nerf_patch_loss_targets(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, n_patches, patch_size, shaded, ambient,
                                       bg_color, bg_width, pixel_loss_weight, w_alpha_mul, w_normal_reg, w_entropy, scratch, g_image, g_alpha,
                                       g_depth, loss5, g_out_extra, tonemap_knots, tonemap_n, nullptr, nullptr, 0.5f, 0.5f, 1.0f, nullptr,
                                       nullptr, nullptr, stream);
}

int mve_nerf_patch_out_normal(const float* alpha, const float* depth, const float* dirs, uint32_t n_patches, uint32_t patch_size,
                              float normal_bg_x, float normal_bg_y, float normal_bg_z, float* scratch, float* out_normal, void* stream) {
    const uint32_t N = n_patches * patch_size * patch_size;
    if (N == 0) return 0;
    MVE_ARG(patch_size >= 2, "nerf_patch_out_normal: patch_size must be >= 2");
    cudaStream_t s = (cudaStream_t)stream;
    LossParams p{};
    p.alpha = alpha; p.depth = depth; p.dirs = dirs; p.P = n_patches; p.ps = patch_size;
    p.normals = scratch; p.fgw = scratch + (size_t)N * 3; p.out_normal = out_normal;
    p.normal_bg[0] = normal_bg_x; p.normal_bg[1] = normal_bg_y; p.normal_bg[2] = normal_bg_z;
    const uint32_t grid = cdiv(N, 256);
    SHIM_LAUNCH(k_normals, grid, 256, p);
    SHIM_LAUNCH(k_out_normal, grid, 256, p);
    MVE_CHECK_LAUNCH("mve_nerf_patch_out_normal");
    return 0;
}

int mve_nerf_patch_out_rgb(const float* image, const float* alpha, const float* depth, const float* dirs, const float* lights,
                           uint32_t n_patches, uint32_t patch_size, int shaded, float ambient, float bg_color, float* scratch,
                           float* out_rgb, const float* tonemap_knots, uint32_t tonemap_n, void* stream) {
    const uint32_t N = n_patches * patch_size * patch_size;
    if (N == 0) return 0;
    MVE_ARG(patch_size >= 2, "nerf_patch_out_rgb: patch_size must be >= 2");
    cudaStream_t s = (cudaStream_t)stream;
    LossParams p{};
    p.image = image; p.alpha = alpha; p.depth = depth; p.dirs = dirs; p.lights = lights; p.P = n_patches; p.ps = patch_size;
    p.shaded = shaded; p.ambient = ambient; p.bg_color = bg_color;
    p.normals = scratch; p.fgw = scratch + (size_t)N * 3; p.out_rgb = out_rgb;
    MVE_ARG(fill_tone_lut(p.tone, tonemap_knots, tonemap_n) == 0, "nerf_patch_out_rgb: tone curve needs 2..32 knots");
    const uint32_t grid = cdiv(N, 256);
    if (shaded) SHIM_LAUNCH(k_normals, grid, 256, p);
    SHIM_LAUNCH(k_out_rgb, grid, 256, p);
    MVE_CHECK_LAUNCH("mve_nerf_patch_out_rgb");
    return 0;
}

}  // extern "C"

// recon.cu -- the host-side glue of one MVEdit3DPipeline.nerf_optim iteration, as two kernels.
//
// The reference spends most of a reconstruction iteration (16 384 rays) in launch latency around its kernels:
//   * BaseNeRF.ray_sample (/root/reference/lib/models/autoencoders/base_nerf.py:245-303) reshuffles ALL V*h*w rays / pixels into
//     patch-major order and indexes the drawn patch; get_ray_directions / get_rays (lib/core/utils/geometry_utils.py:18-55)
//     materialise two (1,V,h,w,3) tensors per nerf_optim call; per-patch weights, lights and dt_gamma are ~10 more tiny ops
//     (lib/pipelines/mvedit_3d_pipeline.py:516-536);
//   * optimizer.zero_grad / torch.optim.Adam.step (:631-633) run ~15 elementwise passes over the 28.7 MB hash table.
// Here:
//   k_patch_rays   one launch: for the drawn patches, ray origins / directions (optionally only a row strip of every patch -- the
//                  data-parallel reconstruction), camera-space directions, target rgb / mask, per-patch weight, light and the
//                  iteration's dt_gamma, straight from (patch ids, poses, intrinsics, images);
//   k_adam_multi   one launch over all parameter tensors: Adam update (torch.optim.Adam defaults: no weight decay, no amsgrad)
//                  AND zeroing of the gradient it just consumed (one read of g, m, v, p and one write of each per step).
#include "cuda_host_shim.h"
#include "mvedit_b200.h"

namespace {

struct PatchParams {
    const long long* inds;   // [P] patch ids, numbered (view, patch row, patch col)
    uint32_t P, V, rs, ps;
    const float* R;          // [V,3,3] camera-to-world rotation
    const float* T;          // [V,3]   camera centre
    const float* K;          // [V,4]   fx fy cx cy at intrinsics_size
    float k_scale;           // render_size / intrinsics_size
    const float* images;     // [V,rs,rs,3]
    const float* masks;      // [V,rs,rs]
    const float* cam_w;      // [V]
    const float* lights;     // [V,3]
    float dt_gamma_scale;
    uint32_t row_lo, row_hi; // rays are emitted for patch rows [row_lo, row_hi)
    float* rays_o; float* rays_d;                    // [P*(row_hi-row_lo)*ps, 3]
    float* dirs; float* tgt_rgb; float* tgt_mask;    // [P*ps*ps, 3|3|1]
    float* patch_w; float* patch_l; float* dt_gamma; // [P], [P,3], [1]
};

__global__ void __launch_bounds__(256) k_patch_rays(const PatchParams p) {
    const uint32_t pp = p.ps * p.ps, n = p.P * pp;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const uint32_t npr = p.rs / p.ps;                                   // patches per image row / column
    if (blockIdx.x == 0 && threadIdx.x < p.P) {
        // per-patch scalars: weight = cam_w[v] / mean(cam_w) (:516,:574), light (:531), dt_gamma of the FIRST patch (:535-536 + one scene)
        float mean = 0.f;
        for (uint32_t v = 0; v < p.V; v++) mean += p.cam_w[v];
        mean /= (float)p.V;
        const uint32_t v = (uint32_t)(p.inds[threadIdx.x] / (npr * npr));
        p.patch_w[threadIdx.x] = p.cam_w[v] / mean;
        p.patch_l[threadIdx.x * 3] = p.lights[v * 3]; p.patch_l[threadIdx.x * 3 + 1] = p.lights[v * 3 + 1];
        p.patch_l[threadIdx.x * 3 + 2] = p.lights[v * 3 + 2];
        if (threadIdx.x == 0) p.dt_gamma[0] = p.dt_gamma_scale / ((p.K[v * 4] + p.K[v * 4 + 1]) * 0.5f * p.k_scale);
    }
    if (i >= n) return;
    const uint32_t patch = i / pp, rem = i % pp, py = rem / p.ps, px = rem % p.ps;
    const uint32_t id = (uint32_t)p.inds[patch];
    const uint32_t v = id / (npr * npr), pr = (id / npr) % npr, pc = id % npr;
    const uint32_t y = pr * p.ps + py, x = pc * p.ps + px;
    const float fx = p.K[v * 4] * p.k_scale, fy = p.K[v * 4 + 1] * p.k_scale, cx = p.K[v * 4 + 2] * p.k_scale, cy = p.K[v * 4 + 3] * p.k_scale;
    const float dx = ((float)x + 0.5f - cx) / fx, dy = ((float)y + 0.5f - cy) / fy;
    p.dirs[i * 3] = dx; p.dirs[i * 3 + 1] = dy; p.dirs[i * 3 + 2] = 1.0f;
    const size_t pix = ((size_t)v * p.rs + y) * p.rs + x;
    p.tgt_rgb[i * 3] = p.images[pix * 3]; p.tgt_rgb[i * 3 + 1] = p.images[pix * 3 + 1]; p.tgt_rgb[i * 3 + 2] = p.images[pix * 3 + 2];
    p.tgt_mask[i] = p.masks[pix];
    if (py >= p.row_lo && py < p.row_hi) {
        const uint32_t j = (patch * (p.row_hi - p.row_lo) + (py - p.row_lo)) * p.ps + px;
        const float* R = p.R + v * 9;
        float wx = R[0] * dx + R[1] * dy + R[2], wy = R[3] * dx + R[4] * dy + R[5], wz = R[6] * dx + R[7] * dy + R[8];
        const float inv = 1.0f / fmaxf(sqrtf(wx * wx + wy * wy + wz * wz), 1e-12f);      // F.normalize
        p.rays_d[j * 3] = wx * inv; p.rays_d[j * 3 + 1] = wy * inv; p.rays_d[j * 3 + 2] = wz * inv;
        p.rays_o[j * 3] = p.T[v * 3]; p.rays_o[j * 3 + 1] = p.T[v * 3 + 1]; p.rays_o[j * 3 + 2] = p.T[v * 3 + 2];
    }
}

constexpr int ADAM_MAX_SEG = 16;
struct AdamParams {
    float* p[ADAM_MAX_SEG]; float* g[ADAM_MAX_SEG]; float* m[ADAM_MAX_SEG]; float* v[ADAM_MAX_SEG];
    const float* lr[ADAM_MAX_SEG];
    uint32_t n[ADAM_MAX_SEG], first_block[ADAM_MAX_SEG + 1];
    uint32_t n_seg;
    float beta1, beta2, eps, grad_scale;
    const int* step;      // device step counter, already incremented for this update
    int zero_grad;
};

__global__ void k_adam_tick(int* step) { step[0] += 1; }

constexpr int ADAM_T = 256, ADAM_V = 4;   // 256 threads x 4 float4 per thread per block
__global__ void __launch_bounds__(ADAM_T) k_adam_multi(const AdamParams a) {
    uint32_t s = 0;
    while (s + 1 < a.n_seg && blockIdx.x >= a.first_block[s + 1]) s++;
    const uint32_t n = a.n[s];
    float* __restrict__ P = a.p[s]; float* __restrict__ G = a.g[s]; float* __restrict__ M = a.m[s]; float* __restrict__ Vv = a.v[s];
    const float t = (float)a.step[0];
    const float bc1 = 1.0f - powf(a.beta1, t), bc2 = 1.0f - powf(a.beta2, t);
    const float step_size = a.lr[s][0] / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const uint32_t base = (blockIdx.x - a.first_block[s]) * (ADAM_T * ADAM_V * 4);
    const bool vec_ok = ((((uintptr_t)P | (uintptr_t)G | (uintptr_t)M | (uintptr_t)Vv) & 15) == 0);
#pragma unroll
    for (int k = 0; k < ADAM_V; k++) {
        const uint32_t i = base + (k * ADAM_T + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec_ok && i + 4 <= n) {
            float4 g = *reinterpret_cast<const float4*>(G + i), m = *reinterpret_cast<const float4*>(M + i);
            float4 v = *reinterpret_cast<const float4*>(Vv + i), p = *reinterpret_cast<const float4*>(P + i);
            float* gp = &g.x; float* mp = &m.x; float* vp = &v.x; float* pp = &p.x;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float gg = gp[c] * a.grad_scale;
                mp[c] = a.beta1 * mp[c] + (1.0f - a.beta1) * gg;                    // exp_avg.lerp_(grad, 1 - beta1)
                vp[c] = a.beta2 * vp[c] + (1.0f - a.beta2) * gg * gg;               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
                pp[c] -= step_size * mp[c] / (sqrtf(vp[c]) * inv_sqrt_bc2 + a.eps); // param.addcdiv_(exp_avg, denom, -step_size)
            }
            *reinterpret_cast<float4*>(M + i) = m; *reinterpret_cast<float4*>(Vv + i) = v; *reinterpret_cast<float4*>(P + i) = p;
            if (a.zero_grad) *reinterpret_cast<float4*>(G + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (uint32_t j = i; j < n && j < i + 4; j++) {
                const float gg = G[j] * a.grad_scale;
                const float m = a.beta1 * M[j] + (1.0f - a.beta1) * gg, v = a.beta2 * Vv[j] + (1.0f - a.beta2) * gg * gg;
                M[j] = m; Vv[j] = v;
                P[j] -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
                if (a.zero_grad) G[j] = 0.f;
            }
        }
    }
}

}  // namespace

extern "C" {

int mve_patch_rays(const int64_t* patch_inds, uint32_t P, uint32_t V, uint32_t render_size, uint32_t patch_size, const float* poses_R,
                   const float* poses_T, const float* intrinsics, float intrinsics_scale, const float* images, const float* masks,
                   const float* cam_weights, const float* cam_lights, float dt_gamma_scale, uint32_t row_lo, uint32_t row_hi,
                   float* rays_o, float* rays_d, float* dirs, float* tgt_rgb, float* tgt_mask, float* patch_w, float* patch_lights,
                   float* dt_gamma, void* stream) {
    if (P == 0) return 0;
    MVE_ARG(P <= 256, "patch_rays: at most 256 patches per iteration");
    MVE_ARG(patch_size > 0 && render_size % patch_size == 0, "patch_rays: render_size must be a multiple of patch_size");
    MVE_ARG(row_lo < row_hi && row_hi <= patch_size, "patch_rays: need 0 <= row_lo < row_hi <= patch_size");
    PatchParams p{(const long long*)patch_inds, P, V, render_size, patch_size, poses_R, poses_T, intrinsics, intrinsics_scale, images, masks,
                  cam_weights, cam_lights, dt_gamma_scale, row_lo, row_hi, rays_o, rays_d, dirs, tgt_rgb, tgt_mask, patch_w, patch_lights,
                  dt_gamma};
    SHIM_LAUNCH(k_patch_rays, cdiv((unsigned long long)P * patch_size * patch_size, 256), 256, p);
    MVE_CHECK_LAUNCH("mve_patch_rays");
    return 0;
}

int mve_adam_step(uint32_t n_tensors, void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                  const uint32_t* numel, const float* const* lr, float beta1, float beta2, float eps, float grad_scale, int32_t* step,
                  int zero_grad, void* stream) {
    if (n_tensors == 0) return 0;
    MVE_ARG(n_tensors <= (uint32_t)ADAM_MAX_SEG, "adam_step: at most 16 tensors per launch");
    AdamParams a{};
    uint32_t blocks = 0;
    for (uint32_t s = 0; s < n_tensors; s++) {
        a.p[s] = (float*)params[s]; a.g[s] = (float*)grads[s]; a.m[s] = (float*)exp_avg[s]; a.v[s] = (float*)exp_avg_sq[s];
        a.lr[s] = lr[s]; a.n[s] = numel[s];
        a.first_block[s] = blocks;
        blocks += cdiv(numel[s], ADAM_T * ADAM_V * 4);
    }
    a.first_block[n_tensors] = blocks;
    a.n_seg = n_tensors; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale; a.step = step; a.zero_grad = zero_grad;
    cudaStream_t st = (cudaStream_t)stream;
    SHIM_LAUNCH(k_adam_tick, 1, 1, step);
    if (blocks) SHIM_LAUNCH(k_adam_multi, blocks, ADAM_T, a);
    MVE_CHECK_LAUNCH("mve_adam_step");
    return 0;
}

}  // extern "C"

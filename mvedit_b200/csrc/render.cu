// render.cu -- fused kernels of the NeRF adapter's render / reconstruct loop.
//
// k_render_rays: the whole inference branch of VolumeRenderer.forward
//   (/root/reference/lib/models/decoders/base_volume_renderer.py:264-329: a Python while-loop of march_rays ->
//   point_decode -> composite_rays -> boolean compaction -> host sync, up to max_steps/n_step rounds) in ONE launch:
//   one thread per ray keeps (t, weight_sum, depth, rgb) in registers, walks the occupancy grid with the shared DDA,
//   evaluates the hash-grid + MLP field in place and stops at T < T_thresh.  No sample ever touches HBM.
//   Ray generation (geometry_utils.get_ray_directions/get_rays, :18-55) is fused in: rays come from (pose, intrinsics,
//   pixel) so the [N,H,W,3] origin/direction tensors of base_nerf.py:489-556 are never materialised (optional).
//
// k_cull_compact: the weight-culling step of the training branch (:222-246: boolean masks, split, cumsum, index):
//   keeps samples with weight > th, warp-per-ray ballot compaction, CTA scan + one atomic for the new offsets.
#include "march_device.cuh"
#include "field_device.cuh"
#include "mlp_mma.cuh"
#include "../../include/mvedit_b200.h"

using namespace march;
using namespace field;

namespace {

__device__ __forceinline__ void slab(const Ray& r, const float* __restrict__ aabb, const float min_near, float& near_o, float& far_o) {
    // raymarching.cu:110-144
    constexpr float kMax = 3.402823466e+38f;
    float near = (aabb[0] - r.ox) * r.rdx, far = (aabb[3] - r.ox) * r.rdx, tmp;
    if (near > far) { tmp = near; near = far; far = tmp; }
    float near_y = (aabb[1] - r.oy) * r.rdy, far_y = (aabb[4] - r.oy) * r.rdy;
    if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
    if (near > far_y || near_y > far) { near_o = far_o = kMax; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = (aabb[2] - r.oz) * r.rdz, far_z = (aabb[5] - r.oz) * r.rdz;
    if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
    if (near > far_z || near_z > far) { near_o = far_o = kMax; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    near_o = near; far_o = far;
}

struct RenderParams {
    const float* rays_o;   // [N,3] or null when cameras are given
    const float* rays_d;
    // camera mode: ray n -> view n / (h*w), pixel (n % (h*w)); direction = R * ((i+0.5-cx)/fx, (j+0.5-cy)/fy, 1) normalised
    const float* poses;       // [V,4,4] c2w row-major
    const float* intrinsics;  // [V,4] fx fy cx cy (already scaled to the render size)
    uint32_t h, w;
    uint32_t N;
    const float* aabb;
    float min_near, T_thresh;
    uint32_t max_steps;
    float* weights_sum; float* depth; float* image;   // [N], [N], [N,3]
    const float* dt_gamma_per_view;  // [V] or null (then march.dt_gamma is used)
};

// Round-based lane scheduling.  Every round (all steps converged except the bounded DDA loop):
//   1. lanes without a ray take the next slots of the warp's current 8x8 pixel tile (one atomic per 64-ray tile, not per ray),
//   2. every lane without a sample advances its DDA by at most DDA_BUDGET voxel steps (retiring the ray if it ends); the loop
//      stops as soon as no lane is still searching, so a warp full of surface rays pays one DDA step per round,
//   3. the lanes that hold a sample shade it together (96 hash-grid gathers + MLP + compositing) -- the expensive part stays converged.
// A ray crossing empty space therefore never stalls its neighbours' shading for more than DDA_BUDGET steps, and finished /
// missing / early-terminated rays are replaced immediately.
constexpr int DDA_BUDGET = 8;
constexpr int CHUNK = 64;

template <int L, bool LEAN>
__global__ void __launch_bounds__(128, 5) k_render_rays(const RenderParams rp, MarchParams mp, const float2* __restrict__ table,
                                                     const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                                     const float* __restrict__ b2, const Levels lv, const FieldCfg cfg,
                                                     unsigned int* __restrict__ next_ray, unsigned long long* __restrict__ stats) {
    using R = Rec<L>;
    using MC = mlpmma::Cfg<L>;
    __shared__ __align__(16) float frags[MC::FRAG_FLOATS];
    __shared__ __align__(16) float stage[4][MC::STAGE_FLOATS];
    __shared__ uint32_t s_lut[LEAN ? 256 : 1];      // Morton bit expansion of 0..255 (dda_step_lean)
    if (LEAN)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = expand_bits((uint32_t)i);
    mlpmma::stage_frags<L>(frags, w1, b1, w2);
    mlpmma::zero_stage_pad<L>(stage[threadIdx.x >> 5]);
    __syncthreads();
    const float ob0 = b2[0], ob1 = b2[1], ob2 = b2[2], ob3 = b2[3];
    const int lane = threadIdx.x & 31;

    bool have_ray = false;
    bool exhausted = false;      // warp-uniform
    // Rays are handed out in chunks of CHUNK consecutive slots per warp; in camera mode a chunk is an 8x8 pixel tile, so the 32
    // lanes of a warp march a narrow frustum in near lock-step and their hash-grid gathers share cache lines on the coarse
    // levels (a single global ray counter leaves every lane on an unrelated ray: 20 L1 wavefronts per gather request,
    // profiles/r01_ncu_k_render_rays.txt).
    uint32_t chunk_cur = 0, chunk_end = 0;      // warp-uniform
    const bool tiled = !rp.rays_o && (rp.h % 8 == 0) && (rp.w % 8 == 0);
    const uint32_t tiles_x = rp.w / 8, tpv = tiles_x * (rp.h / 8);
    uint32_t n = 0, step = 0, shaded = 0, shade_rounds = 0, rounds = 0, dda_trips = 0;
    Ray r;
    RayAux aux = 0;
    float t = 0.f, far = 0.f, ws = 0.f, dsum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    bool terminated = false;
    float cx = 0.f, cy = 0.f, cz = 0.f, dt = 0.f;

    for (;;) {
        // ---------------- 1. fetch: lanes take consecutive slots of the warp's current chunk; an empty chunk is replaced with ONE atomic
#pragma unroll 1
        for (int rep = 0; rep < 2; rep++) {
            const uint32_t need = __ballot_sync(0xffffffffu, !have_ray);
            if (!need || (exhausted && chunk_cur == chunk_end)) break;
            if (chunk_cur == chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(next_ray, (unsigned int)CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= rp.N) { exhausted = true; break; }
                chunk_cur = base; chunk_end = min(base + (uint32_t)CHUNK, rp.N);
            }
            const uint32_t rank = __popc(need & ((1u << lane) - 1u));
            const uint32_t avail = chunk_end - chunk_cur;
            if (!have_ray && rank < avail) {
                const uint32_t slot = chunk_cur + rank;
                if (rp.rays_o) {
                    n = slot;
                    r = load_ray(rp.rays_o, rp.rays_d, n);
                } else {
                    uint32_t v, px, py;
                    if (tiled) {            // slot -> 8x8 pixel tile, row-major inside the tile
                        const uint32_t tile = slot / CHUNK, j = slot % CHUNK, tt = tile % tpv, ty = tt / tiles_x, tx = tt % tiles_x;
                        v = tile / tpv; py = ty * 8 + (j >> 3); px = tx * 8 + (j & 7);
                    } else {
                        const uint32_t hw = rp.h * rp.w, pix = slot % hw;
                        v = slot / hw; py = pix / rp.w; px = pix % rp.w;
                    }
                    n = v * rp.h * rp.w + py * rp.w + px;
                    const float pi = (float)px + 0.5f, pj = (float)py + 0.5f;
                    const float* K = rp.intrinsics + v * 4;
                    const float* P = rp.poses + v * 16;
                    const float cxd = (pi - K[2]) / K[0], cyd = (pj - K[3]) / K[1];
                    const float dx = P[0] * cxd + P[1] * cyd + P[2], dy = P[4] * cxd + P[5] * cyd + P[6], dz = P[8] * cxd + P[9] * cyd + P[10];
                    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
                    r.ox = P[3]; r.oy = P[7]; r.oz = P[11];
                    r.dx = dx * inv; r.dy = dy * inv; r.dz = dz * inv;
                    r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
                    if (rp.dt_gamma_per_view) mp.dt_gamma = rp.dt_gamma_per_view[v];
                }
                if (LEAN) aux = make_aux(r);
                float near;
                slab(r, rp.aabb, rp.min_near, near, far);
                t = near;
                ws = 0.f; dsum = 0.f; cr = 0.f; cg = 0.f; cb = 0.f; step = 0; terminated = false;
                have_ray = true;
            }
            chunk_cur += min(avail, (uint32_t)__popc(need));
        }
        if (!__any_sync(0xffffffffu, have_ray)) break;     // nothing in flight and nothing left to fetch
        rounds++;
        // ---------------- 2. bounded search for the next sample
        bool has = false;
#pragma unroll 1
        for (int k = 0; k < DDA_BUDGET; k++) {
            const bool searching = have_ray && !has;
            if (!__any_sync(0xffffffffu, searching)) break;
            dda_trips++;
            if (searching) {
                if (!terminated && t < far && step < rp.max_steps) {
                    has = LEAN ? dda_step_lean(r, aux, mp, s_lut, t, cx, cy, cz, dt) : dda_step<true>(r, mp, t, cx, cy, cz, dt);
                } else {
                    rp.weights_sum[n] = ws;
                    rp.depth[n] = dsum;
                    rp.image[(size_t)n * 3] = cr; rp.image[(size_t)n * 3 + 1] = cg; rp.image[(size_t)n * 3 + 2] = cb;
                    have_ray = false;
                }
            }
        }
        // ---------------- 3. shade
        if (__any_sync(0xffffffffu, has)) {
            shade_rounds++;
            if (has) t += dt;
            float o[4];
            float* const stg = stage[threadIdx.x >> 5];
            mlpmma::encode_staged<L, 1>(lv, table, (cx + cfg.bound) * cfg.inv2b, (cy + cfg.bound) * cfg.inv2b, (cz + cfg.bound) * cfg.inv2b,
                                                 has, stg);
            mlpmma::mlp_forward_staged<L>(stg, frags, o);     // tensor cores, all 32 lanes participate
            if (has) {
                const float o0 = o[0] + ob0, o1 = o[1] + ob1, o2 = o[2] + ob2, o3 = o[3] + ob3;
                const float sigma = __expf(o0 + blob_of(cfg, cx, cy, cz));
                // kernel_composite_rays (raymarching.cu:878-903): T = 1 - weight_sum, the ray dies after accumulating the sample
                const float alpha = 1.0f - __expf(-sigma * dt);
                const float T = 1 - ws;
                const float weight = alpha * T;
                ws += weight;
                dsum += weight / t;
                cr += weight * fmaf(1.f / (1.f + __expf(-o1)), cfg.sat_scale, cfg.sat_shift);
                cg += weight * fmaf(1.f / (1.f + __expf(-o2)), cfg.sat_scale, cfg.sat_shift);
                cb += weight * fmaf(1.f / (1.f + __expf(-o3)), cfg.sat_scale, cfg.sat_shift);
                step++; shaded++;
                if (T < rp.T_thresh) terminated = true;
            }
        }
    }
    // statistics: samples shaded (one atomic per warp)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) shaded += __shfl_xor_sync(0xffffffffu, shaded, o);
    if (lane == 0) {
        if (shaded) atomicAdd(stats, (unsigned long long)shaded);
        atomicAdd(stats + 1, (unsigned long long)shade_rounds);     // warp-rounds that shaded (x32 = lane slots)
        atomicAdd(stats + 2, (unsigned long long)rounds);
        atomicAdd(stats + 3, (unsigned long long)dda_trips);       // warp-level trips of the DDA search loop
    }
}

// ---------------------------------------------------------------------------------------------
// weight culling: warp per ray, 8 rays per CTA
// ---------------------------------------------------------------------------------------------
constexpr int CC_T = 256;
__global__ void __launch_bounds__(CC_T) k_cull_compact(const float* __restrict__ weights, const float th, const int* __restrict__ rays_in,
                                                       const float* __restrict__ xyzs_in, const float* __restrict__ ts_in, const uint32_t N,
                                                       uint32_t M, const int* __restrict__ M_dev, const uint32_t M_out_cap,
                                                       int* __restrict__ rays_out, float* __restrict__ xyzs_out, float* __restrict__ ts_out,
                                                       int* __restrict__ counter) {
    __shared__ uint32_t s_cnt[CC_T / 32];
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n = blockIdx.x * (CC_T / 32) + warp;
    if (M_dev) M = min(M, (uint32_t)*M_dev);
    uint32_t offset = 0, num = 0;
    if (n < N) {
        offset = rays_in[n * 2]; num = rays_in[n * 2 + 1];
        if ((uint64_t)offset + num > M) num = 0;
    }
    uint32_t kept = 0;
    for (uint32_t i = lane; i < num; i += 32) kept += (weights[offset + i] > th) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if (lane == 0) s_cnt[warp] = kept;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < CC_T / 32; w++) { const uint32_t v = s_cnt[w]; s_cnt[w] = tot; tot += v; }
        s_base = tot ? (uint32_t)atomicAdd(counter, (int)tot) : 0u;
    }
    __syncthreads();
    if (n >= N) return;
    uint32_t out = s_base + s_cnt[warp];
    if ((uint64_t)out + kept > M_out_cap) {      // output buffer full: the ray is dropped (empty) rather than written out of bounds
        if (lane == 0) { rays_out[n * 2] = 0; rays_out[n * 2 + 1] = 0; }
        return;
    }
    if (lane == 0) { rays_out[n * 2] = (int)out; rays_out[n * 2 + 1] = (int)kept; }
    for (uint32_t base = 0; base < num; base += 32) {
        const uint32_t i = base + lane;
        const bool k = (i < num) && (weights[offset + i] > th);
        const uint32_t bal = __ballot_sync(0xffffffffu, k);
        if (k) {
            const uint32_t dst = out + __popc(bal & ((1u << lane) - 1u));
            const size_t src = offset + i;
            xyzs_out[(size_t)dst * 3] = xyzs_in[src * 3]; xyzs_out[(size_t)dst * 3 + 1] = xyzs_in[src * 3 + 1]; xyzs_out[(size_t)dst * 3 + 2] = xyzs_in[src * 3 + 2];
            ts_out[(size_t)dst * 2] = ts_in[src * 2]; ts_out[(size_t)dst * 2 + 1] = ts_in[src * 2 + 1];
        }
        out += __popc(bal);
    }
}

// occupancy-grid EMA (base_volume_renderer.py:163-167): grid = where(grid>=0 & tmp>=0, max(grid*decay, tmp), grid), fp16 grid;
// also accumulates sum(max(grid,0)) for the mean-density threshold (:166).
__global__ void k_grid_ema(__half* __restrict__ grid, const float* __restrict__ sigmas, const int* __restrict__ indices, const uint32_t n,
                           const float decay, float* __restrict__ sum_out) {
    float local = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t cell = indices ? (uint32_t)indices[i] : i;
        const float g = __half2float(grid[cell]);
        // tmp_grid = sigmas.clamp(max=fp16 max).to(fp16) (:139-140)
        const float tmp = __half2float(__float2half(fminf(sigmas[i], 65504.f)));
        float v = g;
        if (g >= 0.f && tmp >= 0.f) v = __half2float(__float2half(fmaxf(__half2float(__float2half(g * decay)), tmp)));
        grid[cell] = __float2half(v);
        local += fmaxf(v, 0.f);
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local != 0.f) atomicAdd(sum_out, local);
}

// packbits with the threshold min(mean, density_thresh) read from the device (:166-175)
__global__ void k_packbits_dev(const __half* __restrict__ grid, const uint32_t N, const float* __restrict__ sum, const float inv_count,
                               const float density_thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float thresh = fminf(__half2float(__float2half(sum[0] * inv_count)), density_thresh);  // torch.mean of an fp16 grid is fp16
    const uint4 raw = reinterpret_cast<const uint4*>(grid)[n];
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 f = __half22float2(h[i]);
        bits |= (f.x >= thresh) ? (1u << (2 * i)) : 0u;
        bits |= (f.y >= thresh) ? (1u << (2 * i + 1)) : 0u;
    }
    bitfield[n] = (uint8_t)bits;
}

// work counter + statistics of the fused renderer: one 48-byte buffer PER DEVICE (a launch zeroes it stream-ordered; concurrent
// mve_render_rays calls on two streams of the same device are not supported and documented so in the header)
static unsigned char* g_render_scratch_dev[16] = {};
#define g_render_scratch (g_render_scratch_dev[render_dev_index()])
static inline int render_dev_index() { int d = 0; cudaGetDevice(&d); return d & 15; }

}  // namespace

extern "C" {

int mve_render_last_sample_count(uint64_t* host_out) {
    MVE_ARG(host_out != nullptr, "render_last_sample_count: null output");
    host_out[0] = host_out[1] = host_out[2] = host_out[3] = 0;
    if (!g_render_scratch) return 0;
    MVE_CUDA(cudaMemcpy(host_out, g_render_scratch + 8, 32, cudaMemcpyDeviceToHost));   // synchronises: statistics only
    return 0;
}

int mve_render_rays(const float* rays_o, const float* rays_d, const float* poses, const float* intrinsics, const float* dt_gamma_per_view,
                    uint32_t h, uint32_t w, uint32_t N, const float* aabb, float min_near, const uint8_t* density_bitfield, float bound,
                    float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, float T_thresh, const float* table, const float* w1,
                    const float* b1, const float* w2, const float* b2, uint32_t n_levels, const float* level_scale,
                    const uint32_t* level_res, const uint32_t* level_size, const uint32_t* level_offset, float blob_density,
                    float blob_radius, float sigmoid_saturation, float* weights_sum, float* depth, float* image, void* stream) {
    if (N == 0) return 0;
    MVE_ARG((rays_o && rays_d) || (poses && intrinsics && h && w), "render_rays: give rays_o/rays_d or poses/intrinsics/h/w");
    MVE_ARG((reinterpret_cast<uintptr_t>(density_bitfield) & 7u) == 0, "render_rays: density_bitfield must be 8-byte aligned");
    Levels lv;
    MVE_ARG(fill_levels(lv, n_levels, level_scale, level_res, level_size, level_offset) == 0, "field: n_levels > 16");
    MVE_ARG(n_levels == 12 || n_levels == 14 || n_levels == 16, "field: n_levels must be 12, 14 or 16");
    const FieldCfg cfg = make_cfg(bound, blob_density, blob_radius, sigmoid_saturation);
    const MarchParams mp = make_params(density_bitfield, bound, false, dt_gamma, max_steps, C, H);
    RenderParams rp{};
    rp.rays_o = rays_o; rp.rays_d = rays_d; rp.poses = poses; rp.intrinsics = intrinsics; rp.h = h; rp.w = w; rp.N = N; rp.aabb = aabb;
    rp.min_near = min_near; rp.T_thresh = T_thresh; rp.max_steps = max_steps; rp.weights_sum = weights_sum; rp.depth = depth;
    rp.image = image; rp.dt_gamma_per_view = dt_gamma_per_view;
    uint32_t grid = cdiv(N, 128);
    if (grid > (uint32_t)(8 * kNumSM)) grid = 8 * kNumSM;
    const float2* t2 = reinterpret_cast<const float2*>(table);
    cudaStream_t s = (cudaStream_t)stream;
    // one process drives one GPU: process-wide scratch = [work counter (u32) | pad | samples shaded (u64)]
    if (!g_render_scratch) MVE_CUDA(cudaMalloc(&g_render_scratch, 48));
    MVE_CUDA(cudaMemsetAsync(g_render_scratch, 0, 48, s));
    unsigned int* next_ray = reinterpret_cast<unsigned int*>(g_render_scratch);
    unsigned long long* stats = reinterpret_cast<unsigned long long*>(g_render_scratch + 8);
    const bool lean = C == 1 && mp.h_pow2 && H >= 4 && H <= 256;      // what every VolumeRenderer call of the reference passes
#define RENDER(LL)                                                                                                        \
    if (lean) k_render_rays<LL, true><<<grid, 128, 0, s>>>(rp, mp, t2, w1, b1, w2, b2, lv, cfg, next_ray, stats);         \
    else k_render_rays<LL, false><<<grid, 128, 0, s>>>(rp, mp, t2, w1, b1, w2, b2, lv, cfg, next_ray, stats);
    if (n_levels == 12) { RENDER(12) } else if (n_levels == 14) { RENDER(14) } else { RENDER(16) }
#undef RENDER
    MVE_CHECK_LAUNCH("mve_render_rays");
    return 0;
}

int mve_cull_samples(const float* weights, float th, const int32_t* rays_in, const float* xyzs_in, const float* ts_in, uint32_t N,
                     uint32_t M, const int32_t* M_dev, uint32_t M_out_cap, int32_t* rays_out, float* xyzs_out, float* ts_out,
                     int32_t* counter, void* stream) {
    if (N == 0) return 0;
    k_cull_compact<<<cdiv(N, CC_T / 32), CC_T, 0, (cudaStream_t)stream>>>(weights, th, rays_in, xyzs_in, ts_in, N, M, M_dev, M_out_cap,
                                                                           rays_out, xyzs_out, ts_out, counter);
    MVE_CHECK_LAUNCH("mve_cull_samples");
    return 0;
}

int mve_density_grid_update(void* grid_half, const float* sigmas, const int32_t* indices, uint32_t n, float decay, float* sum_scratch,
                            uint32_t n_cells, float density_thresh, uint8_t* bitfield, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    MVE_ARG(n_cells % 8 == 0, "density_grid_update: n_cells must be a multiple of 8");
    MVE_CUDA(cudaMemsetAsync(sum_scratch, 0, sizeof(float), s));
    // the mean is over ALL cells: when only a subset (indices) was refreshed the untouched cells are added by a second pass
    if (n > 0) {
        uint32_t g = cdiv(n, 256);
        if (g > (uint32_t)(8 * kNumSM)) g = 8 * kNumSM;
        MVE_ARG(indices == nullptr || n <= n_cells, "density_grid_update: bad index count");
        MVE_ARG(indices != nullptr || n == n_cells, "density_grid_update: dense update needs n == n_cells");
        MVE_ARG(indices == nullptr, "density_grid_update: partial (indexed) update is not implemented; the pipeline always runs the full update");
        k_grid_ema<<<g, 256, 0, s>>>((__half*)grid_half, sigmas, indices, n, decay, sum_scratch);
    }
    k_packbits_dev<<<cdiv(n_cells / 8, 256), 256, 0, s>>>((const __half*)grid_half, n_cells / 8, sum_scratch, 1.0f / (float)n_cells,
                                                          density_thresh, bitfield);
    MVE_CHECK_LAUNCH("mve_density_grid_update");
    return 0;
}

}  // extern "C"

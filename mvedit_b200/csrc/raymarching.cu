// raymarching.cu -- B200 kernels for the ray-marching boundary.
//
// Semantics follow /root/reference/lib/ops/raymarching/src/raymarching.cu (cited per kernel);
// the execution plan does not:
//   * march_train: count + block scan + one atomic per CTA + write in ONE launch (the reference
//     launches twice around a host .item()),
//   * composite fwd/bwd: a sub-warp of G lanes per ray, lane-per-sample with shuffle scans, so every
//     global access is a coalesced run instead of a thread-private strided walk,
//   * packbits reads fp16 or fp32 with 16-byte loads,
//   * everything runs on the caller's stream and never syncs with the host.
#include "march_device.cuh"
#include "../../include/mvedit_b200.h"

using namespace march;

namespace {

// ------------------------------------------------------------------------------------------
// near/far  (raymarching.cu:92-145).  A CTA stages its 3*T floats of o and d through shared
// memory with coalesced loads; each thread then does the slab test for one ray.
// ------------------------------------------------------------------------------------------
constexpr int NF_T = 256;
__global__ void __launch_bounds__(NF_T) k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                   const float* __restrict__ aabb, const uint32_t N, const float min_near,
                                                   float* __restrict__ nears, float* __restrict__ fars) {
    __shared__ float so[NF_T * 3], sd[NF_T * 3];
    const size_t base = (size_t)blockIdx.x * NF_T;
    const uint32_t cnt = min((size_t)NF_T, (size_t)N - base) * 3;
    for (uint32_t i = threadIdx.x; i < cnt; i += NF_T) {
        so[i] = rays_o[base * 3 + i];
        sd[i] = rays_d[base * 3 + i];
    }
    __syncthreads();
    const size_t n = base + threadIdx.x;
    if (n >= N) return;
    const float ox = so[threadIdx.x * 3], oy = so[threadIdx.x * 3 + 1], oz = so[threadIdx.x * 3 + 2];
    const float dx = sd[threadIdx.x * 3], dy = sd[threadIdx.x * 3 + 1], dz = sd[threadIdx.x * 3 + 2];
    const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
    constexpr float kMax = 3.402823466e+38f;

    float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, tmp;
    if (near > far) { tmp = near; near = far; far = tmp; }
    float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
    if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
    if (near > far_y || near_y > far) { nears[n] = fars[n] = kMax; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
    if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
    if (near > far_z || near_z > far) { nears[n] = fars[n] = kMax; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    nears[n] = near;
    fars[n] = far;
}

// morton (raymarching.cu:214-254)
__global__ void k_morton3D(const int* __restrict__ coords, const uint32_t N, int* __restrict__ indices) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    indices[n] = morton3(coords[n * 3], coords[n * 3 + 1], coords[n * 3 + 2]);
}
__global__ void k_morton3D_invert(const int* __restrict__ indices, const uint32_t N, int* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const int ind = indices[n];
    coords[n * 3 + 0] = morton3_inv(ind >> 0);
    coords[n * 3 + 1] = morton3_inv(ind >> 1);
    coords[n * 3 + 2] = morton3_inv(ind >> 2);
}

// packbits (raymarching.cu:268-289): bit i of byte n = grid[8n+i] >= thresh. 8 elements per thread,
// read as one (fp16) or two (fp32) 16-byte loads.
template <bool HALF>
__global__ void k_packbits(const void* __restrict__ grid_, const uint32_t N, const float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    float v[8];
    if (HALF) {
        const uint4 raw = reinterpret_cast<const uint4*>(grid_)[n];
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; i++) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    } else {
        const float4 a = reinterpret_cast<const float4*>(grid_)[2 * (size_t)n], b = reinterpret_cast<const float4*>(grid_)[2 * (size_t)n + 1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) bits |= (v[i] >= thresh) ? (1u << i) : 0u;
    bitfield[n] = (uint8_t)bits;
}

template <bool WRITE>
__device__ __forceinline__ uint32_t march_one(const Ray& r, const MarchParams& p, const float near, const float far, const float noise,
                                              const uint32_t num_steps, float* __restrict__ xyzs, float* __restrict__ dirs,
                                              float* __restrict__ ts) {
    float t = near;
    t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max) * noise;
    uint32_t step = 0;
    float cx, cy, cz, dt;
    while (t < far && step < num_steps) {
        if (dda_step(r, p, t, cx, cy, cz, dt)) {
            step++;
            t += dt;
            if (WRITE) {
                xyzs[0] = cx; xyzs[1] = cy; xyzs[2] = cz;
                if (dirs) { dirs[0] = r.dx; dirs[1] = r.dy; dirs[2] = r.dz; dirs += 3; }   // view-independent fields skip dirs
                ts[0] = t; ts[1] = dt;
                xyzs += 3; ts += 2;
            }
        }
    }
    return step;
}

// Fused train marcher (raymarching.cu:338-475 + raymarching.py:286-300 in one launch).
constexpr int MT_T = 128;
__global__ void __launch_bounds__(MT_T) k_march_train(const float* __restrict__ rays_o, const float* __restrict__ rays_d, MarchParams p,
                                                      const float* __restrict__ dt_gamma_dev, const uint32_t max_steps, const uint32_t N, const float* __restrict__ nears,
                                                      const float* __restrict__ fars, const float* __restrict__ noises,
                                                      float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ ts,
                                                      const uint32_t max_M, int* __restrict__ rays, int* __restrict__ counter) {
    __shared__ uint32_t warp_tot[MT_T / 32];
    __shared__ uint32_t cta_base;
    const uint32_t n = threadIdx.x + blockIdx.x * MT_T;
    const bool live = n < N;
    if (dt_gamma_dev) p.dt_gamma = dt_gamma_dev[0];
    Ray r;
    float near = 0, far = 0, noise = 0;
    uint32_t cnt = 0;
    if (live) {
        r = load_ray(rays_o, rays_d, n);
        near = nears[n]; far = fars[n];
        noise = noises ? noises[n] : 0.0f;
        cnt = march_one<false>(r, p, near, far, noise, max_steps, nullptr, nullptr, nullptr);
    }
    // CTA exclusive scan of cnt
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < MT_T / 32; w++) { const uint32_t v = warp_tot[w]; warp_tot[w] = tot; tot += v; }
        cta_base = tot ? (uint32_t)atomicAdd(counter, (int)tot) : 0u;
    }
    __syncthreads();
    if (!live) return;
    const uint32_t offset = cta_base + warp_tot[warp] + incl - cnt;
    rays[n * 2] = (int)offset;
    rays[n * 2 + 1] = (int)cnt;
    if (xyzs == nullptr || cnt == 0 || (uint64_t)offset + cnt > max_M) return;
    march_one<true>(r, p, near, far, noise, cnt, xyzs + (size_t)offset * 3, dirs ? dirs + (size_t)offset * 3 : nullptr, ts + (size_t)offset * 2);
}

// Single-march variant of the fused train marcher: the counting pass records every sample's t in a scratch row, so the write
// pass no longer walks the occupancy grid a second time -- it is a warp-per-ray, coalesced expansion t -> (xyz, dirs, ts).
// The march itself is one thread per ray (the t sequence of a ray is inherently sequential) and latency-bound: 16 384 rays are
// 3.5 warps per SM, so halving the serial work is the lever.  LEAN: dda_step_lean<false> (same per-cell walk, cheaper steps).
constexpr int MR_T = 64;
template <bool LEAN>
__global__ void __launch_bounds__(MR_T) k_march_record(const float* __restrict__ rays_o, const float* __restrict__ rays_d, MarchParams p,
                                                       const float* __restrict__ dt_gamma_dev, const uint32_t max_steps, const uint32_t N,
                                                       const float* __restrict__ nears, const float* __restrict__ fars,
                                                       const float* __restrict__ noises, float* __restrict__ t_scratch,
                                                       int* __restrict__ rays, int* __restrict__ counter) {
    __shared__ uint32_t warp_tot[MR_T / 32];
    __shared__ uint32_t cta_base;
    __shared__ uint32_t s_lut[LEAN ? 256 : 1];
    if (LEAN) {
        for (int i = threadIdx.x; i < 256; i += MR_T) s_lut[i] = expand_bits((uint32_t)i);
        __syncthreads();
    }
    const uint32_t n = threadIdx.x + blockIdx.x * MR_T;
    const bool live = n < N;
    if (dt_gamma_dev) p.dt_gamma = dt_gamma_dev[0];
    uint32_t cnt = 0;
    if (live) {
        const Ray r = load_ray(rays_o, rays_d, n);
        const RayAux aux = make_aux(r);
        const float far = fars[n];
        float t = nears[n];
        t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max) * (noises ? noises[n] : 0.0f);
        float* __restrict__ row = t_scratch + (size_t)n * max_steps;
        float cx, cy, cz, dt;
        while (t < far && cnt < max_steps) {
            const bool hit = LEAN ? dda_step_lean<false>(r, aux, p, s_lut, t, cx, cy, cz, dt) : dda_step(r, p, t, cx, cy, cz, dt);
            if (hit) {
                row[cnt++] = t;
                t += dt;
            }
        }
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < MR_T / 32; w++) { const uint32_t v = warp_tot[w]; warp_tot[w] = tot; tot += v; }
        cta_base = tot ? (uint32_t)atomicAdd(counter, (int)tot) : 0u;
    }
    __syncthreads();
    if (!live) return;
    rays[n * 2] = (int)(cta_base + warp_tot[warp] + incl - cnt);
    rays[n * 2 + 1] = (int)cnt;
}

// t -> sample: the expressions of dda_step's occupied branch and march_one's write (no contraction on this path)
__global__ void __launch_bounds__(256) k_march_expand(const float* __restrict__ rays_o, const float* __restrict__ rays_d, MarchParams p,
                                                      const float* __restrict__ dt_gamma_dev, const uint32_t max_steps, const uint32_t N,
                                                      const float* __restrict__ t_scratch, const int* __restrict__ rays,
                                                      float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ ts,
                                                      const uint32_t max_M) {
    const uint32_t n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    if (dt_gamma_dev) p.dt_gamma = dt_gamma_dev[0];
    const uint32_t offset = rays[n * 2], cnt = rays[n * 2 + 1];
    if (cnt == 0 || (uint64_t)offset + cnt > max_M) return;
    const Ray r = load_ray(rays_o, rays_d, n);
    const float bound = p.bound;
    const float* __restrict__ row = t_scratch + (size_t)n * max_steps;
    for (uint32_t s = lane; s < cnt; s += 32) {
        const float t = row[s];
        const float x = clampf(r.ox + t * r.dx, -bound, bound);
        const float y = clampf(r.oy + t * r.dy, -bound, bound);
        const float z = clampf(r.oz + t * r.dz, -bound, bound);
        const float dt = clampf(t * p.dt_gamma, p.dt_min, p.dt_max);
        const size_t o = (size_t)offset + s;
        xyzs[o * 3] = x; xyzs[o * 3 + 1] = y; xyzs[o * 3 + 2] = z;
        if (dirs) { dirs[o * 3] = r.dx; dirs[o * 3 + 1] = r.dy; dirs[o * 3 + 2] = r.dz; }
        ts[o * 2] = t + dt; ts[o * 2 + 1] = dt;
    }
}

// second pass only (reference protocol: offsets already in rays)
__global__ void __launch_bounds__(MT_T) k_march_train_write(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const MarchParams p,
                                                            const uint32_t N, const float* __restrict__ nears, const float* __restrict__ fars,
                                                            const float* __restrict__ noises, float* __restrict__ xyzs,
                                                            float* __restrict__ dirs, float* __restrict__ ts, const uint32_t max_M,
                                                            const int* __restrict__ rays) {
    const uint32_t n = threadIdx.x + blockIdx.x * MT_T;
    if (n >= N) return;
    const uint32_t offset = rays[n * 2], cnt = rays[n * 2 + 1];
    if (cnt == 0 || (uint64_t)offset + cnt > max_M) return;
    const Ray r = load_ray(rays_o, rays_d, n);
    march_one<true>(r, p, nears[n], fars[n], noises ? noises[n] : 0.0f, cnt, xyzs + (size_t)offset * 3,
                    dirs ? dirs + (size_t)offset * 3 : nullptr, ts + (size_t)offset * 2);
}

// ------------------------------------------------------------------------------------------
// composite (train) forward  -- raymarching.cu:501-579.
// G lanes per ray, one sample per lane per round.  Transmittance = carried product x in-group
// inclusive product scan of (1-alpha).  A sample is kept iff every earlier T_after >= T_thresh
// (the reference's `break` comes after accumulating the offending sample).
// ------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ float group_scan_mul(float v, const uint32_t gmask, const int gl) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const float u = __shfl_up_sync(gmask, v, o, G);
        if (gl >= o) v *= u;
    }
    return v;
}
template <int G>
__device__ __forceinline__ float group_scan_add(float v, const uint32_t gmask, const int gl) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const float u = __shfl_up_sync(gmask, v, o, G);
        if (gl >= o) v += u;
    }
    return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v, const uint32_t gmask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
    return v;
}

constexpr int CP_T = 256;

template <int G>
__global__ void __launch_bounds__(CP_T) k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                              const float* __restrict__ ts, const int* __restrict__ rays, uint32_t M,
                                                              const int* __restrict__ M_dev, const uint32_t N, const float T_thresh,
                                                              const bool binarize, float* __restrict__ weights,
                                                              float* __restrict__ weights_sum, float* __restrict__ depth,
                                                              float* __restrict__ image) {
    const uint32_t n = (blockIdx.x * CP_T + threadIdx.x) / G;
    if (n >= N) return;  // whole groups leave together
    const int lane = threadIdx.x & 31, gl = lane % G;
    const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - gl));
    if (M_dev) M = min(M, (uint32_t)*M_dev);
    const uint32_t offset = rays[n * 2], num_steps = rays[n * 2 + 1];
    if (num_steps == 0 || (uint64_t)offset + num_steps > M) {
        if (gl == 0) { weights_sum[n] = 0; depth[n] = 0; image[n * 3] = 0; image[n * 3 + 1] = 0; image[n * 3 + 2] = 0; }
        return;
    }
    const float* __restrict__ sg = sigmas + offset;
    const float* __restrict__ cl = rgbs + (size_t)offset * 3;
    const float2* __restrict__ tt = reinterpret_cast<const float2*>(ts) + offset;
    float* __restrict__ wo = weights + offset;

    float T_carry = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0;
    uint32_t base = 0;
    for (; base < num_steps; base += G) {
        const uint32_t i = base + gl;
        const bool valid = i < num_steps;
        float sigma = 0, cr = 0, cg = 0, cb = 0;
        float2 td = make_float2(1.0f, 0.0f);
        if (valid) {
            sigma = sg[i]; td = tt[i];
            if (rgbs) { cr = cl[i * 3]; cg = cl[i * 3 + 1]; cb = cl[i * 3 + 2]; }   // rgbs == NULL: weights-only pass (culling)
        }
        const float real_alpha = 1.0f - __expf(-sigma * td.y);
        const float alpha = valid ? (binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha) : 0.0f;
        const float pin = group_scan_mul<G>(1.0f - alpha, gmask, gl);      // inclusive product within the round
        float pex = __shfl_up_sync(gmask, pin, 1, G);
        if (gl == 0) pex = 1.0f;
        const float T_after = T_carry * pin;
        const float weight = alpha * (T_carry * pex);
        const uint32_t stop = (__ballot_sync(gmask, valid && (T_after < T_thresh)) & gmask) >> (lane - gl);
        const int first = stop ? (__ffs(stop) - 1) : G;
        const bool keep = valid && gl <= first;
        const float w = keep ? weight : 0.0f;
        if (valid) wo[i] = w;
        r += w * cr; g += w * cg; b += w * cb; ws += w;
        d += keep ? w / td.x : 0.0f;
        if (first < G) { base += G; break; }
        T_carry = __shfl_sync(gmask, T_after, G - 1, G);
    }
    for (uint32_t i = base + gl; i < num_steps; i += G) wo[i] = 0.0f;  // early-terminated tail
    r = group_sum<G>(r, gmask); g = group_sum<G>(g, gmask); b = group_sum<G>(b, gmask);
    ws = group_sum<G>(ws, gmask); d = group_sum<G>(d, gmask);
    if (gl == 0) { weights_sum[n] = ws; depth[n] = d; image[n * 3] = r; image[n * 3 + 1] = g; image[n * 3 + 2] = b; }
}

// ------------------------------------------------------------------------------------------
// composite (train) backward -- raymarching.cu:606-695.
// The reference's five running sums collapse into two scans: with q_j = gi.rgb_j + gws + gd/t_j and
// S = sum_j w_j q_j,
//   dsigma_i = dt_i * [ T_i (q_i + gw_i) - (S_final - S_i) - gw_i (ws_final - ws_i) ],
// T_i the transmittance after sample i, S_i / ws_i inclusive running sums, S_final = gi.image + gws*ws + gd*depth.
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(CP_T) k_composite_train_bwd(const float* __restrict__ grad_weights, const float* __restrict__ grad_weights_sum,
                                                              const float* __restrict__ grad_depth, const float* __restrict__ grad_image,
                                                              const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                              const float* __restrict__ ts, const int* __restrict__ rays,
                                                              const float* __restrict__ weights_sum, const float* __restrict__ depth,
                                                              const float* __restrict__ image, uint32_t M, const int* __restrict__ M_dev,
                                                              const uint32_t N, const float T_thresh, const bool binarize,
                                                              const float* __restrict__ ent_w, const float ent_scale,
                                                              float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * CP_T + threadIdx.x) / G;
    if (n >= N) return;
    const int lane = threadIdx.x & 31, gl = lane % G;
    const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - gl));
    if (M_dev) M = min(M, (uint32_t)*M_dev);
    const uint32_t offset = rays[n * 2], num_steps = rays[n * 2 + 1];
    if (num_steps == 0 || (uint64_t)offset + num_steps > M) return;
    // optional fused term: L_ent = -c * sum_i w_i (log max(w_i,1e-6) - log max(dt_i,1e-6))  (mvedit_3d_pipeline.py:595-603, sample part)
    const float ent_c = ent_w ? ent_w[0] * ent_scale : 0.0f;

    const float gws = grad_weights_sum[n], gd = grad_depth[n];
    const float g0 = grad_image[n * 3], g1 = grad_image[n * 3 + 1], g2 = grad_image[n * 3 + 2];
    const float ws_final = weights_sum[n];
    const float S_final = g0 * image[n * 3] + g1 * image[n * 3 + 1] + g2 * image[n * 3 + 2] + gws * ws_final + gd * depth[n];

    const float* __restrict__ gw = grad_weights + offset;
    const float* __restrict__ sg = sigmas + offset;
    const float* __restrict__ cl = rgbs + (size_t)offset * 3;
    const float2* __restrict__ tt = reinterpret_cast<const float2*>(ts) + offset;
    float* __restrict__ gs = grad_sigmas + offset;
    float* __restrict__ gc = grad_rgbs + (size_t)offset * 3;

    float T_carry = 1.0f, S_carry = 0.0f, ws_carry = 0.0f;
    uint32_t base = 0;
    for (; base < num_steps; base += G) {
        const uint32_t i = base + gl;
        const bool valid = i < num_steps;
        float sigma = 0, cr = 0, cg = 0, cb = 0, gwi = 0;
        float2 td = make_float2(1.0f, 0.0f);
        if (valid) { sigma = sg[i]; td = tt[i]; cr = cl[i * 3]; cg = cl[i * 3 + 1]; cb = cl[i * 3 + 2]; gwi = grad_weights ? gw[i] : 0.0f; }
        const float real_alpha = 1.0f - __expf(-sigma * td.y);
        const float alpha = valid ? (binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha) : 0.0f;
        const float pin = group_scan_mul<G>(1.0f - alpha, gmask, gl);
        float pex = __shfl_up_sync(gmask, pin, 1, G);
        if (gl == 0) pex = 1.0f;
        const float T_after = T_carry * pin;
        const float weight = alpha * (T_carry * pex);
        const uint32_t stop = (__ballot_sync(gmask, valid && (T_after < T_thresh)) & gmask) >> (lane - gl);
        const int first = stop ? (__ffs(stop) - 1) : G;
        const bool keep = valid && gl <= first;
        const float w = keep ? weight : 0.0f;
        if (ent_c != 0.0f && valid)
            gwi -= ent_c * ((__logf(fmaxf(w, 1e-6f)) - __logf(fmaxf(td.y, 1e-6f))) + ((w > 1e-6f) ? 1.0f : 0.0f));
        const float q = g0 * cr + g1 * cg + g2 * cb + gws + gd / td.x;
        const float S_i = S_carry + group_scan_add<G>(w * q, gmask, gl);
        const float ws_i = ws_carry + group_scan_add<G>(w, gmask, gl);
        if (valid) {
            const float gsig = keep ? td.y * (T_after * (q + gwi) - (S_final - S_i) - gwi * (ws_final - ws_i)) : 0.0f;
            gs[i] = gsig;
            gc[i * 3] = g0 * w; gc[i * 3 + 1] = g1 * w; gc[i * 3 + 2] = g2 * w;
        }
        if (first < G) { base += G; break; }
        T_carry = __shfl_sync(gmask, T_after, G - 1, G);
        S_carry = __shfl_sync(gmask, S_i, G - 1, G);
        ws_carry = __shfl_sync(gmask, ws_i, G - 1, G);
    }
    for (uint32_t i = base + gl; i < num_steps; i += G) { gs[i] = 0.0f; gc[i * 3] = 0.0f; gc[i * 3 + 1] = 0.0f; gc[i * 3 + 2] = 0.0f; }
}

// ------------------------------------------------------------------------------------------
// inference marcher / compositor (raymarching.cu:714-829, :843-925): reference-protocol versions
// (one thread per alive ray).  The production renderer is the fused persistent kernel in render.cu.
// ------------------------------------------------------------------------------------------
constexpr int MI_T = 128;
__global__ void __launch_bounds__(MI_T) k_march_infer(const uint32_t n_alive, const uint32_t n_step, const int* __restrict__ rays_alive,
                                                      const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                                                      const float* __restrict__ rays_d, const MarchParams p, const float* __restrict__ nears,
                                                      const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                                                      float* __restrict__ ts, const float* __restrict__ noises) {
    const uint32_t n = threadIdx.x + blockIdx.x * MI_T;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float noise = noises ? noises[n] : 0.0f;
    const Ray r = load_ray(rays_o, rays_d, index);
    xyzs += (size_t)n * n_step * 3; dirs += (size_t)n * n_step * 3; ts += (size_t)n * n_step * 2;
    const float far = fars[index];
    float t = rays_t[index];
    t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max) * noise;
    uint32_t step = 0;
    float cx, cy, cz, dt;
    while (t < far && step < n_step) {
        if (dda_step(r, p, t, cx, cy, cz, dt)) {
            xyzs[0] = cx; xyzs[1] = cy; xyzs[2] = cz;
            dirs[0] = r.dx; dirs[1] = r.dy; dirs[2] = r.dz;
            t += dt;
            ts[0] = t; ts[1] = dt;
            xyzs += 3; dirs += 3; ts += 2; step++;
        }
    }
    for (; step < n_step; step++) {  // the reference relies on a zero-filled buffer (raymarching.py:466-468)
        xyzs[0] = 0; xyzs[1] = 0; xyzs[2] = 0; dirs[0] = 0; dirs[1] = 0; dirs[2] = 0; ts[0] = 0; ts[1] = 0;
        xyzs += 3; dirs += 3; ts += 2;
    }
}

__global__ void __launch_bounds__(MI_T) k_composite_infer(const uint32_t n_alive, const uint32_t n_step, const float T_thresh, const bool binarize,
                                                          int* __restrict__ rays_alive, float* __restrict__ rays_t,
                                                          const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                          const float* __restrict__ ts, float* __restrict__ weights_sum,
                                                          float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = threadIdx.x + blockIdx.x * MI_T;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    sigmas += (size_t)n * n_step; rgbs += (size_t)n * n_step * 3; ts += (size_t)n * n_step * 2;
    float t = 0;
    float d = depth[index], r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2], weight_sum = weights_sum[index];
    uint32_t step = 0;
    while (step < n_step) {
        if (ts[0] == 0) break;
        const float real_alpha = 1.0f - __expf(-sigmas[0] * ts[1]);
        const float alpha = binarize ? (real_alpha > 0.5 ? 1.0 : 0.0) : real_alpha;
        const float T = 1 - weight_sum;
        const float weight = alpha * T;
        weight_sum += weight;
        t = ts[0];
        d += weight / t;
        r += weight * rgbs[0]; g += weight * rgbs[1]; b += weight * rgbs[2];
        if (T < T_thresh) break;
        sigmas++; rgbs += 3; ts += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
    weights_sum[index] = weight_sum; depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

// pick lanes-per-ray from the mean sample count
inline int pick_group(uint32_t M, uint32_t N) {
    const double mean = N ? (double)M / N : 0.0;
    if (mean >= 24) return 32;
    if (mean >= 12) return 16;
    return 8;
}

}  // namespace

extern "C" {

int mve_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near, float* nears,
                           float* fars, void* stream) {
    if (N == 0) return 0;
    k_near_far<<<cdiv(N, NF_T), NF_T, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    MVE_CHECK_LAUNCH("mve_near_far_from_aabb");
    return 0;
}

int mve_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream) {
    if (N == 0) return 0;
    k_morton3D<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
    MVE_CHECK_LAUNCH("mve_morton3D");
    return 0;
}

int mve_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream) {
    if (N == 0) return 0;
    k_morton3D_invert<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
    MVE_CHECK_LAUNCH("mve_morton3D_invert");
    return 0;
}

int mve_packbits(const void* grid, int grid_is_half, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream) {
    if (N == 0) return 0;
    MVE_ARG(((uintptr_t)grid & 15) == 0, "packbits: grid must be 16-byte aligned");
    if (grid_is_half) k_packbits<true><<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(grid, N, density_thresh, bitfield);
    else k_packbits<false><<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(grid, N, density_thresh, bitfield);
    MVE_CHECK_LAUNCH("mve_packbits");
    return 0;
}

int mve_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* density_bitfield, float bound, int contract,
                         float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* nears, const float* fars,
                         const float* noises, float* xyzs, float* dirs, float* ts, uint32_t max_M, int32_t* rays, int32_t* counter,
                         const float* dt_gamma_dev, float* t_scratch, void* stream) {
    if (N == 0) return 0;
    MVE_ARG(max_steps > 0 && H > 0 && C > 0, "march_rays_train: max_steps, H, C must be positive");
    const MarchParams p = make_params(density_bitfield, bound, contract != 0, dt_gamma, max_steps, C, H);
    if (t_scratch && xyzs && !contract) {          // single march + coalesced expansion
        cudaStream_t s = (cudaStream_t)stream;
        const bool lean = C == 1 && p.h_pow2 && H >= 4 && H <= 256;
        if (lean) k_march_record<true><<<cdiv(N, MR_T), MR_T, 0, s>>>(rays_o, rays_d, p, dt_gamma_dev, max_steps, N, nears, fars, noises, t_scratch, rays, counter);
        else k_march_record<false><<<cdiv(N, MR_T), MR_T, 0, s>>>(rays_o, rays_d, p, dt_gamma_dev, max_steps, N, nears, fars, noises, t_scratch, rays, counter);
        k_march_expand<<<cdiv(N, 8), 256, 0, s>>>(rays_o, rays_d, p, dt_gamma_dev, max_steps, N, t_scratch, rays, xyzs, dirs, ts, max_M);
        MVE_CHECK_LAUNCH("mve_march_rays_train");
        return 0;
    }
    k_march_train<<<cdiv(N, MT_T), MT_T, 0, (cudaStream_t)stream>>>(rays_o, rays_d, p, dt_gamma_dev, max_steps, N, nears, fars, noises, xyzs, dirs, ts,
                                                                     max_M, rays, counter);
    MVE_CHECK_LAUNCH("mve_march_rays_train");
    return 0;
}

int mve_march_rays_train_write(const float* rays_o, const float* rays_d, const uint8_t* density_bitfield, float bound, int contract,
                               float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* nears,
                               const float* fars, const float* noises, float* xyzs, float* dirs, float* ts, uint32_t max_M,
                               const int32_t* rays, void* stream) {
    if (N == 0) return 0;
    const MarchParams p = make_params(density_bitfield, bound, contract != 0, dt_gamma, max_steps, C, H);
    k_march_train_write<<<cdiv(N, MT_T), MT_T, 0, (cudaStream_t)stream>>>(rays_o, rays_d, p, N, nears, fars, noises, xyzs, dirs, ts, max_M,
                                                                           rays);
    MVE_CHECK_LAUNCH("mve_march_rays_train_write");
    return 0;
}

#define MVE_DISPATCH_G(G_, ...)            \
    switch (G_) {                          \
        case 32: { constexpr int G = 32; __VA_ARGS__; } break; \
        case 16: { constexpr int G = 16; __VA_ARGS__; } break; \
        default: { constexpr int G = 8; __VA_ARGS__; } break;  \
    }

int mve_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays, uint32_t M,
                                     const int32_t* M_dev, uint32_t N, float T_thresh, int binarize, float* weights, float* weights_sum,
                                     float* depth, float* image, void* stream) {
    if (N == 0) return 0;
    const int g = pick_group(M, N);
    MVE_DISPATCH_G(g, k_composite_train_fwd<G><<<cdiv((uint64_t)N * G, CP_T), CP_T, 0, (cudaStream_t)stream>>>(
                          sigmas, rgbs, ts, rays, M, M_dev, N, T_thresh, binarize != 0, weights, weights_sum, depth, image));
    MVE_CHECK_LAUNCH("mve_composite_rays_train_forward");
    return 0;
}

int mve_composite_rays_train_backward(const float* grad_weights, const float* grad_weights_sum, const float* grad_depth,
                                      const float* grad_image, const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays,
                                      const float* weights_sum, const float* depth, const float* image, uint32_t M, const int32_t* M_dev,
                                      uint32_t N, float T_thresh, int binarize, const float* entropy_weight_dev, float entropy_scale,
                                      float* grad_sigmas, float* grad_rgbs, void* stream) {
    if (N == 0) return 0;
    const int g = pick_group(M, N);
    MVE_DISPATCH_G(g, k_composite_train_bwd<G><<<cdiv((uint64_t)N * G, CP_T), CP_T, 0, (cudaStream_t)stream>>>(
                          grad_weights, grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, ts, rays, weights_sum, depth, image, M, M_dev,
                          N, T_thresh, binarize != 0, entropy_weight_dev, entropy_scale, grad_sigmas, grad_rgbs));
    MVE_CHECK_LAUNCH("mve_composite_rays_train_backward");
    return 0;
}

int mve_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                   const float* rays_d, float bound, int contract, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                   const uint8_t* density_bitfield, const float* nears, const float* fars, float* xyzs, float* dirs, float* ts,
                   const float* noises, void* stream) {
    if (n_alive == 0 || n_step == 0) return 0;
    const MarchParams p = make_params(density_bitfield, bound, contract != 0, dt_gamma, max_steps, C, H);
    k_march_infer<<<cdiv(n_alive, MI_T), MI_T, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, p, nears, fars,
                                                                           xyzs, dirs, ts, noises);
    MVE_CHECK_LAUNCH("mve_march_rays");
    return 0;
}

int mve_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int binarize, int32_t* rays_alive, float* rays_t,
                       const float* sigmas, const float* rgbs, const float* ts, float* weights_sum, float* depth, float* image,
                       void* stream) {
    if (n_alive == 0 || n_step == 0) return 0;
    k_composite_infer<<<cdiv(n_alive, MI_T), MI_T, 0, (cudaStream_t)stream>>>(n_alive, n_step, T_thresh, binarize != 0, rays_alive, rays_t,
                                                                               sigmas, rgbs, ts, weights_sum, depth, image);
    MVE_CHECK_LAUNCH("mve_composite_rays");
    return 0;
}

}  // extern "C"

/*
 * oracle/raymarching_oracle.c -- CPU restatement of the reference's in-tree
 * ray-marching kernels.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's .so.
 * The product (mvedit_b200/) never does.
 *
 * Each function restates one kernel of
 *   /root/reference/lib/ops/raymarching/src/raymarching.cu
 * one (ray | element) per loop iteration, fp32 arithmetic, same operation order
 * as the reference's thread body.  Where the reference mixes a double literal
 * into a float expression (0.5 * (...) * H at :401-403, dt * H * 0.5 at :50)
 * the restatement keeps the double arithmetic.
 *
 * Pinning: the reference has no golden vectors.  The pin is
 * (a) tests/golden/raymarching_ref_*.npz -- outputs of the reference's own
 * kernels (oracle/_ref, built by oracle/build_ref.py) captured on a B200 by
 * tests/golden/make_raymarching_golden.py, and (b) the live GPU test
 * tests/test_gpu_ref_parity.py that runs reference kernels, this oracle and our
 * kernels on the same seeded inputs.
 *
 * Differences to expect vs the GPU reference: __expf (reference, :545) vs expf
 * here (<= 2 ulp); nvcc contracts a*b+c to FMA, gcc is told -ffp-contract=off;
 * so float results agree to ~1e-6 relative and march sample counts agree except
 * on rays whose t lands within 1 ulp of a voxel boundary.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

#define SQRT3 1.7320508075688772f

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float signf_(float x) { return copysignf(1.0f, x); }

/* raymarching.cu:42-54 */
static inline int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}
static inline int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}

/* raymarching.cu:56-81 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D_(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert_(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

/* kernel_near_far_from_aabb, raymarching.cu:92-145 */
void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                            uint32_t N, float min_near, float* nears, float* fars) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, tmp;
        if (near > far) { tmp = near; near = far; far = tmp; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* kernel_morton3D / kernel_morton3D_invert, raymarching.cu:214-254 */
void orc_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int32_t)morton3D_((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}
void orc_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
    for (uint32_t n = 0; n < N; n++) {
        const int ind = indices[n];
        coords[n * 3 + 0] = (int32_t)morton3D_invert_((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int32_t)morton3D_invert_((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int32_t)morton3D_invert_((uint32_t)(ind >> 2));
    }
}

/* kernel_packbits, raymarching.cu:268-289: bit i of byte n = grid[8n+i] >= thresh */
void orc_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] >= density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* One DDA step shared by kernel_march_rays_train (:338-475) and kernel_march_rays (:714-829).
 * Returns 1 if the cell at t is occupied (then *cx,*cy,*cz,*dt_out describe the sample and t
 * is NOT yet advanced), else advances *t past the empty voxel and returns 0. */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, rH, H3, bound, dt_gamma, dt_min, dt_max;
    int contract; uint32_t C, H; const uint8_t* grid;
} march_ctx;

static inline int march_step(const march_ctx* c, float* t_io, float* cx_o, float* cy_o, float* cz_o, float* dt_o) {
    float t = *t_io;
    const float bound = c->bound;
    const uint32_t H = c->H;
    const float x = clampf(c->ox + t * c->dx, -bound, bound);
    const float y = clampf(c->oy + t * c->dy, -bound, bound);
    const float z = clampf(c->oz + t * c->dz, -bound, bound);
    float dt = clampf(t * c->dt_gamma, c->dt_min, c->dt_max);
    const int l1 = mip_from_pos(x, y, z, (float)c->C), l2 = mip_from_dt(dt, (float)H, (float)c->C);
    const int level = l1 > l2 ? l1 : l2;
    const float mip_bound = fminf(scalbnf(1.0f, level), bound);
    const float mip_rbound = 1 / mip_bound;
    float cx = x, cy = y, cz = z;
    const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    if (c->contract && mag > 1) {
        const float s = (2 - 1 / mag) / mag;
        cx *= s; cy *= s; cz *= s;
    }
    const int nx = (int)clampf((float)(0.5 * (double)(cx * mip_rbound + 1) * (double)H), 0.0f, (float)(H - 1));
    const int ny = (int)clampf((float)(0.5 * (double)(cy * mip_rbound + 1) * (double)H), 0.0f, (float)(H - 1));
    const int nz = (int)clampf((float)(0.5 * (double)(cz * mip_rbound + 1) * (double)H), 0.0f, (float)(H - 1));
    const uint32_t index = (uint32_t)((float)level * c->H3 + (float)morton3D_((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const int occ = c->grid[index / 8] & (1 << (index % 8));
    if (occ) {
        *cx_o = cx; *cy_o = cy; *cz_o = cz; *dt_o = dt;
        return 1;
    } else if (c->contract && mag > 1) {
        *t_io = t + dt;
        return 0;
    } else {
        const float tx = (((nx + 0.5f + 0.5f * signf_(c->dx)) * c->rH * 2 - 1) * mip_bound - cx) * c->rdx;
        const float ty = (((ny + 0.5f + 0.5f * signf_(c->dy)) * c->rH * 2 - 1) * mip_bound - cy) * c->rdy;
        const float tz = (((nz + 0.5f + 0.5f * signf_(c->dz)) * c->rH * 2 - 1) * mip_bound - cz) * c->rdz;
        const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        do {
            dt = clampf(t * c->dt_gamma, c->dt_min, c->dt_max);
            t += dt;
        } while (t < tt);
        *t_io = t;
        return 0;
    }
}

static inline void march_ctx_init(march_ctx* c, const float* o, const float* d, const uint8_t* grid, float bound,
                                  int contract, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    c->ox = o[0]; c->oy = o[1]; c->oz = o[2];
    c->dx = d[0]; c->dy = d[1]; c->dz = d[2];
    c->rdx = 1 / c->dx; c->rdy = 1 / c->dy; c->rdz = 1 / c->dz;
    c->rH = 1 / (float)H;
    c->H3 = (float)(H * H * H);
    c->bound = bound; c->dt_gamma = dt_gamma; c->contract = contract; c->C = C; c->H = H; c->grid = grid;
    c->dt_min = 2 * SQRT3 / max_steps;
    c->dt_max = 2 * SQRT3 * bound / H;
}

/* kernel_march_rays_train (:338-475), both passes folded: count pass over all rays, exclusive
 * prefix sum in ray order (the reference hands out offsets by atomicAdd, i.e. in arbitrary order;
 * any order is a valid reference output), then the write pass.  rays[n] = (offset, count).
 * If xyzs == NULL only the counts/offsets are produced and the total M is returned. */
int64_t orc_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, int contract,
                             float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                             const float* nears, const float* fars, const float* noises,
                             float* xyzs, float* dirs, float* ts, int32_t* rays, int64_t max_M) {
    /* pass 1 */
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        march_ctx c;
        march_ctx_init(&c, rays_o + n * 3, rays_d + n * 3, grid, bound, contract, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n];
        float t = near;
        t += clampf(t * dt_gamma, c.dt_min, c.dt_max) * noises[n];
        uint32_t step = 0;
        float cx, cy, cz, dt;
        while (t < far && step < max_steps) {
            if (march_step(&c, &t, &cx, &cy, &cz, &dt)) { step++; t += dt; }
        }
        rays[n * 2 + 1] = (int32_t)step;
    }
    int64_t M = 0;
    for (uint32_t n = 0; n < N; n++) { rays[n * 2] = (int32_t)M; M += rays[n * 2 + 1]; }
    if (xyzs == NULL) return M;
    if (M > max_M) return -M;
    /* pass 2 */
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        march_ctx c;
        march_ctx_init(&c, rays_o + n * 3, rays_d + n * 3, grid, bound, contract, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n];
        const uint32_t num_steps = (uint32_t)rays[n * 2 + 1];
        size_t p = (size_t)rays[n * 2];
        float t = near;
        t += clampf(t * dt_gamma, c.dt_min, c.dt_max) * noises[n];
        uint32_t step = 0;
        float cx, cy, cz, dt;
        while (t < far && step < num_steps) {
            if (march_step(&c, &t, &cx, &cy, &cz, &dt)) {
                step++; t += dt;
                xyzs[p * 3] = cx; xyzs[p * 3 + 1] = cy; xyzs[p * 3 + 2] = cz;
                dirs[p * 3] = c.dx; dirs[p * 3 + 1] = c.dy; dirs[p * 3 + 2] = c.dz;
                ts[p * 2] = t; ts[p * 2 + 1] = dt;
                p++;
            }
        }
    }
    return M;
}

/* kernel_composite_rays_train_forward (:501-579) */
void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays,
                                      uint32_t M, uint32_t N, float T_thresh, int binarize,
                                      float* weights, float* weights_sum, float* depth, float* image) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const uint32_t offset = (uint32_t)rays[n * 2], num_steps = (uint32_t)rays[n * 2 + 1];
        if (num_steps == 0 || offset + num_steps > M) {
            weights_sum[n] = 0; depth[n] = 0; image[n * 3] = image[n * 3 + 1] = image[n * 3 + 2] = 0;
            continue;
        }
        const float* s = sigmas + offset; const float* c = rgbs + (size_t)offset * 3; const float* tt = ts + (size_t)offset * 2;
        float* w = weights + offset;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0;
        for (uint32_t step = 0; step < num_steps; step++) {
            const float real_alpha = 1.0f - expf(-s[step] * tt[step * 2 + 1]);
            const float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
            const float weight = alpha * T;
            w[step] = weight;
            r += weight * c[step * 3]; g += weight * c[step * 3 + 1]; b += weight * c[step * 3 + 2];
            ws += weight;
            d += weight / tt[step * 2];
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
        }
        weights_sum[n] = ws; depth[n] = d; image[n * 3] = r; image[n * 3 + 1] = g; image[n * 3 + 2] = b;
    }
}

/* kernel_composite_rays_train_backward (:606-695) */
void orc_composite_rays_train_backward(const float* grad_weights, const float* grad_weights_sum, const float* grad_depth,
                                       const float* grad_image, const float* sigmas, const float* rgbs, const float* ts,
                                       const int32_t* rays, const float* weights_sum, const float* depth, const float* image,
                                       uint32_t M, uint32_t N, float T_thresh, int binarize,
                                       float* grad_sigmas, float* grad_rgbs) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const uint32_t offset = (uint32_t)rays[n * 2], num_steps = (uint32_t)rays[n * 2 + 1];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float* gw = grad_weights + offset;
        const float gws = grad_weights_sum[n], gd = grad_depth[n];
        const float* gi = grad_image + n * 3;
        const float* s = sigmas + offset; const float* c = rgbs + (size_t)offset * 3; const float* tt = ts + (size_t)offset * 2;
        float* gs = grad_sigmas + offset; float* gc = grad_rgbs + (size_t)offset * 3;
        const float r_final = image[n * 3], g_final = image[n * 3 + 1], b_final = image[n * 3 + 2];
        const float ws_final = weights_sum[n], d_final = depth[n];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0;
        for (uint32_t step = 0; step < num_steps; step++) {
            const float real_alpha = 1.0f - expf(-s[step] * tt[step * 2 + 1]);
            const float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
            const float weight = alpha * T;
            r += weight * c[step * 3]; g += weight * c[step * 3 + 1]; b += weight * c[step * 3 + 2];
            ws += weight;
            d += weight / tt[step * 2];
            T *= 1.0f - alpha;
            gc[step * 3] = gi[0] * weight; gc[step * 3 + 1] = gi[1] * weight; gc[step * 3 + 2] = gi[2] * weight;
            gs[step] = tt[step * 2 + 1] * (
                gi[0] * (T * c[step * 3] - (r_final - r)) +
                gi[1] * (T * c[step * 3 + 1] - (g_final - g)) +
                gi[2] * (T * c[step * 3 + 2] - (b_final - b)) +
                (gws + gw[step]) * (T - (ws_final - ws)) +
                gd * (T / tt[step * 2] - (d_final - d)));
            if (T < T_thresh) break;
        }
    }
}

/* kernel_march_rays (:714-829), inference: n_step samples for each alive ray.
 * xyzs/dirs/ts must come in zero-filled (raymarching.py:466-468). */
void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                    const float* rays_o, const float* rays_d, float bound, int contract, float dt_gamma,
                    uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                    const float* fars, float* xyzs, float* dirs, float* ts, const float* noises) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t n = 0; n < (int64_t)n_alive; n++) {
        const int index = rays_alive[n];
        march_ctx c;
        march_ctx_init(&c, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, contract, dt_gamma, max_steps, C, H);
        const float far = fars[index];
        float* px = xyzs + (size_t)n * n_step * 3; float* pd = dirs + (size_t)n * n_step * 3; float* pt = ts + (size_t)n * n_step * 2;
        float t = rays_t[index];
        t += clampf(t * dt_gamma, c.dt_min, c.dt_max) * noises[n];
        uint32_t step = 0;
        float cx, cy, cz, dt;
        while (t < far && step < n_step) {
            if (march_step(&c, &t, &cx, &cy, &cz, &dt)) {
                px[0] = cx; px[1] = cy; px[2] = cz;
                pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt;
                pt[0] = t; pt[1] = dt;
                px += 3; pd += 3; pt += 2; step++;
            }
        }
    }
}

/* kernel_composite_rays (:843-925), inference, in place */
void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int binarize, int32_t* rays_alive,
                        float* rays_t, const float* sigmas, const float* rgbs, const float* ts,
                        float* weights_sum, float* depth, float* image) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t n = 0; n < (int64_t)n_alive; n++) {
        const int index = rays_alive[n];
        const float* s = sigmas + (size_t)n * n_step; const float* c = rgbs + (size_t)n * n_step * 3; const float* tt = ts + (size_t)n * n_step * 2;
        float t = 0;
        float d = depth[index], r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2], weight_sum = weights_sum[index];
        uint32_t step = 0;
        while (step < n_step) {
            if (tt[0] == 0) break;
            const float real_alpha = 1.0f - expf(-s[0] * tt[1]);
            const float alpha = binarize ? (real_alpha > 0.5f ? 1.0f : 0.0f) : real_alpha;
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t = tt[0];
            d += weight / t;
            r += weight * c[0]; g += weight * c[1]; b += weight * c[2];
            if (T < T_thresh) break;
            s++; c += 3; tt += 2; step++;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = weight_sum; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

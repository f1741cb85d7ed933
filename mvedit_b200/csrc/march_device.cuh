// march_device.cuh -- DDA occupancy-grid traversal shared by raymarching.cu and render.cu.
// Arithmetic follows /root/reference/lib/ops/raymarching/src/raymarching.cu:42-81,375-461 expression by expression
// (bit-exact against the reference kernels, tests/test_gpu_raymarching.py).
#pragma once
#include "common.cuh"

namespace march {

__device__ __forceinline__ float clampf(const float x, const float lo, const float hi) { return fminf(hi, fmaxf(lo, x)); }
__device__ __forceinline__ float sgnf(const float x) { return copysignf(1.0f, x); }

// raymarching.cu:42-54
__device__ __forceinline__ int mip_from_pos(const float x, const float y, const float z, const float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return fminf(max_cascade - 1, fmaxf(0, e));
}
__device__ __forceinline__ int mip_from_dt(const float dt, const float H, const float max_cascade) {
    const float mx = dt * H * 0.5;  // double literal on purpose: same promotion as the reference (:50)
    int e;
    frexpf(mx, &e);
    return fminf(max_cascade - 1, fmaxf(0, e));
}
// raymarching.cu:56-81
__host__ __device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t morton3_inv(uint32_t x) {
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

// ------------------------------------------------------------------------------------------
// DDA traversal shared by the train and inference marchers (raymarching.cu:375-461 / :753-828).
// Same arithmetic, expression by expression, so results are bit-identical to the reference.
// ------------------------------------------------------------------------------------------
struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
};
struct MarchParams {
    const uint8_t* __restrict__ grid;
    float bound, dt_gamma, dt_min, dt_max, rH, H3;
    uint32_t C, H;
    bool contract, h_pow2;
};

// Visit the cell at t. Occupied: returns true with the (contracted) sample position and dt; t is not advanced.
// Empty: advances t to the first step past the current voxel and returns false.
__device__ __forceinline__ bool dda_step(const Ray& r, const MarchParams& p, float& t, float& cx, float& cy, float& cz, float& dt) {
    const float bound = p.bound;
    const uint32_t H = p.H;
    const float x = clampf(r.ox + t * r.dx, -bound, bound);
    const float y = clampf(r.oy + t * r.dy, -bound, bound);
    const float z = clampf(r.oz + t * r.dz, -bound, bound);

    dt = clampf(t * p.dt_gamma, p.dt_min, p.dt_max);

    // get mip level.  C == 1 (every VolumeRenderer call of the reference passes 1, base_volume_renderer.py:216,300) => level 0.
    const int level = (p.C == 1) ? 0 : max(mip_from_pos(x, y, z, p.C), mip_from_dt(dt, H, p.C));
    const float mip_bound = (p.C == 1) ? fminf(1.0f, bound) : fminf(scalbnf(1.0f, level), bound);
    const float mip_rbound = 1 / mip_bound;

    cx = x; cy = y; cz = z;
    const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    if (p.contract && mag > 1) {
        const float Linf_scale = (2 - 1 / mag) / mag;
        cx *= Linf_scale; cy *= Linf_scale; cz *= Linf_scale;
    }
    int nx, ny, nz;
    if (p.h_pow2) {
        // 0.5 * v * H with H a power of two only rescales the exponent: the fp32 product is bit-identical to the reference's
        // double-promoted expression (:401-403) and avoids 3 x (F2F, 2 DMUL, F2F) on the conversion pipe per step
        const float hH = 0.5f * (float)H;
        nx = clampf((cx * mip_rbound + 1) * hH, 0.0f, (float)(H - 1));
        ny = clampf((cy * mip_rbound + 1) * hH, 0.0f, (float)(H - 1));
        nz = clampf((cz * mip_rbound + 1) * hH, 0.0f, (float)(H - 1));
    } else {
        // 0.5 is a double literal in the reference (:401-403): keep the promotion.
        nx = clampf(0.5 * (cx * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
        ny = clampf(0.5 * (cy * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
        nz = clampf(0.5 * (cz * mip_rbound + 1) * H, 0.0f, (float)(H - 1));
    }

    const uint32_t index = level * p.H3 + morton3(nx, ny, nz);
    const bool occ = p.grid[index / 8] & (1 << (index % 8));
    if (occ) return true;
    if (p.contract && mag > 1) { t += dt; return false; }
    const float tx = (((nx + 0.5f + 0.5f * sgnf(r.dx)) * p.rH * 2 - 1) * mip_bound - cx) * r.rdx;
    const float ty = (((ny + 0.5f + 0.5f * sgnf(r.dy)) * p.rH * 2 - 1) * mip_bound - cy) * r.rdy;
    const float tz = (((nz + 0.5f + 0.5f * sgnf(r.dz)) * p.rH * 2 - 1) * mip_bound - cz) * r.rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do {
        dt = clampf(t * p.dt_gamma, p.dt_min, p.dt_max);
        t += dt;
    } while (t < tt);
    return false;
}

__device__ __forceinline__ Ray load_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const size_t n) {
    Ray r;
    r.ox = rays_o[n * 3]; r.oy = rays_o[n * 3 + 1]; r.oz = rays_o[n * 3 + 2];
    r.dx = rays_d[n * 3]; r.dy = rays_d[n * 3 + 1]; r.dz = rays_d[n * 3 + 2];
    r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
    return r;
}

__host__ __device__ inline MarchParams make_params(const uint8_t* grid, float bound, bool contract, float dt_gamma, uint32_t max_steps,
                                                   uint32_t C, uint32_t H) {
    MarchParams p;
    p.grid = grid; p.bound = bound; p.contract = contract; p.dt_gamma = dt_gamma; p.C = C; p.H = H;
    p.dt_min = 2 * 1.7320508075688772f / max_steps;
    p.dt_max = 2 * 1.7320508075688772f * bound / H;
    p.rH = 1 / (float)H;
    p.h_pow2 = (H & (H - 1)) == 0;
    p.H3 = H * H * H;
    return p;
}


}  // namespace march

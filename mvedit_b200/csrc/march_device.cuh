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

// "step until next voxel" (raymarching.cu:452-455 / :818-821):  do { t += clamp(t * dt_gamma, dt_min, dt_max); } while (t < tt);
// The same floating-point operations in the same order (so t is bit-identical), but unrolled: the original loop is one
// dependent FMUL-FMNMX-FMNMX-FADD-FSETP-BRA chain per step (~50 cycles), and a ray crossing empty space takes hundreds of them,
// which made the empty-space walk of the fused renderer latency-bound (7 search trips per shading round cost more than the
// shading itself).  When tt * dt_gamma <= dt_min every step of this skip is exactly dt_min (rounding is monotone and t < tt
// before each step), and the chain collapses to one FADD per step.
__device__ __forceinline__ void step_until(float& t, const float tt, const MarchParams& p, float& dt) {
    if (p.dt_gamma >= 0.0f && tt * p.dt_gamma <= p.dt_min && p.dt_min <= p.dt_max) {
        dt = p.dt_min;
        for (;;) {
            const float t1 = t + dt, t2 = t1 + dt, t3 = t2 + dt, t4 = t3 + dt, t5 = t4 + dt, t6 = t5 + dt, t7 = t6 + dt, t8 = t7 + dt;
            if (!(t8 < tt)) {
                t = !(t1 < tt) ? t1 : !(t2 < tt) ? t2 : !(t3 < tt) ? t3 : !(t4 < tt) ? t4 : !(t5 < tt) ? t5 : !(t6 < tt) ? t6 : !(t7 < tt) ? t7 : t8;
                return;
            }
            t = t8;
        }
    }
    for (;;) {
        const float d1 = clampf(t * p.dt_gamma, p.dt_min, p.dt_max), t1 = t + d1;
        const float d2 = clampf(t1 * p.dt_gamma, p.dt_min, p.dt_max), t2 = t1 + d2;
        const float d3 = clampf(t2 * p.dt_gamma, p.dt_min, p.dt_max), t3 = t2 + d3;
        const float d4 = clampf(t3 * p.dt_gamma, p.dt_min, p.dt_max), t4 = t3 + d4;
        if (!(t4 < tt)) {
            if (!(t1 < tt)) { t = t1; dt = d1; }
            else if (!(t2 < tt)) { t = t2; dt = d2; }
            else if (!(t3 < tt)) { t = t3; dt = d3; }
            else { t = t4; dt = d4; }
            return;
        }
        t = t4;
    }
}

// Visit the cell at t. Occupied: returns true with the (contracted) sample position and dt; t is not advanced.
// Empty: advances t to the first step past the current voxel and returns false.
//
// BLOCK_SKIP (fused renderer only): 64 consecutive Morton cells are a 4x4x4 block and 8 aligned bytes of the bitfield, so ONE
// 64-bit load answers "is the whole block empty"; if so the ray steps to the block's exit plane instead of the cell's.  The
// t sequence (t += dt while t < exit) is the same sequence of increments the per-cell walk takes through those all-empty cells,
// so the next sample is the same one up to the rounding of the exit distance (computed from another point of the same ray):
// a sample that lies within an ulp of a cell face may move by one dt.  The training marcher keeps the per-cell walk and stays
// bit-identical to the reference.
template <bool BLOCK_SKIP = false>
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
    if (BLOCK_SKIP && p.h_pow2 && H >= 4 && !p.contract) {
        const unsigned long long blk = *reinterpret_cast<const unsigned long long*>(p.grid + ((index >> 6) << 3));
        if ((blk >> (index & 63u)) & 1ull) return true;
        if (blk == 0ull) {
            const float px = r.dx > 0 ? (float)((nx | 3) + 1) : (r.dx < 0 ? (float)(nx & ~3) : nx + 0.5f);
            const float py = r.dy > 0 ? (float)((ny | 3) + 1) : (r.dy < 0 ? (float)(ny & ~3) : ny + 0.5f);
            const float pz = r.dz > 0 ? (float)((nz | 3) + 1) : (r.dz < 0 ? (float)(nz & ~3) : nz + 0.5f);
            const float tx = ((px * p.rH * 2 - 1) * mip_bound - cx) * r.rdx;
            const float ty = ((py * p.rH * 2 - 1) * mip_bound - cy) * r.rdy;
            const float tz = ((pz * p.rH * 2 - 1) * mip_bound - cz) * r.rdz;
            const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
            step_until(t, tt, p, dt);
            return false;
        }
    } else {
        const bool occ = p.grid[index / 8] & (1 << (index % 8));
        if (occ) return true;
    }
    if (p.contract && mag > 1) { t += dt; return false; }
    const float tx = (((nx + 0.5f + 0.5f * sgnf(r.dx)) * p.rH * 2 - 1) * mip_bound - cx) * r.rdx;
    const float ty = (((ny + 0.5f + 0.5f * sgnf(r.dy)) * p.rH * 2 - 1) * mip_bound - cy) * r.rdy;
    const float tz = (((nz + 0.5f + 0.5f * sgnf(r.dz)) * p.rH * 2 - 1) * mip_bound - cz) * r.rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    step_until(t, tt, p, dt);
    return false;
}

// Lean traversal step of the fused renderer for the configuration every VolumeRenderer call of the reference uses
// (C == 1, no contraction, H a power of two, 4 <= H <= 256): same arithmetic as dda_step<true> on that path, expression by
// expression, minus the per-step work that does not depend on the step (mip level, contraction, sign selects) and with the
// Morton code taken from a 256-entry shared-memory table.  ~6 search trips run per shading round of the renderer, so the
// trip's instruction count is a first-order term of the render time.
// per-ray constants packed in one register: bits 0-2 direction component > 0 (x, y, z), bits 3-5 component != 0
typedef uint32_t RayAux;
__device__ __forceinline__ RayAux make_aux(const Ray& r) {
    return (r.dx > 0.f ? 1u : 0u) | (r.dy > 0.f ? 2u : 0u) | (r.dz > 0.f ? 4u : 0u) | (r.dx != 0.f ? 8u : 0u) | (r.dy != 0.f ? 16u : 0u) |
           (r.dz != 0.f ? 32u : 0u);
}
template <bool BLOCK = true>
__device__ __forceinline__ bool dda_step_lean(const Ray& r, const RayAux a, const MarchParams& p, const uint32_t* __restrict__ lut,
                                              float& t, float& cx, float& cy, float& cz, float& dt) {
    const float bound = p.bound;
    const float x = clampf(r.ox + t * r.dx, -bound, bound);
    const float y = clampf(r.oy + t * r.dy, -bound, bound);
    const float z = clampf(r.oz + t * r.dz, -bound, bound);
    const float mip_bound = fminf(1.0f, bound);
    const float mip_rbound = 1 / mip_bound;
    const float hH = 0.5f * (float)p.H, top = (float)(p.H - 1);
    const int nx = clampf((x * mip_rbound + 1) * hH, 0.0f, top);
    const int ny = clampf((y * mip_rbound + 1) * hH, 0.0f, top);
    const int nz = clampf((z * mip_rbound + 1) * hH, 0.0f, top);
    const uint32_t index = lut[nx] | (lut[ny] << 1) | (lut[nz] << 2);
    bool occ, e = false;
    if (BLOCK) {
        const unsigned long long blk = *reinterpret_cast<const unsigned long long*>(p.grid + ((index >> 6) << 3));
        occ = (blk >> (index & 63u)) & 1ull;
        e = blk == 0ull;             // the whole 4x4x4 block is empty: leave through the block's exit planes
    } else {
        occ = p.grid[index >> 3] & (1 << (index & 7u));   // per-cell walk only: bit-identical to the reference (training marcher)
    }
    if (occ) {
        cx = x; cy = y; cz = z;
        dt = clampf(t * p.dt_gamma, p.dt_min, p.dt_max);
        return true;
    }
    // exit plane per axis: cell face nx + (d > 0) [+ 0.5 when d == 0: the reference's nx + 0.5 + 0.5 * sign(d)], block face when skipping
    const int sx = a & 1u, sy = (a >> 1) & 1u, sz = (a >> 2) & 1u;
    const bool nzx = a & 8u, nzy = a & 16u, nzz = a & 32u;
    const float px = (float)((e && nzx) ? (nx & ~3) + 4 * sx : nx + sx) + (nzx ? 0.f : 0.5f);
    const float py = (float)((e && nzy) ? (ny & ~3) + 4 * sy : ny + sy) + (nzy ? 0.f : 0.5f);
    const float pz = (float)((e && nzz) ? (nz & ~3) + 4 * sz : nz + sz) + (nzz ? 0.f : 0.5f);
    const float tx = ((px * p.rH * 2 - 1) * mip_bound - x) * r.rdx;
    const float ty = ((py * p.rH * 2 - 1) * mip_bound - y) * r.rdy;
    const float tz = ((pz * p.rH * 2 - 1) * mip_bound - z) * r.rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    step_until(t, tt, p, dt);
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

// mesh_loss.cu -- the per-pixel part of mesh_optim's objective as three launches (forward + gradient in one call), in place of the
// ~60 eager torch ops and their autograd of /root/reference/lib/pipelines/mvedit_3d_pipeline.py:745-774:
//   rgb  = rgba.xyz / max(a, 1e-3);  rgb' = rgb m_erode + target (1 - m_erode)            (:746-748)
//   L_rgb   = 4.5 L1(rgb', target; w) ;  L_alpha = 2 L1(a, m_blur; w)                      (:761-770, L1LossMod weighted mean)
//   n_fg = (n - n_bg (1 - a)) / max(a, 1e-3), gradient to n scaled by the view-cosine gate  (:757-759)
//   L_tv = 2 w_reg mean_c,pixels || (d_h n_fg min(a, a_down), d_w n_fg min(a, a_right)) ||^1.5   (:771-774; tv_loss.py:7-42)
// k_mesh_loss_fwd    per pixel: rgb' (kept for the LPIPS patch term), n_fg, the two L1 sums (warp-reduced atomics)
// k_mesh_loss_tv     per pixel: the TV term of the pixel's own forward differences; its gradient goes to n_fg of the pixel and of its
//                    down / right neighbours with red.add (3 addresses per channel: no contention)
// k_mesh_loss_bwd    per pixel: everything folded back to d rgba (premultiplied rgb, alpha) and d normal
// HBM-bound elementwise / stencil work: ~100 B read + 40 B written per pixel over the three passes.  Opt-in from mesh_optim
// (fused_objective=True); like mesh_raster.cu the file also compiles as plain C++ for the CPU test-suite (host_dual.cuh).
#include "host_dual.cuh"

namespace {

struct MeshLossP {
    const float4* rgba;         // [n] antialiased render: premultiplied rgb, alpha
    const float* normal;        // [n,3] antialiased camera-space normal map (background = n_bg)
    const float* gate;          // [n] view-cosine gate of the normal gradient, or NULL (= 1)
    const float* tgt_rgb;       // [n,3]
    const float* m_erode;       // [n] 5x5-eroded target mask
    const float* m_blur;        // [n] softened target alpha
    const float* w_view;        // [bs] camera weight / mean camera weight
    float nbg[3];
    uint32_t bs, h, w;
    float c_rgb, c_alpha, c_tv; // term weight / element count (x the data-parallel share)
    float* out_rgb;             // [n,3] rgb'
    float* nfg;                 // [n,3]
    float* g_nfg;               // [n,3] zeroed by the caller of the tv pass
    const float* g_rgb_extra;   // [n,3] d(patch term) / d rgb', or NULL
    float* loss;                // [3]: rgb, alpha, tv sums (zeroed)
    float4* g_rgba;             // [n]
    float* g_normal;            // [n,3]
};

MVE_HD float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

MVE_HD void mesh_loss_fwd(const MeshLossP& p, uint32_t i, float* acc) {
    float4 c = p.rgba[i];
    float ac = c.w > 1e-3f ? c.w : 1e-3f, me = p.m_erode[i], wv = p.w_view[i / (p.h * p.w)];
    const float* t = p.tgt_rgb + (size_t)i * 3;
    const float* n = p.normal + (size_t)i * 3;
    float rgb[3] = {c.x / ac, c.y / ac, c.z / ac};
    float l = 0.f;
    for (int k = 0; k < 3; ++k) {
        float v = rgb[k] * me + t[k] * (1.f - me);
        p.out_rgb[(size_t)i * 3 + k] = v;
        l += fabsf(v - t[k]);
        p.nfg[(size_t)i * 3 + k] = (n[k] - p.nbg[k] * (1.f - c.w)) / ac;
    }
    acc[0] = p.c_rgb * wv * l;
    acc[1] = p.c_alpha * wv * fabsf(c.w - p.m_blur[i]);
}

MVE_HD void mesh_loss_tv(const MeshLossP& p, uint32_t i, float* acc) {
    uint32_t x = i % p.w, y = (i / p.w) % p.h;
    bool has_d = y + 1 < p.h, has_r = x + 1 < p.w;
    float a = p.rgba[i].w;
    float wh = has_d ? fminf(a, p.rgba[i + p.w].w) : 0.f, ww = has_r ? fminf(a, p.rgba[i + 1].w) : 0.f;
    float l = 0.f;
    for (int k = 0; k < 3; ++k) {
        float v = p.nfg[(size_t)i * 3 + k];
        float dh = has_d ? (p.nfg[(size_t)(i + p.w) * 3 + k] - v) * wh : 0.f;
        float dw = has_r ? (p.nfg[(size_t)(i + 1) * 3 + k] - v) * ww : 0.f;
        float t2 = dh * dh + dw * dw;
        if (t2 <= 0.f) continue;                              // || . || has the zero sub-gradient at 0
        float t = sqrtf(t2), rt = sqrtf(t);
        l += t * rt;                                          // t^1.5
        float kf = p.c_tv * 1.5f / rt;                        // d (c t^1.5) / d (dh, dw) = 1.5 c t^-0.5 (dh, dw)
        float gh = kf * dh * wh, gw = kf * dw * ww;
        atomic_add_f(p.g_nfg + (size_t)i * 3 + k, -(gh + gw));
        if (has_d) atomic_add_f(p.g_nfg + (size_t)(i + p.w) * 3 + k, gh);
        if (has_r) atomic_add_f(p.g_nfg + (size_t)(i + 1) * 3 + k, gw);
    }
    acc[0] = p.c_tv * l;
}

MVE_HD void mesh_loss_bwd(const MeshLossP& p, uint32_t i) {
    float4 c = p.rgba[i];
    bool live = c.w >= 1e-3f;                                 // clamp(min=1e-3) passes the gradient from the bound upwards
    float ac = c.w > 1e-3f ? c.w : 1e-3f, me = p.m_erode[i], wv = p.w_view[i / (p.h * p.w)];
    const float* t = p.tgt_rgb + (size_t)i * 3;
    float rgb[3] = {c.x / ac, c.y / ac, c.z / ac};
    float g[3], ga = p.c_alpha * wv * sgn(c.w - p.m_blur[i]), gate = p.gate ? p.gate[i] : 1.f;
    for (int k = 0; k < 3; ++k) {
        float v = rgb[k] * me + t[k] * (1.f - me);
        float gv = p.c_rgb * wv * sgn(v - t[k]) + (p.g_rgb_extra ? p.g_rgb_extra[(size_t)i * 3 + k] : 0.f);
        float grgb = gv * me;
        g[k] = grgb / ac;
        if (live) ga -= grgb * rgb[k] / ac;
        float gn = p.g_nfg[(size_t)i * 3 + k];
        p.g_normal[(size_t)i * 3 + k] = gn / ac * gate;
        ga += gn * p.nbg[k] / ac;
        if (live) ga -= gn * p.nfg[(size_t)i * 3 + k] / ac;
    }
    p.g_rgba[i] = make_float4(g[0], g[1], g[2], ga);
}

MVE_REDUCE_KERNEL(k_mesh_loss_fwd, MeshLossP, mesh_loss_fwd, 2, loss)
MVE_ELEMENT_KERNEL(k_mesh_loss_bwd, MeshLossP, mesh_loss_bwd)

// the tv pass adds its sum to loss[2]
struct MeshLossTvP { MeshLossP q; float* loss_tv; };
MVE_HD void mesh_loss_tv_w(const MeshLossTvP& p, uint32_t i, float* acc) { mesh_loss_tv(p.q, i, acc); }
MVE_REDUCE_KERNEL(k_mesh_loss_tv, MeshLossTvP, mesh_loss_tv_w, 1, loss_tv)

}  // namespace

/* Forward pass: out_rgb [n,3], nfg [n,3] and loss[0..1] (rgb, alpha sums; loss must be zeroed).  n = bs*h*w.
 * c_rgb / c_alpha / c_tv = term weight / element count, e.g. 1.2 * 4.5 / (n * 3), 1.2 * 2 / n, w_reg * 2 / (n * 3). */
MVE_EXPORT int mve_mesh_loss_forward(const float* rgba, const float* normal, const float* tgt_rgb, const float* m_erode, const float* m_blur,
                                     const float* w_view, const float* nbg_host3, uint32_t bs, uint32_t h, uint32_t w, float c_rgb, float c_alpha,
                                     float* out_rgb, float* nfg, float* loss, void* stream) {
    MVE_ARG(rgba && normal && tgt_rgb && m_erode && m_blur && w_view && nbg_host3 && out_rgb && nfg && loss, "mve_mesh_loss_forward: NULL pointer");
    MVE_ARG((unsigned long long)bs * h * w < (1ull << 30), "mve_mesh_loss_forward: too many pixels");
    MeshLossP p = {};
    p.rgba = (const float4*)rgba; p.normal = normal; p.tgt_rgb = tgt_rgb; p.m_erode = m_erode; p.m_blur = m_blur; p.w_view = w_view;
    p.nbg[0] = nbg_host3[0]; p.nbg[1] = nbg_host3[1]; p.nbg[2] = nbg_host3[2];
    p.bs = bs; p.h = h; p.w = w; p.c_rgb = c_rgb; p.c_alpha = c_alpha; p.out_rgb = out_rgb; p.nfg = nfg; p.loss = loss;
    MVE_LAUNCH(k_mesh_loss_fwd, p, bs * h * w, stream);
    return 0;
}

/* TV pass + backward: adds the TV sum to loss[2], writes g_rgba [n,4] and g_normal [n,3] (gradient of the sum of the three terms, plus
 * g_rgb_extra chained through rgb').  g_nfg [n,3] is scratch (cleared here).  gate / g_rgb_extra may be NULL. */
MVE_EXPORT int mve_mesh_loss_backward(const float* rgba, const float* tgt_rgb, const float* m_erode, const float* m_blur, const float* w_view,
                                      const float* gate, const float* nbg_host3, uint32_t bs, uint32_t h, uint32_t w, float c_rgb, float c_alpha,
                                      float c_tv, const float* nfg, const float* g_rgb_extra, float* g_nfg, float* loss, float* g_rgba,
                                      float* g_normal, void* stream) {
    MVE_ARG(rgba && tgt_rgb && m_erode && m_blur && w_view && nbg_host3 && nfg && g_nfg && loss && g_rgba && g_normal,
            "mve_mesh_loss_backward: NULL pointer");
    MeshLossTvP t = {};
    MeshLossP& p = t.q;
    p.rgba = (const float4*)rgba; p.tgt_rgb = tgt_rgb; p.m_erode = m_erode; p.m_blur = m_blur; p.w_view = w_view; p.gate = gate;
    p.nbg[0] = nbg_host3[0]; p.nbg[1] = nbg_host3[1]; p.nbg[2] = nbg_host3[2];
    p.bs = bs; p.h = h; p.w = w; p.c_rgb = c_rgb; p.c_alpha = c_alpha; p.c_tv = c_tv; p.nfg = (float*)nfg; p.g_rgb_extra = g_rgb_extra;
    p.g_nfg = g_nfg; p.g_rgba = (float4*)g_rgba; p.g_normal = g_normal;
    t.loss_tv = loss + 2;
    MVE_MEMSET(g_nfg, 0, (size_t)bs * h * w * 3 * 4, stream);
    MVE_LAUNCH(k_mesh_loss_tv, t, bs * h * w, stream);
    MVE_LAUNCH(k_mesh_loss_bwd, p, bs * h * w, stream);
    return 0;
}

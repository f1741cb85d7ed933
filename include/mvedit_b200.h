/*
 * mvedit_b200.h -- C ABI of libmvedit_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for the MVEdit multi-view denoise -> reconstruct -> render hot
 * path.  Every entry point takes raw DEVICE pointers, sizes, and
 * a cudaStream_t passed as void*; returns 0 on success or a cudaError_t value
 * (negative values: argument errors detected on the host).  No torch / ATen /
 * pybind types appear here.  All buffers are caller-allocated; kernels write in
 * place, exactly like the reference's native layer
 * (/root/reference/lib/ops/raymarching/src/raymarching.h:8-18), except that
 *   - work is enqueued on the given stream (the reference uses the legacy
 *     default stream, raymarching.cu:483),
 *   - random noise is an explicit input (the reference draws it in Python,
 *     raymarching.py:279-282),
 *   - nothing synchronises with the host.
 *
 * Each declaration cites the reference interface it replaces.
 */
#ifndef MVEDIT_B200_H
#define MVEDIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* library / device info ------------------------------------------------------ */
int mve_version(void);                 /* ABI version, currently 1 */
const char* mve_last_error(void);      /* human-readable text of the last non-zero return */

/* ---------------------------------------------------------------------------
 * B3: ray-marching ops  (replaces lib/ops/raymarching/src/raymarching.h:8-18,
 *     Python wrappers lib/ops/raymarching/raymarching.py:31-524)
 * ------------------------------------------------------------------------- */

/* raymarching.h:8 near_far_from_aabb ; kernel raymarching.cu:92-145.
 * rays_o/rays_d [N,3] f32, aabb [6] f32, nears/fars [N] f32 out. */
int mve_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                           uint32_t N, float min_near, float* nears, float* fars, void* stream);

/* raymarching.h:10-11 morton3D / morton3D_invert ; kernels raymarching.cu:214-254. */
int mve_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream);
int mve_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream);

/* raymarching.h:12 packbits ; kernel raymarching.cu:268-289.
 * grid [8*N] f32 (or f16 when grid_is_half != 0: the reference's density grid is fp16,
 * base_nerf.py:208-211, and is cast to f32 by custom_fwd before the kernel), bitfield [N] u8 out. */
int mve_packbits(const void* grid, int grid_is_half, uint32_t N, float density_thresh,
                 uint8_t* bitfield, void* stream);

/* raymarching.h:15 march_rays_train ; kernel raymarching.cu:338-475 (the reference runs it twice
 * around a host .item(), raymarching.py:286-300).  Here ONE launch does count -> block scan ->
 * one atomic per block on *counter -> write.  counter [1] i32 must be zeroed by the caller and holds
 * the total M afterwards.  Samples of a ray are contiguous; rays[n] = (offset, count).  Rays whose
 * samples would not fit in max_M keep their (offset,count) but write nothing (composite treats
 * offset+count > M as an empty ray, raymarching.cu:523).  If xyzs == NULL only rays/counter are written
 * (the reference's first pass).  noises [N] f32 may be NULL (= zeros, perturb=False).
 * dirs may be NULL (view-independent fields never read it).
 * dt_gamma_dev (optional, device, 1 float) overrides dt_gamma so a captured CUDA graph can change it between replays.
 * t_scratch (optional, device, N * max_steps floats): the counting pass records every sample's t there and the write pass
 * becomes a coalesced warp-per-ray expansion instead of a second walk of the grid (two launches; same samples, bit for bit). */
int mve_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* density_bitfield,
                         float bound, int contract, float dt_gamma, uint32_t max_steps,
                         uint32_t N, uint32_t C, uint32_t H,
                         const float* nears, const float* fars, const float* noises,
                         float* xyzs, float* dirs, float* ts, uint32_t max_M,
                         int32_t* rays, int32_t* counter, const float* dt_gamma_dev, float* t_scratch, void* stream);
/* second pass of the reference protocol: rays[n] already holds (offset,count); write the samples. */
int mve_march_rays_train_write(const float* rays_o, const float* rays_d, const uint8_t* density_bitfield,
                               float bound, int contract, float dt_gamma, uint32_t max_steps,
                               uint32_t N, uint32_t C, uint32_t H,
                               const float* nears, const float* fars, const float* noises,
                               float* xyzs, float* dirs, float* ts, uint32_t max_M,
                               const int32_t* rays, void* stream);

/* raymarching.h:16 composite_rays_train_forward ; kernel raymarching.cu:501-579.
 * M may also be given on the device: if M_dev != NULL the kernel uses min(M, *M_dev)
 * (lets the caller skip the host sync after mve_march_rays_train; M is then the buffer capacity).
 * weights [M] is fully written for every sample covered by a ray (zeros after early termination),
 * so the caller need not pre-zero it.  rgbs may be NULL (weights / weights_sum / depth only: the culling pre-pass). */
int mve_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays,
                                     uint32_t M, const int32_t* M_dev, uint32_t N, float T_thresh, int binarize,
                                     float* weights, float* weights_sum, float* depth, float* image, void* stream);

/* raymarching.h:17 composite_rays_train_backward ; kernel raymarching.cu:606-695.
 * grad_sigmas [M], grad_rgbs [M,3] are fully written for every sample covered by a ray.  grad_weights may be NULL (= 0).
 * entropy_weight_dev (optional, device scalar c) adds the gradient of the reference's sample-entropy regulariser
 * -c*entropy_scale*sum_i w_i (log max(w_i,1e-6) - log max(dt_i,1e-6))  (mvedit_3d_pipeline.py:595-603) to grad_weights on the fly. */
int mve_composite_rays_train_backward(const float* grad_weights, const float* grad_weights_sum, const float* grad_depth,
                                      const float* grad_image, const float* sigmas, const float* rgbs, const float* ts,
                                      const int32_t* rays, const float* weights_sum, const float* depth, const float* image,
                                      uint32_t M, const int32_t* M_dev, uint32_t N, float T_thresh, int binarize,
                                      const float* entropy_weight_dev, float entropy_scale,
                                      float* grad_sigmas, float* grad_rgbs, void* stream);

/* raymarching.h:19 march_rays (inference) ; kernel raymarching.cu:714-829.
 * xyzs/dirs/ts [n_alive*n_step, 3|3|2] are fully written (zeros past a ray's end), noises may be NULL. */
int mve_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                   const float* rays_o, const float* rays_d, float bound, int contract, float dt_gamma,
                   uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* density_bitfield,
                   const float* nears, const float* fars, float* xyzs, float* dirs, float* ts,
                   const float* noises, void* stream);

/* raymarching.h:20 composite_rays (inference, in place) ; kernel raymarching.cu:843-925. */
int mve_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int binarize,
                       int32_t* rays_alive, float* rays_t, const float* sigmas, const float* rgbs, const float* ts,
                       float* weights_sum, float* depth, float* image, void* stream);

/* ---------------------------------------------------------------------------
 * B2: denoiser building blocks (tcgen05 tensor cores).  These replace the cuBLAS / cuDNN calls that
 * diffusers' UNet2DConditionModel / ControlNetModel make under Adapter3DMixin.get_noise_pred*
 * (lib/pipelines/adapter3d_mixin.py:68-317; lib/models/architecture/diffusers.py:57-164).
 * All activations are bf16, NHWC / row-major; accumulation is fp32 in TMEM.
 * ------------------------------------------------------------------------- */

/* C[M,N] = act(A[M,K] . B[N,K]^T + bias[N] + row_bias[row / rows_per_group, N]) * alpha + residual[M,N]
 * A, B, C, residual bf16 (lda/ldb/ldc/ldr in elements, multiples of 8); bias/row_bias f32 or NULL (row_bias row stride ldrb, 0 = N); K % 64 == 0.
 * act: 0 none, 1 SiLU, 2 GELU(erf), 3 GEGLU: every 256 columns of B hold [128 value | 128 gate] rows of a diffusers GEGLU
 * projection and C gets the N/2 products value * gelu(gate) (ldc >= N/2; no row_bias / residual; N % 256 == 0);
 * 4 ReLU; 5 ReLU backward gate: C = residual > 0 ? (acc + bias) * alpha : 0 (residual = the forward activation, not added);
 * 6 PReLU with per-column slopes act_param [N] (f32; NULL otherwise) -- mve_conv3x3_bf16 only.
 * Replaces torch.nn.functional.linear / 1x1 conv (cuBLAS) on the UNet path. */
int mve_gemm_bf16(const void* A, const void* B, void* C, uint32_t M, uint32_t N, uint32_t K,
                  uint32_t lda, uint32_t ldb, uint32_t ldc,
                  const float* bias, const float* row_bias, uint32_t rows_per_group, uint32_t ldrb,
                  const void* residual, uint32_t ldr, int act, float alpha, const float* act_param, void* stream);

/* 3x3 convolution, stride 1, pad 1, as an implicit GEMM (no im2col buffer): X [B,H,W,Cin] bf16 NHWC,
 * Wt [Cout,3,3,Cin] bf16, Y [B*H*W, ldy] bf16.  Epilogue as mve_gemm_bf16 with rows_per_group = H*W
 * (row_bias [B,Cout] = the time-embedding projection of a ResnetBlock2D).  Cin % 64 == 0.
 * act | 0x100 allows split-K for few-tile problems with K >= 1024 (the 8^2 / 16^2 levels of the LPIPS VGG: one CTA per output tile
 * is bound by what a single SM pulls from L2): CTAs reduce K slices into an fp32 workspace with atomics and a second kernel applies
 * the epilogue -- the sums are then not bit-reproducible run to run, which is why it is opt-in.
 * Replaces torch.nn.functional.conv2d (cuDNN) on the UNet path. */
int mve_conv3x3_bf16(const void* X, const void* Wt, void* Y, uint32_t B, uint32_t H, uint32_t W,
                     uint32_t Cin, uint32_t Cout, uint32_t ldy,
                     const float* bias, const float* row_bias, uint32_t ldrb, const void* residual, uint32_t ldr,
                     int act, float alpha, const float* act_param, void* stream);

/* 3x3 convolution (pad 1, stride 1 or 2) for FEW channels on the CUDA cores: the front of diffusers' ControlNetConditioningEmbedding
 * (3->16, 16->16, 16->32 s2, 32->32, 32->96 s2 on the 512^2 condition images), where a tensor-core tile would be
 * mostly channel padding.  x: x_format 0 = NHWC bf16 [B,H,W,Cin], 1 = NCHW f32, 2 = NCHW bf16 [B,Cin,H,W]; w_packed f32
 * [9 taps][Cin][Cout]; y bf16 NHWC [B,H/stride,W/stride, ldy >= Cout] (columns >= Cout untouched); act 0 none / 1 SiLU.
 * Instantiated (Cin, Cout, stride): (3,16,1) (16,16,1) (16,32,2) (32,32,1) (32,96,2) and the test sizes (3,8,1) (8,8,1) (8,16,2) (16,32,1). */
int mve_conv3x3_direct_bf16(const void* x, int x_format, const float* w_packed, const float* bias, void* y, uint32_t B, uint32_t H,
                            uint32_t W, uint32_t Cin, uint32_t Cout, uint32_t stride, uint32_t ldy, int act, void* stream);

/* ---------------------------------------------------------------------------
 * B4: Instant-NGP field = tcnn-style HashGrid (Smoothstep, 2 features/level, fp32 table) + Linear(2L,64)+ReLU+Linear(64,4)
 * + trunc_exp(h0 + blob) / sigmoid saturation, fused.  Replaces tinycudann.Encoding fwd/bwd + torch Linear x2
 * under iNGPDecoder.point_decode / point_density_decode (lib/models/decoders/ingp_decoder.py:62-74,101-125;
 * lib/ops/activation.py:8-23).  level_* are HOST arrays of n_levels entries (scale, resolution, entries, entry offset);
 * table [n_entries,2] f32; w1 [64,2L], b1 [64], w2 [4,64], b2 [4] f32 (nn.Linear layout).
 * M_dev (optional, device): actual sample count <= M.
 * density_only: 0 = sigma+rgb (fp32 FFMA MLP), 1 = sigma only (fp32); 2 = sigma only / 3 = sigma+rgb with the MLP on tensor
 * cores in TF32 -- the reference's own matmul precision (torch allow_tf32, lib/apis/adapter3d.py:51-61).
 * ------------------------------------------------------------------------- */
int mve_field_forward(const float* xyz, uint32_t M, const int32_t* M_dev, const float* table,
                      const float* w1, const float* b1, const float* w2, const float* b2,
                      uint32_t n_levels, const float* level_scale, const uint32_t* level_res,
                      const uint32_t* level_size, const uint32_t* level_offset,
                      float bound, float blob_density, float blob_radius, float sigmoid_saturation,
                      int density_only, float* sigma, float* rgb, void* stream);

/* Backward of mve_field_forward w.r.t. table (atomically ACCUMULATED into grad_table: caller zeroes / owns .grad),
 * MLP parameters (written, or added when accumulate_mlp != 0; deterministic two-stage reduction through `workspace`
 * of mve_field_backward_workspace_floats(n_levels) floats) and optionally xyz (grad_xyz [M,3] or NULL).
 * grad_rgb may be NULL (density-only graph).  mlp_tf32 != 0: the MLP backward runs on tensor cores in TF32 (see above). */
uint32_t mve_field_backward_workspace_floats(uint32_t n_levels);
int mve_field_backward(const float* xyz, uint32_t M, const int32_t* M_dev, const float* table,
                       const float* w1, const float* b1, const float* w2, const float* b2,
                       uint32_t n_levels, const float* level_scale, const uint32_t* level_res,
                       const uint32_t* level_size, const uint32_t* level_offset,
                       float bound, float blob_density, float blob_radius, float sigmoid_saturation,
                       const float* grad_sigma, const float* grad_rgb,
                       float* grad_table, float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2,
                       int accumulate_mlp, int mlp_tf32, float* workspace, float* grad_xyz, void* stream);

/* The bare hash-grid encoding as a differentiable op: `tcnn.Encoding(n_input_dims=3, HashGrid ...)(x)` (seam B4; ingp_decoder.py:62-74,112,
 * triplane_ingp_decoder.py:102-114,150).  x01 [M,3] in [0,1]; out [M, 2*n_levels] (feature 2l+f).  The backward ACCUMULATES into
 * grad_table [n_entries,2] (atomics; may be NULL) and into grad_x [M,3] (zeroed by the caller; may be NULL). */
int mve_hashgrid_forward(const float* x01, uint32_t M, const float* table, uint32_t n_levels, const float* level_scale,
                         const uint32_t* level_res, const uint32_t* level_size, const uint32_t* level_offset, float* out, void* stream);
int mve_hashgrid_backward(const float* x01, uint32_t M, const float* table, uint32_t n_levels, const float* level_scale,
                          const uint32_t* level_res, const uint32_t* level_size, const uint32_t* level_offset,
                          const float* grad_out, float* grad_table, float* grad_x, void* stream);

/* ---------------------------------------------------------------------------
 * a-6 / a-9: fused NeRF-adapter kernels (no reference native counterpart: they replace Python loops)
 * ------------------------------------------------------------------------- */

/* Whole inference branch of VolumeRenderer.forward (lib/models/decoders/base_volume_renderer.py:264-329) + BaseNeRF.render's ray
 * generation (lib/models/autoencoders/base_nerf.py:489-556; lib/core/utils/geometry_utils.py:18-55) in one launch.
 * Rays: either rays_o/rays_d [N,3], or cameras (poses [V,4,4] c2w, intrinsics [V,4]=fx,fy,cx,cy at the render size, h, w; N = V*h*w,
 * optional dt_gamma_per_view [V]).  perturb=False semantics.  Outputs weights_sum [N], depth [N] (sum w/t), image [N,3].
 * One launch at a time per device (the work counter / statistics block is per device, zeroed stream-ordered by the launch). */
int mve_render_rays(const float* rays_o, const float* rays_d, const float* poses, const float* intrinsics,
                    const float* dt_gamma_per_view, uint32_t h, uint32_t w, uint32_t N,
                    const float* aabb, float min_near, const uint8_t* density_bitfield, float bound, float dt_gamma,
                    uint32_t max_steps, uint32_t C, uint32_t H, float T_thresh,
                    const float* table, const float* w1, const float* b1, const float* w2, const float* b2,
                    uint32_t n_levels, const float* level_scale, const uint32_t* level_res,
                    const uint32_t* level_size, const uint32_t* level_offset,
                    float blob_density, float blob_radius, float sigmoid_saturation,
                    float* weights_sum, float* depth, float* image, void* stream);

/* Diagnostic: the L2 gather ceiling the fused renderer is measured against.  blocks x 256 threads each issue per_thread (multiple of 8)
 * random 8-byte gathers from table [n_entries,2] f32 (in-register LCG indices, 8 loads in flight) and write one float to out
 * [blocks*256].  bench.py times it with CUDA events: gathers / s = blocks * 256 * per_thread / t. */
int mve_gather_ceiling(const float* table, uint32_t n_entries, uint32_t blocks, uint32_t per_thread, float* out, void* stream);

/* Statistics of the most recent mve_render_rays launch into host_out[4]: samples shaded, warp-rounds that shaded, warp-rounds total,
 * warp-level trips of the occupancy-grid search loop (host-synchronous; diagnostics / bench only). */
int mve_render_last_sample_count(uint64_t* host_out);

/* Post-render shading of denoise P2's hint images, two launches for all views (replaces ~40 elementwise torch kernels):
 * BaseNeRF.render's tail (inverse-z depth, depth / alpha, depth_to_normal: base_nerf.py:536-556, geometry_utils.py:119-148),
 * Lambert shading + background compositing (mvedit_3d_pipeline.py:1352-1380) and normalize_depth (geometry_utils.py:151-168).
 * weights_sum / depth [V,h,w], image [V,h,w,3] are mve_render_rays' outputs; intrinsics [V,4] at the render size; lights [V,3].
 * out_images / out_depths: bf16 [V,3,h,w] (NCHW), both clamped to [0,1]; reduce_scratch: [V,2] i32 of device scratch.
 * out_normals_fg (optional): f32 [V,h,w,3], depth_to_normal(depth_fg) in the opengl [0,1] encoding (BaseNeRF.render's
 * compute_normal output, base_nerf.py:549-553).  With out_images == NULL only the normals are produced (lights / scratch unused).
 * tonemap_knots (HOST array [lut_x (n) | lut_y (n)], 2 <= n <= 32, or NULL): the reference's Tonemapping module
 * (lib/models/decoders/tonemapping.py:5-52) -- shading is then applied in tone-mapped space,
 * lut(inverse_lut(rgb / alpha) + log2(shading)) * alpha + bg (1 - alpha) (mvedit_3d_pipeline.py:1377-1384). */
int mve_shade_views(const float* weights_sum, const float* depth, const float* image, const float* intrinsics, const float* lights,
                    uint32_t V, uint32_t h, uint32_t w, float ambient, float bg_color, float far_depth, float alpha_clip, float eps,
                    int32_t* reduce_scratch, void* out_images, void* out_depths, float* out_normals_fg,
                    const float* tonemap_knots, uint32_t tonemap_n, void* stream);

/* Weight culling of the training branch (base_volume_renderer.py:222-246): keep samples with weight > th, compact xyzs/ts,
 * rebuild rays (offset,count); *counter (zeroed by the caller) receives the kept total.  Rays whose kept samples would not fit
 * in the M_out_cap-sample output buffers are emitted as empty. */
int mve_cull_samples(const float* weights, float th, const int32_t* rays_in, const float* xyzs_in, const float* ts_in,
                     uint32_t N, uint32_t M, const int32_t* M_dev, uint32_t M_out_cap,
                     int32_t* rays_out, float* xyzs_out, float* ts_out, int32_t* counter, void* stream);

/* Occupancy-grid refresh tail of VolumeRenderer.update_extra_state (base_volume_renderer.py:163-175): fp16 EMA
 * grid = where(grid>=0 & s>=0, max(grid*decay, s), grid), mean of clamp(grid,0), packbits with min(mean, density_thresh).
 * sigmas [n_cells] are the freshly decoded densities in Morton order (indices must be NULL: full update, the only mode
 * the pipelines reach -- SURVEY.md Appendix F).  sum_scratch: 1 float of device scratch. */
int mve_density_grid_update(void* grid_half, const float* sigmas, const int32_t* indices, uint32_t n, float decay,
                            float* sum_scratch, uint32_t n_cells, float density_thresh, uint8_t* bitfield, void* stream);

/* Flash attention (tcgen05): O[b, i, h, :] = softmax_j(scale * Q[b,i,h,:].K[b,j,h,:]) V[b,j,h,:].
 * Q [batch, q_len, heads, d] with row stride ldq (elements; e.g. the q slice of a fused qkv GEMM output), K/V [batch, kv_len, heads, d]
 * (ldk/ldv), O [batch, q_len, heads*d] (ldo).  bf16, d in {40,64,80,128,160}.  Replaces F.scaled_dot_product_attention under
 * AttnProcessor2_0 / CrossImageAttnProcWrapper (lib/models/architecture/joint_attn.py:11-37). */
int mve_attention_bf16(const void* Q, const void* K, const void* V, void* O, uint32_t batch, uint32_t heads,
                       uint32_t q_len, uint32_t kv_len, uint32_t d, uint32_t ldq, uint32_t ldk, uint32_t ldv, uint32_t ldo,
                       float scale, void* stream);

/* HBM-bound glue of the UNet path (bf16, NHWC / row-major).  Replace torch.nn.GroupNorm / LayerNorm / GEGLU / F.interpolate /
 * strided-conv unfolding inside diffusers ResnetBlock2D, Transformer2DModel, Upsample2D, Downsample2D. */
/* x,y [B,HW,C]; G groups (32); stats_scratch: B*G*2 floats of device scratch. */
int mve_groupnorm_bf16(const void* x, void* y, uint32_t B, uint32_t HW, uint32_t C, uint32_t G,
                       const float* gamma, const float* beta, float eps, int silu_act, float* stats_scratch, void* stream);
int mve_layernorm_bf16(const void* x, void* y, uint32_t rows, uint32_t C, const float* gamma, const float* beta, float eps, void* stream);
/* h [M,2F] -> y [M,F] = h[:, :F] * gelu(h[:, F:]) */
int mve_geglu_bf16(const void* h, void* y, uint64_t M, uint32_t F, void* stream);
int mve_upsample2x_bf16(const void* x, void* y, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* stream);
/* x [B,H,W,C] -> y [B*(H/2)*(W/2), 9*C] patches of a 3x3 stride-2 conv (K index = tap*C + c).  pad_lo = 1: pad 1 on every side
 * (UNet Downsample2D); pad_lo = 0: zero pad on the right / bottom only (AutoencoderKL encoder Downsample2D, F.pad (0,1,0,1)). */
int mve_im2col3x3s2_bf16(const void* x, void* y, uint32_t B, uint32_t H, uint32_t W, uint32_t C, int pad_lo, void* stream);
/* y[r, :cols] = softmax(scale * x[r, :cols]) per row (bf16 in / out, fp32 inside; in place allowed).  The AutoencoderKL mid-block
 * attention (one head, d = 512: lib/pipelines/mvedit_3d_pipeline.py:1119-1120,1258-1263 -> diffusers Attention) runs as
 * score GEMM -> this -> value GEMM. */
int mve_softmax_rows_bf16(const void* x, void* y, uint32_t rows, uint32_t cols, uint32_t ldx, uint32_t ldy, float scale, void* stream);
/* ---- LPIPS(net='vgg') patch loss of the reconstruction objective, forward AND gradient w.r.t. the rendered patch
 * (lib/models/losses/lpips_loss.py:14-43 -> lpips==0.1.4 LPIPS.forward; lib/pipelines/mvedit_3d_pipeline.py:611-617,796-801).
 * The 13 VGG16 convolutions and their 13 input-gradient convolutions are mve_conv3x3_bf16 calls (act 4 = ReLU; act 5 = ReLU
 * backward gate: out = residual > 0 ? acc * alpha : 0, residual = the forward activation); the rest of the graph is these kernels.
 *   prep:        pred / target [n_pix_each,3] f32 in [0,1] -> out [2*n_pix_each,64] bf16 = ((2x-1) - shift) / scale, channels 3.. zero
 *   input_grad:  g64 [n_pix,64] bf16 (gradient of prep's output, pred half) -> g_pred [n_pix,3] f32 (times 2 / scale)
 *   maxpool2x2:  x [B,H,W,C] -> y [B,H/2,W/2,C]
 *   maxpool2x2_relu_backward: g_x [B,H,W,C] = (x > 0) ? g_x + (g_y routed to each window's first maximum) : 0   (in place)
 *   lpips_layer: feat [2P,HW,C] bf16 (pred images then target images), lin_w [C] f32 >= 0, gscale [P] f32 (d total / d lpips[img]):
 *                loss[img] += mean_pixels sum_c lin_w[c] (f_p/(|f_p|+1e-10) - f_t/(|f_t|+1e-10))_c^2   (atomicAdd; caller zeroes),
 *                g_feat [P,HW,C] bf16 = gscale[img] * d loss[img] / d f_p, gated by f_p > 0. */
int mve_lpips_prep(const float* pred, const float* target, uint32_t n_pix_each, void* out, void* stream);
int mve_lpips_input_grad(const void* g64, uint32_t n_pix, float* g_pred, void* stream);
int mve_maxpool2x2_bf16(const void* x, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* y, void* stream);
int mve_maxpool2x2_relu_backward_bf16(const void* x, const void* g_y, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* g_x,
                                      void* stream);
int mve_lpips_layer(const void* feat, uint32_t P, uint32_t HW, uint32_t C, const float* lin_w, const float* gscale, float* loss,
                    void* g_feat, void* stream);
/* Tail of SRVGGNetCompact.forward, the render enhancer below 512^2 (lib/models/decoders/image_space_ss.py:66-75;
 * mvedit_3d_pipeline.py:1398-1401): out [B,C,rH,rW] bf16 NCHW = PixelShuffle(r)(y) + nearest-upsampled x, with y [B,H,W,ldy] bf16 NHWC
 * holding channel c r^2 + i r + j of the last convolution and x [B,H,W,ldx] bf16 NHWC the network input (first C channels).  The 34
 * convolutions before it are mve_conv3x3_bf16 calls with the PReLU epilogue (act 6). */
int mve_pixel_shuffle_add_bf16(const void* y, const void* x, uint32_t B, uint32_t H, uint32_t W, uint32_t C, uint32_t r,
                               uint32_t ldy, uint32_t ldx, void* out, void* stream);
/* x [B,C,HW] (f32 or bf16) -> y [B,HW,Cpad] bf16, channels >= C zero-filled */
int mve_nchw_to_nhwc_pad_bf16(const void* x, int x_is_f32, void* y, uint32_t B, uint32_t C, uint32_t HW, uint32_t Cpad, void* stream);

/* Fused objective of one nerf_optim iteration on P patches of ps x ps rays (mvedit_3d_pipeline.py:541-603 without target normals /
 * depths / tonemapping): depth->normals, Lambert shading, L1 rgb, L1 alpha, TV^1.5 normal regulariser, background entropy.
 * Produces the 5 loss terms (total, rgb, alpha, normal_reg, bg-entropy) and d/d(image, weights_sum, depth).
 * w_* are device scalars (schedule dependent).  scratch: mve_nerf_patch_loss_scratch_floats(N) floats.
 * g_out_extra [N,3] or NULL: the gradient of further terms w.r.t. the shaded, composited rgb (out = image * shading + bg (1 - alpha))
 * -- the LPIPS patch term (:611-617), evaluated by the caller on mve_nerf_patch_out_rgb's output -- chained through shading /
 * compositing together with the L1 term.
 * tonemap_knots / tonemap_n as in mve_shade_views: with shading on, the shaded colour is lut(inverse_lut(image / alpha) + log2(shading))
 * (mvedit_3d_pipeline.py:564-570) and the gradients follow the piecewise-linear curve. */
uint32_t mve_nerf_patch_loss_scratch_floats(uint32_t n_rays);
int mve_nerf_patch_loss(const float* image, const float* alpha, const float* depth, const float* tgt_rgb, const float* tgt_mask,
                        const float* dirs, const float* patch_w, const float* lights, uint32_t n_patches, uint32_t patch_size,
                        int shaded, float ambient, float bg_color, float bg_width, float pixel_loss_weight,
                        const float* w_alpha_mul, const float* w_normal_reg, const float* w_entropy,
                        float* scratch, float* g_image, float* g_alpha, float* g_depth, float* loss5, const float* g_out_extra,
                        const float* tonemap_knots, uint32_t tonemap_n, void* stream);
/* The same objective with the optional targets of image-to-3D runs (mvedit_3d_pipeline.py:462-463; every pointer may be NULL):
 *   tgt_normal [N,3]      opengl normals in [0,1] (:518-519, :527): the TV term becomes diff(normal_fg) - diff(tgt_normal) (:582-585,
 *                         lib/models/losses/tv_loss.py:27)
 *   g_normal_extra [N,3]  gradient of further terms w.r.t. out_normal = normal_fg * alpha + normal_bg * (1 - alpha) (:553-554) -- the
 *                         high-passed normal patch term (:619-626), evaluated by the caller on mve_nerf_patch_out_normal's output
 *   tgt_depth [N]         target 1/z (:520-521, :529): L1 on depth * |dir| weighted like the rgb term, x w_depth[0] (device scalar,
 *                         depth_weight); its value is written to loss_depth[0] and added to loss5[0] (:586-592) */
int mve_nerf_patch_loss_targets(const float* image, const float* alpha, const float* depth, const float* tgt_rgb, const float* tgt_mask,
                                const float* dirs, const float* patch_w, const float* lights, uint32_t n_patches, uint32_t patch_size,
                                int shaded, float ambient, float bg_color, float bg_width, float pixel_loss_weight,
                                const float* w_alpha_mul, const float* w_normal_reg, const float* w_entropy, float* scratch, float* g_image,
                                float* g_alpha, float* g_depth, float* loss5, const float* g_out_extra, const float* tonemap_knots,
                                uint32_t tonemap_n, const float* tgt_normal, const float* g_normal_extra, float normal_bg_x,
                                float normal_bg_y, float normal_bg_z, const float* tgt_depth, const float* w_depth, float* loss_depth,
                                void* stream);
/* out_normal [N,3]: depth -> normal_fg (geometry_utils.depth_to_normal :119-148) composited over normal_bg with alpha (:553-554). */
int mve_nerf_patch_out_normal(const float* alpha, const float* depth, const float* dirs, uint32_t n_patches, uint32_t patch_size,
                              float normal_bg_x, float normal_bg_y, float normal_bg_z, float* scratch, float* out_normal, void* stream);
/* out_rgb [N,3]: what the pixel and patch losses compare with the target (mvedit_3d_pipeline.py:558-571); same scratch. */
int mve_nerf_patch_out_rgb(const float* image, const float* alpha, const float* depth, const float* dirs, const float* lights,
                           uint32_t n_patches, uint32_t patch_size, int shaded, float ambient, float bg_color, float* scratch,
                           float* out_rgb, const float* tonemap_knots, uint32_t tonemap_n, void* stream);

/* ---------------------------------------------------------------------------
 * a-5: glue of one nerf_optim iteration (lib/pipelines/mvedit_3d_pipeline.py:507-536, :631-633)
 * ------------------------------------------------------------------------- */

/* Rays and targets of the drawn patches in one launch: replaces BaseNeRF.ray_sample's whole-image patch reshuffle
 * (lib/models/autoencoders/base_nerf.py:245-303), get_ray_directions / get_rays (lib/core/utils/geometry_utils.py:18-55) and the
 * per-patch weight / light / dt_gamma ops (mvedit_3d_pipeline.py:516-536).
 * patch_inds [P] int64: patch ids numbered (view, patch row, patch col) as ray_sample numbers them; poses_R [V,3,3], poses_T [V,3]
 * (c2w), intrinsics [V,4] at intrinsics_size, intrinsics_scale = render_size / intrinsics_size; images [V,rs,rs,3], masks [V,rs,rs].
 * Outputs for the FULL patches (n = P*ps*ps, patch-major, row-major inside a patch): dirs [n,3] camera-space directions (z = 1),
 * tgt_rgb [n,3], tgt_mask [n], patch_w [P] = cam_w / mean(cam_w), patch_lights [P,3], dt_gamma [1] (first patch's view, as the
 * one-scene reference uses only element 0: base_volume_renderer.py:212-218).
 * rays_o / rays_d [P*(row_hi-row_lo)*ps, 3]: world-space rays of patch rows [row_lo, row_hi) only -- the strip this rank marches
 * in the data-parallel reconstruction (row_lo = 0, row_hi = ps on one GPU). */
int mve_patch_rays(const int64_t* patch_inds, uint32_t P, uint32_t V, uint32_t render_size, uint32_t patch_size,
                   const float* poses_R, const float* poses_T, const float* intrinsics, float intrinsics_scale,
                   const float* images, const float* masks, const float* cam_weights, const float* cam_lights,
                   float dt_gamma_scale, uint32_t row_lo, uint32_t row_hi,
                   float* rays_o, float* rays_d, float* dirs, float* tgt_rgb, float* tgt_mask,
                   float* patch_w, float* patch_lights, float* dt_gamma, void* stream);

/* torch.optim.Adam.step (defaults: no weight decay, no amsgrad) + optimizer.zero_grad for up to 16 tensors in one launch
 * (mvedit_3d_pipeline.py:631-633): params / grads / exp_avg / exp_avg_sq / lr are HOST arrays of n_tensors DEVICE pointers
 * (lr[i] points at a device float: schedulable inside a captured graph), numel a host array.  *step (device int32) is incremented
 * first and used for the bias corrections.  grads are scaled by grad_scale before use and zeroed afterwards when zero_grad != 0. */
int mve_adam_step(uint32_t n_tensors, void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                  const uint32_t* numel, const float* const* lr, float beta1, float beta2, float eps, float grad_scale,
                  int32_t* step, int zero_grad, void* stream);

/* ---------------------------------------------------------------------------
 * a-12: 3D Gaussian-splatting rasteriser -- tile binning + alpha blending.  The reference snapshot does not
 * contain its 3DGS code (README.md:121 names ashawkey/diff-gaussian-rasterization as the upstream): these follow that public
 * algorithm (restated in oracle/gs_oracle.py).  Per-Gaussian inputs come from the projection step (mvedit_b200.gs_renderer):
 * xy [P,2] projected means in pixel-index coordinates, conic_opacity [P,4] = (A, B, C of the inverse 2-D covariance, opacity),
 * feat [P,4] = (r, g, b, camera depth), rect [P,4] i32 = tile rect (x0, y0, x1, y1) of the 3-sigma radius (empty = culled).
 * ------------------------------------------------------------------------- */

/* One (tile << 32 | depth bits, gaussian id) pair per touched 16x16 tile.  offsets [P] = inclusive prefix sum of the tile counts
 * (int64); keys / vals have offsets[P-1] entries.  grid_x = tiles per image row. */
int mve_gs_duplicate_keys(const int32_t* rect, const float* depth, const int64_t* offsets, uint32_t P, uint32_t grid_x,
                          int64_t* keys, int32_t* vals, void* stream);
/* ranges [n_tiles,2] (zeroed by the caller) <- [start, end) of every tile in the SORTED key list. */
int mve_gs_tile_ranges(const int64_t* keys_sorted, uint32_t L, int32_t* ranges, void* stream);
/* Front-to-back blending, one 256-thread CTA per tile: alpha = min(0.99, o exp(-0.5 d^T conic d)), alpha < 1/255 skipped, stop before
 * T < 1e-4.  bg_host3: HOST array of 3 floats.  Outputs: out_color [H,W,3] (background composited with the final T), out_depth [H,W]
 * (sum alpha T depth), out_alpha [H,W] = 1 - T, final_T [H,W], n_contrib [H,W] (index of the last contributor, for the backward). */
int mve_gs_blend_forward(const int32_t* ranges, const int32_t* point_list, const float* xy, const float* conic_opacity, const float* feat,
                         const float* bg_host3, uint32_t W, uint32_t H, float* out_color, float* out_depth, float* out_alpha,
                         float* final_T, int32_t* n_contrib, void* stream);
/* Backward of the blend w.r.t. xy, conic_opacity, feat (ACCUMULATED with warp-reduced atomics into zeroed buffers).
 * g_depth / g_alpha may be NULL. */
int mve_gs_blend_backward(const int32_t* ranges, const int32_t* point_list, const float* xy, const float* conic_opacity, const float* feat,
                          const float* bg_host3, uint32_t W, uint32_t H, const float* final_T, const int32_t* n_contrib,
                          const float* g_color, const float* g_depth, const float* g_alpha,
                          float* d_xy, float* d_conic_opacity, float* d_feat, void* stream);

/* ---------------------------------------------------------------------------
 * B5: triangle-mesh rasteriser -- the nvdiffrast.torch ops MeshRenderer uses
 * (lib/models/decoders/mesh_renderer/base_mesh_renderer.py:204 RasterizeCudaContext, :241-242 dr.rasterize, :247-278 dr.interpolate,
 * :296-298 dr.antialias; texture-space rasterisation :407-410, :442, :521).  nvdiffrast is an un-vendored dependency
 * (requirements.txt:3): semantics per SURVEY.md Appendix C, restated in oracle/raster_oracle.py.  Instance mode only
 * (one topology, B views; the reference's "range mode" needs num_scenes > 1, which MVEdit never uses).
 * pos [B,V,4] (pos_batched != 0) or [V,4] clip-space f32; tri [F,3] i32; images are [B,H,W,*] f32, row y <-> NDC y = (2y+1)/H - 1.
 * ------------------------------------------------------------------------- */

/* dr.rasterize: rast [B,H,W,4] = (u, v, z/w, triangle id + 1; 0 = empty), u / v = perspective-correct barycentrics of vertex 0 / 1;
 * rast_db [B,H,W,4] = (du/dX, du/dY, dv/dX, dv/dY) per pixel, or NULL to skip it.  A pixel is covered when its centre is inside or on
 * the triangle (homogeneous edge functions, both windings) AND inside the triangle's fp32 screen bounding box padded by one pixel (this
 * pins down zero-area triangles, whose edge functions are rounding noise).  Nearest z/w in [-1, 1] wins, ties go to the
 * lower triangle id (deterministic).  Triangles with a vertex at w <= 0 are dropped (no near-plane clipping).
 * Scratch (caller-allocated, contents irrelevant): zbuf [B*H*W] u64, queue [1 + B*F] u32.  Three launches, no host sync. */
int mve_rasterize_fwd(const float* pos, const int32_t* tri, uint32_t B, uint32_t V, uint32_t F, uint32_t H, uint32_t W,
                      int pos_batched, void* zbuf, uint32_t* queue, float* rast, float* rast_db, void* stream);
/* d rast (u, v, z/w; the id channel and rast_db carry no gradient) -> g_pos [B or 1, V, 4], ACCUMULATED into a zeroed buffer. */
int mve_rasterize_bwd(const float* pos, const int32_t* tri, uint32_t B, uint32_t V, uint32_t F, uint32_t H, uint32_t W,
                      int pos_batched, const float* rast, const float* g_rast, float* g_pos, void* stream);

/* dr.interpolate: out [B,H,W,C] = u a0 + v a1 + (1-u-v) a2 over attr [B or 1, Va, C] with the attribute triangles tri [F,3]
 * (zeros on empty pixels); out_da [B,H,W,2C] = (d/dX, d/dY) of every attribute (diff_attrs='all'; needs rast_db) or NULL. */
int mve_interpolate_fwd(const float* attr, const int32_t* tri, const float* rast, const float* rast_db, uint32_t B, uint32_t H,
                        uint32_t W, uint32_t Va, uint32_t F, uint32_t C, int attr_batched, float* out, float* out_da, void* stream);
/* g_out -> g_attr (ACCUMULATED into a zeroed buffer) and g_rast [B,H,W,4] = (d/du, d/dv, 0, 0) (written; may be NULL). */
int mve_interpolate_bwd(const float* attr, const int32_t* tri, const float* rast, uint32_t B, uint32_t H, uint32_t W, uint32_t Va,
                        uint32_t F, uint32_t C, int attr_batched, const float* g_out, float* g_attr, float* g_rast, void* stream);

/* dr.antialias: out = color + silhouette blends.  opp [F,3] i32: for edge k of a triangle (the edge facing its vertex k) the vertex
 * opposite to that edge in the adjacent triangle, -1 on an open edge (mvedit_b200.mesh_raster.edge_opposites; it stands in for
 * nvdiffrast's internal edge hash). */
int mve_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri, const int32_t* opp, uint32_t B,
                      uint32_t H, uint32_t W, uint32_t C, uint32_t V, uint32_t F, int pos_batched, float* out, void* stream);
/* g_out -> g_color [B,H,W,C] (written) and g_pos [B or 1, V, 4] (ACCUMULATED into a zeroed buffer; may be NULL). */
int mve_antialias_bwd(const float* color, const float* rast, const float* pos, const int32_t* tri, const int32_t* opp, uint32_t B,
                      uint32_t H, uint32_t W, uint32_t C, uint32_t V, uint32_t F, int pos_batched, const float* g_out,
                      float* g_color, float* g_pos, void* stream);

/* dr.texture(tex, uv, uv_da=, filter_mode='linear-mipmap-linear' | 'linear'), boundary mode 'wrap' (nvdiffrast's default; the reference
 * never passes another): base_mesh_renderer.py:263-264 (albedo), :470-475 / :547-552 (coverage of a ones map through the texture
 * GRADIENT), :499-500 / :576-577 (images resampled into texture space).  The mip pyramid is ONE flat f32 buffer: level l holds
 * [Bt, th >> l, tw >> l, C] at the offset mve_texture_pyramid_floats(Bt, th, tw, C, l) (2x2 box averages; every halved level needs
 * even dimensions).  The caller copies the texture into level 0 and calls mve_texture_mip_build once per texture. */
unsigned long long mve_texture_pyramid_floats(uint32_t Bt, uint32_t th, uint32_t tw, uint32_t C, uint32_t n_levels);
int mve_texture_mip_build(float* pyr, uint32_t Bt, uint32_t th, uint32_t tw, uint32_t C, uint32_t n_levels, void* stream);
/* out [B,H,W,C]; uv [B,H,W,2] in texture units ((0,0) = corner of texel (0,0)); uv_da [B,H,W,4] = (du/dX, du/dY, dv/dX, dv/dY) from
 * mve_interpolate_fwd's out_da, or NULL (level 0 only).  Level = clamp(0.5 log2(major axis^2 of the pixel footprint in texels), 0,
 * n_levels - 1), trilinear between floor(level) and the next.  Bt = 1 (one texture for all images) or B. */
int mve_texture_fwd(const float* pyr, uint32_t Bt, uint32_t th, uint32_t tw, uint32_t C, uint32_t n_levels, const float* uv,
                    const float* uv_da, uint32_t B, uint32_t H, uint32_t W, float* out, void* stream);
/* g_out -> g_pyr (same layout as the pyramid, zeroed by the caller): taps scattered with atomics, then the coarse levels' gradients
 * folded down so that level 0 of g_pyr is d/d(texture).  No gradient to uv / uv_da (MVEdit optimises either geometry or texture). */
int mve_texture_bwd(uint32_t Bt, uint32_t th, uint32_t tw, uint32_t C, uint32_t n_levels, const float* uv, const float* uv_da,
                    uint32_t B, uint32_t H, uint32_t W, const float* g_out, float* g_pyr, void* stream);

/* ---------------------------------------------------------------------------
 * a-10: the per-pixel part of mesh_optim's objective (lib/pipelines/mvedit_3d_pipeline.py:745-774), loss and gradient in two calls
 * (three launches).  Images are [bs,h,w,*] f32: rgba = the antialiased render (premultiplied rgb, alpha), normal = its camera-space
 * normal map, tgt_rgb / m_erode / m_blur the target colours, the 5x5-eroded target mask and the softened target alpha, w_view [bs] =
 * camera weight / mean weight, nbg_host3 = HOST array of the 3 background-normal values.  c_* = term weight / element count
 * (x the data-parallel share): L1LossMod(1.2) * 4.5 / (n*3), 1.2 * 2 / n, normal_reg_weight * 2 / (n*3).
 * ------------------------------------------------------------------------- */
/* out_rgb [n,3] = rgb' (what the LPIPS patch term looks at), nfg [n,3] (kept for the backward), loss[0..1] += the rgb / alpha sums. */
int mve_mesh_loss_forward(const float* rgba, const float* normal, const float* tgt_rgb, const float* m_erode, const float* m_blur,
                          const float* w_view, const float* nbg_host3, uint32_t bs, uint32_t h, uint32_t w, float c_rgb, float c_alpha,
                          float* out_rgb, float* nfg, float* loss, void* stream);
/* loss[2] += the TV-normal sum; g_rgba [n,4], g_normal [n,3] written.  gate [n] (view-cosine gate of the normal gradient) and
 * g_rgb_extra [n,3] (d patch term / d rgb') may be NULL; g_nfg [n,3] is scratch. */
int mve_mesh_loss_backward(const float* rgba, const float* tgt_rgb, const float* m_erode, const float* m_blur, const float* w_view,
                           const float* gate, const float* nbg_host3, uint32_t bs, uint32_t h, uint32_t w, float c_rgb, float c_alpha,
                           float c_tv, const float* nfg, const float* g_rgb_extra, float* g_nfg, float* loss, float* g_rgba,
                           float* g_normal, void* stream);

/* opp [F,3] of mve_antialias_* by hashing instead of sorting (opt-in; mvedit_b200.mesh_raster.edge_opposites(method='hash')): the
 * undirected edges of the 3F (triangle, edge) entries go into an open-addressing table of `slots` 64-bit keys (a power of two
 * >= 6 F; keys [slots] u64, first / second [slots] u32, slot_of [3F] u32 are scratch), the two smallest entry indices per edge pair
 * up -- the same rule as the stable sort, so both methods return the same table.  Vertex indices must be >= 0. */
int mve_edge_opposites(const int32_t* tri, uint32_t F, uint32_t slots, void* keys, uint32_t* first, uint32_t* second,
                       uint32_t* slot_of, int32_t* opp, void* stream);

/* ---------------------------------------------------------------------------
 * a-10: decimation of the final DMTet mesh (lib/pipelines/mvedit_3d_pipeline.py:829-844)
 * ------------------------------------------------------------------------- */

/* Quadric-error edge-collapse decimation ON THE HOST (all pointers are HOST pointers; this entry launches no kernel): replaces
 * open3d's TriangleMesh.simplify_quadric_decimation(target, boundary_weight=0) that mesh_optim calls on the CPU when mesh_reduction < 1
 * at the last step.  verts [V,3], faces [F,3] int32 -> out_verts (room for V), out_faces (room for F), out_counts = {n_verts, n_faces}.
 * Stops at target_faces or when no legal collapse is left (link condition + normal-flip rejection keep a closed manifold closed). */
int mve_mesh_simplify(const float* verts, uint32_t V, const int32_t* faces, uint32_t F, uint32_t target_faces, float* out_verts,
                      int32_t* out_faces, uint32_t* out_counts);

#ifdef __cplusplus
}
#endif
#endif /* MVEDIT_B200_H */

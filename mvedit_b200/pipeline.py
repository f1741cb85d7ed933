"""One iteration of the MVEdit hot loop -- denoise P1 -> (decode) -> NeRF reconstruction -> render all views -> denoise P2 -> solver
step -- as the reference runs it at /root/reference/lib/pipelines/mvedit_3d_pipeline.py:1141-1479 (t != None branch, NeRF stage,
'2-pass' mode, blend_weight 0), built on the B200 components of this package.

What is and is not inside ``MVEdit3DStep.step``:
  * inside: scale_model_input, get_noise_pred_p1 (UNet enc+dec, all views one batch), pred_x0, nerf_optim (n_inverse_steps Adam
    iterations of march/field/composite fwd+bwd), nerf.render of all views + Lambert shading + normalize_depth, get_noise_pred_p2
    (tile+depth ControlNets on the fresh renders + UNet dec), Euler-ancestral solver step;
  * hooks (neighbours of the path): ``decode_fn(pred_x0) -> (tgt_images, tgt_masks)`` stands where
    vae.decode + TRACER masks are in the reference (:1258-1266); LPIPS ``patch_loss`` on the NeRF; the SRVGG enhancer (only
    active below 512^2 renders).  bench.py passes a synthetic decode_fn and says so in its JSON.

View sharding: with ``torch.distributed`` initialised every rank denoises / renders its slice of the views and the
decoded targets are exchanged with ONE all_gather per step (``view_shard.gather_views``); the reconstruction runs replicated and
rank 0's field is broadcast afterwards so that replicas cannot drift through atomic-order noise.
"""
import math

import numpy as np
import torch

from .adapter3d_mixin import Adapter3DMixin
from .nerf import nerf_optim, normalize_depth
from . import view_shard
from ._lib import call, ptr, stream, c_u32, c_f32


def get_noise_scales(alphas_bar, t, num_timesteps, dtype=torch.float32):
    """lib/core/diffusion.py:4-21."""
    alphas_bar = t.new_tensor(alphas_bar, dtype=torch.float32)
    if t.is_floating_point():
        int_t = t.long()
        frac_t = t - int_t
        a0 = alphas_bar[int_t]
        a1 = alphas_bar[(int_t + 1).clamp(max=num_timesteps - 1)]
        s0, s1 = torch.sqrt((1 - a0) / a0), torch.sqrt((1 - a1) / a1)
        ve = s0 * (1 - frac_t) + s1 * frac_t
        return torch.sqrt(1 / (1 + ve ** 2)).to(dtype), torch.sqrt(ve ** 2 / (1 + ve ** 2)).to(dtype)
    a = alphas_bar[t]
    return torch.sqrt(a).to(dtype), torch.sqrt(1 - a).to(dtype)


class EulerAncestralScheduler:
    """diffusers EulerAncestralDiscreteScheduler as the reference configures it for SD1.5 (scaled_linear betas 0.00085 -> 0.012,
    1000 train steps, epsilon prediction, timestep_spacing='trailing': lib/apis/adapter3d.py:280-300).
    The ancestral noise is an explicit argument of ``step`` so that oracle and kernels see identical draws."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        self.num_train_timesteps = num_train_timesteps
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
        self.alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(self.alphas)
        self.init_noise_sigma = None

    def set_timesteps(self, num_inference_steps, device='cpu'):
        step_ratio = self.num_train_timesteps / num_inference_steps
        ts = np.round(np.arange(self.num_train_timesteps, 0, -step_ratio)) - 1
        sig = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        sigmas = np.interp(ts, np.arange(len(sig)), sig)
        self.sigmas = torch.tensor(np.concatenate([sigmas, [0.0]]), dtype=torch.float32, device=device)
        self.timesteps = torch.tensor(ts, dtype=torch.float32, device=device)
        self.init_noise_sigma = float(self.sigmas.max())
        self._step_index = 0

    def scale_model_input(self, sample, i):
        return sample / ((self.sigmas[i] ** 2 + 1) ** 0.5)

    def step(self, model_output, i, sample, noise):
        sigma = self.sigmas[i]
        pred_original_sample = sample - sigma * model_output
        sigma_to = self.sigmas[i + 1]
        sigma_up = (sigma_to ** 2 * (sigma ** 2 - sigma_to ** 2) / sigma ** 2) ** 0.5
        sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
        derivative = (sample - pred_original_sample) / sigma
        return sample + derivative * (sigma_down - sigma) + noise * sigma_up


class MVEdit3DStep(Adapter3DMixin):
    """The loop body of MVEdit3DPipeline.__call__ (NeRF stage) on B200 components."""

    def __init__(self, unet, controlnet, nerf, scheduler, tonemapping=None, normal_bg=(0.5, 0.5, 1.0)):
        self.unet, self.controlnet, self.nerf, self.scheduler = unet, controlnet, nerf, scheduler
        self.tonemapping = tonemapping
        self.normal_bg = list(normal_bg)

    # ------------------------------------------------------------------ render all (local) views, mvedit_3d_pipeline.py:1341-1395
    def render_views(self, density_bitfield, camera_poses, intrinsics, intrinsics_size, render_size, cam_lights, ambient_light,
                     testmode_dt_gamma_scale):
        """-> (ctrl_images, ctrl_depths), bf16 [V,3,rs,rs] in [0,1].  One fused render launch + mve_shade_views (two launches) when no
        tone mapping is configured; ``render_views_torch`` is the op-by-op restatement of the reference it is tested against."""
        if self.tonemapping is not None:
            return self.render_views_torch(density_bitfield, camera_poses, intrinsics, intrinsics_size, render_size, cam_lights,
                                           ambient_light, testmode_dt_gamma_scale)
        nerf = self.nerf
        K = (intrinsics * (render_size / intrinsics_size)).float().contiguous()
        dt_gamma = float(testmode_dt_gamma_scale * 2 / (K[:, 0] + K[:, 1]).mean())
        ws, depth, image = nerf.decoder.render_cameras(camera_poses, K, render_size, render_size, density_bitfield, nerf.grid_size,
                                                       dt_gamma=dt_gamma)
        V, dev = K.shape[0], K.device
        images = torch.empty(V, 3, render_size, render_size, dtype=torch.bfloat16, device=dev)
        depths = torch.empty_like(images)
        scratch = torch.empty(V, 2, dtype=torch.int32, device=dev)
        call('mve_shade_views', ptr(ws), ptr(depth), ptr(image), ptr(K), ptr(cam_lights.float().contiguous()), c_u32(V), c_u32(render_size),
             c_u32(render_size), c_f32(float(ambient_light)), c_f32(float(nerf.bg_color)), c_f32(0.25), c_f32(0.5), c_f32(1e-5),
             ptr(scratch), ptr(images), ptr(depths), stream())
        return images, depths

    def render_views_torch(self, density_bitfield, camera_poses, intrinsics, intrinsics_size, render_size, cam_lights, ambient_light,
                           testmode_dt_gamma_scale):
        nerf = self.nerf
        rgba, depth, normal, normal_fg = nerf.render(
            nerf.decoder, None, density_bitfield, render_size, render_size, intrinsics[None] * (render_size / intrinsics_size),
            camera_poses[None], cfg=dict(return_rgba=True, compute_normal=True, dt_gamma_scale=testmode_dt_gamma_scale),
            perturb=False, normal_bg=self.normal_bg)
        normal_fg_opencv = torch.cat([normal_fg[..., :1] * 2 - 1, -normal_fg[..., 1:3] * 2 + 1], dim=-1)
        shading = ((cam_lights[:, None, None, None, :] @ normal_fg_opencv[..., :, None]).clamp(min=0) * (1 - ambient_light)
                   + ambient_light).squeeze(-1)
        if self.tonemapping is None:
            image = rgba[..., :3] * shading + nerf.bg_color * (1 - rgba[..., 3:])
        else:
            image = self.tonemapping.lut(self.tonemapping.inverse_lut(rgba[..., :3] / rgba[..., 3:].clamp(min=1e-6))
                                         + shading.clamp(min=1e-6).log2()) * rgba[..., 3:] + nerf.bg_color * (1 - rgba[..., 3:])
        images = image.squeeze(0).to(torch.bfloat16).permute(0, 3, 1, 2).clamp(min=0, max=1)
        alphas = rgba[..., 3:].squeeze(0)
        depths = normalize_depth(depth.squeeze(0), alphas).to(torch.bfloat16).unsqueeze(1).repeat(1, 3, 1, 1)
        return images, depths

    # ------------------------------------------------------------------ one loop iteration (t != None)
    def step(self, i, latents, prompt_embeds, decode_fn, density_grid, density_bitfield, optimizer, camera_poses, intrinsics,
             intrinsics_size, cam_weights, cam_lights, ancestral_noise, guidance_scale=7.0, render_size=512, n_inverse_steps=96,
             n_inverse_rays=2 ** 14, lr=0.01, alpha_soften=0.02, normal_reg_weight=0.1, entropy_weight=0.01, patch_rgb_weight=0.0,
             patch_normal_weight=0.0, bg_width=0.015, ambient_light=0.2, dt_gamma_scale=1.0, testmode_dt_gamma_scale=0.25,
             is_init=False, tile_weight=1.0, depth_weight=1.0, phase_events=None):
        """latents (n_local,4,L,L) fp32; prompt_embeds (2*n_local,T,D) as [neg ; pos]; cameras are the GLOBAL set (all views);
        with view sharding the local slice is ``view_shard.local_range``.  Returns (new latents, ctrl_images, ctrl_depths)."""
        def mark(name):
            if phase_events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                phase_events.append((name, e))

        sch = self.scheduler
        t = sch.timesteps[i]
        mark('start')
        lo, hi = view_shard.local_range(camera_poses.shape[0])
        n_local = hi - lo
        assert latents.shape[0] == n_local
        sqrt_ab, sqrt_1mab = get_noise_scales(sch.alphas_cumprod, t, sch.num_train_timesteps)
        # ---- denoise P1 (:1224-1256)
        latents_scaled = sch.scale_model_input(latents, i)
        latent_batches = [torch.cat([latents_scaled] * 2, dim=0)]
        prompt_batches = [prompt_embeds]
        noise_pred, dec_args, dec_kwargs = self.get_noise_pred_p1(latent_batches, prompt_batches, t, guidance_scale)
        pred_x0 = ((latents_scaled - sqrt_1mab * noise_pred.float()) / sqrt_ab)
        mark('denoise_p1')
        # ---- decode (neighbour hook) + exchange of the decoded views (:1258-1266)
        tgt_images, tgt_masks = decode_fn(pred_x0, lo, hi)              # (n_local, rs, rs, 3), (n_local, rs, rs, 1) fp32
        tgt_images = view_shard.gather_views(tgt_images)[None]
        tgt_masks = view_shard.gather_views(tgt_masks)[None]
        mark('decode_hook+gather')
        # ---- reconstruct (:1296-1305)
        nerf_optim(self.nerf, tgt_images, tgt_masks, None, optimizer, lr, n_inverse_steps, n_inverse_rays, patch_rgb_weight,
                   patch_normal_weight, alpha_soften, normal_reg_weight, entropy_weight, None, density_grid, density_bitfield,
                   render_size, intrinsics, intrinsics_size, camera_poses, cam_weights, cam_lights, self.nerf.patch_size, is_init,
                   bg_width, ambient_light, dt_gamma_scale, init_shaded=False, tonemapping=self.tonemapping)
        view_shard.broadcast_field(self.nerf.decoder, density_grid, density_bitfield)
        mark('nerf_optim')
        # ---- render my views (:1341-1395)
        ctrl_images, ctrl_depths = self.render_views(density_bitfield, camera_poses[lo:hi], intrinsics[lo:hi], intrinsics_size,
                                                     render_size, cam_lights[lo:hi], ambient_light, testmode_dt_gamma_scale)
        if render_size != 512:
            ctrl_images = torch.nn.functional.interpolate(ctrl_images.float(), size=(512, 512), mode='bilinear').clamp(0, 1).to(torch.bfloat16)
            ctrl_depths = torch.nn.functional.interpolate(ctrl_depths.float(), size=(512, 512), mode='bilinear').to(torch.bfloat16)
        mark('render_views')
        # ---- denoise P2 (:1413-1426)
        noise_pred = self.get_noise_pred_p2(latent_batches, prompt_batches, dec_args, dec_kwargs, t, guidance_scale,
                                            [torch.cat([ctrl_images] * 2, dim=0)], tile_weight, [torch.cat([ctrl_depths] * 2, dim=0)],
                                            depth_weight, ctrl_is_cfg_duplicate=True)
        # ---- solver step (:1438-1461, blend_weight 0)
        latents = sch.step(noise_pred.float(), i, latents, ancestral_noise)
        mark('denoise_p2+solver')
        return latents, ctrl_images, ctrl_depths

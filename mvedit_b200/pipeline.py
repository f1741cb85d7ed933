"""One iteration of the MVEdit hot loop -- denoise P1 -> (decode) -> NeRF reconstruction -> render all views -> denoise P2 -> solver
step -- as the reference runs it at /root/reference/lib/pipelines/mvedit_3d_pipeline.py:1141-1479 (t != None branch, NeRF stage,
'2-pass' mode, blend_weight 0), built on the B200 components of this package.

What is and is not inside ``MVEdit3DStep.step``:
  * inside: scale_model_input, get_noise_pred_p1 (UNet enc+dec, all views one batch), pred_x0, nerf_optim (n_inverse_steps Adam
    iterations of march/field/composite fwd+bwd), nerf.render of all views + Lambert shading + normalize_depth, get_noise_pred_p2
    (tile+depth ControlNets on the fresh renders + UNet dec), Euler-ancestral solver step;
  * vae.decode of pred_x0 (:1258-1263) runs on the tcgen05 kernels (mvedit_b200.vae) when a ``vae`` is given; the target masks come
    from a ``segmentation`` callable (TRACER in the reference: a neighbour that is not built).  A
    ``decode_fn(pred_x0, lo, hi) -> (tgt_images, tgt_masks)`` overrides both (bench.py uses it to decode for real and then hand the
    reconstruction analytic targets -- a random-init VAE decodes noise -- and says so in its JSON).  LPIPS and the SRVGG enhancer
    (only active below 512^2 renders) are not built.

Sharding (mvedit_b200.view_shard): every rank denoises / decodes / renders its slice of the views; the decoded
targets are exchanged with ONE all_gather per step; the reconstruction is data-parallel over rays (``nerf.data_parallel``: one
all_gather of per-ray outputs + one all_reduce of the flat gradient per iteration) or, without it, replicated + one flat broadcast.
"""
import math

import numpy as np
import torch

from .adapter3d_mixin import Adapter3DMixin
from .nerf import nerf_optim
from . import view_shard
from .schedulers import EulerAncestralScheduler, DPMSolverMultistepScheduler, DDIMScheduler      # noqa: F401 (re-exported)
from ._lib import call, ptr, stream, c_u32, c_f32
from .tonemapping import tone_args


def segment(images, seg_model, padding=0, bg_color=None, color_threshold=0.25):
    """The tensor path of ``do_segmentation`` (lib/pipelines/utils.py:73-96; what ``get_tgt_masks`` runs on the decoded views every step,
    adapter3d_mixin.py:14-19) around the segmentation network: replicate padding "helps to detect foreground objects", and every pixel that is
    not within ``color_threshold`` of the background colour in all channels is foreground whatever the network says.
    images (N,3,H,W) in [0,1] -> masks (N,1,H,W)."""
    if padding > 0:
        masks = seg_model(torch.nn.functional.pad(images, (padding, padding, padding, padding), mode='replicate'))[:, :, padding:-padding, padding:-padding]
    else:
        masks = seg_model(images)
    if bg_color is not None:
        bg = images.new_tensor(bg_color)[..., None, None]
        non_fg = torch.all(bg - color_threshold <= images, dim=1) & torch.all(images <= bg + color_threshold, dim=1)
        masks = torch.where(non_fg.unsqueeze(1), masks, torch.ones_like(masks))
    return masks


class MVEdit3DStep(Adapter3DMixin):
    """The loop body of MVEdit3DPipeline.__call__ (NeRF stage) on B200 components."""

    def __init__(self, unet, controlnet, nerf, scheduler, tonemapping=None, normal_bg=(0.5, 0.5, 1.0), vae=None, segmentation=None):
        self.unet, self.controlnet, self.nerf, self.scheduler = unet, controlnet, nerf, scheduler
        self.vae, self.segmentation = vae, segmentation
        # a mvedit_b200.tonemapping.Tonemapping (or any module with its lut_x / lut_y buffers): shading then happens in tone-mapped
        # space inside the fused objective and shading kernels (mvedit_3d_pipeline.py:564-570, :1377-1384)
        if tonemapping is not None and not hasattr(tonemapping, 'knots'):
            from .tonemapping import Tonemapping
            tm = Tonemapping(lut_steps=tonemapping.lut_x.numel())
            tm.load_state_dict({'lut_x': tonemapping.lut_x.detach().float().cpu(), 'lut_y': tonemapping.lut_y.detach().float().cpu()})
            tonemapping = tm
        self.tonemapping = tonemapping
        self.normal_bg = list(normal_bg)

    # ------------------------------------------------------------------ decode + masks (mvedit_3d_pipeline.py:1258-1266)
    def decode_targets(self, pred_x0, seg_padding=0):
        """pred_x0 (n,4,L,L) -> (tgt_images (n,8L,8L,3), tgt_masks (n,8L,8L,1)) fp32.  vae.decode runs on the tcgen05 kernels
        (mvedit_b200.vae); the masks come from ``self.segmentation`` (TRACER in the reference, adapter3d_mixin.py:14-19 -- a
        neighbour that is not built: any callable images (n,3,H,W) in [0,1] -> masks (n,1,H,W))."""
        if self.vae is None:
            raise RuntimeError('MVEdit3DStep: no vae -- pass vae= (mvedit_b200.vae.AutoencoderKL) or a decode_fn')
        imgs = self.vae.decode_images(pred_x0)
        if self.segmentation is None:
            raise RuntimeError('MVEdit3DStep: no segmentation callable for the target masks (TRACER is not built)')
        masks = segment(imgs.permute(0, 3, 1, 2), self.segmentation, padding=seg_padding, bg_color=getattr(self, 'bg_color', self.nerf.bg_color))
        return imgs, masks.permute(0, 2, 3, 1).float()

    # ------------------------------------------------------------------ render all (local) views, mvedit_3d_pipeline.py:1341-1395
    def render_views(self, density_bitfield, camera_poses, intrinsics, intrinsics_size, render_size, cam_lights, ambient_light,
                     testmode_dt_gamma_scale, render_bs=None, view_offset=0, all_intrinsics=None):
        """-> (ctrl_images, ctrl_depths), bf16 [V,3,rs,rs] in [0,1]: one fused render launch + mve_shade_views (two launches).
        The reference renders ``render_bs`` views per call and derives dt_gamma from the mean focal length OF THAT BATCH
        (base_nerf.py:501-502); with ``render_bs`` set, every view gets the dt_gamma of its reference batch (batches are cut from the
        GLOBAL view list: ``all_intrinsics`` / ``view_offset`` under view sharding), so the result does not depend on the shard layout."""
        nerf = self.nerf
        scale = render_size / intrinsics_size
        K = (intrinsics * scale).float().contiguous()
        V, dev = K.shape[0], K.device
        if render_bs is None:
            dt_gamma, per_view = float(testmode_dt_gamma_scale * 2 / (K[:, 0] + K[:, 1]).mean()), None
        else:
            Kg = (all_intrinsics if all_intrinsics is not None else intrinsics).float() * scale
            f2 = Kg[:, 0] + Kg[:, 1]
            batch_mean = torch.cat([c.mean().expand(c.numel()) for c in f2.split(render_bs)])
            dt_gamma, per_view = 0.0, (testmode_dt_gamma_scale * 2 / batch_mean)[view_offset:view_offset + V].contiguous()
        ws, depth, image = nerf.decoder.render_cameras(camera_poses, K, render_size, render_size, density_bitfield, nerf.grid_size,
                                                       dt_gamma=dt_gamma, dt_gamma_per_view=per_view)
        images = torch.empty(V, 3, render_size, render_size, dtype=torch.bfloat16, device=dev)
        depths = torch.empty_like(images)
        scratch = torch.empty(V, 2, dtype=torch.int32, device=dev)
        call('mve_shade_views', ptr(ws), ptr(depth), ptr(image), ptr(K), ptr(cam_lights.float().contiguous()), c_u32(V), c_u32(render_size),
             c_u32(render_size), c_f32(float(ambient_light)), c_f32(float(nerf.bg_color)), c_f32(0.25), c_f32(0.5), c_f32(1e-5),
             ptr(scratch), ptr(images), ptr(depths), ptr(None), *tone_args(self.tonemapping), stream())
        return images, depths

    # ------------------------------------------------------------------ one loop iteration (t != None)
    def step(self, i, latents, prompt_embeds, decode_fn, density_grid, density_bitfield, optimizer, camera_poses, intrinsics,
             intrinsics_size, cam_weights, cam_lights, ancestral_noise, guidance_scale=7.0, render_size=512, n_inverse_steps=96,
             n_inverse_rays=2 ** 14, lr=0.01, alpha_soften=0.02, normal_reg_weight=0.1, entropy_weight=0.01, patch_rgb_weight=0.0,
             patch_normal_weight=0.0, bg_width=0.015, ambient_light=0.2, dt_gamma_scale=1.0, testmode_dt_gamma_scale=0.25,
             is_init=False, tile_weight=1.0, depth_weight=1.0, phase_events=None, render_bs=None):
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
        sqrt_ab, sqrt_1mab = sch.noise_scales(t)
        # ---- denoise P1 (:1224-1256)
        latents_scaled = sch.scale_model_input(latents, i)
        latent_batches = [torch.cat([latents_scaled] * 2, dim=0)]
        prompt_batches = [prompt_embeds]
        noise_pred, dec_args, dec_kwargs = self.get_noise_pred_p1(latent_batches, prompt_batches, t, guidance_scale)
        pred_x0 = ((latents_scaled - sqrt_1mab * noise_pred.float()) / sqrt_ab)
        mark('denoise_p1')
        # ---- decode (neighbour hook) + exchange of the decoded views (:1258-1266)
        if decode_fn is None:
            tgt_images, tgt_masks = self.decode_targets(pred_x0)        # (n_local, rs, rs, 3), (n_local, rs, rs, 1) fp32
        else:
            tgt_images, tgt_masks = decode_fn(pred_x0, lo, hi)
        mark('decode')
        tgt_images, tgt_masks = view_shard.gather_views(tgt_images, tgt_masks, camera_poses.shape[0])     # ONE collective
        tgt_images, tgt_masks = tgt_images[None], tgt_masks[None]
        if render_size != tgt_images.shape[2]:              # the fit runs at the render size (:1284-1288)
            rs_ = lambda x: torch.nn.functional.interpolate(x.squeeze(0).permute(0, 3, 1, 2), size=render_size, mode='bilinear').permute(0, 2, 3, 1)[None]
            tgt_images, tgt_masks = rs_(tgt_images), rs_(tgt_masks)
        mark('gather')
        # ---- reconstruct (:1296-1305)
        nerf_optim(self.nerf, tgt_images, tgt_masks, None, optimizer, lr, n_inverse_steps, n_inverse_rays, patch_rgb_weight,
                   patch_normal_weight, alpha_soften, normal_reg_weight, entropy_weight, None, density_grid, density_bitfield,
                   render_size, intrinsics, intrinsics_size, camera_poses, cam_weights, cam_lights, self.nerf.patch_size, is_init,
                   bg_width, ambient_light, dt_gamma_scale, init_shaded=False, tonemapping=self.tonemapping)
        if not getattr(self.nerf, 'data_parallel', False):
            view_shard.broadcast_field(self.nerf.decoder, density_grid, density_bitfield)    # replicated mode: rank 0's field wins
        mark('nerf_optim')
        # ---- render my views (:1341-1395)
        ctrl_images, ctrl_depths = self.render_views(density_bitfield, camera_poses[lo:hi], intrinsics[lo:hi], intrinsics_size,
                                                     render_size, cam_lights[lo:hi], ambient_light, testmode_dt_gamma_scale,
                                                     render_bs=render_bs, view_offset=lo, all_intrinsics=intrinsics)
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

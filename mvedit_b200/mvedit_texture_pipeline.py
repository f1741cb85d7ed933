"""``MVEditTexturePipeline`` -- the reference's text-guided re-texturing pipeline (``lib/pipelines/mvedit_texture_pipeline.py``) over
the same kernels as the 3D pipeline: per diffusion timestep the denoiser (UNet + tile / depth ControlNets, 1- or 2-pass, optional
reference-image joint attention) predicts all views, ``vae.decode`` turns them into target images, ``MeshRenderer.bake_multiview``
blends them into the mesh's UV texture (the mesh stage's rasterize / interpolate / texture kernels), the textured mesh is re-rendered
as the next step's tile condition; after the last step ``texture_optim`` fits the hash-grid field on the fixed mesh and
``bake_xyz_shading_fun`` bakes it (``:175-544``).  Constructor / ``__call__`` argument names and defaults are the reference's
(``:55-83,175-220``); ``prompt_embeds`` is the same extension as in ``MVEdit3DPipeline`` (no CLIP weights offline).

Host-side helpers with the reference's definitions: ``default_patch_rgb_weight`` / ``default_max_num_views`` (``:32-37``) and
``camera_dense_weighting`` (``:40-50``), pinned against the reference's functions by tests/test_mesh_pins.py.

``ip_adapter`` is any object with the reference IPAdapter's ``get_prompt_embeds`` (as in the 3D pipeline).
"""
import traceback
from copy import deepcopy

import numpy as np
import torch
import torch.nn.functional as F

from .mesh_optim import make_nerf_albedo_shading_fun, normalize_depth, texture_optim
from .mesh_renderer import view_cosine
from .mvedit_3d_pipeline import MVEdit3DPipeline, get_camera_dists, join_prompts, prune_cameras
from .nerf import pixel_directions
from .optim import FusedAdam


def default_patch_rgb_weight(progress, start_weight=0.1, end_weight=0.1):
    return start_weight + (end_weight - start_weight) * progress


def default_max_num_views(progress, start_num=32, end_num=7, power=2):
    return (start_num - end_num) * (1 - progress) ** power + end_num


def camera_dense_weighting(intrinsics, intrinsics_size, render_size, alphas, depths, cos_weight_pow=1.0):
    """Per-pixel confidence of every view (``:40-50``): frontalness of the rendered surface (normals from the inverse-depth map) x alpha,
    eroded by a 5x5 minimum and smoothed by a 5x5 box.  intrinsics [n,4], alphas [n,s,s,1], depths [n,s,s] -> [n,s,s,1]."""
    dirs = pixel_directions(intrinsics * (render_size / intrinsics_size), render_size, render_size)
    weight = view_cosine(depths, dirs) ** cos_weight_pow * alphas
    pooled = F.avg_pool2d(F.max_pool2d(-weight.permute(0, 3, 1, 2), 5, stride=1, padding=2), 5, stride=1, padding=2)
    return -pooled.permute(0, 2, 3, 1)


class MVEditTexturePipeline(MVEdit3DPipeline):
    def __init__(self, vae, text_encoder, tokenizer, unet, controlnet, scheduler, nerf, mesh_renderer):
        super().__init__(vae, text_encoder, tokenizer, unet, controlnet, scheduler, nerf, mesh_renderer=mesh_renderer)
        self.bg_color = 0.5

    def texture_optim(self, *args, **kwargs):
        return texture_optim(self, *args, **kwargs)

    def render_albedo_views(self, in_mesh, camera_poses, intrinsics, intrinsics_size, render_size, render_bs, shading_fun=None):
        """The per-step render (``:455-468``): textured (or field-shaded) mesh over the background colour -> [n,3,s,s] in [0,1]."""
        images = []
        for p_b, i_b in zip(camera_poses.split(render_bs, dim=0), intrinsics.split(render_bs, dim=0)):
            rgba = self.mesh_renderer([in_mesh], p_b[None], i_b[None] * (render_size / intrinsics_size), render_size, render_size,
                                      shading_fun)['rgba'].squeeze(0)
            images.append(rgba[..., :3] + self.bg_color * (1 - rgba[..., 3:]))
        return torch.cat(images, dim=0).permute(0, 3, 1, 2).clamp(min=0, max=1)

    def __call__(self, prompt='', negative_prompt='', in_model=None, ingp_states=None, init_images=None, cond_images=None,
                 extra_control_images=None, nerf_code=None, camera_poses=None, intrinsics=None, intrinsics_size=256, use_reference=True,
                 cam_weights=None, weighted_cam_pruning=False, keep_views=None, guidance_scale=7, num_inference_steps=24,
                 denoising_strength=0.6, diff_size=512, patch_size=512, patch_bs=1, diff_bs=12, render_bs=8, n_inverse_steps=512,
                 ip_adapter=None, lr=0.01, max_num_views=default_max_num_views, patch_rgb_weight=default_patch_rgb_weight,
                 optim_only=False, debug=False, out_dir=None, save_interval=None, save_all_interval=None,
                 default_prompt='high-res, best quality, 3d model, cg, rendering, extremely detailed, photorealistic, RAW photo, 4k uhd, '
                                'dslr, high quality',
                 default_neg_prompt='depth of field, out of focus, lowres, worst quality, low quality, drawing, illustration, painting, '
                                    'blurry, jpeg artifacts, macro',
                 bake_texture=True, bake_texture_kwargs=None, mode='1-pass', prog_bar=None, prompt_embeds=None):
        """-> (textured mesh | None, ingp state dict | None), as mvedit_texture_pipeline.py:175-544."""
        assert in_model is not None
        assert optim_only or mode in ('1-pass', '2-pass')
        nerf, dec, sch = self.nerf, self.nerf.decoder, self.scheduler
        device = next(dec.parameters()).device
        grad_mode = torch.is_grad_enabled()
        torch.set_grad_enabled(False)
        render_size = diff_size
        if dec.state_dict_bak is None:
            dec.backup_state_dict()
        out_mesh = output_state = None
        try:
            if ingp_states is not None:
                dec.load_state_dict(ingp_states if isinstance(ingp_states, dict) else torch.load(ingp_states, map_location='cpu'), strict=False)
            camera_poses = (camera_poses if torch.is_tensor(camera_poses) else torch.from_numpy(np.stack(camera_poses, axis=0))
                            ).to(device=device, dtype=torch.float32)
            num_cameras = len(camera_poses)
            intrinsics = intrinsics.to(device=device, dtype=torch.float32)
            if intrinsics.dim() == 1:
                intrinsics = intrinsics[None].expand(num_cameras, -1)
            albedo_fun = make_nerf_albedo_shading_fun(dec, nerf_code)
            in_mesh, images, in_masks, depths = self.load_init_mesh(in_model, camera_poses, intrinsics, intrinsics_size, render_bs,
                                                                    None if ingp_states is None else albedo_fun, diff_size=diff_size)
            in_images = images.permute(0, 3, 1, 2).float()
            ctrl_depths = normalize_depth(depths, in_masks).unsqueeze(1).repeat(1, 3, 1, 1).to(torch.bfloat16)
            if init_images is not None:
                in_images = self.load_init_images(init_images, ret_masks=False, diff_size=diff_size)
            cam_weights = camera_poses.new_tensor([1.0] * num_cameras if cam_weights is None else cam_weights)
            cam_weights_dense = cam_weights[:, None, None, None] * camera_dense_weighting(intrinsics, intrinsics_size, render_size, in_masks, depths)
            prompt = prompt if isinstance(prompt, list) else [prompt] * num_cameras
            negative_prompt = negative_prompt if isinstance(negative_prompt, list) else [negative_prompt] * num_cameras
            cond_images, extra_control_images = self.load_cond_images(in_images, cond_images, extra_control_images)
            if not optim_only:
                sch.set_timesteps(num_inference_steps, device=device)
                timesteps = sch.timesteps
                if denoising_strength is not None:
                    timesteps = timesteps[min(int(round(len(timesteps) * (1 - denoising_strength) / sch.order)) * sch.order,
                                              len(timesteps) - 1):]
                pe = self.get_prompt_embeds(in_images, [join_prompts(p, default_prompt) for p in prompt],
                                            [join_prompts(p, default_neg_prompt) for p in negative_prompt], ip_adapter=ip_adapter,
                                            cond_images=cond_images, prompt_embeds=prompt_embeds)
                encode = lambda x: torch.cat([self.vae.encode(b * 2 - 1).latent_dist.sample() * self.vae.config.scaling_factor
                                              for b in x.split(diff_bs, dim=0)], dim=0)
                init_latents = encode(in_images)
                ref_latents = None
                if use_reference:
                    ref_latents = init_latents if cond_images is None else encode(
                        torch.cat([F.interpolate(c, size=(diff_size, diff_size), mode='bilinear') for c in cond_images], dim=0))
                L = init_latents.shape[-1]
            optimizer = FusedAdam(dec.parameters(), lr=0.01)
            total_steps = num_inference_steps if optim_only else len(timesteps)
            steps = [None] * (num_inference_steps + 1) if optim_only else [None] + list(timesteps)
            it = prog_bar(steps) if prog_bar is not None else steps
            latents = ctrl_images = None
            for i, t in enumerate(it):
                progress = i / total_steps
                # ---- cameras (:327-380)
                if i == 0:
                    keep_views = list(keep_views or [])
                    num_keep_views = len(keep_views)
                    order = torch.tensor(keep_views + [c for c in range(num_cameras) if c not in keep_views], device=device)
                    in_images, in_masks, ctrl_depths = in_images[order], in_masks[order], ctrl_depths[order]
                    camera_poses, intrinsics, cam_weights_dense = camera_poses[order], intrinsics[order], cam_weights_dense[order]
                    extra_control_images = [e[order] for e in extra_control_images]
                    if not optim_only:
                        init_latents = init_latents[order]
                        ref_latents = ref_latents[order] if use_reference else None
                        pe = pe[torch.cat([order, order + num_cameras])]
                    # as the reference (:349-350): the weights are NOT re-ordered with the cameras here
                    dists = get_camera_dists(camera_poses, cam_weights if weighted_cam_pruning else None)
                else:
                    max_num_cameras = max(int(round(max_num_views(progress))), num_keep_views)
                    if max_num_cameras < num_cameras:
                        keep_ids, dists = prune_cameras(dists, num_keep_views, max_num_cameras)
                        in_images, in_masks, ctrl_depths = in_images[keep_ids], in_masks[keep_ids], ctrl_depths[keep_ids]
                        camera_poses, intrinsics, cam_weights_dense = camera_poses[keep_ids], intrinsics[keep_ids], cam_weights_dense[keep_ids]
                        extra_control_images = [e[keep_ids] for e in extra_control_images]
                        if not optim_only:
                            if mode == '1-pass':
                                ctrl_images = ctrl_images[keep_ids]
                            latents = latents[keep_ids]
                            ref_latents = ref_latents[keep_ids] if use_reference else None
                            pe = pe[torch.cat([keep_ids, keep_ids + num_cameras])]
                            sch.prune(keep_ids)
                        num_cameras = max_num_cameras
                # ---- denoise P1 + decode (:382-433)
                if not optim_only:
                    sqrt_ab, sqrt_1mab = sch.noise_scales(timesteps[0] if t is None else t)
                if t is not None and not optim_only:
                    latents_scaled = sch.scale_model_input(latents, t)
                    if use_reference:
                        lat_b, pe_b = [latents_scaled[:, :, -L:], latents_scaled], [pe[:num_cameras], pe[-num_cameras:]]
                        dup = lambda x: [x, x]
                    else:
                        lat_b, pe_b = [torch.cat([latents_scaled] * 2, dim=0)], [pe]
                        dup = lambda x: [torch.cat([x] * 2, dim=0)]
                    extra_b = [dup(e.to(torch.bfloat16)) for e in extra_control_images]
                    if mode == '1-pass':
                        noise_pred = self.get_noise_pred(lat_b, pe_b, dup(ctrl_images.to(torch.bfloat16)), dup(ctrl_depths), t, float(sqrt_ab), 1.0,
                                                         guidance_scale, extra_control_batches=extra_b)
                    else:
                        noise_pred, dec_args, dec_kwargs = self.get_noise_pred_p1(lat_b, pe_b, t, guidance_scale, dup(ctrl_depths), 1.0,
                                                                                  extra_control_batches=extra_b)
                    pred_x0 = (latents_scaled[:, :, -L:] - sqrt_1mab * noise_pred.float()) / sqrt_ab
                    tgt_images = torch.cat([(self.vae.decode(b / self.vae.config.scaling_factor, return_dict=False)[0].float() / 2 + 0.5
                                             ).clamp(min=0, max=1).permute(0, 2, 3, 1) for b in pred_x0.split(diff_bs, dim=0)], dim=0)[None]
                else:
                    tgt_images = in_images.permute(0, 2, 3, 1)[None].float()
                # ---- texture update (:435-449): the final step fits the field, every other denoising step bakes the views into the UV map
                if i == total_steps:
                    self.texture_optim(tgt_images, optimizer, lr, n_inverse_steps, render_bs, patch_bs, patch_rgb_weight(progress), nerf_code,
                                       in_mesh, render_size, intrinsics, intrinsics_size, camera_poses, cam_weights_dense, patch_size, debug=debug)
                elif t is not None:
                    in_mesh = self.mesh_renderer.bake_multiview([in_mesh], tgt_images, cam_weights_dense[None], camera_poses[None],
                                                                intrinsics[None] * (render_size / intrinsics_size), cos_weight_pow=0.0,
                                                                render_bs=render_bs)[0]
                # ---- render (:451-470)
                if i < total_steps and not optim_only and (t is not None or mode == '1-pass'):
                    ctrl_images = self.render_albedo_views(in_mesh, camera_poses, intrinsics, intrinsics_size, render_size, render_bs)
                if i >= total_steps or optim_only:
                    continue
                # ---- denoise P2 (:474-486) + solver (:497-527)
                if mode == '2-pass' and t is not None:
                    noise_pred = self.get_noise_pred_p2(lat_b, pe_b, dec_args, dec_kwargs, t, guidance_scale, dup(ctrl_images.to(torch.bfloat16)), 1,
                                                        ctrl_is_cfg_duplicate=not use_reference)
                if t is not None:
                    merged = noise_pred.float()
                    if use_reference:
                        merged = torch.cat([(latents_scaled[:, :, :L] - ref_latents * sqrt_ab) / sqrt_1mab, merged], dim=2)
                    latents = sch.step(merged, t, latents, torch.randn(latents.shape, device=device))
                elif denoising_strength is None:
                    shared = lambda: torch.randn_like(init_latents[0]).expand(init_latents.size(0), -1, -1, -1) * sch.init_noise_sigma
                    latents = shared()
                    if use_reference:        # as the reference (:511-514): ``ref_latents`` is REPLACED by the noise the reference half starts from
                        ref_latents = shared()
                        latents = torch.cat([ref_latents, latents], dim=2)
                else:
                    latents = torch.cat([ref_latents, init_latents], dim=2) if use_reference else init_latents
                    latents = sch.add_noise(latents, torch.randn_like(latents[0]).expand(latents.size(0), -1, -1, -1), timesteps[0:1])
            kw = dict(map_size=2048, force_auto_uv=False)
            kw.update(bake_texture_kwargs or {})
            out_mesh = self.mesh_renderer.bake_xyz_shading_fun([in_mesh.detach()], albedo_fun, **kw)[0] if bake_texture else in_mesh.detach()
            output_state = deepcopy(dec.state_dict())
        except NotImplementedError:
            dec.restore_state_dict()
            raise
        except Exception:
            print(traceback.format_exc())
            out_mesh = output_state = None
        finally:
            torch.set_grad_enabled(grad_mode)
        dec.restore_state_dict()
        return out_mesh, output_state


class MVEditTextureSuperResPipeline(MVEditTexturePipeline):
    """The texture super-resolution pipeline (``lib/pipelines/mvedit_texture_superres_pipeline.py``; BASELINE configs[3]'s second half):
    the views rendered from the input mesh are re-noised and denoised with the tile ControlNet at full weight on the ORIGINAL renders
    (the condition never changes: no per-step baking, no pruning, 1-pass only), the last step's prediction plus a set of
    regulariser views (``reg_camera_poses``: plain renders of the input, weight 0.5) are fitted by ``texture_optim``, and the baked
    field is blended with the original texture by per-texel camera confidence (``get_cam_weights_uv``, cos^4, original weight 0.2^4).
    Returns the mesh only, as the reference (``:493-496``)."""

    def get_prompt_embeds(self, in_images, rgb_prompt, rgb_negative_prompt, ip_adapter=None, ip_adapter_use_cond_idx=None, cond_images=None,
                          prompt_embeds=None):
        """mvedit_texture_superres_pipeline.py:62-87: as the 3D pipeline's, but the IP-Adapter looks at the INPUT renders, replaced by the
        conditioning image only for the views listed in ``ip_adapter_use_cond_idx``."""
        if ip_adapter is None:
            return super().get_prompt_embeds(in_images, rgb_prompt, rgb_negative_prompt, prompt_embeds=prompt_embeds)
        size = (self.clip_img_size, self.clip_img_size)
        ipa = F.interpolate(in_images, size=size, mode='bilinear')
        if isinstance(ip_adapter_use_cond_idx, list) and cond_images is not None:
            ipa = ipa.clone()
            for idx, c in enumerate(cond_images):
                if idx in ip_adapter_use_cond_idx:
                    ipa[idx] = F.interpolate(c, size=size, mode='bilinear')[0]
        return ip_adapter.get_prompt_embeds((ipa - ipa.new_tensor(self.clip_img_mean)[:, None, None]) / ipa.new_tensor(self.clip_img_std)[:, None, None],
                                            prompt=rgb_prompt, negative_prompt=rgb_negative_prompt)

    def __call__(self, prompt='', negative_prompt='', in_model=None, ingp_states=None, init_images=None, cond_images=None,
                 extra_control_images=None, nerf_code=None, camera_poses=None, reg_camera_poses=None, intrinsics=None, intrinsics_size=256,
                 use_reference=True, cam_weights=None, reg_cam_weights=None, guidance_scale=7, num_inference_steps=26, denoising_strength=0.5,
                 diff_size=512, patch_size=512, patch_bs=1, diff_bs=12, render_bs=8, n_inverse_steps=512, ip_adapter=None,
                 ip_adapter_use_cond_idx=None, lr=0.01, patch_rgb_weight=default_patch_rgb_weight, optim_only=False, debug=False,
                 out_dir=None, save_interval=None, save_all_interval=None,
                 default_prompt='best quality, sharp focus, photorealistic, extremely detailed',
                 default_neg_prompt='worst quality, low quality, depth of field, blurry, out of focus, low-res, illustration, painting, drawing',
                 bake_texture_kwargs=None, prog_bar=None, prompt_embeds=None):
        assert in_model is not None
        from .mesh_renderer import edge_dilation
        nerf, dec, sch = self.nerf, self.nerf.decoder, self.scheduler
        device = next(dec.parameters()).device
        grad_mode = torch.is_grad_enabled()
        torch.set_grad_enabled(False)
        render_size = diff_size
        patch_size = render_size // round(render_size / patch_size)               # make sure patch_size divides render_size (:219)
        if dec.state_dict_bak is None:
            dec.backup_state_dict()
        out_mesh = None
        try:
            if ingp_states is not None:
                dec.load_state_dict(ingp_states if isinstance(ingp_states, dict) else torch.load(ingp_states, map_location='cpu'), strict=False)
            to_poses = lambda p: (p if torch.is_tensor(p) else torch.from_numpy(np.stack(p, axis=0))).to(device=device, dtype=torch.float32)
            camera_poses = to_poses(camera_poses)
            num_cameras = len(camera_poses)
            reg_camera_poses = to_poses(reg_camera_poses) if reg_camera_poses is not None else camera_poses.new_zeros((0,) + tuple(camera_poses.shape[1:]))
            num_reg = len(reg_camera_poses)
            all_poses = torch.cat([camera_poses, reg_camera_poses], dim=0)
            intrinsics = intrinsics.to(device=device, dtype=torch.float32)
            if intrinsics.dim() == 1:
                intrinsics = intrinsics[None].expand(len(all_poses), -1)
            albedo_fun = make_nerf_albedo_shading_fun(dec, nerf_code)
            in_mesh, images, alphas, depths = self.load_init_mesh(in_model, all_poses, intrinsics, intrinsics_size, render_bs,
                                                                  None if ingp_states is None else albedo_fun, diff_size=diff_size)
            if init_images is None:
                in_images, reg_images = images[:num_cameras].permute(0, 3, 1, 2).float(), images[num_cameras:]
            else:
                init = self.load_init_images(init_images, ret_masks=False, diff_size=diff_size)
                in_images, reg_images = init[:num_cameras], init[num_cameras:].permute(0, 2, 3, 1).float()
            ctrl_images = in_images.to(torch.bfloat16)
            ctrl_depths = normalize_depth(depths[:num_cameras], alphas[:num_cameras]).unsqueeze(1).repeat(1, 3, 1, 1).to(torch.bfloat16)
            cam_weights = camera_poses.new_tensor(list(cam_weights if cam_weights is not None else [1.0] * num_cameras)
                                                  + list(reg_cam_weights if reg_cam_weights is not None else [0.5] * num_reg))
            cam_weights_dense = cam_weights[:, None, None, None] * camera_dense_weighting(intrinsics, intrinsics_size, render_size, alphas, depths)
            prompt = prompt if isinstance(prompt, list) else [prompt] * num_cameras
            negative_prompt = negative_prompt if isinstance(negative_prompt, list) else [negative_prompt] * num_cameras
            cond_images, extra_control_images = self.load_cond_images(in_images, cond_images, extra_control_images)
            if not optim_only:
                sch.set_timesteps(num_inference_steps, device=device)
                timesteps = sch.timesteps
                if denoising_strength is not None:
                    timesteps = timesteps[min(int(round(len(timesteps) * (1 - denoising_strength) / sch.order)) * sch.order,
                                              len(timesteps) - 1):]
                pe = self.get_prompt_embeds(in_images, [join_prompts(p, default_prompt) for p in prompt],
                                            [join_prompts(p, default_neg_prompt) for p in negative_prompt], ip_adapter=ip_adapter,
                                            ip_adapter_use_cond_idx=ip_adapter_use_cond_idx, cond_images=cond_images, prompt_embeds=prompt_embeds)
                encode = lambda x: torch.cat([self.vae.encode(b * 2 - 1).latent_dist.sample() * self.vae.config.scaling_factor
                                              for b in x.split(diff_bs, dim=0)], dim=0)
                init_latents = encode(in_images)
                ref_latents = None
                if use_reference:
                    ref_latents = init_latents if cond_images is None else encode(
                        torch.cat([F.interpolate(c, size=(diff_size, diff_size), mode='bilinear') for c in cond_images], dim=0))
                L = init_latents.shape[-1]
            optimizer = FusedAdam(dec.parameters(), lr=0.01)
            total_steps = num_inference_steps if optim_only else len(timesteps)
            steps = [None] * (num_inference_steps + 1) if optim_only else [None] + list(timesteps)
            latents = None
            tgt_images = None
            for i, t in enumerate(prog_bar(steps) if prog_bar is not None else steps):
                progress = i / total_steps
                if not optim_only:
                    sqrt_ab, sqrt_1mab = sch.noise_scales(timesteps[0] if t is None else t)
                if t is not None and not optim_only:
                    latents_scaled = sch.scale_model_input(latents, t)
                    if use_reference:
                        lat_b, pe_b = [latents_scaled[:, :, -L:], latents_scaled], [pe[:num_cameras], pe[-num_cameras:]]
                        dup = lambda x: [x, x]
                    else:
                        lat_b, pe_b = [torch.cat([latents_scaled] * 2, dim=0)], [pe]
                        dup = lambda x: [torch.cat([x] * 2, dim=0)]
                    noise_pred = self.get_noise_pred(lat_b, pe_b, dup(ctrl_images), dup(ctrl_depths), t, 1.0, 1.0, guidance_scale,
                                                     extra_control_batches=[dup(e.to(torch.bfloat16)) for e in extra_control_images])
                    if i == total_steps:
                        pred_x0 = (latents_scaled[:, :, -L:] - sqrt_1mab * noise_pred.float()) / sqrt_ab
                        tgt_images = torch.cat([(self.vae.decode(b / self.vae.config.scaling_factor, return_dict=False)[0].float() / 2 + 0.5
                                                 ).clamp(min=0, max=1).permute(0, 2, 3, 1) for b in pred_x0.split(diff_bs, dim=0)], dim=0)[None]
                elif i == total_steps:
                    tgt_images = in_images.permute(0, 2, 3, 1)[None].float()
                if i == total_steps:                                   # optimisation only at the final step (:398-406)
                    tgt_images = torch.cat([tgt_images, reg_images[None].float()], dim=1)
                    self.texture_optim(tgt_images, optimizer, lr, n_inverse_steps, render_bs, patch_bs, patch_rgb_weight(progress), nerf_code,
                                       in_mesh, render_size, intrinsics, intrinsics_size, all_poses, cam_weights_dense, patch_size, debug=debug,
                                       patch_views=num_cameras)
                if i >= total_steps or optim_only:
                    continue
                if t is not None:
                    merged = noise_pred.float()
                    if use_reference:
                        merged = torch.cat([(latents_scaled[:, :, :L] - ref_latents * sqrt_ab) / sqrt_1mab, merged], dim=2)
                    latents = sch.step(merged, t, latents, torch.randn(latents.shape, device=device))
                elif denoising_strength is None:
                    shared = lambda: torch.randn_like(init_latents[0]).expand(init_latents.size(0), -1, -1, -1) * sch.init_noise_sigma
                    latents = shared()
                    if use_reference:        # as the reference (:511-514): ``ref_latents`` is REPLACED by the noise the reference half starts from
                        ref_latents = shared()
                        latents = torch.cat([ref_latents, latents], dim=2)
                else:
                    latents = torch.cat([ref_latents, init_latents], dim=2) if use_reference else init_latents
                    latents = sch.add_noise(latents, torch.randn_like(latents[0]).expand(latents.size(0), -1, -1, -1), timesteps[0:1])
            kw = dict(map_size=2048, force_auto_uv=False)
            kw.update(bake_texture_kwargs or {})
            ori_albedo = in_mesh.albedo
            keep_original = not (ori_albedo is None or in_mesh.textureless)
            if keep_original:
                kw.update(dilation_iters=0)
            out_mesh = self.mesh_renderer.bake_xyz_shading_fun([in_mesh], albedo_fun, **kw)[0]
            if keep_original:                                          # blend with the input texture where the cameras saw little (:468-487)
                cos_weight_pow, map_size = 4.0, kw['map_size']
                ori_weight = 0.2 ** cos_weight_pow
                ori = F.interpolate(ori_albedo.permute(2, 0, 1)[None], size=map_size, mode='bilinear').squeeze(0).permute(1, 2, 0)
                w_uv, valid = self.mesh_renderer.get_cam_weights_uv([in_mesh], all_poses[None], intrinsics[None] * (render_size / intrinsics_size),
                                                                    render_size=render_size, map_size=map_size, render_bs=render_bs,
                                                                    cos_weight_pow=cos_weight_pow)
                w_uv = (w_uv.squeeze(0) * cam_weights[:, None, None, None]).sum(dim=0)
                albedo = (ori[..., :3] * ori_weight + out_mesh.albedo[..., :3] * w_uv) / (ori_weight + w_uv).clamp(min=1e-6)
                out_mesh.albedo = edge_dilation(albedo.permute(2, 0, 1)[None], valid[None].float()).squeeze(0).permute(1, 2, 0)
                out_mesh.textureless = False
        except NotImplementedError:
            dec.restore_state_dict()
            raise
        except Exception:
            print(traceback.format_exc())
            out_mesh = None
        finally:
            torch.set_grad_enabled(grad_mode)
        dec.restore_state_dict()
        return out_mesh

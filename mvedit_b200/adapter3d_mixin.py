"""Host-side mirror of ``Adapter3DMixin.get_noise_pred{,_p1,_p2}`` -- the denoiser seam B2 of SURVEY.md §8b
(/root/reference/lib/pipelines/adapter3d_mixin.py:68-317).  Same argument lists and return values; ``self.unet`` is a
``mvedit_b200.unet.UNet`` and ``self.controlnet`` a ``MultiControlNet`` whose ``nets`` are ordered [tile, depth, *extra] as in the
reference.  The lists of ``diff_bs``-sized chunks the pipeline passes in are fused: all chunks of equal shape run as ONE batch
(the reference loops chunk by chunk only to bound VRAM, adapter3d.py:546); results are identical per sample.

``dec_args`` / ``dec_kwargs`` stay opaque to the caller (one entry per fused group instead of one per chunk).
IP-Adapter: ``unet.set_ip_adapter(...)`` / ``controlnet.set_cn_attn_processor()`` (mvedit_b200.unet) switch the cross-attentions to
IPAttnProcessor2_0 / CNAttnProcessor2_0; prompt embeddings then carry the image tokens (T = 77 + 16) and ``adapter_scale`` selects the
adapter-only guidance.  Extra ControlNets (ip2p, nets[2:]) take ``extra_control_batches`` with weight 1.0 as in the reference.
Not carried over: ``cond_noisy_latent_batches`` (the texture pipelines' attention read/write modes) and ``added_cond_kwargs``
(SDXL-style conditioning, unused by SD1.5) -- they raise.
"""
from copy import copy

import torch

from .unet import MultiControlNet


def _groups(batches):
    """indices of consecutive chunks with identical shape -> [(start, end)]"""
    out, s = [], 0
    for i in range(1, len(batches) + 1):
        if i == len(batches) or batches[i].shape[1:] != batches[s].shape[1:]:
            out.append((s, i))
            s = i
    return out


def _cat(lst, a, b):
    if lst is None or lst[a] is None:
        return None
    return torch.cat(list(lst[a:b]), dim=0) if b - a > 1 else lst[a]


def _interleave_zero(t):
    return torch.stack([torch.zeros_like(t), t], dim=1).view(-1, *t.shape[1:])


class Adapter3DMixin:
    unet = None
    controlnet = None

    def load_init_mesh(self, in_model, camera_poses, intrinsics, intrinsics_size, render_bs, shading_fun=None, diff_size=512):
        """Render an input mesh from every camera at 2x supersampling (adapter3d_mixin.py:21-66) -> (mesh, images [N,s,s,3] composited on
        ``self.bg_color``, alphas [N,s,s,1], inverse depths [N,s,s]).  ``in_model`` is a ``mesh_renderer.Mesh`` or a path to an ``.obj`` / ``.glb`` / binary
        ``.ply`` (``Mesh.load``)."""
        if isinstance(in_model, str):
            from .mesh_renderer import Mesh
            in_mesh = Mesh.load(in_model, flip_yz=in_model.endswith(('.obj', '.glb'))).to(camera_poses.device)
        else:
            in_mesh = in_model.detach().to(camera_poses.device)
        if in_mesh.vn is None:
            in_mesh.auto_normal()
        renderer = copy(self.mesh_renderer)
        renderer.ssaa = 2
        pose_b, intr_b = camera_poses.split(render_bs, dim=0), intrinsics.split(render_bs, dim=0)
        funs = shading_fun if isinstance(shading_fun, list) else [shading_fun] * len(pose_b)
        images, alphas, depths = [], [], []
        for p_b, i_b, fun in zip(pose_b, intr_b, funs):
            out = renderer([in_mesh], p_b[None], i_b[None] * (diff_size / intrinsics_size), diff_size, diff_size, fun)
            rgba = out['rgba'].squeeze(0)
            images.append(rgba[..., :3] + (1 - rgba[..., 3:]) * self.bg_color)
            alphas.append(rgba[..., 3:])
            depths.append(out['depth'].squeeze(0))
        return in_mesh, torch.cat(images, dim=0).clamp(min=0, max=1), torch.cat(alphas, dim=0), torch.cat(depths, dim=0)

    @staticmethod
    def _split_ref(lat, latent_size, pe):
        shp = lat.shape
        if shp[2] == 2 * shp[3]:
            unet_in = lat.reshape(*shp[:2], 2, shp[3], shp[3]).permute(0, 2, 1, 3, 4).reshape(shp[0] * 2, shp[1], shp[3], shp[3])
            return dict(num_cross_attn_imgs=2), unet_in, lat[:, :, -latent_size:], \
                pe.unsqueeze(1).expand(-1, 2, -1, -1).reshape(-1, *pe.shape[1:]), pe
        return None, lat, lat, pe, pe

    @staticmethod
    def _cfg(noise_pred, guidance_scale, adapter_scale=None):
        uncond, cond = noise_pred.chunk(2)
        if adapter_scale is not None:
            return adapter_scale * (cond - uncond)
        return guidance_scale * cond + (1 - guidance_scale) * uncond

    def get_noise_pred(self, latent_batches, prompt_embeds_batches, ctrl_images_batches, ctrl_depths_batches,
                       t, tile_weight, depth_weight, guidance_scale, extra_control_batches=None,
                       added_cond_kwargs_batches=None, adapter_scale=None):
        """adapter3d_mixin.py:68-135 (1-pass)."""
        if added_cond_kwargs_batches is not None:
            raise NotImplementedError('added_cond_kwargs (SDXL conditioning) is not built')
        extra_control_batches = list(extra_control_batches or [])
        latent_size = latent_batches[0].size(-1)
        noise_pred = []
        for a, b in _groups(latent_batches):
            lat, pe = _cat(latent_batches, a, b), _cat(prompt_embeds_batches, a, b)
            ci = _cat(ctrl_images_batches, a, b)
            cd = _cat(ctrl_depths_batches, a, b) if ctrl_depths_batches is not None else None
            extra = [_cat(e, a, b) for e in extra_control_batches]
            cak, unet_in, cn_in, unet_pe, cn_pe = self._split_ref(lat, latent_size, pe)
            assert cd is not None or not extra, 'MultiControlNet order is [tile, depth, *extra]'
            nets = MultiControlNet(self.controlnet.nets[:2 + len(extra)] if cd is not None else self.controlnet.nets[:1])
            down, mid = nets(cn_in, t, cn_pe, ([ci, cd] if cd is not None else [ci]) + extra,
                             ([tile_weight, depth_weight] if cd is not None else [tile_weight]) + [1.0] * len(extra))
            if cak is not None:
                down, mid = [_interleave_zero(d) for d in down], _interleave_zero(mid)
            o = self.unet(unet_in, t, unet_pe, cak, down, mid)
            if cak is not None:
                o = o.reshape(lat.shape[0], 2, lat.shape[1], lat.shape[3], lat.shape[3])[:, 1]
            noise_pred.append(o)
        return self._cfg(torch.cat(noise_pred, dim=0), guidance_scale, adapter_scale)

    def get_noise_pred_p1(self, latent_batches, prompt_embeds_batches, t, guidance_scale,
                          ctrl_depths_batches=None, depth_weight=None, extra_control_batches=None,
                          cond_noisy_latent_batches=None, added_cond_kwargs_batches=None):
        """adapter3d_mixin.py:137-237: encoder once, decoder once without the tile ControlNet (depth / extra nets if given)."""
        if cond_noisy_latent_batches is not None or added_cond_kwargs_batches is not None:
            raise NotImplementedError('cond_noisy_latent_batches (attention read/write modes) / added_cond_kwargs are not built')
        extra_control_batches = list(extra_control_batches or [])
        latent_size = latent_batches[0].size(-1)
        noise_pred, dec_args, dec_kwargs = [], [], []
        for a, b in _groups(latent_batches):
            lat, pe = _cat(latent_batches, a, b), _cat(prompt_embeds_batches, a, b)
            cd = _cat(ctrl_depths_batches, a, b) if ctrl_depths_batches is not None else None
            extra = [_cat(e, a, b) for e in extra_control_batches]
            cak, unet_in, cn_in, unet_pe, cn_pe = self._split_ref(lat, latent_size, pe)
            controlnet_skip = 2 if cd is None else 1
            if len(self.controlnet.nets) > controlnet_skip and (cd is not None or extra):
                nets = MultiControlNet(self.controlnet.nets[controlnet_skip:controlnet_skip + (cd is not None) + len(extra)])
                down, mid = nets(cn_in, t, cn_pe, ([cd] if cd is not None else []) + extra,
                                 ([depth_weight] if cd is not None else []) + [1.0] * len(extra))
                if cak is not None:
                    down, mid = [_interleave_zero(d) for d in down], _interleave_zero(mid)
            else:
                down = mid = None
            emb, res, s = self.unet.enc(unet_in, t, unet_pe, cak)
            dec_args.append((emb, res, s))
            dec_kwargs.append(dict(encoder_hidden_states=unet_pe, cross_attention_kwargs=cak,
                                   down_block_additional_residuals=down, mid_block_additional_residual=mid))
            o = self.unet.dec(*dec_args[-1], **dec_kwargs[-1])
            if cak is not None:
                o = o.reshape(lat.shape[0], 2, lat.shape[1], lat.shape[3], lat.shape[3])[:, 1]
            noise_pred.append(o)
        return self._cfg(torch.cat(noise_pred, dim=0), guidance_scale), dec_args, dec_kwargs

    def get_noise_pred_p2(self, latent_batches, prompt_embeds_batches, dec_args, dec_kwargs, t, guidance_scale,
                          ctrl_images_batches, tile_weight, ctrl_depths_batches=None, depth_weight=None,
                          added_cond_kwargs_batches=None, guess_mode=False, adapter_scale=None, ctrl_text_embedding=True,
                          ctrl_is_cfg_duplicate=False):
        """adapter3d_mixin.py:239-317: ControlNets on the fresh renders + decoder only."""
        assert added_cond_kwargs_batches is None and not guess_mode and ctrl_text_embedding
        latent_size = latent_batches[0].size(-1)
        noise_pred = []
        for (a, b), da, dk in zip(_groups(latent_batches), dec_args, dec_kwargs):
            lat, pe = _cat(latent_batches, a, b), _cat(prompt_embeds_batches, a, b)
            ci = _cat(ctrl_images_batches, a, b)
            cd = _cat(ctrl_depths_batches, a, b) if ctrl_depths_batches is not None else None
            ref = lat.shape[2] == 2 * lat.shape[3]
            cn_in = lat[:, :, -latent_size:] if ref else lat
            nets = MultiControlNet(self.controlnet.nets[:1 if cd is None else 2])
            # ctrl_is_cfg_duplicate (extension): the caller built ctrl_* as torch.cat([x] * 2) for the CFG halves
            down, mid = nets(cn_in, t, pe, [ci] if cd is None else [ci, cd], [tile_weight] if cd is None else [tile_weight, depth_weight],
                             cond_repeat=2 if (ctrl_is_cfg_duplicate and not ref) else 1)
            if ref:
                down, mid = [_interleave_zero(d) for d in down], _interleave_zero(mid)
            if dk['down_block_additional_residuals'] is not None:
                down = [x + y for x, y in zip(down, dk['down_block_additional_residuals'])]
                mid = mid + dk['mid_block_additional_residual']
            dk_ = copy(dk)
            dk_.update(dict(down_block_additional_residuals=down, mid_block_additional_residual=mid))
            o = self.unet.dec(*da, **dk_)
            if ref:
                o = o.reshape(lat.shape[0], 2, lat.shape[1], lat.shape[3], lat.shape[3])[:, 1]
            noise_pred.append(o)
        return self._cfg(torch.cat(noise_pred, dim=0), guidance_scale, adapter_scale)

"""Host-side mirror of ``Adapter3DMixin.get_noise_pred{,_p1,_p2}`` -- the denoiser seam B2 of SURVEY.md §8b
(/root/reference/lib/pipelines/adapter3d_mixin.py:68-317).  Same argument lists and return values; ``self.unet`` is a
``mvedit_b200.unet.UNet`` and ``self.controlnet`` a ``MultiControlNet`` whose ``nets`` are ordered [tile, depth, *extra] as in the
reference.  The lists of ``diff_bs``-sized chunks the pipeline passes in are fused: all chunks of equal shape run as ONE batch
(the reference loops chunk by chunk only to bound VRAM, adapter3d.py:546); results are identical per sample.

``dec_args`` / ``dec_kwargs`` stay opaque to the caller (one entry per fused group instead of one per chunk).
Not carried over in this round: ``cond_noisy_latent_batches`` (reference-only attention read/write modes), ``added_cond_kwargs``
(SDXL-style conditioning, unused by SD1.5), IP-Adapter ``adapter_scale`` tokens -- they raise.
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
        assert added_cond_kwargs_batches is None and not extra_control_batches
        latent_size = latent_batches[0].size(-1)
        noise_pred = []
        for a, b in _groups(latent_batches):
            lat, pe = _cat(latent_batches, a, b), _cat(prompt_embeds_batches, a, b)
            ci = _cat(ctrl_images_batches, a, b)
            cd = _cat(ctrl_depths_batches, a, b) if ctrl_depths_batches is not None else None
            cak, unet_in, cn_in, unet_pe, cn_pe = self._split_ref(lat, latent_size, pe)
            nets = MultiControlNet(self.controlnet.nets[:2] if cd is not None else self.controlnet.nets[:1])
            down, mid = nets(cn_in, t, cn_pe, [ci, cd] if cd is not None else [ci],
                             [tile_weight, depth_weight] if cd is not None else [tile_weight])
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
        assert cond_noisy_latent_batches is None and added_cond_kwargs_batches is None and not extra_control_batches
        latent_size = latent_batches[0].size(-1)
        noise_pred, dec_args, dec_kwargs = [], [], []
        for a, b in _groups(latent_batches):
            lat, pe = _cat(latent_batches, a, b), _cat(prompt_embeds_batches, a, b)
            cd = _cat(ctrl_depths_batches, a, b) if ctrl_depths_batches is not None else None
            cak, unet_in, cn_in, unet_pe, cn_pe = self._split_ref(lat, latent_size, pe)
            controlnet_skip = 2 if cd is None else 1
            if len(self.controlnet.nets) > controlnet_skip:
                nets = MultiControlNet(self.controlnet.nets[controlnet_skip:])
                assert cd is not None, 'extra ControlNets (ip2p) are not wired in this round'
                down, mid = nets(cn_in, t, cn_pe, [cd], [depth_weight])
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

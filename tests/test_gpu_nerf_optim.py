"""nerf_optim (mvedit_3d_pipeline.py:452-656) end to end on the GPU: fitting the hash-grid NeRF to synthetic multi-view
targets must reduce the losses and make BaseNeRF.render reproduce the targets (rendered RGBA after k recon iterations)."""
import math

import numpy as np
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu


def _targets(poses, K, size):
    """Analytic target: a sphere of radius 0.5 with position-dependent colour, white background."""
    from oracle.nerf_oracle import get_ray_directions, get_rays
    d = get_ray_directions(size, size, K[None], device='cuda')
    ro, rd = get_rays(d, poses[None], norm=True)
    b = (ro * rd).sum(-1)
    c = (ro * ro).sum(-1) - 0.25
    disc = b * b - c
    hit = disc > 0
    t = -b - disc.clamp(min=0).sqrt()
    p = ro + t[..., None] * rd
    col = 0.5 + 0.5 * torch.sin(p * 6)
    img = torch.where(hit[..., None], col, torch.ones_like(col))
    return img, hit[..., None].float()


@pytest.mark.parametrize('mode', ['adam', 'fused', 'graph'])
def test_nerf_optim_fits_targets(mode):
    """adam: torch.optim.Adam + autograd-accumulated gradients; fused: FusedAdam (flat gradient sink, one-launch update);
    graph: FusedAdam + one CUDA graph per iteration (rays, forward, objective, backward, update)."""
    from mvedit_b200.optim import FusedAdam
    from mvedit_b200.nerf import BaseNeRF, nerf_optim
    from mvedit_b200.ingp_decoder import iNGPDecoder
    torch.manual_seed(0)
    V, size, ps = 6, 64, 32
    poses = torch.from_numpy(synth.surround_poses(V, seed=3)).cuda()
    f = 0.5 * size / math.tan(math.radians(15))
    K = torch.tensor([[f, f, size / 2, size / 2]] * V, device='cuda')
    tgt_images, tgt_masks = _targets(poses, K, size)
    nerf = BaseNeRF(grid_size=64, decoder=iNGPDecoder(max_steps=256, weight_culling_th=0.001), patch_size=ps).cuda()
    grid = nerf.get_init_density_grid(1, 'cuda')
    bitfield = nerf.get_init_density_bitfield(1, 'cuda')
    nerf.decoder.sample_capacity = ps * ps * 2 * 192
    nerf.use_cuda_graph = mode == 'graph'
    opt = torch.optim.Adam(nerf.decoder.parameters(), lr=0.01) if mode == 'adam' else FusedAdam(nerf.decoder.parameters(), lr=0.01)
    cam_w = torch.ones(V, device='cuda')
    lights = torch.nn.functional.normalize(torch.randn(V, 3, device='cuda'), dim=-1)
    kw = dict(optimizer=opt, lr=0.01, n_inverse_rays=ps * ps * 2, patch_rgb_weight=0.0, patch_normal_weight=0.0, alpha_soften=0.02,
              normal_reg_weight=0.1, entropy_weight=0.01, nerf_code=None, density_grid=grid, density_bitfield=bitfield, render_size=size,
              intrinsics=K, intrinsics_size=size, camera_poses=poses, cam_weights=cam_w, cam_lights=lights, patch_size=ps, is_init=True,
              bg_width=0.015, ambient_light=0.2, dt_gamma_scale=0.5, init_shaded=False, debug=(mode != 'graph'))
    log1 = nerf_optim(nerf, tgt_images, tgt_masks, None, inverse_steps=48, **kw)
    log2 = nerf_optim(nerf, tgt_images, tgt_masks, None, inverse_steps=150, **kw)
    if mode != 'graph':
        first = np.mean([l['pixel_rgb'] + l['alpha'] for l in log1[:8]])
        last = np.mean([l['pixel_rgb'] + l['alpha'] for l in log2[-8:]])
        assert last < 0.35 * first, (first, last)
    assert bitfield.sum() > 0
    assert 0 < nerf.decoder.check_sample_overflow(sync=True) <= nerf.decoder.sample_capacity
    img, depth = nerf.render(nerf.decoder, None, bitfield, size, size, K[None], poses[None], cfg=dict(dt_gamma_scale=0.5, return_rgba=True))
    assert img.shape == (1, V, size, size, 4)
    alpha_err = (img[..., 3:] - tgt_masks).abs().mean().item()
    rgb = img[..., :3] + (1 - img[..., 3:])
    rgb_err = (rgb - tgt_images).abs().mean().item()
    assert alpha_err < 0.08 and rgb_err < 0.08, (alpha_err, rgb_err)
    # state-dict contract
    sd = nerf.decoder.state_dict()
    assert set(sd.keys()) == {'aabb', 'encoder.params', 'mlp.net.0.weight', 'mlp.net.0.bias', 'mlp.net.1.weight', 'mlp.net.1.bias'}
    nerf.decoder.backup_state_dict()
    with torch.no_grad():
        nerf.decoder.encoder.params.zero_()
    nerf.decoder.restore_state_dict()
    assert nerf.decoder.encoder.params.abs().sum() > 0


def test_fused_shade_views_matches_torch_restatement():
    """mve_shade_views (inverse-z depth, depth_to_normal, Lambert shading, background compositing, normalize_depth in two launches)
    == the op-by-op torch restatement of base_nerf.py:536-556 / geometry_utils.py:119-168 / mvedit_3d_pipeline.py:1352-1380 in
    oracle/nerf_oracle.py, applied to the same raw render.  Outputs are bf16 in [0,1]: a differently rounded fp32 intermediate may
    flip one bf16 rounding (2^-8 near 1).  Also: BaseNeRF.render(compute_normal=True)'s kernel normals == oracle depth_to_normal."""
    from oracle import nerf_oracle as no
    from mvedit_b200.nerf import BaseNeRF, nerf_optim
    from mvedit_b200.ingp_decoder import iNGPDecoder
    from mvedit_b200.pipeline import MVEdit3DStep
    from mvedit_b200.optim import FusedAdam
    torch.manual_seed(0)
    V, size, ps = 4, 64, 32
    poses = torch.from_numpy(synth.surround_poses(V, seed=3)).cuda()
    f = 0.5 * size / math.tan(math.radians(15))
    K = torch.tensor([[f, f * 1.02, size / 2, size / 2 - 1]] * V, device='cuda') * torch.linspace(1.0, 1.15, V, device='cuda')[:, None]
    tgt_images, tgt_masks = _targets(poses, K, size)
    nerf = BaseNeRF(grid_size=64, decoder=iNGPDecoder(max_steps=256, weight_culling_th=0.001), patch_size=ps).cuda()
    grid, bitfield = nerf.get_init_density_grid(1, 'cuda'), nerf.get_init_density_bitfield(1, 'cuda')
    opt = FusedAdam(nerf.decoder.parameters(), lr=0.01)
    lights = torch.nn.functional.normalize(torch.randn(V, 3, device='cuda'), dim=-1)
    nerf_optim(nerf, tgt_images, tgt_masks, None, optimizer=opt, lr=0.01, inverse_steps=120, n_inverse_rays=ps * ps * 2, patch_rgb_weight=0.0,
               patch_normal_weight=0.0, alpha_soften=0.02, normal_reg_weight=0.1, entropy_weight=0.01, nerf_code=None, density_grid=grid,
               density_bitfield=bitfield, render_size=size, intrinsics=K, intrinsics_size=size, camera_poses=poses,
               cam_weights=torch.ones(V, device='cuda'), cam_lights=lights, patch_size=ps, is_init=True, bg_width=0.015, ambient_light=0.2,
               dt_gamma_scale=0.5, init_shaded=False)
    tone_o = no.Tonemapping().cuda()
    with torch.no_grad():
        for rs, tone in ((64, None), (96, None), (64, tone_o)):     # 96: intrinsics rescaled, size not a multiple of the 32 x 8 pixel CTA tile
            step = MVEdit3DStep(None, None, nerf, None, tonemapping=tone)       # an external module with lut_x / lut_y buffers is adopted
            for render_bs in (None, 3):
                img_f, dep_f = step.render_views(bitfield, poses, K, size, rs, lights, 0.2, 0.5, render_bs=render_bs)
                # oracle chain on the product's raw render, batch by batch as the reference renders (render_bs views per call)
                imgs, alphas, depths = [], [], []
                for pb, kb, lb in zip(poses.split(render_bs or V), K.split(render_bs or V), lights.split(render_bs or V)):
                    rgba, depth, normal, normal_fg = nerf.render(nerf.decoder, None, bitfield, rs, rs, kb[None] * (rs / size), pb[None],
                                                                 cfg=dict(return_rgba=True, compute_normal=True, dt_gamma_scale=0.5))
                    dirs = no.get_ray_directions(rs, rs, kb[None] * (rs / size), device='cuda')
                    nfg_o = no.depth_to_normal(depth / rgba[..., 3].clamp(min=1e-6), dirs)
                    assert float((normal_fg - nfg_o).abs().max()) < 2e-4
                    ncv = torch.cat([normal_fg[..., :1] * 2 - 1, -normal_fg[..., 1:3] * 2 + 1], dim=-1)
                    sh = ((lb[:, None, None, None, :] @ ncv[..., :, None]).clamp(min=0) * 0.8 + 0.2).squeeze(-1)
                    if tone is None:
                        imgs.append((rgba[..., :3] * sh + nerf.bg_color * (1 - rgba[..., 3:])).squeeze(0))
                    else:         # shading in tone-mapped space (mvedit_3d_pipeline.py:1377-1384)
                        imgs.append((tone.lut(tone.inverse_lut(rgba[..., :3] / rgba[..., 3:].clamp(min=1e-6)) + sh.clamp(min=1e-6).log2())
                                     * rgba[..., 3:] + nerf.bg_color * (1 - rgba[..., 3:])).squeeze(0))
                    alphas.append(rgba[..., 3:].squeeze(0)); depths.append(depth.squeeze(0))
                img_t = torch.cat(imgs).to(torch.bfloat16).permute(0, 3, 1, 2).clamp(0, 1)
                al = torch.cat(alphas)
                dep_t = no.normalize_depth(torch.cat(depths), al).to(torch.bfloat16).unsqueeze(1).repeat(1, 3, 1, 1)
                assert img_f.shape == img_t.shape == (V, 3, rs, rs) and dep_f.shape == dep_t.shape
                assert float(dep_t.float().max()) > 0.5 and float((img_t.float() < 0.99).float().mean()) > 0.05      # the object is there
                for a, b in ((img_f, img_t), (dep_f, dep_t)):
                    d = (a.float() - b.float()).abs()
                    assert float(d.max()) <= 2.0 ** -7 and float((d > 0).float().mean()) < 0.02, (float(d.max()), float((d > 0).float().mean()))

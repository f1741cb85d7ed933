"""Host logic of ``MVEditTexturePipeline.__call__`` (mvedit_texture_pipeline.py:175-544) on the CPU: the whole loop -- initial mesh render,
dense camera weights, latents, per-step denoise (stubbed) -> decode (stubbed) -> ``bake_multiview`` -> textured re-render -> solver step,
camera pruning, final ``texture_optim`` and UV bake -- with the rasteriser / texture kernels' code running through tests/host_harness.py
and an analytic field.  The denoiser and the VAE are stand-ins (their kernels are CUDA-only and have their own parity tests)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests import host_harness, synth_mesh
from tests.test_pipeline_mesh_stage_host import AdamLike, ToyDecoder
from mvedit_b200 import mesh_raster as dr
from mvedit_b200 import mvedit_texture_pipeline as TP
from mvedit_b200.mesh_renderer import Mesh, MeshRenderer
from mvedit_b200.nerf import L1LossMod
from mvedit_b200.schedulers import DDIMScheduler, DPMSolverMultistepScheduler


@pytest.fixture(autouse=True)
def _route():
    with host_harness.routed(dr):
        yield


class ToyVAE:
    """8x average-pool 'encoder' / nearest-upsample 'decoder' with diffusers' call shapes."""
    config = SimpleNamespace(scaling_factor=0.5)

    def encode(self, x):
        z = F.avg_pool2d(torch.cat([x, x[:, :1]], dim=1), 8)
        return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda: z, mean=z))

    def decode(self, z, return_dict=False):
        return (F.interpolate(z[:, :3], scale_factor=8, mode='nearest'),)


def _pipeline(scheduler, n, size):
    dec = ToyDecoder()
    nerf = nn.Module()
    nerf.decoder, nerf.bg_color, nerf.grid_size, nerf.pixel_loss, nerf.patch_loss = dec, 1.0, 8, L1LossMod(loss_weight=1.2), None
    unet = nn.Module()
    unet.device = torch.device('cpu')
    pipe = TP.MVEditTexturePipeline(ToyVAE(), None, None, unet, None, scheduler, nerf, MeshRenderer(near=0.01, far=100))
    calls = []

    def fake_noise_pred(lat_b, pe_b, ci_b, cd_b, t, tile_w, depth_w, g, extra_control_batches=None):
        calls.append((lat_b[0].shape, ci_b[0].shape, cd_b[0].shape, float(tile_w)))
        lat = lat_b[-1][..., -lat_b[0].shape[-1]:, :] if lat_b[-1].shape[2] != lat_b[-1].shape[3] else lat_b[-1]
        k = lat.shape[0] // 2 if len(lat_b) == 1 else lat.shape[0]
        return 0.1 * lat[:k].float()
    pipe.get_noise_pred = fake_noise_pred
    v, f = synth_mesh.icosphere(1)
    mesh = Mesh(v=torch.from_numpy(v).float() * 0.5, f=torch.from_numpy(f).int())
    mesh.auto_normal()
    mesh.auto_uv()
    mesh.albedo = torch.full((64, 64, 4), 0.5)
    poses = torch.from_numpy(synth_mesh.surround_poses(n, 1)).float()
    intr = torch.from_numpy(synth_mesh.intrinsics(size)).float()
    return pipe, dec, mesh, poses, intr, calls


@pytest.mark.parametrize('sched', ['ddim', 'dpm'])
def test_texture_pipeline_call_runs_the_whole_loop(monkeypatch, sched):
    monkeypatch.setattr(TP, 'FusedAdam', AdamLike)
    n, size = 4, 32
    scheduler = DDIMScheduler() if sched == 'ddim' else DPMSolverMultistepScheduler()
    pipe, dec, mesh, poses, intr, calls = _pipeline(scheduler, n, size)
    w0 = dec.w.detach().clone()
    bakes = []
    real_bake = pipe.mesh_renderer.bake_multiview
    pipe.mesh_renderer.bake_multiview = lambda *a, **k: (bakes.append((a[1].shape, k.get('cos_weight_pow'))), real_bake(*a, **k))[1]
    out_mesh, state = pipe(in_model=mesh, camera_poses=poses, intrinsics=intr, intrinsics_size=size, use_reference=False, diff_size=size,
                           patch_size=16, render_bs=2, n_inverse_steps=2, num_inference_steps=6, denoising_strength=0.6,
                           max_num_views=lambda p: 4 if p < 0.5 else 3, patch_rgb_weight=lambda p: 0.0, mode='1-pass',
                           prompt_embeds=torch.zeros(2 * n, 77, 8), bake_texture_kwargs=dict(map_size=64))
    assert out_mesh is not None and state is not None
    n_t = len(calls)
    # one multi-view bake per denoising step but the last, whose targets go to texture_optim instead; cos_weight_pow 0
    assert n_t >= 3 and len(bakes) == n_t - 1 and all(b[1] == 0.0 for b in bakes)
    assert calls[0][0] == (2 * 4, 4, 4, 4) and calls[-1][0] == (2 * 3, 4, 4, 4)          # CFG batch; pruned from 4 to 3 views on the way
    assert calls[0][1] == (2 * 4, 3, size, size) and calls[0][2] == (2 * 4, 3, size, size)
    assert 0 < calls[0][3] < 1                                                        # tile weight = sqrt(alpha_bar_t)
    assert out_mesh.albedo.shape == (64, 64, 4) and (out_mesh.albedo[..., :3] - 0.5).abs().max() > 1e-3    # baked from the fitted field
    assert (state['w'] - w0).abs().max() > 1e-5 and torch.equal(dec.w.detach(), w0)   # texture_optim moved the field; weights restored


def test_texture_pipeline_reference_mode_and_two_pass(monkeypatch):
    monkeypatch.setattr(TP, 'FusedAdam', AdamLike)
    n, size = 3, 32
    pipe, dec, mesh, poses, intr, _ = _pipeline(DDIMScheduler(), n, size)
    seen = dict(p1=0, p2=0)

    def p1(lat_b, pe_b, t, g, cd_b=None, dw=None, extra_control_batches=None):
        seen['p1'] += 1
        assert len(lat_b) == 2 and lat_b[0].shape == (n, 4, 4, 4) and lat_b[1].shape == (n, 4, 8, 4) and dw == 1.0   # view-only uncond, ref||view cond
        return 0.1 * lat_b[0].float(), ['args'], ['kwargs']

    def p2(lat_b, pe_b, dec_args, dec_kwargs, t, g, ci_b, tile_w, **kw):
        seen['p2'] += 1
        assert dec_args == ['args'] and ci_b[0].shape == (n, 3, size, size) and kw.get('ctrl_is_cfg_duplicate') is False
        return 0.1 * lat_b[0].float()
    pipe.get_noise_pred_p1, pipe.get_noise_pred_p2 = p1, p2
    out_mesh, state = pipe(in_model=mesh, camera_poses=poses, intrinsics=intr, intrinsics_size=size, use_reference=True, diff_size=size,
                           patch_size=16, render_bs=2, n_inverse_steps=1, num_inference_steps=4, denoising_strength=0.75,
                           max_num_views=lambda p: 3, patch_rgb_weight=lambda p: 0.0, mode='2-pass',
                           prompt_embeds=torch.zeros(2 * n, 77, 8), bake_texture=False)
    assert out_mesh is not None and seen['p1'] == seen['p2'] + 1 >= 3      # the last step runs P1 only: its targets feed texture_optim
    assert out_mesh.albedo.shape == (1024, 1024, 4)                                    # the multi-view-baked map at bake_multiview's default size (no final field bake)


def test_camera_dense_weighting_shape_and_range():
    n, size = 2, 32
    _, _, mesh, poses, intr, _ = _pipeline(DDIMScheduler(), n, size)
    r = MeshRenderer(near=0.01, far=100)
    with torch.no_grad():
        out = r([mesh], poses[None], intr[None].expand(1, n, -1), size, size)
    w = TP.camera_dense_weighting(intr[None].expand(n, -1), size, size, out['rgba'].squeeze(0)[..., 3:], out['depth'].squeeze(0))
    assert w.shape == (n, size, size, 1) and w.min() >= 0 and 0.3 < w.max() <= 1.0 + 1e-5
    assert (w[out['rgba'].squeeze(0)[..., 3:] == 0] == 0).all()


def test_superres_pipeline_runs_and_keeps_unseen_texels_of_the_original(monkeypatch):
    monkeypatch.setattr(TP, 'FusedAdam', AdamLike)
    n, size = 3, 32
    base, dec, mesh, poses, intr, _ = _pipeline(DDIMScheduler(), n, size)
    pipe = TP.MVEditTextureSuperResPipeline(base.vae, None, None, base.unet, None, base.scheduler, base.nerf, base.mesh_renderer)
    calls = []

    def fake(lat_b, pe_b, ci_b, cd_b, t, tile_w, depth_w, g, extra_control_batches=None):
        calls.append((lat_b[0].shape[0], ci_b[0].shape, float(tile_w)))
        return 0.1 * lat_b[0][: lat_b[0].shape[0] // 2].float()
    pipe.get_noise_pred = fake
    seen = {}
    real_optim = pipe.texture_optim
    pipe.texture_optim = lambda *a, **k: (seen.update(n_tgt=a[0].shape[1], n_poses=a[12].shape[0], patch_views=k.get('patch_views')), real_optim(*a, **k))[1]
    mesh.albedo = torch.rand(64, 64, 4, generator=torch.Generator().manual_seed(1))
    ori = mesh.albedo.clone()
    reg_poses = torch.from_numpy(synth_mesh.surround_poses(2, 9)).float()
    out = pipe(in_model=mesh, camera_poses=poses, reg_camera_poses=reg_poses, intrinsics=intr, intrinsics_size=size, use_reference=False,
               diff_size=size, patch_size=20, render_bs=2, n_inverse_steps=2, num_inference_steps=4, denoising_strength=0.5,
               patch_rgb_weight=lambda p: 0.0, prompt_embeds=torch.zeros(2 * n, 77, 8), bake_texture_kwargs=dict(map_size=64))
    assert isinstance(out, Mesh) and out.albedo.shape[:2] == (64, 64)
    assert len(calls) >= 2 and all(c[0] == 2 * n and c[1] == (2 * n, 3, size, size) and c[2] == 1.0 for c in calls)     # fixed tile condition, weight 1
    assert seen == dict(n_tgt=n + 2, n_poses=n + 2, patch_views=n)                     # regulariser views join the fit, not the patch term
    # texels no camera looks at frontally keep (a dilated copy of) the original texture; well-seen texels moved to the fitted field
    d = (out.albedo[..., :3] - ori[..., :3]).abs().max(dim=-1).values
    assert (d < 0.02).float().mean() > 0.1 and (d > 0.05).float().mean() > 0.05

"""Image-to-3D targets of ``nerf_optim`` on the GPU (sorted last on purpose: written after the round's GPU budget was spent, so these
run for the first time in the driver's round-end pass): ``mve_nerf_patch_loss_targets`` / ``mve_nerf_patch_out_normal`` against the
torch chain of tests/test_nerf_loss_host.py (the same checks pass on the CPU build of the kernel source), and ``nerf_optim`` with target
normals, the high-passed normal patch term and target depths fitting an analytic sphere."""
import math

import numpy as np
import pytest
import torch

from tests import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_simplify_mesh_from_device_tensors():
    """``simplify_mesh`` (host C++) fed from CUDA tensors, and the decimated mesh through the rasteriser."""
    from mvedit_b200.mesh_renderer import DMTet, Mesh, MeshRenderer, make_tet_grid, simplify_mesh
    grid = make_tet_grid(24, device='cuda')
    tv = -grid['vertices'] * 2 * 0.9
    sdf = 0.5 - tv.norm(dim=-1) + 0.06 * torch.sin(7 * tv[:, 0]) * torch.sin(6 * tv[:, 1])
    mv, mf = DMTet('cuda')(tv, sdf, grid['indices'])
    v2, f2 = simplify_mesh(mv, mf, mf.shape[0] // 2)
    assert v2.is_cuda and f2.is_cuda and f2.dtype == torch.int64 and mf.shape[0] // 2 - 1 <= f2.shape[0] <= mf.shape[0] // 2
    assert int(f2.max()) == v2.shape[0] - 1
    mesh = Mesh(v=v2, f=f2.int(), device='cuda')
    mesh.auto_normal()
    poses = torch.from_numpy(synth.surround_poses(2, seed=0)).cuda()
    f = 0.5 * 128 / math.tan(math.radians(15))
    out = MeshRenderer(near=0.01, far=100)([mesh], poses[None], torch.tensor([[f, f, 64.0, 64.0]] * 2, device='cuda')[None], 128, 128,
                                           lambda world_pos=None, **kw: torch.full_like(world_pos, 0.5))
    cover = (out['rgba'][0, ..., 3] > 0.5).float().mean().item()
    assert 0.02 < cover < 0.6, cover


@pytest.mark.parametrize('which', ['normal', 'depth', 'patch_normal', 'all', 'all_tone'])
def test_patch_loss_targets_match_torch_chain(which):
    from tests.test_nerf_loss_host import make_inputs, torch_chain, check, NORMAL_BG
    from mvedit_b200.nerf import patch_loss, patch_out_normal
    from mvedit_b200.tonemapping import Tonemapping
    from oracle.nerf_oracle import Tonemapping as OracleTonemapping
    P, ps = 3, 16
    image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, tgt_normal, tgt_depth, probe = [t.cuda() for t in make_inputs(P, ps, 11)]
    tone_o, tone_p = (OracleTonemapping().cuda(), Tonemapping()) if which == 'all_tone' else (None, None)
    use_n, use_d, use_p = which in ('normal', 'all', 'all_tone'), which in ('depth', 'all', 'all_tone'), which in ('patch_normal', 'all', 'all_tone')
    sc = [torch.tensor(v, device='cuda') for v in (1.0, 1.3, 0.02)]
    inp = [t.clone().requires_grad_(True) for t in (image, alpha, depth)]
    ref, ref_normals = torch_chain(*inp, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, True, 0.2, 1.0, 0.015, 1.2, 1.0, 1.3, 0.02, tone=tone_o,
                                   tgt_normal=tgt_normal if use_n else None, tgt_depth=tgt_depth if use_d else None, w_depth=0.7,
                                   normal_probe=probe if use_p else None)
    ref[0].backward()
    out, *grads = patch_loss(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, True, 0.2, 1.0, 0.015, 1.2, *sc,
                             tonemapping=tone_p, tgt_normal=tgt_normal.reshape(-1, 3) if use_n else None,
                             g_normal_extra=probe.reshape(-1, 3) if use_p else None, normal_bg=NORMAL_BG,
                             tgt_depth=tgt_depth.reshape(-1) if use_d else None, w_depth=torch.tensor(0.7, device='cuda') if use_d else None)
    normals = patch_out_normal(alpha, depth, dirs, ps, NORMAL_BG)
    torch.testing.assert_close(normals.view_as(ref_normals), ref_normals.detach(), rtol=1e-4, atol=5e-5)
    expect = ref.detach().clone()
    if use_p:
        expect[0] -= (ref_normals.detach() * probe).sum()
    torch.testing.assert_close(out, expect, rtol=2e-4, atol=1e-6)
    check(grads, inp, alpha)


def test_nerf_optim_with_target_normals_and_depths():
    from oracle.nerf_oracle import get_ray_directions, get_rays
    from mvedit_b200.optim import FusedAdam
    from mvedit_b200.nerf import BaseNeRF, nerf_optim
    from mvedit_b200.ingp_decoder import iNGPDecoder
    from mvedit_b200.lpips import LPIPSLoss, random_lpips_state_dict
    torch.manual_seed(0)
    V, size, ps = 6, 64, 32
    poses = torch.from_numpy(synth.surround_poses(V, seed=3)).cuda()
    f = 0.5 * size / math.tan(math.radians(15))
    K = torch.tensor([[f, f, size / 2, size / 2]] * V, device='cuda')
    # analytic sphere (radius 0.5): colour, mask, camera-space normals (opengl, [0,1]) and 1/z
    d = get_ray_directions(size, size, K[None], device='cuda')
    ro, rd = get_rays(d, poses[None], norm=True)
    b, c = (ro * rd).sum(-1), (ro * ro).sum(-1) - 0.25
    disc = b * b - c
    hit = (disc > 0)[..., None].float()
    t = -b - disc.clamp(min=0).sqrt()
    p = ro + t[..., None] * rd
    tgt_images = torch.where(hit > 0, 0.5 + 0.5 * torch.sin(p * 6), torch.ones_like(p))
    n_world = torch.nn.functional.normalize(p, dim=-1)
    n_cam = torch.einsum('vji,nvhwj->nvhwi', poses[:, :3, :3], n_world)                    # world -> camera (opencv)
    n_gl = torch.stack([n_cam[..., 0], -n_cam[..., 1], -n_cam[..., 2]], -1)
    tgt_normals = torch.where(hit > 0, n_gl / 2 + 0.5, n_gl.new_tensor([0.5, 0.5, 1.0]).expand_as(n_gl)).contiguous()
    z = t * (rd * poses[None, :, None, None, :3, 2]).sum(-1)                               # depth along the optical axis
    tgt_depths = (hit[..., 0] / z.clamp(min=1e-3))[..., None].contiguous()
    nerf = BaseNeRF(grid_size=64, decoder=iNGPDecoder(max_steps=256, weight_culling_th=0.001), patch_size=ps,
                    patch_loss=LPIPSLoss(random_lpips_state_dict(0, 'cuda'), loss_weight=1.2, device='cuda')).cuda()
    grid, bitfield = nerf.get_init_density_grid(1, 'cuda'), nerf.get_init_density_bitfield(1, 'cuda')
    opt = FusedAdam(nerf.decoder.parameters(), lr=0.01)
    lights = torch.nn.functional.normalize(torch.randn(V, 3, device='cuda'), dim=-1)
    kw = dict(optimizer=opt, lr=0.01, n_inverse_rays=ps * ps * 2, patch_rgb_weight=0.1, patch_normal_weight=0.5, alpha_soften=0.02,
              normal_reg_weight=0.1, entropy_weight=0.01, nerf_code=None, density_grid=grid, density_bitfield=bitfield, render_size=size,
              intrinsics=K, intrinsics_size=size, camera_poses=poses, cam_weights=torch.ones(V, device='cuda'), cam_lights=lights,
              patch_size=ps, is_init=True, bg_width=0.015, ambient_light=0.2, dt_gamma_scale=0.5, init_shaded=False, debug=True,
              tgt_depths=tgt_depths, depth_weight=0.5)
    log1 = nerf_optim(nerf, tgt_images, hit, tgt_normals, inverse_steps=40, **kw)
    log2 = nerf_optim(nerf, tgt_images, hit, tgt_normals, inverse_steps=160, **kw)
    for k in ('pixel_rgb', 'alpha', 'depth', 'patch_normal', 'normal_reg'):
        assert all(np.isfinite(l[k]) for l in log1 + log2), k
    first = lambda k: np.mean([l[k] for l in log1[:8]])
    last = lambda k: np.mean([l[k] for l in log2[-8:]])
    assert last('pixel_rgb') + last('alpha') < 0.4 * (first('pixel_rgb') + first('alpha'))
    assert last('depth') < 0.7 * first('depth'), (first('depth'), last('depth'))           # the depth term pulls 1/z to the target
    assert log2[-1]['patch_normal'] > 0
    # graph mode: the same configuration replays as one CUDA graph per iteration
    nerf.use_cuda_graph = True
    kw['debug'] = False
    nerf_optim(nerf, tgt_images, hit, tgt_normals, inverse_steps=6, **kw)
    torch.cuda.synchronize()
    assert all(torch.isfinite(p_).all() for p_ in nerf.decoder.parameters())


# ---- the pipeline call with the options this round added after its GPU budget was spent (first GPU runs) -------------------------------
from tests.test_gpu_pipeline_call import parts, make_pipe, call, _schedulers, N, IMG      # noqa: E402,F401  (``parts`` is a fixture)


def test_call_with_target_normals_and_depths(parts):
    """``use_normal`` with maps handed in + ``depths``: enable_normals / load_depths, then every nerf_optim of the run carries the target
    terms (TV target, depth L1, and -- with an LPIPS patch loss and the default weight schedule -- the high-passed normal patch term)."""
    sch = _schedulers()['euler']()
    pipe, dec = make_pipe(parts, sch, lpips=True)
    before = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    g = torch.Generator(device='cuda').manual_seed(7)
    normals = [torch.nn.functional.normalize(torch.rand(3, 64, 64, device='cuda', generator=g) - 0.5 + torch.tensor([0.0, 0.0, 1.0], device='cuda')[:, None, None],
                                             dim=0) / 2 + 0.5 for _ in range(N)]
    depths = [0.2 + 0.2 * torch.rand(64, 64, device='cuda', generator=g) for _ in range(N)]
    mesh, state = call(pipe, parts, mode='2-pass', use_normal=True, normals=normals, depths=depths, depth_weight=0.3)
    assert mesh is None and state is not None, 'the run raised inside __call__ (traceback printed above)'
    assert all(torch.isfinite(v).all() for v in state.values() if torch.is_floating_point(v))
    assert max(float((state[k].float() - before[k].float()).abs().max()) for k in before if torch.is_floating_point(before[k])) > 1e-3


def test_call_initialises_from_the_field(parts):
    """Without ``init_images`` and ``in_model`` the initial targets are renders of the field handed in (``load_init_nerf``,
    mvedit_3d_pipeline.py:138-173,1060-1064)."""
    sch = _schedulers()['euler']()
    pipe, dec = make_pipe(parts, sch)
    mesh0, state0 = call(pipe, parts, patch_rgb_weight=lambda p: 0.0)                    # a fitted field to start from
    assert state0 is not None
    mesh, state = call(pipe, parts, init_images=None, ingp_states=state0, patch_rgb_weight=lambda p: 0.0, num_inference_steps=4)
    assert mesh is None and state is not None, 'the run raised inside __call__ (traceback printed above)'
    assert all(torch.isfinite(v).all() for v in state.values() if torch.is_floating_point(v))

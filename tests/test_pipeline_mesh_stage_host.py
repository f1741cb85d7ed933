"""Host logic of ``MVEdit3DPipeline.__call__`` past ``progress_to_dmtet`` (mvedit_3d_pipeline.py:1306-1334,1480-1487 of the reference):
NeRF stage -> ``init_tet`` -> two-group optimiser -> ``mesh_optim`` per step -> baked, textured mesh + state dict, and the state-dict
restore.  CPU: ``optim_only=True`` (no denoiser), the NeRF fit stubbed out, an analytic field instead of the CUDA hash grid, the
rasteriser through tests/host_harness.py.  What it checks is the control flow and the argument plumbing of the mesh branch."""
import copy

import pytest
import torch
import torch.nn as nn

from tests import host_harness, synth_mesh
from mvedit_b200 import mesh_raster as dr
from mvedit_b200 import mvedit_3d_pipeline as P
from mvedit_b200.mesh_renderer import MeshRenderer, make_tet_grid
from mvedit_b200.nerf import L1LossMod


@pytest.fixture(autouse=True)
def _route():
    with host_harness.routed(dr):
        yield


class ToyDecoder(nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.w = nn.Parameter(torch.randn(3, 3, generator=g))
        self.b = nn.Parameter(torch.randn(3, generator=g) * 0.1)
        self.grad_sink = None
        self.state_dict_bak = None

    def backup_state_dict(self):
        self.state_dict_bak = copy.deepcopy(self.state_dict())

    def restore_state_dict(self):
        self.load_state_dict(self.state_dict_bak)

    def point_decode(self, xyzs, dirs, code, density_only=False, **kw):
        x = xyzs[0]
        sigma = 40 * (0.45 - x.norm(dim=-1))
        return sigma, (None if density_only else torch.sigmoid(x @ self.w + self.b)), [len(x)]

    def point_density_decode(self, xyzs, code, **kw):
        s, _, n = self.point_decode(xyzs, None, code, density_only=True)
        return s, n


class AdamLike(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, **kw):
        super().__init__(params, lr=lr)

    def set_lr(self, lr, group=None):
        for i, g in enumerate(self.param_groups):
            if group is None or group == i:
                g['lr'] = float(lr)


def test_call_runs_the_dmtet_stage_and_returns_a_baked_mesh(monkeypatch):
    n, size = 4, 32
    calls = dict(nerf=0, mesh=[])
    monkeypatch.setattr(P, 'FusedAdam', AdamLike)
    monkeypatch.setattr(P, 'nerf_optim', lambda *a, **k: calls.__setitem__('nerf', calls['nerf'] + 1))
    real_mesh_optim = P.mesh_optim

    def spy(self, *a, **k):
        calls['mesh'].append((a[6], a[5]))              # (inverse_steps, lr_multiplier)
        return real_mesh_optim(self, *a, **k)
    monkeypatch.setattr(P, 'mesh_optim', spy)

    dec = ToyDecoder()
    w0 = dec.w.detach().clone()
    nerf = nn.Module()
    nerf.decoder, nerf.bg_color, nerf.grid_size, nerf.pixel_loss, nerf.patch_loss = dec, 1.0, 8, L1LossMod(loss_weight=1.2), None
    nerf.get_init_density_grid = lambda ns, device=None: torch.zeros(ns, 8 ** 3, dtype=torch.float16)
    nerf.get_init_density_bitfield = lambda ns, device=None: torch.zeros(ns, 8 ** 3 // 8, dtype=torch.uint8)
    unet = nn.Module()
    unet.device = torch.device('cpu')
    pipe = P.MVEdit3DPipeline(None, None, None, unet, None, None, nerf, mesh_renderer=MeshRenderer(near=0.01, far=100))
    poses = torch.from_numpy(synth_mesh.surround_poses(n, 1)).float()
    intr = torch.from_numpy(synth_mesh.intrinsics(size)).float()
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    init = torch.cat([torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(2)), disc[None, None].expand(n, -1, -1, -1)], dim=1)
    mesh, state = pipe(init_images=init, camera_poses=poses, intrinsics=intr, intrinsics_size=size, use_reference=False, use_normal=False,
                       optim_only=True, num_inference_steps=4, progress_to_dmtet=0.4, tet_resolution=12, tets=make_tet_grid(12),
                       n_inverse_steps=2, init_inverse_steps=2, tet_init_inverse_steps=3, render_bs=2, patch_size=16,
                       render_size_p=lambda p: size, patch_rgb_weight=lambda p: 0.0, mesh_simplify_texture_steps=0,
                       bake_texture=True, bake_texture_kwargs=dict(map_size=64))
    # progress = i / 4 for i = 0..4: NeRF stage at 0, 0.25; DMTet at 0.5 (init: tet_init_inverse_steps), 0.75, 1.0
    assert calls['nerf'] == 2
    assert [c[0] for c in calls['mesh']] == [3, 2, 2]
    assert calls['mesh'][0][1] == pytest.approx(min((1 - 0.5) / (1 - 0.4), 1)) and calls['mesh'][2][1] == pytest.approx(0.0)
    assert mesh is not None and mesh.v.shape[0] > 50 and mesh.f.shape[1] == 3 and not mesh.v.requires_grad
    assert mesh.albedo.shape == (64, 64, 4) and mesh.vt is not None and mesh.textureless is False
    assert set(state.keys()) == {'w', 'b'} and (state['w'] - w0).abs().max() > 1e-4        # the returned state is the optimised field ...
    assert torch.equal(dec.w.detach(), w0)                                                # ... and the module's own weights are restored


def test_call_without_mesh_renderer_raises_cleanly(monkeypatch):
    monkeypatch.setattr(P, 'FusedAdam', AdamLike)
    monkeypatch.setattr(P, 'nerf_optim', lambda *a, **k: None)
    dec = ToyDecoder()
    nerf = nn.Module()
    nerf.decoder, nerf.bg_color, nerf.grid_size, nerf.pixel_loss, nerf.patch_loss = dec, 1.0, 8, L1LossMod(loss_weight=1.2), None
    nerf.get_init_density_grid = lambda ns, device=None: torch.zeros(ns, 8 ** 3, dtype=torch.float16)
    nerf.get_init_density_bitfield = lambda ns, device=None: torch.zeros(ns, 8 ** 3 // 8, dtype=torch.uint8)
    unet = nn.Module()
    unet.device = torch.device('cpu')
    pipe = P.MVEdit3DPipeline(None, None, None, unet, None, None, nerf, mesh_renderer=None)
    poses = torch.from_numpy(synth_mesh.surround_poses(2, 1)).float()
    init = torch.rand(2, 4, 16, 16)
    mesh, state = pipe(init_images=init, camera_poses=poses, intrinsics=torch.tensor([30.0, 30.0, 8.0, 8.0]), intrinsics_size=16,
                       use_reference=False, use_normal=False, optim_only=True, num_inference_steps=2, progress_to_dmtet=0.4)
    assert mesh is None and state is None               # the reference's behaviour on any failure: print the traceback, return (None, None)
    assert dec.state_dict_bak is not None


def test_call_initialises_from_an_input_mesh(monkeypatch):
    """3D-to-3D: without init_images the targets of the initial fit are renders of ``in_model`` shaded by the sampled lights
    (mvedit_3d_pipeline.py:1052-1066), with ``init_shaded`` set."""
    from mvedit_b200.mesh_renderer import Mesh
    monkeypatch.setattr(P, 'FusedAdam', AdamLike)
    seen = []
    monkeypatch.setattr(P, 'nerf_optim', lambda nerf, tgt_images, tgt_masks, *a, **k: seen.append((tgt_images, tgt_masks, a[24])))    # a[24] = init_shaded
    dec = ToyDecoder()
    nerf = nn.Module()
    nerf.decoder, nerf.bg_color, nerf.grid_size, nerf.pixel_loss, nerf.patch_loss = dec, 1.0, 8, L1LossMod(loss_weight=1.2), None
    nerf.get_init_density_grid = lambda ns, device=None: torch.zeros(ns, 8 ** 3, dtype=torch.float16)
    nerf.get_init_density_bitfield = lambda ns, device=None: torch.zeros(ns, 8 ** 3 // 8, dtype=torch.uint8)
    unet = nn.Module()
    unet.device = torch.device('cpu')
    pipe = P.MVEdit3DPipeline(None, None, None, unet, None, None, nerf, mesh_renderer=MeshRenderer(near=0.01, far=100))
    v, f = synth_mesh.icosphere(1)
    mesh = Mesh(v=torch.from_numpy(v).float() * 0.5, f=torch.from_numpy(f).int(), vc=torch.cat([torch.rand(1, len(v), 3), torch.ones(1, len(v), 1)], dim=-1))
    poses = torch.from_numpy(synth_mesh.surround_poses(2, 1)).float()
    pipe(in_model=mesh, camera_poses=poses, intrinsics=torch.from_numpy(synth_mesh.intrinsics(32)).float(), intrinsics_size=32,
         use_reference=False, use_normal=False, optim_only=True, num_inference_steps=1, progress_to_dmtet=0.9, tet_resolution=8, render_bs=2,
         render_size_p=lambda p: 512, max_num_views=lambda p, q: 2, bake_texture=False, tet_init_inverse_steps=0, n_inverse_steps=0)
    tgt_images, tgt_masks, init_shaded = seen[0]
    assert tgt_images.shape == (1, 2, 512, 512, 3) and tgt_masks.shape == (1, 2, 512, 512, 1) and init_shaded is True
    fg = tgt_masks[0, ..., 0] > 0.99
    assert 0.05 < fg.float().mean() < 0.5 and (tgt_images[0][~(tgt_masks[0, ..., 0] > 0)] == 1.0).all()       # composited on the background colour
    assert tgt_images[0][fg].std() > 0.02                                                                   # shaded vertex colours, not flat


def test_call_threads_target_normals_and_depths_through_both_stages(monkeypatch):
    """``use_normal`` with maps handed in and ``depths``: re-shaded inputs (``init_shaded``), targets re-ordered with ``keep_views``,
    resized to the render size and passed to ``nerf_optim`` (with the depth weight) and to ``mesh_optim`` (which runs here, incl. the
    high-passed normal patch term) -- mvedit_3d_pipeline.py:1082-1090,1166-1169,1272-1295,1298-1332."""
    from tests.test_mesh_stage_host import _FakePatchLoss
    n, size = 4, 32
    monkeypatch.setattr(P, 'FusedAdam', AdamLike)
    seen = dict(nerf=[], mesh=[])
    monkeypatch.setattr(P, 'nerf_optim', lambda nerf, ti, tm, tn, *a, **k: seen['nerf'].append((ti, tn, a[5], a[-1], k['tgt_depths'], k['depth_weight'])))
    real_mesh_optim = P.mesh_optim

    def spy(self, ti, tm, tn, *a, **k):
        seen['mesh'].append((tn, a[8]))                  # (tgt_normals, patch_normal_weight)
        return real_mesh_optim(self, ti, tm, tn, *a, **k)
    monkeypatch.setattr(P, 'mesh_optim', spy)
    dec = ToyDecoder()
    nerf = nn.Module()
    nerf.decoder, nerf.bg_color, nerf.grid_size, nerf.pixel_loss, nerf.patch_loss = dec, 1.0, 8, L1LossMod(loss_weight=1.2), _FakePatchLoss()
    nerf.get_init_density_grid = lambda ns, device=None: torch.zeros(ns, 8 ** 3, dtype=torch.float16)
    nerf.get_init_density_bitfield = lambda ns, device=None: torch.zeros(ns, 8 ** 3 // 8, dtype=torch.uint8)
    unet = nn.Module()
    unet.device = torch.device('cpu')
    pipe = P.MVEdit3DPipeline(None, None, None, unet, None, None, nerf, mesh_renderer=MeshRenderer(near=0.01, far=100))
    poses = torch.from_numpy(synth_mesh.surround_poses(n, 1)).float()
    intr = torch.from_numpy(synth_mesh.intrinsics(size)).float()
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    init = torch.cat([torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(2)), disc[None, None].expand(n, -1, -1, -1)], dim=1)
    colours = torch.tensor([[0.9, 0.5, 0.8], [0.5, 0.9, 0.8], [0.2, 0.5, 0.8], [0.5, 0.2, 0.8]])
    normals = [c[:, None, None].expand(3, size, size) for c in colours]                   # one flat normal map per view, all different
    depths = [torch.full((size, size), 0.1 * (k + 1)) for k in range(n)]
    with pytest.raises(NotImplementedError):                                              # no maps and no normal model
        pipe(init_images=init, camera_poses=poses, intrinsics=intr, intrinsics_size=size, use_reference=False, optim_only=True,
             num_inference_steps=1)
    mesh, state = pipe(init_images=init, camera_poses=poses, intrinsics=intr, intrinsics_size=size, use_reference=False, use_normal=True,
                       normals=normals, depths=depths, depth_weight=0.3, keep_views=[2], optim_only=True, num_inference_steps=2,
                       progress_to_dmtet=0.4, tet_resolution=12, tets=make_tet_grid(12), n_inverse_steps=1, init_inverse_steps=1,
                       tet_init_inverse_steps=1, render_bs=2, patch_bs=2, patch_size=16, render_size_p=lambda p: size,
                       patch_rgb_weight=lambda p: 0.0, patch_normal_weight=lambda p: 0.5, mesh_simplify_texture_steps=0, bake_texture=False)
    assert mesh is not None and state is not None
    ti, tn, pnw, init_shaded, td, dw = seen['nerf'][0]
    assert tn.shape == (1, n, size, size, 3) and td.shape == (1, n, size, size, 1) and pnw == 0.5 and dw == 0.3 and init_shaded is True
    fg = disc > 0.5
    # keep_views=[2] moves view 2 first: its (normalised) flat normal and its depth come first
    first = torch.nn.functional.normalize(colours[2] * 2 - 1, dim=0) / 2 + 0.5
    assert torch.allclose(tn[0, 0][fg].mean(0), first, atol=2e-3) and torch.allclose(td[0, 0][fg], torch.tensor(0.3), atol=1e-5)
    assert torch.allclose(tn[0, 0][~fg & (disc < 0.01)].mean(0), torch.tensor([0.5, 0.5, 1.0]), atol=1e-3)      # background normal
    assert len(seen['mesh']) == 2 and all(m[0].shape == (1, n, size, size, 3) and m[1] == 0.5 for m in seen['mesh'])

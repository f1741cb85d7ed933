"""CPU checks of the mesh stage against oracle/mesh_oracle.py: ``MeshRenderer.forward`` and a few ``mesh_optim``
iterations, with the rasteriser kernels' code running through tests/host_harness.py and an analytic stand-in for the hash-grid field
(the field kernel is CUDA-only; its own parity tests are tests/test_gpu_field.py).  The GPU version of the same comparison with the
real field is tests/test_gpu_zmesh.py."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import mesh_oracle as mo
from oracle.nerf_oracle import L1LossMod
from tests import host_harness, synth_mesh
from mvedit_b200 import mesh_raster as dr
from mvedit_b200 import mesh_optim as mopt
from mvedit_b200.mesh_renderer import DMTet, Mesh, MeshRenderer, make_tet_grid
from types import SimpleNamespace


@pytest.fixture(autouse=True)
def _route():
    with host_harness.routed(dr):
        yield


class ToyField(nn.Module):
    """Stand-in for iNGPDecoder on the CPU: a smooth density blob and an albedo that depends on position through two parameters."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.w = nn.Parameter(torch.randn(3, 3, generator=g))
        self.b = nn.Parameter(torch.randn(3, generator=g) * 0.1)
        self.grad_sink = None

    def point_decode(self, xyzs, dirs, code, density_only=False, **kw):
        x = xyzs[0]
        sigma = 40 * (0.45 - x.norm(dim=-1)) + 4 * torch.sin(5 * x[:, 0]) * torch.sin(4 * x[:, 1])
        rgb = None if density_only else torch.sigmoid(x @ self.w + self.b)
        return sigma, rgb, [len(x)]

    def point_density_decode(self, xyzs, code, **kw):
        s, _, n = self.point_decode(xyzs, None, code, density_only=True)
        return s, n


def _cameras(n, size, seed=0):
    poses = torch.from_numpy(synth_mesh.surround_poses(n, seed)).float()
    intr = torch.from_numpy(synth_mesh.intrinsics(size)).float()[None].expand(n, -1).contiguous()
    return poses, intr


def test_mesh_renderer_forward_and_gradients_match_oracle():
    v, f = synth_mesh.icosphere(2)
    size, n = 40, 3
    poses, intr = _cameras(n, size)
    field = ToyField()
    lights = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(1)), dim=-1)
    lights_px = lights[:, None, None, :].expand(-1, size, size, -1)
    out = {}
    for name in ('product', 'oracle'):
        vt = (torch.from_numpy(v).float() * 0.55).requires_grad_(True)
        field.zero_grad()
        if name == 'product':
            mesh = Mesh(v=vt, f=torch.from_numpy(f).int())
            mesh.auto_normal()
            r = MeshRenderer(near=0.01, far=100)([mesh], poses[None], intr[None], size, size,
                                                 mopt.make_nerf_shading_fun(field, None, lights_px, 0.2), normal_bg=[0.5, 0.5, 1.0])
        else:
            mesh = mo.make_mesh(vt, torch.from_numpy(f).int())
            r = mo.mesh_renderer_forward(mesh, poses[None], intr[None], size, size,
                                         mopt.make_nerf_shading_fun(field, None, lights_px, 0.2), normal_bg=(0.5, 0.5, 1.0))
        g = torch.Generator().manual_seed(2)
        loss = sum((r[k] * torch.randn(r[k].shape, generator=g)).sum() for k in ('rgba', 'depth', 'normal'))
        loss.backward()
        out[name] = dict(r={k: x.detach() for k, x in r.items()}, gv=vt.grad.clone(), gw=field.w.grad.clone())
    p, o = out['product'], out['oracle']
    for k in ('rgba', 'depth', 'normal'):
        assert p['r'][k].shape == o['r'][k].shape
        torch.testing.assert_close(p['r'][k], o['r'][k], rtol=1e-4, atol=2e-5)
    assert p['r']['rgba'].shape == (1, n, size, size, 4) and 0.05 < p['r']['rgba'][..., 3].mean() < 0.6
    assert (p['gv'] - o['gv']).abs().max() <= 2e-3 * o['gv'].abs().max()
    assert (p['gw'] - o['gw']).abs().max() <= 2e-3 * o['gw'].abs().max()


def test_vertex_colour_path_and_no_grad_render():
    v, f = synth_mesh.icosphere(1)
    poses, intr = _cameras(2, 32)
    vt = torch.from_numpy(v).float() * 0.5
    vc = torch.rand(1, vt.shape[0], 4, generator=torch.Generator().manual_seed(3))
    mesh = Mesh(v=vt, f=torch.from_numpy(f).int(), vc=vc)
    mesh.auto_normal()
    with torch.no_grad():
        r = MeshRenderer(near=0.01, far=100)([mesh], poses[None], intr[None], 32, 32)
    om = mo.make_mesh(vt, torch.from_numpy(f).int())
    om.vc = vc
    ro_ = mo.mesh_renderer_forward(om, poses[None], intr[None], 32, 32)
    for k in ('rgba', 'depth', 'normal'):
        torch.testing.assert_close(r[k], ro_[k], rtol=1e-4, atol=2e-5)


def test_textured_mesh_forward_matches_oracle_and_gradient_reaches_the_texture():
    v, f = synth_mesh.icosphere(2)
    poses, intr = _cameras(2, 40)
    vt_v = torch.from_numpy(v).float() * 0.5
    mesh = Mesh(v=vt_v, f=torch.from_numpy(f).int())
    mesh.auto_normal()
    mesh.auto_uv()
    assert mesh.vt.shape == (3 * len(f), 2) and mesh.ft.shape == (len(f), 3) and mesh.vt.min() >= 0 and mesh.vt.max() <= 1
    albedo = torch.rand(32, 32, 4, generator=torch.Generator().manual_seed(7)).requires_grad_(True)
    mesh.albedo = albedo
    r = MeshRenderer(near=0.01, far=100)([mesh], poses[None], intr[None], 40, 40)
    om = mo.make_mesh(vt_v, torch.from_numpy(f).int())
    om.vt, om.ft = mesh.vt, mesh.ft
    om.albedo = albedo.detach().clone().requires_grad_(True)
    r_o = mo.mesh_renderer_forward(om, poses[None], intr[None], 40, 40)
    for k in ('rgba', 'depth', 'normal'):
        torch.testing.assert_close(r[k].detach(), r_o[k].detach(), rtol=1e-4, atol=2e-5)
    w = torch.randn(r['rgba'].shape, generator=torch.Generator().manual_seed(8))
    (r['rgba'] * w).sum().backward()
    (r_o['rgba'] * w).sum().backward()
    assert albedo.grad[..., :3].abs().sum() > 0 and (albedo.grad[..., 3] == 0).all()
    torch.testing.assert_close(albedo.grad, om.albedo.grad, rtol=1e-4, atol=1e-4)


def test_range_mode_raises():
    v, f = synth_mesh.icosphere(0)
    poses, intr = _cameras(1, 16)
    mesh = Mesh(v=torch.from_numpy(v).float() * 0.5, f=torch.from_numpy(f).int())
    mesh.auto_normal()
    with pytest.raises(NotImplementedError):
        MeshRenderer()([mesh, mesh], poses[None].expand(2, -1, -1, -1), intr[None].expand(2, -1, -1), 16, 16)


def test_init_tet_and_mesh_optim_match_oracle():
    torch.manual_seed(0)
    n, size, steps = 4, 32, 3
    poses, intr = _cameras(n, size, seed=2)
    lights = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(4)), dim=-1)
    cam_weights = torch.tensor([1.0, 0.5, 1.0, 2.0])
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    tgt_masks = disc[None, None, :, :, None].expand(1, n, -1, -1, -1).contiguous()
    tgt_images = (torch.rand(1, n, size, size, 3, generator=torch.Generator().manual_seed(5)) * 0.5 + 0.25) * tgt_masks + (1 - tgt_masks)
    noise = dict(camera_perm=torch.tensor([2, 0, 3, 1]), jitter=torch.rand(steps, 2, 2, generator=torch.Generator().manual_seed(6)))
    grid = make_tet_grid(12)
    results = {}
    for name in ('product', 'oracle'):
        field = ToyField()
        nerf = SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=None)
        tet_verts, tet_indices, tet_sdf = mopt.init_tet(nerf, None, density_thresh=5.0, tets=grid)
        assert (tet_sdf > 0).sum() > 20 and (tet_sdf < 0).sum() > 20
        deform = torch.zeros_like(tet_verts).requires_grad_(True)
        tet_sdf.requires_grad_(True)
        opt = torch.optim.Adam([{'params': list(field.parameters())}, {'params': [tet_sdf, deform], 'lr': 1e-3}], lr=0.01)
        if name == 'product':
            dm = DMTet('cpu')
            with torch.enable_grad():
                mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
                mesh = Mesh(v=mv, f=mf.int())
                mesh.auto_normal()
            pipe = SimpleNamespace(nerf=nerf, mesh_renderer=MeshRenderer(near=0.01, far=100), normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
            mesh = mopt.mesh_optim(pipe, tgt_images, tgt_masks, None, opt, 0.01, 0.8, steps, 2, 8, 24, 0.0, 0.0, 0.02, 0.1, 5.0, None,
                                   tet_verts, deform, tet_sdf, tet_indices, dm, mesh, size, intr, size, poses, cam_weights, lights, 16,
                                   False, 0.2, 1.0, noise=noise)
        else:
            dm = mo.DMTetOracle()
            mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
            mesh, _ = mo.mesh_optim(field, tgt_images, tgt_masks, opt, 0.01, 0.8, steps, 2, 8, 0.0, 0.02, 0.1, 5.0, None, tet_verts, deform,
                                    tet_sdf, tet_indices, dm, mo.make_mesh(mv, mf.int()), size, intr, size, poses, cam_weights, lights, 16,
                                    0.2, noise)
        results[name] = dict(sdf=tet_sdf.detach().clone(), deform=deform.detach().clone(), w=field.w.detach().clone(), f=mesh.f.clone(),
                             v=mesh.v.detach().clone())
    p, o = results['product'], results['oracle']
    assert (p['sdf'] - o['sdf']).abs().max() < 2e-5 and (p['deform'] - o['deform']).abs().max() < 2e-5
    assert (p['w'] - o['w']).abs().max() < 2e-5
    assert (p['deform'].abs().max() > 1e-4) and (p['w'] - ToyField().w.detach()).abs().max() > 1e-3      # everything moved
    assert p['f'].shape == o['f'].shape and (p['f'] == o['f']).all()
    torch.testing.assert_close(p['v'], o['v'], rtol=1e-4, atol=2e-5)


def _textured_sphere(subdiv=2):
    v, f = synth_mesh.icosphere(subdiv)
    mesh = Mesh(v=torch.from_numpy(v).float() * 0.5, f=torch.from_numpy(f).int())
    mesh.auto_normal()
    mesh.auto_uv()
    om = mo.make_mesh(mesh.v, mesh.f)
    om.vt, om.ft = mesh.vt, mesh.ft
    return mesh, om


def test_bake_multiview_and_cam_weights_match_oracle():
    mesh, om = _textured_sphere()
    n, size, ms = 3, 32, 64
    poses, intr = _cameras(n, size, seed=3)
    g = torch.Generator().manual_seed(9)
    images = torch.rand(1, n, size, size, 3, generator=g)
    alphas = (torch.rand(1, n, size, size, 1, generator=g) > 0.1).float()
    r = MeshRenderer(near=0.01, far=100)
    wts, valid = r.get_cam_weights_uv([mesh], poses[None], intr[None], alphas=alphas[0], render_size=size, map_size=ms, render_bs=2, cos_weight_pow=1.0)
    wts_o, valid_o = mo.bake_multiview(om, images, alphas, poses[None], intr[None], map_size=ms, cos_weight_pow=1.0, render_bs=2, weights_only=True)
    assert wts.shape == (1, n, ms, ms, 1) and (valid == valid_o).all() and 0.2 < valid.float().mean() < 0.9
    torch.testing.assert_close(wts, wts_o, rtol=1e-3, atol=1e-4)
    assert wts.sum() > 1
    baked = r.bake_multiview([mesh], images, alphas, poses[None], intr[None], map_size=ms, cos_weight_pow=8.0, render_bs=2)[0]
    albedo_o = mo.bake_multiview(om, images, alphas, poses[None], intr[None], map_size=ms, cos_weight_pow=8.0, render_bs=2)
    assert baked.albedo.shape == (ms, ms, 4) and baked.textureless is False
    d = (baked.albedo - albedo_o).abs()
    assert d.mean() < 1e-4 and (d > 1e-2).float().mean() < 2e-3        # a texel whose total weight is ~1e-8 may round differently


def test_bake_xyz_shading_fun_matches_oracle_and_renders_back():
    mesh, om = _textured_sphere(1)
    field = ToyField()
    fun = mopt.make_nerf_albedo_shading_fun(field, None)
    r = MeshRenderer(near=0.01, far=100)
    with torch.no_grad():
        baked = r.bake_xyz_shading_fun([mesh], fun, map_size=256)[0]
        albedo_o = mo.bake_xyz_shading_fun(om, fun, map_size=256)
    torch.testing.assert_close(baked.albedo, albedo_o, rtol=1e-4, atol=2e-5)
    # rendering the baked texture reproduces the field's albedo at the visible surface points
    poses, intr = _cameras(2, 48, seed=5)
    with torch.no_grad():
        out = r([baked], poses[None], intr[None], 48, 48, aa=False)
        ref = r([Mesh(v=mesh.v, f=mesh.f, vn=mesh.vn, fn=mesh.fn)], poses[None], intr[None], 48, 48, shading_fun=lambda world_pos=None, **kw: fun(world_pos=world_pos), aa=False)
    fg = out['rgba'][..., 3] > 0
    assert (out['rgba'][..., :3][fg] - ref['rgba'][..., :3][fg]).abs().mean() < 0.03


def test_texture_optim_matches_oracle():
    mesh, om = _textured_sphere(2)
    n, size, steps = 4, 32, 3
    poses, intr = _cameras(n, size, seed=4)
    g = torch.Generator().manual_seed(12)
    tgt = torch.rand(1, n, size, size, 3, generator=g)
    w_dense = torch.rand(n, size, size, 1, generator=g)
    noise = dict(camera_perm=torch.tensor([1, 3, 0, 2]), jitter=torch.rand(steps, 2, 2, generator=g))
    res = {}
    for name in ('product', 'oracle'):
        field = ToyField()
        opt = torch.optim.Adam(field.parameters(), lr=0.01)
        if name == 'product':
            nerf = SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=None)
            pipe = SimpleNamespace(nerf=nerf, mesh_renderer=MeshRenderer(near=0.01, far=100), bg_color=0.5)
            mopt.texture_optim(pipe, tgt, opt, 0.02, steps, 2, 8, 0.0, None, mesh, size, intr, size, poses, w_dense, 16, noise=noise)
        else:
            mo.texture_optim(field, tgt, opt, 0.02, steps, 2, None, om, size, intr, size, poses, w_dense, noise)
        res[name] = (field.w.detach().clone(), field.b.detach().clone())
    for a, b in zip(res['product'], res['oracle']):
        assert (a - b).abs().max() < 2e-5
    assert (res['product'][0] - ToyField().w.detach()).abs().max() > 1e-2


def test_render_mesh_views_matches_oracle_composition():
    """The mesh branch of the per-step render (mvedit_3d_pipeline.py:1341-1360,1391-1396): shaded rgb over the background colour and the
    normalised inverse depth, as bf16 NCHW ControlNet inputs."""
    from oracle.nerf_oracle import normalize_depth
    v, f = synth_mesh.icosphere(2)
    n, size = 3, 32
    poses, intr = _cameras(n, size, seed=6)
    lights = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(1)), dim=-1)
    field = ToyField()
    vt = torch.from_numpy(v).float() * 0.5
    mesh = Mesh(v=vt, f=torch.from_numpy(f).int())
    mesh.auto_normal()
    pipe = SimpleNamespace(nerf=SimpleNamespace(decoder=field, bg_color=1.0), mesh_renderer=MeshRenderer(near=0.01, far=100),
                           normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
    with torch.no_grad():
        images, depths = mopt.render_mesh_views(pipe, mesh, None, poses, intr, size * 2, size, lights, 0.2, render_bs=2)
        lp = lights[:, None, None, :].expand(-1, size, size, -1)
        r_o = mo.mesh_renderer_forward(mo.make_mesh(vt, torch.from_numpy(f).int()), poses[None], intr[None] * 0.5, size, size,
                                       mopt.make_nerf_shading_fun(field, None, lp, 0.2))
    assert images.shape == depths.shape == (n, 3, size, size) and images.dtype == depths.dtype == torch.bfloat16
    rgba = r_o['rgba'].squeeze(0)
    img_o = (rgba[..., :3] + 1.0 * (1 - rgba[..., 3:])).permute(0, 3, 1, 2).clamp(0, 1)
    dep_o = normalize_depth(r_o['depth'].squeeze(0), rgba[..., 3:]).unsqueeze(1).repeat(1, 3, 1, 1)
    assert (images.float() - img_o).abs().max() < 1e-2 and (depths.float() - dep_o).abs().max() < 1e-2     # bf16 storage
    assert depths.float().max() > 0.9 and (depths.float()[:, 0][rgba[..., 3] == 0] == 0).all()


class _FakePatchLoss:
    """Stands in for ``LPIPSLoss`` (CUDA-only) with the same two entry points: ``loss_and_grad`` on NHWC patches (what the product
    calls, gradient returned instead of recorded) and the reference-style ``__call__`` on NCHW with autograd (what the oracle calls).
    The metric is a weighted MSE, so both sides must agree exactly."""
    loss_weight = 1.2

    def loss_and_grad(self, pred, target, weight=None, scale=1.0):
        P = pred.shape[0]
        w = torch.ones(P) if weight is None else weight
        per = (pred - target).square().flatten(1).mean(1)
        loss = (per * w).mean() * scale * self.loss_weight
        g = 2 * (pred - target) / pred[0].numel() * (w * scale * self.loss_weight / P)[:, None, None, None]
        return loss, g, per

    def __call__(self, pred, target, weight=None, avg_factor=None):
        per = (pred - target).square().flatten(1).mean(1)
        return ((per if weight is None else per * weight).mean() * self.loss_weight)


def test_mesh_optim_patch_term_plumbing_matches_oracle():
    n, size, steps, ps = 2, 32, 2, 16
    poses, intr = _cameras(n, size, seed=2)
    lights = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(4)), dim=-1)
    cam_weights = torch.tensor([1.0, 2.0])
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    tgt_masks = disc[None, None, :, :, None].expand(1, n, -1, -1, -1).contiguous()
    tgt_images = (torch.rand(1, n, size, size, 3, generator=torch.Generator().manual_seed(5)) * 0.5 + 0.25) * tgt_masks + (1 - tgt_masks)
    g = torch.Generator().manual_seed(6)
    noise = dict(camera_perm=torch.tensor([1, 0]), jitter=torch.rand(steps, 2, 2, generator=g),
                 patch_perm=torch.stack([torch.randperm(n * (size // ps) ** 2, generator=g) for _ in range(steps)]))
    grid = make_tet_grid(10)
    res = {}
    for name in ('product', 'oracle'):
        field = ToyField()
        nerf = SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss())
        tet_verts, tet_indices, tet_sdf = mopt.init_tet(nerf, None, density_thresh=5.0, tets=grid)
        deform = torch.zeros_like(tet_verts).requires_grad_(True)
        tet_sdf.requires_grad_(True)
        opt = torch.optim.Adam([{'params': list(field.parameters())}, {'params': [tet_sdf, deform], 'lr': 1e-3}], lr=0.01)
        if name == 'product':
            dm = DMTet('cpu')
            with torch.enable_grad():
                mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
                mesh = Mesh(v=mv, f=mf.int())
                mesh.auto_normal()
            pipe = SimpleNamespace(nerf=nerf, mesh_renderer=MeshRenderer(near=0.01, far=100), normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
            mopt.mesh_optim(pipe, tgt_images, tgt_masks, None, opt, 0.01, 0.8, steps, 2, 3, 24, 0.7, 0.0, 0.02, 0.1, 5.0, None,
                            tet_verts, deform, tet_sdf, tet_indices, dm, mesh, size, intr, size, poses, cam_weights, lights, ps,
                            False, 0.2, 1.0, noise=noise)
        else:
            dm = mo.DMTetOracle()
            mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
            mo.mesh_optim(field, tgt_images, tgt_masks, opt, 0.01, 0.8, steps, 2, 3, 0.7, 0.02, 0.1, 5.0, None, tet_verts, deform, tet_sdf,
                          tet_indices, dm, mo.make_mesh(mv, mf.int()), size, intr, size, poses, cam_weights, lights, ps, 0.2, noise,
                          patch_loss=nerf.patch_loss)
        res[name] = (tet_sdf.detach().clone(), deform.detach().clone(), field.w.detach().clone())
    for a, b in zip(res['product'], res['oracle']):
        assert (a - b).abs().max() < 2e-5


def test_ssaa_and_edge_dilation_branches_match_oracle():
    v, f = synth_mesh.icosphere(1)
    poses, intr = _cameras(2, 24)
    vt = torch.from_numpy(v).float() * 0.5
    vc = torch.rand(1, vt.shape[0], 4, generator=torch.Generator().manual_seed(3)) * 0.5 + 0.5
    mesh = Mesh(v=vt, f=torch.from_numpy(f).int(), vc=vc)
    mesh.auto_normal()
    om = mo.make_mesh(vt, torch.from_numpy(f).int())
    om.vc = vc
    with torch.no_grad():
        r = MeshRenderer(near=0.01, far=100, ssaa=2)([mesh], poses[None], intr[None], 24, 24)
        r_o = mo.mesh_renderer_forward(om, poses[None], intr[None], 24, 24, ssaa=2)
        d = MeshRenderer(near=0.01, far=100)([mesh], poses[None], intr[None], 24, 24, dilate_edges=2, aa=False)
        d_o = mo.mesh_renderer_forward(om, poses[None], intr[None], 24, 24, dilate_edges=2, aa=False)
        plain = MeshRenderer(near=0.01, far=100)([mesh], poses[None], intr[None], 24, 24, aa=False)
    for k in ('rgba', 'depth', 'normal'):
        assert r[k].shape[2:4] == (24, 24)
        torch.testing.assert_close(r[k], r_o[k], rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(d[k], d_o[k], rtol=1e-4, atol=2e-5)
    assert ((r['rgba'][..., 3] > 0.05) & (r['rgba'][..., 3] < 0.95)).sum() > 8        # area-averaged 2x2 supersamples: soft outline
    grown = (d['rgba'][..., 3] > 0).sum() - (plain['rgba'][..., 3] > 0).sum()
    assert grown > 20                                                               # the coverage grew by the dilation ring


def _objective_inputs(bs=2, size=24, seed=0):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 11.5) ** 2 + (yy - 12.5) ** 2).float().sqrt() < 8).float()
    a = (disc * (0.3 + 0.7 * torch.rand(bs, size, size, generator=g)))[..., None]
    a[0, 3, 3] = 5e-4                                                       # below the 1e-3 clamp: no gradient through the division
    rgba = torch.cat([torch.rand(bs, size, size, 3, generator=g) * a, a], dim=-1)
    normal = torch.nn.functional.normalize(torch.randn(bs, size, size, 3, generator=g), dim=-1) / 2 + 0.5
    normal = normal * a + torch.tensor([0.5, 0.5, 1.0]) * (1 - a)
    gate = torch.rand(bs, size, size, 1, generator=g)
    tgt = torch.rand(bs, size, size, 3, generator=g)
    m = (((xx - 12) ** 2 + (yy - 12) ** 2).float().sqrt() < 9).float()[None, :, :, None].expand(bs, -1, -1, -1).contiguous()
    return rgba, normal, gate, tgt, mopt.min_pool(m), (m * 0.96 + 0.02), torch.tensor([0.7, 1.3][:bs])


@pytest.mark.parametrize('with_patch', [False, True])
def test_fused_mesh_objective_matches_the_eager_composition(with_patch):
    """csrc/mesh_loss.cu (through the host harness) against the torch ops + autograd it replaces: value and d/d(rgba, normal)."""
    rgba0, normal0, gate, tgt, m_erode, m_blur, w_view = _objective_inputs()
    bs, size = rgba0.shape[0], rgba0.shape[1]
    nbg = [0.5, 0.5, 1.0]
    ps, patch_w, reg_w = 8, 0.7, 0.1
    pl = _FakePatchLoss()
    pick = torch.tensor([4, 11, 0])
    w_px = w_view[:, None, None, None].expand(-1, size, size, 1)
    w_pick = mopt._patches(w_px, size, ps)[pick, 0, 0, 0]
    n_px = bs * size * size
    res = {}
    with host_harness.routed(mopt):
        for mode in ('fused', 'eager'):
            rgba, normal = rgba0.clone().requires_grad_(True), normal0.clone().requires_grad_(True)
            if mode == 'fused':
                val = mopt._MeshObjectiveFn.apply(rgba, normal, gate.squeeze(-1), tgt, m_erode.squeeze(-1), m_blur.squeeze(-1), w_view, nbg,
                                                  1.2 * 4.5 / (n_px * 3), 1.2 * 2.0 / n_px, reg_w * 2 / (n_px * 3),
                                                  (pl, pick, ps, w_pick, patch_w) if with_patch else None)
            else:
                a = rgba[..., 3:]
                rgb = rgba[..., :3] / a.clamp(min=1e-3)
                rgb = rgb * m_erode + tgt * (1 - m_erode)
                n = normal * gate + normal.detach() * (1 - gate)
                nfg = (n - torch.tensor(nbg) * (1 - a)) / a.clamp(min=1e-3)
                l1 = L1LossMod(loss_weight=1.2)
                val = l1(rgb, tgt, weight=w_px) * 4.5 + l1(a, m_blur, weight=w_px) * 2.0 \
                    + mopt.tv_normal_loss(nfg.permute(0, 3, 1, 2), a.detach().permute(0, 3, 1, 2)) * (reg_w * 2)
                if with_patch:
                    val = val + mopt.lpips_patch_loss(pl, mopt._patches(rgb, size, ps)[pick], mopt._patches(tgt, size, ps)[pick], w_pick) * patch_w
            (val * 1.7).backward()
            res[mode] = (val.detach(), rgba.grad.clone(), normal.grad.clone())
    f, e = res['fused'], res['eager']
    assert abs(float(f[0]) - float(e[0])) < 1e-5 * max(1.0, abs(float(e[0])))
    for a_, b_ in zip(f[1:], e[1:]):
        assert b_.abs().max() > 0 and (a_ - b_).abs().max() <= 1e-4 * b_.abs().max() + 1e-9


def test_mesh_optim_with_the_fused_objective_matches_the_eager_one():
    n, size, steps, ps = 2, 32, 2, 16
    poses, intr = _cameras(n, size, seed=2)
    lights = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(4)), dim=-1)
    cam_weights = torch.tensor([1.0, 2.0])
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    tgt_masks = disc[None, None, :, :, None].expand(1, n, -1, -1, -1).contiguous()
    tgt_images = (torch.rand(1, n, size, size, 3, generator=torch.Generator().manual_seed(5)) * 0.5 + 0.25) * tgt_masks + (1 - tgt_masks)
    g = torch.Generator().manual_seed(6)
    noise = dict(camera_perm=torch.tensor([1, 0]), jitter=torch.rand(steps, 2, 2, generator=g),
                 patch_perm=torch.stack([torch.randperm(n * (size // ps) ** 2, generator=g) for _ in range(steps)]))
    grid = make_tet_grid(10)
    res = {}
    with host_harness.routed(mopt):
        for fused in (True, False):
            field = ToyField()
            nerf = SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss())
            tet_verts, tet_indices, tet_sdf = mopt.init_tet(nerf, None, density_thresh=5.0, tets=grid)
            deform = torch.zeros_like(tet_verts).requires_grad_(True)
            tet_sdf.requires_grad_(True)
            opt = torch.optim.Adam([{'params': list(field.parameters())}, {'params': [tet_sdf, deform], 'lr': 1e-3}], lr=0.01)
            dm = DMTet('cpu')
            with torch.enable_grad():
                mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
                mesh = Mesh(v=mv, f=mf.int())
                mesh.auto_normal()
            pipe = SimpleNamespace(nerf=nerf, mesh_renderer=MeshRenderer(near=0.01, far=100), normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
            mopt.mesh_optim(pipe, tgt_images, tgt_masks, None, opt, 0.01, 0.8, steps, 2, 3, 24, 0.7, 0.0, 0.02, 0.1, 5.0, None,
                            tet_verts, deform, tet_sdf, tet_indices, dm, mesh, size, intr, size, poses, cam_weights, lights, ps,
                            False, 0.2, 1.0, noise=noise, fused_objective=fused)
            res[fused] = (tet_sdf.detach().clone(), deform.detach().clone(), field.w.detach().clone())
    for a_, b_ in zip(res[True], res[False]):
        assert (a_ - b_).abs().max() < 2e-5


# ------------------------------------------------------------------------------------------------ decimation of the final mesh
def _manifold_stats(v, f):
    import numpy as np
    und = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    _, c = np.unique(und, axis=0, return_counts=True)
    _, cd = np.unique(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=0, return_counts=True)
    vol = np.einsum('ij,ij->i', v[f[:, 0]], np.cross(v[f[:, 1]], v[f[:, 2]])).sum() / 6
    return bool((c == 2).all()), bool((cd == 1).all()), len(v) - len(c) + len(f), float(vol)


def test_simplify_mesh_properties():
    """``mve_mesh_simplify`` (host C++ in place of open3d's quadric decimation -- parity unpinned, so properties): exact face budget, a
    closed oriented 2-manifold stays one (Euler characteristic 2), the enclosed volume and the surface are preserved."""
    import numpy as np
    from scipy.spatial import cKDTree
    from mvedit_b200.mesh_renderer import simplify_mesh
    grid = make_tet_grid(32)
    tv = -grid['vertices'] * 2 * 0.9
    sdf = 0.5 - tv.norm(dim=-1) + 0.08 * torch.sin(7 * tv[:, 0]) * torch.sin(6 * tv[:, 1])
    mv, mf = DMTet('cpu')(tv, sdf, grid['indices'])
    closed, oriented, euler, vol = _manifold_stats(mv.numpy(), mf.numpy())
    assert closed and oriented and euler == 2
    for frac in (0.5, 0.25, 0.05):
        target = round(mf.shape[0] * frac)
        v2, f2 = simplify_mesh(mv, mf, target)
        assert target - 1 <= f2.shape[0] <= target and f2.dtype == torch.int64 and v2.dtype == mv.dtype        # a collapse removes two faces
        c2, o2, e2, vol2 = _manifold_stats(v2.numpy(), f2.numpy())
        assert c2 and o2 and e2 == 2 and abs(vol2 - vol) < 0.01 * vol / frac ** 0.5, (frac, c2, o2, e2, vol2, vol)
        d, _ = cKDTree(mv.numpy()).query(v2.numpy())
        assert d.max() < 0.03 / frac ** 0.5                                    # vertices stay on the input surface (grid spacing 0.056)
        assert int(f2.max()) == v2.shape[0] - 1 and len(torch.unique(f2)) == v2.shape[0]          # compact vertex list
    v3, f3 = simplify_mesh(mv, mf, mf.shape[0] + 10)                           # nothing to do
    assert f3.shape == mf.shape and torch.equal(torch.sort(v3.norm(dim=-1)).values, torch.sort(mv.norm(dim=-1)).values)
    v4, f4 = simplify_mesh(mv, mf, 0)                                          # as far as legal collapses go: never below a tetrahedron
    assert 4 <= f4.shape[0] < 100 and _manifold_stats(v4.numpy(), f4.numpy())[:3] == (True, True, 2)


def test_mesh_optim_decimates_at_the_end_and_then_fits_only_the_texture():
    """``mesh_reduction < 1`` on the last call (mvedit_3d_pipeline.py:829-844): geometry steps, one decimation once the remaining steps fit in
    ``mesh_simplify_texture_steps``, then colour-only steps on the fixed mesh under a fresh optimiser."""
    n, size = 4, 32
    poses, intr = _cameras(n, size, seed=2)
    lights = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(4)), dim=-1)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    tgt_masks = disc[None, None, :, :, None].expand(1, n, -1, -1, -1).contiguous()
    tgt_images = (torch.rand(1, n, size, size, 3, generator=torch.Generator().manual_seed(5)) * 0.5 + 0.25) * tgt_masks + (1 - tgt_masks)
    field = ToyField()
    nerf = SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss())
    with host_harness.routed(dr):
        tet_verts, tet_indices, tet_sdf = mopt.init_tet(nerf, None, density_thresh=5.0, tets=make_tet_grid(16))
        deform = torch.zeros_like(tet_verts).requires_grad_(True)
        tet_sdf.requires_grad_(True)
        opt = torch.optim.Adam([{'params': list(field.parameters())}, {'params': [tet_sdf, deform], 'lr': 1e-3}], lr=0.01)
        dm = DMTet('cpu')
        with torch.enable_grad():
            mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
            mesh = Mesh(v=mv, f=mf.int())
            mesh.auto_normal()
        pipe = SimpleNamespace(nerf=nerf, mesh_renderer=MeshRenderer(near=0.01, far=100), normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
        snaps = []
        real_simplify = mopt.simplify_mesh

        def spy(v, f, target):
            snaps.append((tet_sdf.detach().clone(), deform.detach().clone(), field.w.detach().clone(), f.shape[0], target,
                          _manifold_stats(v.numpy(), f.numpy())[:3]))
            return real_simplify(v, f, target)
        mopt.simplify_mesh = spy
        try:
            out = mopt.mesh_optim(pipe, tgt_images, tgt_masks, None, opt, 0.01, 0.8, 2, 2, 2, 3, 0.5, 0.0, 0.02, 0.1, 5.0, None, tet_verts, deform,
                                  tet_sdf, tet_indices, dm, mesh, size, intr, size, poses, torch.ones(n), lights, 16, True, 0.2, 0.5)
        finally:
            mopt.simplify_mesh = real_simplify
    # inverse_steps = max(2, 3) = 3 and 3 - (step + 1) <= 3 at step 0: one geometry step, the decimation, two texture steps
    assert len(snaps) == 1
    sdf1, deform1, w1, n_faces, target, topo = snaps[0]
    assert target == round(n_faces * 0.5) and target - 1 <= out.f.shape[0] <= target and not out.v.requires_grad
    assert torch.equal(tet_sdf.detach(), sdf1) and torch.equal(deform.detach(), deform1)       # the geometry is frozen afterwards ...
    assert (field.w.detach() - w1).abs().max() > 1e-3                                          # ... the colour field keeps fitting
    assert _manifold_stats(out.v.numpy(), out.f.long().numpy())[:3] == topo                     # same topology as the marching-tets mesh

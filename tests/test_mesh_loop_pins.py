"""The mesh stage's Python against THE REFERENCE'S OWN CODE: tests/golden/make_mesh_loop_pins.py ran the reference's ``MeshRenderer``
(forward, get_cam_weights_uv, bake_multiview, bake_xyz_shading_fun), ``mesh_optim`` and ``texture_optim`` unmodified in the build
container, with only the four nvdiffrast ops served by oracle/raster_oracle.py, and recorded their random draws.  Here the product
(rasteriser kernels through tests/host_harness.py) and oracle/mesh_oracle.py are held to those outputs."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import mesh_oracle as mo
from oracle.nerf_oracle import L1LossMod
from tests import host_harness
from tests.test_mesh_stage_host import ToyField, _FakePatchLoss
from mvedit_b200 import mesh_raster as dr
from mvedit_b200 import mesh_optim as mopt
from mvedit_b200.mesh_renderer import DMTet, Mesh, MeshRenderer, make_tet_grid

PINS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mesh_loop_pins.npz'))
T = lambda k: torch.from_numpy(PINS[k])


@pytest.fixture(autouse=True)
def _route():
    with host_harness.routed(dr):
        yield


def _close(a, b, rtol=1e-4, atol=2e-5):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize('impl', ['product', 'oracle'])
def test_forward_with_field_shading_matches_the_reference_renderer(impl):
    size = 40
    poses, intr, lights = T('fw_poses'), T('fw_intr'), T('fw_lights')
    lp = lights[:, None, None, :].expand(-1, size, size, -1)
    field = ToyField()
    v = T('fw_v').clone().requires_grad_(True)
    fun = mopt.make_nerf_shading_fun(field, None, lp, 0.2)
    if impl == 'product':
        mesh = Mesh(v=v, f=T('fw_f'))
        mesh.auto_normal()
        r = MeshRenderer(near=0.01, far=100)([mesh], poses[None], intr[None], size, size, fun, normal_bg=[0.5, 0.5, 1.0])
    else:
        r = mo.mesh_renderer_forward(mo.make_mesh(v, T('fw_f')), poses[None], intr[None], size, size, fun)
    for k in ('rgba', 'depth', 'normal'):
        _close(r[k].detach(), T('fw_' + k))
    sum((r[k] * T('fw_w_' + k)).sum() for k in ('rgba', 'depth', 'normal')).backward()
    assert (v.grad - T('fw_g_v')).abs().max() <= 2e-3 * T('fw_g_v').abs().max()
    assert (field.w.grad - T('fw_g_w')).abs().max() <= 2e-3 * T('fw_g_w').abs().max()


def _textured_mesh():
    mesh = Mesh(v=T('fw_v'), f=T('fw_f'))
    mesh.auto_normal()
    mesh.vt, mesh.ft, mesh.albedo = T('tx_vt'), T('tx_ft'), T('tx_albedo').clone()
    return mesh


def test_textured_render_and_bakers_match_the_reference_renderer():
    size = 40
    poses, intr = T('fw_poses'), T('fw_intr')
    r = MeshRenderer(near=0.01, far=100)
    mesh = _textured_mesh()
    assert torch.equal(T('tx_vt'), (lambda m: (m.auto_uv(), m.vt)[1])(Mesh(v=T('fw_v'), f=T('fw_f'))))      # the atlas the fixture was made with
    with torch.no_grad():
        _close(r([mesh], poses[None], intr[None], size, size)['rgba'], T('tx_rgba'))
        wts, valid = r.get_cam_weights_uv([mesh], poses[None], intr[None], alphas=T('bk_alphas')[0], render_size=size, map_size=64, render_bs=2,
                                          cos_weight_pow=1.0)
        assert torch.equal(valid, T('bk_valid'))
        _close(wts, T('bk_weights'), rtol=1e-3, atol=1e-4)
        baked = r.bake_multiview([mesh], T('bk_images'), T('bk_alphas'), poses[None], intr[None], map_size=64, cos_weight_pow=8.0, base_weight=0.3,
                                 render_bs=2)[0]
        d = (baked.albedo - T('bk_multiview')).abs()
        assert d.mean() < 1e-4 and (d > 1e-2).float().mean() < 2e-3
        field = ToyField()
        xyz = r.bake_xyz_shading_fun([_textured_mesh()], mopt.make_nerf_albedo_shading_fun(field, None), map_size=64)[0]
        _close(xyz.albedo, T('bk_xyz'))


def _mesh_optim_setup():
    grid = make_tet_grid(12)
    tet_verts = (-grid['vertices'] * 2 * 0.9).contiguous()
    sdf = T('mo_sdf0').clone().requires_grad_(True)
    deform = torch.zeros_like(tet_verts).requires_grad_(True)
    field = ToyField()
    opt = torch.optim.Adam([{'params': list(field.parameters())}, {'params': [sdf, deform], 'lr': 1e-3}], lr=0.01)
    noise = dict(camera_perm=T('mo_camera_perm'), jitter=T('mo_jitter'), patch_perm=T('mo_patch_perm'))
    return grid, tet_verts, sdf, deform, field, opt, noise


@pytest.mark.parametrize('impl', ['product', 'product-fused', 'oracle'])
def test_mesh_optim_matches_the_reference_method(impl):
    """Two iterations of the reference's own ``mesh_optim`` (incl. a patch term): same SDF, deformation, field and mesh afterwards."""
    grid, tet_verts, sdf, deform, field, opt, noise = _mesh_optim_setup()
    size, steps, ps = 32, 2, 16
    args = dict(tgt_images=T('mo_tgt_images'), tgt_masks=T('mo_tgt_masks'), intr=T('mo_intr'), poses=T('mo_poses'), cw=T('mo_cam_weights'),
                lights=T('mo_lights'))
    pl = _FakePatchLoss()
    if impl.startswith('product'):
        dm = DMTet('cpu')
        with torch.enable_grad():
            mv, mf = dm(tet_verts + deform, sdf, grid['indices'])
            mesh = Mesh(v=mv, f=mf.int())
            mesh.auto_normal()
        pipe = SimpleNamespace(nerf=SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=pl),
                               mesh_renderer=MeshRenderer(near=0.01, far=100), normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
        with host_harness.routed(mopt):
            out = mopt.mesh_optim(pipe, args['tgt_images'], args['tgt_masks'], None, opt, 0.01, 0.8, steps, 2, 3, 24, 0.7, 0.0, 0.02, 0.1, 5.0, None,
                                  tet_verts, deform, sdf, grid['indices'], dm, mesh, size, args['intr'], size, args['poses'], args['cw'],
                                  args['lights'], ps, False, 0.2, 1.0, noise=noise, fused_objective=impl.endswith('fused'))
    else:
        dm = mo.DMTetOracle()
        mv, mf = dm(tet_verts + deform, sdf, grid['indices'])
        out, _ = mo.mesh_optim(field, args['tgt_images'], args['tgt_masks'], opt, 0.01, 0.8, steps, 2, 3, 0.7, 0.02, 0.1, 5.0, None, tet_verts, deform,
                               sdf, grid['indices'], dm, mo.make_mesh(mv, mf.int()), size, args['intr'], size, args['poses'], args['cw'],
                               args['lights'], ps, 0.2, noise, patch_loss=pl)
    for got, key in ((sdf, 'mo_sdf'), (deform, 'mo_deform'), (field.w, 'mo_w'), (field.b, 'mo_b')):
        assert (got.detach() - T(key)).abs().max() < 2e-5, key
    assert torch.equal(out.f.long(), T('mo_faces').long())
    _close(out.v.detach(), T('mo_verts'))
    assert (T('mo_sdf') - T('mo_sdf0')).abs().max() > 1e-4


@pytest.mark.parametrize('impl', ['product', 'oracle'])
def test_mesh_optim_with_target_normals_matches_the_reference_method(impl):
    """The reference's ``mesh_optim`` with ``tgt_normals`` (TV term against the target's differences, geometry lr without the multiplier,
    the high-passed normal patch term with its own patch draw): same SDF, deformation, field and mesh after two iterations."""
    grid, tet_verts, sdf, deform, field, opt, _ = _mesh_optim_setup()
    noise = dict(camera_perm=T('mn_camera_perm'), jitter=T('mn_jitter'), patch_perm=T('mn_patch_perm'), patch_perm_normal=T('mn_patch_perm_normal'))
    size, steps, ps = 32, 2, 16
    pl = _FakePatchLoss()
    if impl == 'product':
        dm = DMTet('cpu')
        with torch.enable_grad():
            mv, mf = dm(tet_verts + deform, sdf, grid['indices'])
            mesh = Mesh(v=mv, f=mf.int())
            mesh.auto_normal()
        pipe = SimpleNamespace(nerf=SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=pl),
                               mesh_renderer=MeshRenderer(near=0.01, far=100), normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
        with host_harness.routed(mopt):
            out = mopt.mesh_optim(pipe, T('mo_tgt_images'), T('mo_tgt_masks'), T('mn_tgt_normals'), opt, 0.01, 0.8, steps, 2, 3, 24, 0.7, 0.9, 0.02,
                                  0.1, 5.0, None, tet_verts, deform, sdf, grid['indices'], dm, mesh, size, T('mo_intr'), size, T('mo_poses'),
                                  T('mo_cam_weights'), T('mo_lights'), ps, False, 0.2, 1.0, noise=noise)
            with pytest.raises(NotImplementedError):
                mopt.mesh_optim(pipe, T('mo_tgt_images'), T('mo_tgt_masks'), T('mn_tgt_normals'), opt, 0.01, 0.8, steps, 2, 3, 24, 0.7, 0.9, 0.02,
                                0.1, 5.0, None, tet_verts, deform, sdf, grid['indices'], dm, mesh, size, T('mo_intr'), size, T('mo_poses'),
                                T('mo_cam_weights'), T('mo_lights'), ps, False, 0.2, 1.0, noise=noise, fused_objective=True)
    else:
        dm = mo.DMTetOracle()
        mv, mf = dm(tet_verts + deform, sdf, grid['indices'])
        out, _ = mo.mesh_optim(field, T('mo_tgt_images'), T('mo_tgt_masks'), opt, 0.01, 0.8, steps, 2, 3, 0.7, 0.02, 0.1, 5.0, None, tet_verts, deform,
                               sdf, grid['indices'], dm, mo.make_mesh(mv, mf.int()), size, T('mo_intr'), size, T('mo_poses'), T('mo_cam_weights'),
                               T('mo_lights'), ps, 0.2, noise, patch_loss=pl, tgt_normals=T('mn_tgt_normals'), patch_normal_weight=0.9)
    for got, key in ((sdf, 'mn_sdf'), (deform, 'mn_deform'), (field.w, 'mn_w'), (field.b, 'mn_b')):
        assert (got.detach() - T(key)).abs().max() < 2e-5, key
    assert torch.equal(out.f.long(), T('mn_faces').long())
    _close(out.v.detach(), T('mn_verts'))
    assert (T('mn_sdf') - T('mo_sdf')).abs().max() > 1e-5                      # the target normals did change the trajectory


def test_texture_optim_matches_the_reference_method():
    mesh = Mesh(v=T('fw_v'), f=T('fw_f'))
    mesh.auto_normal()
    field = ToyField()
    opt = torch.optim.Adam(field.parameters(), lr=0.01)
    pipe = SimpleNamespace(nerf=SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss()),
                           mesh_renderer=MeshRenderer(near=0.01, far=100), bg_color=0.5)
    noise = dict(camera_perm=T('to_camera_perm'), jitter=T('to_jitter'), patch_perm=T('to_patch_perm'))
    mopt.texture_optim(pipe, T('to_tgt'), opt, 0.02, 3, 2, 2, 0.6, None, mesh, 32, T('mo_intr'), 32, T('mo_poses'), T('to_w'), 16, noise=noise)
    assert (field.w.detach() - T('to_w_out')).abs().max() < 2e-5 and (field.b.detach() - T('to_b_out')).abs().max() < 2e-5
    assert (field.w.detach() - ToyField().w.detach()).abs().max() > 1e-2


def test_superres_texture_optim_matches_the_reference_method():
    """The super-resolution pipeline's own ``texture_optim`` (mvedit_texture_superres_pipeline.py:89-168, ``num_cameras`` = 1 of the two views
    per batch): the product's ``texture_optim(patch_views=)``."""
    mesh = Mesh(v=T('fw_v'), f=T('fw_f'))
    mesh.auto_normal()
    field = ToyField()
    opt = torch.optim.Adam(field.parameters(), lr=0.01)
    pipe = SimpleNamespace(nerf=SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss()),
                           mesh_renderer=MeshRenderer(near=0.01, far=100), bg_color=0.5)
    noise = dict(camera_perm=T('ts_camera_perm'), jitter=T('ts_jitter'), patch_perm=T('ts_patch_perm'))
    mopt.texture_optim(pipe, T('to_tgt'), opt, 0.02, 3, 2, 2, 0.6, None, mesh, 32, T('mo_intr'), 32, T('mo_poses'), T('to_w'), 16, noise=noise,
                       patch_views=1)
    assert (field.w.detach() - T('ts_w_out')).abs().max() < 2e-5 and (field.b.detach() - T('ts_b_out')).abs().max() < 2e-5
    assert (T('ts_w_out') - T('to_w_out')).abs().max() > 1e-4                  # not the plain variant's trajectory


@pytest.mark.parametrize('impl', ['product', 'oracle'])
def test_renderer_options_match_the_reference_renderer(impl):
    """2x supersampling (what ``load_init_mesh`` renders with), vertex colours with edge dilation, and antialiasing switched off."""
    size = 40
    poses, intr, lights = T('fw_poses'), T('fw_intr'), T('fw_lights')
    lp2 = lights[:, None, None, :].expand(-1, 2 * size, 2 * size, -1)
    field = ToyField()
    fun = mopt.make_nerf_shading_fun(field, None, lp2, 0.2)
    with torch.no_grad():
        if impl == 'product':
            mesh = Mesh(v=T('fw_v'), f=T('fw_f'), vc=T('op_vc'))
            mesh.auto_normal()
            rend = MeshRenderer(near=0.01, far=100, ssaa=2)
            o1 = rend([mesh], poses[None], intr[None], size, size, None, dilate_edges=2, normal_bg=[0.5, 0.5, 1.0], render_vc=True)
            o2 = rend([mesh], poses[None], intr[None], size, size, fun, normal_bg=[0.5, 0.5, 1.0], aa=False)
        else:
            m = mo.make_mesh(T('fw_v'), T('fw_f'))
            m.vc = T('op_vc')
            o1 = mo.mesh_renderer_forward(m, poses[None], intr[None], size, size, None, dilate_edges=2, ssaa=2)
            o2 = mo.mesh_renderer_forward(m, poses[None], intr[None], size, size, fun, aa=False, ssaa=2)
    for k in ('rgba', 'depth', 'normal'):
        _close(o1[k], T('op1_' + k))
        _close(o2[k], T('op2_' + k))

"""Generates tests/golden/mesh_loop_pins.npz by RUNNING the reference's own mesh-stage Python -- ``MeshRenderer`` (forward, bake_multiview,
get_cam_weights_uv, bake_xyz_shading_fun: base_mesh_renderer.py:191-603), ``MVEdit3DPipeline.mesh_optim`` / ``make_nerf_shading_fun``
(mvedit_3d_pipeline.py:425-442,658-872) and ``MVEditTexturePipeline.texture_optim`` (mvedit_texture_pipeline.py:93-172) -- cut out by
AST and executed unmodified in the build container, with exactly ONE substitution: the four ``nvdiffrast.torch`` ops (an absent CUDA
dependency) are served by oracle/raster_oracle.py.  Everything around them -- projection, compositing, the objective, the optimiser
steps, DMTet re-extraction, the bakers' weighting -- is the reference's code.  The fixture therefore pins the product's (and
oracle/mesh_oracle.py's) restatements of those loops against the reference itself, modulo the rasteriser ops.

Other stand-ins, each pinned elsewhere: ``TVLoss`` = the reference's ``tv_loss`` body (tests/golden/make_reference_pins.py pins it) under
mmgen's weighted-mean reduction; ``L1LossMod`` from oracle/nerf_oracle.py; ``Mesh`` = a namespace with the reference's ``auto_normal``;
the field is the analytic ``ToyField`` of tests/test_mesh_stage_host.py (tcnn is absent).  ``torch.randperm`` / ``torch.rand_like`` are
recorded while the reference runs so that the tests can hand the same draws to the product.

Run:  python tests/golden/make_mesh_loop_pins.py      (CPU, ~1 min)
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mesh_loop_pins.npz')


def extract(rel, names, env):
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names and node.name not in found:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, rel, 'exec'), env)
            found[node.name] = env[node.name]
    assert not (set(names) - set(found)), set(names) - set(found)
    return found


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class TorchRecorder:
    """``torch`` for the extracted loops: every attribute is torch's, but randperm / rand_like results are kept in order."""

    def __init__(self):
        self.draws = []

    def __getattr__(self, k):
        return getattr(torch, k)

    def randperm(self, n, **kw):
        kw.pop('device', None)
        r = torch.randperm(n, **kw)
        self.draws.append(('randperm', r.clone()))
        return r

    def rand_like(self, t, **kw):
        r = torch.rand_like(t, **kw)
        self.draws.append(('rand_like', r.clone()))
        return r


def main():
    from oracle import raster_oracle as ro, mesh_oracle as mo
    from oracle.nerf_oracle import L1LossMod
    from tests import synth_mesh
    from tests.test_mesh_stage_host import ToyField, _FakePatchLoss
    from mvedit_b200.mesh_renderer import Mesh as ProductMesh, make_tet_grid
    torch.Tensor.cuda = lambda self, *a, **k: self
    for stub in ('mcubes', 'skimage'):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    sys.modules['skimage'].morphology = types.ModuleType('morphology')
    gu = load_by_path('ref_geometry_utils', 'lib/core/utils/geometry_utils.py')
    ed = load_by_path('ref_edge_dilation', 'lib/ops/edge_dilation.py')

    class dr:                                    # the one substitution: nvdiffrast.torch -> oracle/raster_oracle.py
        class RasterizeCudaContext:
            pass
        RasterizeGLContext = RasterizeCudaContext

        @staticmethod
        def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
            assert ranges is None
            return mo.dr_rasterize(pos, tri, tuple(resolution))

        @staticmethod
        def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
            out, da = ro.interpolate(attr, rast, tri, rast_db=rast_db, diff_attrs=diff_attrs)
            return out, (da if da is not None else out.new_zeros(out.shape[:-1] + (0,)))

        @staticmethod
        def texture(tex, uv, uv_da=None, filter_mode='auto', **kw):
            return ro.texture(tex, uv, uv_da, filter_mode='linear-mipmap-linear' if uv_da is not None else 'linear')

        @staticmethod
        def antialias(color, rast, pos, tri):
            return ro.antialias(color, rast, pos, np.asarray(tri))

    import math
    env = dict(torch=torch, nn=nn, F=F, math=math, dr=dr, get_ray_directions=gu.get_ray_directions, depth_to_normal=gu.depth_to_normal,
               edge_dilation=ed.edge_dilation)
    R = extract('lib/models/decoders/mesh_renderer/base_mesh_renderer.py',
                ['MeshRenderer', 'DMTet', 'interpolate_hwc', 'compute_edge_to_face_mapping', 'normal_consistency', 'laplacian_uniform',
                 'laplacian_smooth_loss'], env)
    ref_auto_normal = extract('lib/models/decoders/mesh_renderer/mesh_utils.py', ['auto_normal'], dict(torch=torch, F=F))['auto_normal']

    class RefMesh:                               # Mesh(v=, f=, device=) + the reference's auto_normal (mesh_utils.py:39-66,359-382)
        def __init__(self, v=None, f=None, device=None, **kw):
            self.v, self.f, self.device = v, f, device
            self.vn = self.fn = self.vt = self.ft = self.vc = self.albedo = self.face_normals = None
            self.textureless = False
        auto_normal = ref_auto_normal

        def detach(self):
            self.v = self.v.detach()
            return self

    out = {}
    renderer = R['MeshRenderer'](near=0.01, far=100, ssaa=1)

    # ---- MeshRenderer.forward: vertex colours, field shading (with gradients), textured; the bakers
    v, f = synth_mesh.icosphere(2)
    n, size = 3, 40
    poses = torch.from_numpy(synth_mesh.surround_poses(n, 0)).float()
    intr = torch.from_numpy(synth_mesh.intrinsics(size)).float()[None].expand(n, -1).contiguous()
    field = ToyField()
    lights = F.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(1)), dim=-1)
    lp = lights[:, None, None, :].expand(-1, size, size, -1)

    def shading_fun(world_pos=None, albedo=None, world_normal=None, fg_mask=None, **kw):
        base = field.point_decode([world_pos], None, None)[1]
        sh = ((lp[fg_mask.squeeze(0)][:, None, :] @ world_normal[:, :, None]).clamp(min=0) * 0.8 + 0.2).squeeze(-1)
        return base * sh

    vt_v = (torch.from_numpy(v).float() * 0.55).requires_grad_(True)
    m = RefMesh(v=vt_v, f=torch.from_numpy(f).int())
    m.auto_normal()
    with torch.enable_grad():
        r = renderer([m], poses[None], intr[None], size, size, shading_fun, normal_bg=[0.5, 0.5, 1.0])
        gen = torch.Generator().manual_seed(2)
        ws = {k: torch.randn(r[k].shape, generator=gen) for k in ('rgba', 'depth', 'normal')}
        sum((r[k] * ws[k]).sum() for k in ws).backward()
    out.update(fw_v=vt_v.detach().numpy(), fw_f=f.astype(np.int32), fw_poses=poses.numpy(), fw_intr=intr.numpy(), fw_lights=lights.numpy(),
               fw_rgba=r['rgba'].detach().numpy(), fw_depth=r['depth'].detach().numpy(), fw_normal=r['normal'].detach().numpy(),
               fw_g_v=vt_v.grad.numpy(), fw_g_w=field.w.grad.numpy(), **{'fw_w_' + k: t.numpy() for k, t in ws.items()})

    # ---- the renderer's options: 2x supersampling (load_init_mesh renders with it), vertex colours, edge dilation, antialiasing off
    vc = torch.cat([torch.rand(1, len(v), 3, generator=torch.Generator().manual_seed(3)), torch.ones(1, len(v), 1)], dim=-1)
    mo_ = RefMesh(v=vt_v.detach(), f=torch.from_numpy(f).int())
    mo_.vc = vc
    mo_.auto_normal()
    r2 = R['MeshRenderer'](near=0.01, far=100, ssaa=2)
    with torch.no_grad():
        o1 = r2([mo_], poses[None], intr[None], size, size, None, dilate_edges=2, normal_bg=[0.5, 0.5, 1.0], render_vc=True)
        lp2 = lights[:, None, None, :].expand(-1, 2 * size, 2 * size, -1)      # per-pixel lights at the supersampled size, as load_init_mesh's callers pass them

        def shading_fun2(world_pos=None, albedo=None, world_normal=None, fg_mask=None, **kw):
            base = field.point_decode([world_pos], None, None)[1]
            return base * ((lp2[fg_mask.squeeze(0)][:, None, :] @ world_normal[:, :, None]).clamp(min=0) * 0.8 + 0.2).squeeze(-1)
        o2 = r2([mo_], poses[None], intr[None], size, size, shading_fun2, normal_bg=[0.5, 0.5, 1.0], aa=False)
    out.update(op_vc=vc.numpy(), **{'op1_' + k: o1[k].detach().numpy() for k in ('rgba', 'depth', 'normal')},
               **{'op2_' + k: o2[k].detach().numpy() for k in ('rgba', 'depth', 'normal')})

    pm = ProductMesh(v=vt_v.detach(), f=torch.from_numpy(f).int())           # the per-triangle atlas is an INPUT here (xatlas is absent)
    pm.auto_uv()
    tm = RefMesh(v=vt_v.detach(), f=torch.from_numpy(f).int())
    tm.auto_normal()
    tm.vt, tm.ft = pm.vt, pm.ft
    tm.albedo = torch.rand(32, 32, 4, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        rt = renderer([tm], poses[None], intr[None], size, size)
        gimg = torch.Generator().manual_seed(9)
        images = torch.rand(1, n, size, size, 3, generator=gimg)
        alphas = (torch.rand(1, n, size, size, 1, generator=gimg) > 0.1).float()
        wts, valid = renderer.get_cam_weights_uv([tm], poses[None], intr[None], alphas=alphas[0], render_size=size, map_size=64, render_bs=2,
                                                 cos_weight_pow=1.0)
        albedo_in = tm.albedo.clone()
        baked = renderer.bake_multiview([tm], images, alphas, poses[None], intr[None], map_size=64, cos_weight_pow=8.0, base_weight=0.3, render_bs=2)[0]
        bake_mv = baked.albedo.clone()
        field2 = ToyField()
        baked2 = renderer.bake_xyz_shading_fun([tm], lambda world_pos=None, **kw: field2.point_decode([world_pos], None, None)[1], map_size=64)[0]
    out.update(tx_vt=pm.vt.numpy(), tx_ft=pm.ft.numpy(), tx_albedo=albedo_in.numpy(), tx_rgba=rt['rgba'].numpy(), bk_images=images.numpy(),
               bk_alphas=alphas.numpy(), bk_weights=wts.numpy(), bk_valid=valid.numpy(), bk_multiview=bake_mv.numpy(), bk_xyz=baked2.albedo.numpy())

    # ---- mesh_optim (mvedit_3d_pipeline.py:658-872) and texture_optim (mvedit_texture_pipeline.py:93-172)
    tv_body = extract('lib/models/losses/tv_loss.py', ['tv_loss'], dict(torch=torch))['tv_loss']

    class TVLoss(nn.Module):                     # tv_loss.py:45-61 under mmgen's weighted_loss (mean reduction, no element weight here)
        def __init__(self, dims=[-2, -1], power=1, loss_weight=1.0):
            super().__init__()
            self.dims, self.power, self.loss_weight = dims, power, loss_weight

        def forward(self, pred, target=None, weight=None, avg_factor=None):
            return tv_body(pred, target, self.dims, power=self.power, dense_weight=weight).mean() * self.loss_weight

    import torchvision.transforms.v2.functional as F_t
    rec = TorchRecorder()
    penv = dict(torch=rec, F=F, F_t=F_t, np=np, TVLoss=TVLoss, get_module_device=lambda mod: 'cpu', get_ray_directions=gu.get_ray_directions,
                depth_to_normal=gu.depth_to_normal, laplacian_smooth_loss=R['laplacian_smooth_loss'], normal_consistency=R['normal_consistency'],
                Mesh=RefMesh, o3d=None)
    penv['highpass'] = extract('lib/pipelines/utils.py', ['highpass'], dict(F_t=F_t))['highpass']
    P = extract('lib/pipelines/mvedit_3d_pipeline.py', ['mesh_optim', 'make_nerf_shading_fun'], penv)
    n, size, steps, ps = 4, 32, 2, 16
    poses = torch.from_numpy(synth_mesh.surround_poses(n, 2)).float()
    intr = torch.from_numpy(synth_mesh.intrinsics(size)).float()[None].expand(n, -1).contiguous()
    lights = F.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(4)), dim=-1)
    cam_weights = torch.tensor([1.0, 0.5, 1.0, 2.0])
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    tgt_masks = disc[None, None, :, :, None].expand(1, n, -1, -1, -1).contiguous()
    tgt_images = (torch.rand(1, n, size, size, 3, generator=torch.Generator().manual_seed(5)) * 0.5 + 0.25) * tgt_masks + (1 - tgt_masks)
    grid = make_tet_grid(12)
    field = ToyField()
    tet_verts = (-grid['vertices'] * 2 * 0.9).contiguous()
    tet_indices = grid['indices']
    tet_sdf = (0.45 - tet_verts.norm(dim=-1) + 0.05 * torch.sin(6 * tet_verts[:, 0]) * torch.sin(5 * tet_verts[:, 1])).clamp(-1, 1)
    # keep the field away from zero AT the grid vertices: a crossing edge with |sdf| ~ 1e-8 at one end has d vertex / d sdf ~ 1e8, and the
    # comparison would measure fp32 cancellation noise (seen: gradient entries of 4e10 next to entries of 1) instead of the loop's logic
    tet_sdf = torch.where(tet_sdf.abs() < 0.02, torch.where(tet_sdf >= 0, 0.02, -0.02), tet_sdf).requires_grad_(True)
    sdf0 = tet_sdf.detach().clone()
    deform = torch.zeros_like(tet_verts).requires_grad_(True)
    opt = torch.optim.Adam([{'params': list(field.parameters())}, {'params': [tet_sdf, deform], 'lr': 1e-3}], lr=0.01)
    dm = R['DMTet']('cpu')
    with torch.enable_grad():
        mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
        in_mesh = RefMesh(v=mv, f=mf.int())
        in_mesh.auto_normal()
    self_ = types.SimpleNamespace(nerf=types.SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss()),
                                  mesh_renderer=renderer, normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
    self_.make_nerf_shading_fun = lambda *a: P['make_nerf_shading_fun'](self_, *a)
    torch.manual_seed(11)
    out_mesh = P['mesh_optim'](self_, tgt_images, tgt_masks, None, opt, 0.01, 0.8, steps, 2, 3, 24, 0.7, 0.0, 0.02, 0.1, 5.0, [None], tet_verts,
                               deform, tet_sdf, tet_indices, dm, in_mesh, size, intr, size, poses, cam_weights, lights, ps, False, 0.2, 1.0)
    draws = rec.draws
    kinds = [k for k, _ in draws]
    assert kinds == ['randperm'] + ['rand_like', 'randperm'] * steps, kinds
    out.update(mo_poses=poses.numpy(), mo_intr=intr.numpy(), mo_lights=lights.numpy(), mo_cam_weights=cam_weights.numpy(), mo_tgt_images=tgt_images.numpy(),
               mo_tgt_masks=tgt_masks.numpy(), mo_sdf0=sdf0.numpy(), mo_camera_perm=draws[0][1].numpy(),
               mo_jitter=np.stack([draws[1 + 2 * s][1].numpy() for s in range(steps)]),
               mo_patch_perm=np.stack([draws[2 + 2 * s][1].numpy() for s in range(steps)]),
               mo_sdf=tet_sdf.detach().numpy(), mo_deform=deform.detach().numpy(), mo_w=field.w.detach().numpy(), mo_b=field.b.detach().numpy(),
               mo_faces=out_mesh.f.numpy(), mo_verts=out_mesh.v.detach().numpy())

    # ---- the same loop with target normals: TV target, lr of the geometry group without the multiplier, high-passed normal patch term
    rec.draws = []
    field_n = ToyField()
    sdf_n = sdf0.clone().requires_grad_(True)
    deform_n = torch.zeros_like(tet_verts).requires_grad_(True)
    opt_n = torch.optim.Adam([{'params': list(field_n.parameters())}, {'params': [sdf_n, deform_n], 'lr': 1e-3}], lr=0.01)
    nx, ny = (xx - 15.5) / 12.0, -(yy - 15.5) / 12.0
    nrm = F.normalize(torch.stack([nx, ny, (1 - nx ** 2 - ny ** 2).clamp(min=0.05).sqrt()], -1)
                      + 0.05 * torch.randn(size, size, 3, generator=torch.Generator().manual_seed(6)), dim=-1)
    tgt_normals = (nrm / 2 + 0.5)[None, None].expand(1, n, -1, -1, -1).contiguous()
    with torch.enable_grad():
        mv, mf = dm(tet_verts + deform_n, sdf_n, tet_indices)
        mesh_n = RefMesh(v=mv, f=mf.int())
        mesh_n.auto_normal()
    self_n = types.SimpleNamespace(nerf=types.SimpleNamespace(decoder=field_n, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss()),
                                   mesh_renderer=renderer, normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
    self_n.make_nerf_shading_fun = lambda *a: P['make_nerf_shading_fun'](self_n, *a)
    torch.manual_seed(12)
    out_n = P['mesh_optim'](self_n, tgt_images, tgt_masks, tgt_normals, opt_n, 0.01, 0.8, steps, 2, 3, 24, 0.7, 0.9, 0.02, 0.1, 5.0, [None], tet_verts,
                            deform_n, sdf_n, tet_indices, dm, mesh_n, size, intr, size, poses, cam_weights, lights, ps, False, 0.2, 1.0)
    dn = rec.draws
    assert [k for k, _ in dn] == ['randperm'] + ['rand_like', 'randperm', 'randperm'] * steps
    out.update(mn_tgt_normals=tgt_normals.numpy(), mn_camera_perm=dn[0][1].numpy(),
               mn_jitter=np.stack([dn[1 + 3 * s][1].numpy() for s in range(steps)]),
               mn_patch_perm=np.stack([dn[2 + 3 * s][1].numpy() for s in range(steps)]),
               mn_patch_perm_normal=np.stack([dn[3 + 3 * s][1].numpy() for s in range(steps)]),
               mn_sdf=sdf_n.detach().numpy(), mn_deform=deform_n.detach().numpy(), mn_w=field_n.w.detach().numpy(), mn_b=field_n.b.detach().numpy(),
               mn_faces=out_n.f.numpy(), mn_verts=out_n.v.detach().numpy())

    rec2 = TorchRecorder()
    tenv = dict(torch=rec2, F=F, np=np, get_module_device=lambda mod: 'cpu')
    T = extract('lib/pipelines/mvedit_texture_pipeline.py', ['texture_optim'], tenv)
    A = extract('lib/pipelines/mvedit_3d_pipeline.py', ['make_nerf_albedo_shading_fun'], dict(torch=torch))
    field3 = ToyField()
    opt3 = torch.optim.Adam(field3.parameters(), lr=0.01)
    fixed = RefMesh(v=vt_v.detach(), f=torch.from_numpy(f).int())
    fixed.auto_normal()
    self3 = types.SimpleNamespace(nerf=types.SimpleNamespace(decoder=field3, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss()),
                                  mesh_renderer=renderer, bg_color=0.5)
    self3.make_nerf_albedo_shading_fun = lambda *a: A['make_nerf_albedo_shading_fun'](self3, *a)
    gt = torch.Generator().manual_seed(12)
    t_tgt = torch.rand(1, n, size, size, 3, generator=gt)
    t_w = torch.rand(n, size, size, 1, generator=gt)
    torch.manual_seed(13)
    T['texture_optim'](self3, t_tgt, opt3, 0.02, 3, 2, 2, 0.6, [None], fixed, size, intr, size, poses, t_w, ps)
    d2 = rec2.draws
    assert [k for k, _ in d2] == ['randperm'] + ['rand_like', 'randperm'] * 3
    out.update(to_tgt=t_tgt.numpy(), to_w=t_w.numpy(), to_camera_perm=d2[0][1].numpy(), to_jitter=np.stack([d2[1 + 2 * s][1].numpy() for s in range(3)]),
               to_patch_perm=np.stack([d2[2 + 2 * s][1].numpy() for s in range(3)]), to_w_out=field3.w.detach().numpy(), to_b_out=field3.b.detach().numpy())
    # ---- the super-resolution pipeline's own texture_optim (mvedit_texture_superres_pipeline.py:89-168): patch term on the first num_cameras
    # views of every rendered batch only ("ignore regularization views")
    rec3 = TorchRecorder()
    T2 = extract('lib/pipelines/mvedit_texture_superres_pipeline.py', ['texture_optim'], dict(torch=rec3, F=F, np=np, get_module_device=lambda mod: 'cpu'))
    field4 = ToyField()
    opt4 = torch.optim.Adam(field4.parameters(), lr=0.01)
    self4 = types.SimpleNamespace(nerf=types.SimpleNamespace(decoder=field4, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=_FakePatchLoss()),
                                  mesh_renderer=renderer, bg_color=0.5)
    self4.make_nerf_albedo_shading_fun = lambda *a: A['make_nerf_albedo_shading_fun'](self4, *a)
    torch.manual_seed(14)
    T2['texture_optim'](self4, t_tgt, 1, opt4, 0.02, 3, 2, 2, 0.6, [None], fixed, size, intr, size, poses, t_w, ps)
    d3 = rec3.draws
    assert [k for k, _ in d3] == ['randperm'] + ['rand_like', 'randperm'] * 3
    out.update(ts_camera_perm=d3[0][1].numpy(), ts_jitter=np.stack([d3[1 + 2 * s][1].numpy() for s in range(3)]),
               ts_patch_perm=np.stack([d3[2 + 2 * s][1].numpy() for s in range(3)]), ts_w_out=field4.w.detach().numpy(), ts_b_out=field4.b.detach().numpy())
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, sum(v.nbytes for v in out.values()) // 1024, 'KiB')


if __name__ == '__main__':
    main()

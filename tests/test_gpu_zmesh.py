"""Mesh stage on the GPU, through the C ABI: the rasteriser kernels against oracle/raster_oracle.py (ids exact,
barycentrics to fp32 rounding, gradients against the oracle's autograd), ``MeshRenderer.forward`` and ``mesh_optim`` with the real
hash-grid field against oracle/mesh_oracle.py driving the plain-torch field (oracle/field_oracle.py) on the CPU.

(The file sorts after the other GPU tests on purpose: these kernels were written in a round whose GPU budget was already spent, so
their first run on a B200 is the driver's; the same arithmetic is checked bit for bit on the CPU by tests/test_mesh_raster_host.py.)

Tolerances: triangle ids exact; (u, v, z/w), rast_db rtol 1e-5 / atol 1e-6; interpolate / antialias outputs rtol 1e-4 / atol 2e-5;
gradients 2e-3 of the largest entry (fp32 atomics in a different order); renderer outputs with the fp32 field rtol 1e-3 / atol 2e-3;
mesh_optim after 2 Adam steps: |d sdf|, |d deform| <= 1e-4 on all but 0.2 % of the entries (Adam's first steps are sign-like, so an
entry whose gradient is pure rounding noise may move by a full step in either direction).
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests import synth_mesh

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = 'cuda'


def _scene(kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == 'sphere':
        v, f = synth_mesh.icosphere(2)
        pos = synth_mesh.project(v * 0.6, synth_mesh.surround_poses(3, seed), fov_deg=30.0)
        return pos.astype(np.float32), f.astype(np.int32), (48, 40)
    if kind == 'soup':
        n = 300
        c = rng.uniform(-1.2, 1.2, (n, 1, 2))
        xy = c + rng.normal(0, 0.08, (n, 3, 2))
        z = rng.uniform(-1.3, 1.3, (n, 1, 1)) + rng.normal(0, 0.05, (n, 3, 1))
        w = rng.uniform(0.5, 2.0, (n, 3, 1))
        w[:5] = -w[:5]
        w[5:8, 0] = 0.0
        v = np.concatenate([xy * w, z * w, w], axis=-1).reshape(1, n * 3, 4)
        f = np.arange(n * 3).reshape(n, 3)
        f[10] = f[10][[0, 0, 1]]
        return v.astype(np.float32), f.astype(np.int32), (37, 53)
    v = np.array([[[-1.5, -1.5, 0.2, 1], [1.5, -1.5, 0.2, 1], [1.5, 1.5, 0.2, 1], [-1.5, 1.5, 0.2, 1], [-0.9, -0.8, -0.5, 1], [0.9, -0.7, 0.9, 1],
                   [0.0, 0.95, 0.1, 1], [-3, -3, 0.5, 2], [3, -3, 0.5, 2], [0, 3, -0.5, 2]]], np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [7, 8, 9], [0, 1, 2]], np.int32)
    return v, f, (64, 64)


def _rand(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).normal(size=shape).astype(np.float32))


@pytest.mark.parametrize('kind', ['sphere', 'soup', 'large'])
def test_rasterize_matches_oracle(kind):
    from oracle import raster_oracle as ro
    from mvedit_b200 import mesh_raster as dr
    pos, tri, res = _scene(kind)
    r_o, db_o = ro.rasterize(pos, tri, res)
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), torch.from_numpy(pos).cuda(), torch.from_numpy(tri).cuda(), res)
    rast, db = rast.cpu().numpy(), db.cpu().numpy()
    assert (rast[..., 3] == r_o[..., 3]).all() and (rast[..., 3] > 0).sum() > 50
    np.testing.assert_allclose(rast[..., :3], r_o[..., :3], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(db, db_o, rtol=1e-4, atol=1e-6)


def test_rasterize_is_deterministic_and_scales():
    """512^2 x 4 views of a 20 k-triangle sphere (the DMTet regime: triangles a few pixels wide): two runs agree bit for bit, the closed
    surface has no holes, and the large-triangle queue path (a screen-filling quad behind the sphere) composes with the small path."""
    from mvedit_b200 import mesh_raster as dr
    v, f = synth_mesh.icosphere(5)
    quad_v = np.array([[-0.9, -0.9, -0.9], [0.9, -0.9, -0.9], [0.9, 0.9, -0.9], [-0.9, 0.9, -0.9]])
    vv = np.concatenate([v * 0.6, quad_v])
    ff = np.concatenate([f, [[len(v), len(v) + 1, len(v) + 2], [len(v), len(v) + 2, len(v) + 3]]]).astype(np.int32)
    pos = torch.from_numpy(synth_mesh.project(vv, synth_mesh.surround_poses(4, 0), fov_deg=30.0).astype(np.float32)).cuda()
    tri = torch.from_numpy(ff).cuda()
    ctx = dr.RasterizeCudaContext()
    a, _ = dr.rasterize(ctx, pos, tri, (512, 512))
    b, _ = dr.rasterize(ctx, pos, tri, (512, 512))
    assert torch.equal(a, b)
    ids = a[..., 3].long() - 1
    sphere = (ids >= 0) & (ids < len(f))
    assert 0.05 < sphere.float().mean() < 0.5
    for bi in range(a.shape[0]):
        rows = sphere[bi].any(dim=1).nonzero().flatten()
        m = sphere[bi][rows]
        first = m.float().argmax(dim=1)
        last = m.shape[1] - 1 - m.flip(1).float().argmax(dim=1)
        assert (m.sum(dim=1) == last - first + 1).all()          # every row of the convex silhouette is one run


def test_interpolate_and_rasterize_backward_vs_oracle():
    from oracle import raster_oracle as ro
    from mvedit_b200 import mesh_raster as dr
    pos, tri, res = _scene('sphere')
    tri_t = torch.from_numpy(tri)
    pos_g = torch.from_numpy(pos.copy()).cuda().requires_grad_(True)
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), pos_g, tri_t.cuda(), res)
    attr = _rand((1, pos.shape[1], 5), 3)
    attr_g = attr.clone().cuda().requires_grad_(True)
    out, da = dr.interpolate(attr_g, rast, tri_t.cuda(), rast_db=db, diff_attrs='all')
    wgt = _rand(out.shape, 4)
    (out * wgt.cuda()).sum().backward()
    ids = rast.detach().cpu()[..., 3].long() - 1
    pos_o = torch.from_numpy(pos).double().requires_grad_(True)
    attr_o = attr.detach().double().requires_grad_(True)
    rast_o = torch.cat([ro.barycentrics(pos_o, tri_t, ids), rast.detach().cpu()[..., 3:].double()], dim=-1)
    out_o, da_o = ro.interpolate(attr_o, rast_o, tri_t, rast_db=db.cpu().double(), diff_attrs='all')
    (out_o * wgt.double()).sum().backward()
    torch.testing.assert_close(out.detach().cpu(), out_o.detach().float(), rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(da.cpu(), da_o.detach().float(), rtol=1e-3, atol=1e-5)
    for g, g_o in ((attr_g.grad.cpu(), attr_o.grad), (pos_g.grad.cpu(), pos_o.grad)):
        assert g_o.abs().max() > 0 and (g - g_o.float()).abs().max() <= 2e-3 * g_o.abs().max()


@pytest.mark.parametrize('kind', ['sphere', 'large'])
def test_antialias_forward_backward_vs_oracle(kind):
    from oracle import raster_oracle as ro
    from mvedit_b200 import mesh_raster as dr
    pos, tri, res = _scene(kind)
    tri_g = torch.from_numpy(tri).cuda()
    pos_g = torch.from_numpy(pos.copy()).cuda().requires_grad_(True)
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), pos_g.detach(), tri_g, res)
    B, H, W, _ = rast.shape
    fg = (rast[..., 3:] > 0).float().cpu()
    color = torch.cat([_rand((B, H, W, 3), 9).abs() * fg, fg, _rand((B, H, W, 4), 10)], dim=-1)
    color_g = color.clone().cuda().requires_grad_(True)
    out = dr.antialias(color_g, rast, pos_g, tri_g)
    color_o = color.detach().double().requires_grad_(True)
    pos_o = torch.from_numpy(pos).double().requires_grad_(True)
    out_o = ro.antialias(color_o, rast.cpu().double(), pos_o, tri)
    torch.testing.assert_close(out.detach().cpu(), out_o.detach().float(), rtol=1e-4, atol=2e-5)
    assert ((out.detach().cpu() - color).abs().sum(-1) > 1e-6).float().mean() > 0.004
    g = _rand(out.shape, 11)
    out.backward(g.cuda())
    out_o.backward(g.double())
    torch.testing.assert_close(color_g.grad.cpu(), color_o.grad.float(), rtol=1e-4, atol=2e-5)
    assert pos_o.grad.abs().max() > 0 and (pos_g.grad.cpu() - pos_o.grad.float()).abs().max() <= 2e-3 * pos_o.grad.abs().max()


# ---- the renderer and the optimisation loop with the real field ----------------------------------------------------------------------

def _fields(seed=0):
    """The product decoder (CUDA, fp32 MLP) and the oracle decoder (CPU, plain torch) with the same non-trivial parameters."""
    from mvedit_b200.ingp_decoder import iNGPDecoder
    from oracle.nerf_oracle import OracleDecoder
    dec = iNGPDecoder(max_resolution=64).cuda()
    dec.mlp_tf32 = False
    g = torch.Generator().manual_seed(seed)
    table = (torch.rand(dec.encoder.params.numel(), generator=g) - 0.5)
    with torch.no_grad():
        dec.encoder.params.copy_(table.cuda())
    od = OracleDecoder(None, max_resolution=64)
    od.load_state_dict({k: v.detach().cpu() for k, v in dec.state_dict().items()}, strict=False)
    return dec, od


def _cameras(n, size, seed=0):
    poses = torch.from_numpy(synth_mesh.surround_poses(n, seed)).float()
    intr = torch.from_numpy(synth_mesh.intrinsics(size)).float()[None].expand(n, -1).contiguous()
    return poses, intr


def test_mesh_renderer_forward_with_field_vs_oracle():
    from oracle import mesh_oracle as mo
    from mvedit_b200 import mesh_optim as mopt
    from mvedit_b200.mesh_renderer import Mesh, MeshRenderer
    dec, od = _fields()
    v, f = synth_mesh.icosphere(2)
    size, n = 40, 3
    poses, intr = _cameras(n, size)
    lights = torch.nn.functional.normalize(_rand((n, 3), 1), dim=-1)
    lp = lights[:, None, None, :].expand(-1, size, size, -1)
    vt = (torch.from_numpy(v).float() * 0.55)
    v_g = vt.clone().cuda().requires_grad_(True)
    mesh = Mesh(v=v_g, f=torch.from_numpy(f).int().cuda())
    mesh.auto_normal()
    r = MeshRenderer(near=0.01, far=100)([mesh], poses[None].cuda(), intr[None].cuda(), size, size,
                                         mopt.make_nerf_shading_fun(dec, None, lp.cuda(), 0.2), normal_bg=[0.5, 0.5, 1.0])
    v_o = vt.clone().requires_grad_(True)
    r_o = mo.mesh_renderer_forward(mo.make_mesh(v_o, torch.from_numpy(f).int()), poses[None], intr[None], size, size,
                                   mopt.make_nerf_shading_fun(od, None, lp, 0.2))
    for k in ('rgba', 'depth', 'normal'):
        torch.testing.assert_close(r[k].detach().cpu(), r_o[k].detach(), rtol=1e-3, atol=2e-3)
    gen = torch.Generator().manual_seed(2)
    ws = {k: torch.randn(r_o[k].shape, generator=gen) for k in ('rgba', 'depth', 'normal')}
    sum((r[k] * ws[k].cuda()).sum() for k in ws).backward()
    sum((r_o[k] * ws[k]).sum() for k in ws).backward()
    assert (v_g.grad.cpu() - v_o.grad).abs().max() <= 2e-2 * v_o.grad.abs().max()
    g_t, g_o = dec.encoder.params.grad.cpu(), od.encoder.params.grad
    assert g_o.abs().max() > 0 and (g_t - g_o).abs().max() <= 2e-2 * g_o.abs().max()


def test_mesh_optim_steps_vs_oracle():
    from oracle import mesh_oracle as mo
    from oracle.nerf_oracle import L1LossMod
    from mvedit_b200 import mesh_optim as mopt
    from mvedit_b200.mesh_renderer import DMTet, Mesh, MeshRenderer, make_tet_grid
    from mvedit_b200.optim import FusedAdam
    dec, od = _fields(1)
    n, size, steps = 4, 32, 2
    poses, intr = _cameras(n, size, seed=2)
    lights = torch.nn.functional.normalize(_rand((n, 3), 4), dim=-1)
    cam_weights = torch.tensor([1.0, 0.5, 1.0, 2.0])
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    tgt_masks = disc[None, None, :, :, None].expand(1, n, -1, -1, -1).contiguous()
    tgt_images = (torch.rand(1, n, size, size, 3, generator=torch.Generator().manual_seed(5)) * 0.5 + 0.25) * tgt_masks + (1 - tgt_masks)
    noise = dict(camera_perm=torch.tensor([2, 0, 3, 1]), jitter=torch.rand(steps, 2, 2, generator=torch.Generator().manual_seed(6)))
    grid = make_tet_grid(12)
    tet_verts0 = -grid['vertices'] * 2 * 0.9
    sdf0 = (0.45 - tet_verts0.norm(dim=-1) + 0.05 * torch.sin(6 * tet_verts0[:, 0]) * torch.sin(5 * tet_verts0[:, 1])).clamp(-1, 1)
    res = {}
    for name in ('product', 'oracle'):
        dev = DEV if name == 'product' else 'cpu'
        field = dec if name == 'product' else od
        tet_verts, tet_indices = tet_verts0.to(dev), grid['indices'].to(dev)
        tet_sdf = sdf0.clone().to(dev).requires_grad_(True)
        deform = torch.zeros_like(tet_verts).requires_grad_(True)
        groups = [{'params': list(field.parameters())}, {'params': [tet_sdf, deform], 'lr': 1e-3}]
        to = lambda x: x.to(dev)
        if name == 'product':
            opt = FusedAdam(groups, lr=0.01)
            dm = DMTet(dev)
            with torch.enable_grad():
                mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
                mesh = Mesh(v=mv, f=mf.int())
                mesh.auto_normal()
            nerf = SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=None)
            pipe = SimpleNamespace(nerf=nerf, mesh_renderer=MeshRenderer(near=0.01, far=100), normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
            mesh = mopt.mesh_optim(pipe, to(tgt_images), to(tgt_masks), None, opt, 0.01, 0.8, steps, 2, 8, 24, 0.0, 0.0, 0.02, 0.1, 5.0, None,
                                   tet_verts, deform, tet_sdf, tet_indices, dm, mesh, size, to(intr), size, to(poses), to(cam_weights),
                                   to(lights), 16, False, 0.2, 1.0, noise=noise)
        else:
            opt = torch.optim.Adam(groups, lr=0.01)
            dm = mo.DMTetOracle()
            mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
            mesh, _ = mo.mesh_optim(field, tgt_images, tgt_masks, opt, 0.01, 0.8, steps, 2, 8, 0.0, 0.02, 0.1, 5.0, None, tet_verts, deform,
                                    tet_sdf, tet_indices, dm, mo.make_mesh(mv, mf.int()), size, intr, size, poses, cam_weights, lights, 16, 0.2, noise)
        res[name] = dict(sdf=tet_sdf.detach().cpu(), deform=deform.detach().cpu(), table=field.encoder.params.detach().cpu().clone(),
                         nf=mesh.f.shape[0])
    p, o = res['product'], res['oracle']
    for k in ('sdf', 'deform'):
        d = (p[k] - o[k]).abs()
        assert (d > 1e-4).float().mean() < 2e-3, (k, d.max())
    assert (p['deform'].abs().max() > 1e-4) and (p['sdf'] - sdf0).abs().max() > 1e-4            # the geometry moved
    dt = (p['table'] - o['table']).abs()
    assert (dt > 1e-3).float().mean() < 2e-3
    assert abs(p['nf'] - o['nf']) <= max(4, o['nf'] // 100)


def test_init_tet_from_field():
    from mvedit_b200 import mesh_optim as mopt
    from mvedit_b200.mesh_renderer import make_tet_grid
    dec, _ = _fields(2)
    nerf = SimpleNamespace(decoder=dec)
    grid = make_tet_grid(16)
    with torch.no_grad():
        sig = dec.point_density_decode([(-grid['vertices'] * 2).cuda()], None)[0]
    thr = float(sig.quantile(0.7))
    verts, idx, sdf = mopt.init_tet(nerf, None, density_thresh=thr, tets=grid)
    assert verts.shape == (17 ** 3, 3) and idx.shape == (6 * 16 ** 3, 4) and sdf.shape == (17 ** 3,)
    assert sdf.min() >= -1 and sdf.max() <= 1 and (sdf > 0).any() and (sdf < 0).any()
    outside = (verts.abs() > 1).any(dim=-1)
    assert (sdf[outside] == -1).all()


# ---- texture side (row a-11) -------------------------------------------------------------------------------------------------------------

def test_texture_fetch_and_gradient_vs_oracle():
    from oracle import raster_oracle as ro
    from mvedit_b200 import mesh_raster as dr
    v = np.array([[[-0.9, -0.7, 0.3, 1.0], [0.9, -0.8, 0.6, 1.6], [0.8, 0.9, 0.9, 2.6], [-0.8, 0.8, 0.2, 1.2]]], np.float32)
    v[..., :3] *= v[..., 3:]
    tri = torch.tensor([[0, 1, 2], [0, 2, 3]], dtype=torch.int32).cuda()
    vt = torch.tensor([[[0.05, 0.1], [2.3, 0.0], [2.1, 1.7], [-0.4, 1.2]]]).cuda()
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), torch.from_numpy(v).cuda(), tri, (40, 48))
    uv, uv_da = dr.interpolate(vt, rast, tri, rast_db=db, diff_attrs='all')
    for mode in ('linear', 'linear-mipmap-linear'):
        tex = _rand((1, 32, 64, 3), 21)
        tex_g = tex.clone().cuda().requires_grad_(True)
        out = dr.texture(tex_g, uv, uv_da=uv_da if mode != 'linear' else None, filter_mode=mode)
        tex_o = tex.detach().double().requires_grad_(True)
        out_o = ro.texture(tex_o, uv.cpu().double(), uv_da.cpu().double() if mode != 'linear' else None, filter_mode=mode)
        # log2 of the footprint is not bit-identical between the GPU and the CPU: the blend weight between two mip levels may differ by ~1e-6
        torch.testing.assert_close(out.detach().cpu(), out_o.detach().float(), rtol=1e-4, atol=5e-5)
        g = _rand(out.shape, 22)
        out.backward(g.cuda())
        out_o.backward(g.double())
        torch.testing.assert_close(tex_g.grad.cpu(), tex_o.grad.float(), rtol=1e-3, atol=2e-4)


def test_bake_and_textured_render_vs_oracle():
    from oracle import mesh_oracle as mo
    from mvedit_b200.mesh_renderer import Mesh, MeshRenderer
    v, f = synth_mesh.icosphere(2)
    mesh = Mesh(v=(torch.from_numpy(v).float() * 0.5).cuda(), f=torch.from_numpy(f).int().cuda())
    mesh.auto_normal()
    mesh.auto_uv()
    om = mo.make_mesh(mesh.v.cpu(), mesh.f.cpu())
    om.vt, om.ft = mesh.vt.cpu(), mesh.ft.cpu()
    n, size, ms = 3, 32, 64
    poses, intr = _cameras(n, size, seed=3)
    g = torch.Generator().manual_seed(9)
    images = torch.rand(1, n, size, size, 3, generator=g)
    alphas = (torch.rand(1, n, size, size, 1, generator=g) > 0.1).float()
    r = MeshRenderer(near=0.01, far=100)
    baked = r.bake_multiview([mesh], images.cuda(), alphas.cuda(), poses[None].cuda(), intr[None].cuda(), map_size=ms, cos_weight_pow=8.0, render_bs=2)[0]
    albedo_o = mo.bake_multiview(om, images, alphas, poses[None], intr[None], map_size=ms, cos_weight_pow=8.0, render_bs=2)
    d = (baked.albedo.cpu() - albedo_o).abs()
    assert d.mean() < 2e-4 and (d > 1e-2).float().mean() < 5e-3
    with torch.no_grad():
        out = r([baked], poses[None].cuda(), intr[None].cuda(), size, size)
    om.albedo = baked.albedo.cpu()
    out_o = mo.mesh_renderer_forward(om, poses[None], intr[None], size, size)
    torch.testing.assert_close(out['rgba'].cpu(), out_o['rgba'], rtol=1e-3, atol=1e-3)


def test_fused_mesh_objective_vs_eager_on_gpu():
    """csrc/mesh_loss.cu against the torch ops + autograd it replaces (the same comparison runs on the CPU through the host harness)."""
    from mvedit_b200 import mesh_optim as mopt
    from mvedit_b200.nerf import L1LossMod
    g = torch.Generator().manual_seed(0)
    bs, size = 3, 64
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 31.5) ** 2 + (yy - 30.5) ** 2).float().sqrt() < 22).float()
    a = (disc * (0.3 + 0.7 * torch.rand(bs, size, size, generator=g)))[..., None]
    rgba0 = torch.cat([torch.rand(bs, size, size, 3, generator=g) * a, a], dim=-1).to(DEV)
    normal0 = (torch.rand(bs, size, size, 3, generator=g) * a + torch.tensor([0.5, 0.5, 1.0]) * (1 - a)).to(DEV)
    gate, tgt = torch.rand(bs, size, size, 1, generator=g).to(DEV), torch.rand(bs, size, size, 3, generator=g).to(DEV)
    m = (((xx - 32) ** 2 + (yy - 32) ** 2).float().sqrt() < 24).float()[None, :, :, None].expand(bs, -1, -1, -1).contiguous().to(DEV)
    m_erode, m_blur = mopt.min_pool(m), m * 0.96 + 0.02
    w_view = torch.tensor([0.7, 1.3, 1.0]).to(DEV)
    nbg, n_px = [0.5, 0.5, 1.0], bs * size * size
    res = {}
    for mode in ('fused', 'eager'):
        rgba, normal = rgba0.clone().requires_grad_(True), normal0.clone().requires_grad_(True)
        if mode == 'fused':
            val = mopt._MeshObjectiveFn.apply(rgba, normal, gate.squeeze(-1), tgt, m_erode.squeeze(-1), m_blur.squeeze(-1), w_view, nbg,
                                              1.2 * 4.5 / (n_px * 3), 1.2 * 2.0 / n_px, 0.1 * 2 / (n_px * 3), None)
        else:
            al = rgba[..., 3:]
            rgb = rgba[..., :3] / al.clamp(min=1e-3) * m_erode + tgt * (1 - m_erode)
            n = normal * gate + normal.detach() * (1 - gate)
            nfg = (n - torch.tensor(nbg, device=DEV) * (1 - al)) / al.clamp(min=1e-3)
            w_px = w_view[:, None, None, None].expand(-1, size, size, 1)
            l1 = L1LossMod(loss_weight=1.2)
            val = l1(rgb, tgt, weight=w_px) * 4.5 + l1(al, m_blur, weight=w_px) * 2.0 \
                + mopt.tv_normal_loss(nfg.permute(0, 3, 1, 2), al.detach().permute(0, 3, 1, 2)) * (0.1 * 2)
        val.backward()
        res[mode] = (float(val), rgba.grad.cpu(), normal.grad.cpu())
    f, e = res['fused'], res['eager']
    assert abs(f[0] - e[0]) < 1e-4 * max(1.0, abs(e[0]))
    for x, y in zip(f[1:], e[1:]):
        assert torch.isfinite(x).all() and (x - y).abs().max() <= 2e-4 * y.abs().max() + 1e-8

"""CPU checks of the mesh rasteriser (seam B5) without a GPU: the kernel source compiled as plain C++ (tests/host_harness.py) runs
the same per-triangle / per-pixel code and is compared with oracle/raster_oracle.py -- ids and barycentrics bit for bit, gradients
against the oracle's autograd -- through the product's Python mirror (mvedit_b200/mesh_raster.py) routed to the harness."""
import numpy as np
import pytest
import torch

from oracle import raster_oracle as ro
from tests import host_harness, synth_mesh
from mvedit_b200 import mesh_raster as dr


@pytest.fixture(autouse=True)
def _route():
    with host_harness.routed(dr):
        yield


def _scene(kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == 'sphere':
        v, f = synth_mesh.icosphere(2)
        pos = synth_mesh.project(v * 0.6, synth_mesh.surround_poses(3, seed), fov_deg=30.0)
        return pos.astype(np.float32), f.astype(np.int32), (48, 40)
    if kind == 'soup':                       # random small triangles at random depths, both windings, some behind the camera / off screen
        n = 300
        c = rng.uniform(-1.2, 1.2, (n, 1, 2))
        xy = c + rng.normal(0, 0.08, (n, 3, 2))
        z = rng.uniform(-1.3, 1.3, (n, 1, 1)) + rng.normal(0, 0.05, (n, 3, 1))
        w = rng.uniform(0.5, 2.0, (n, 3, 1))
        w[:5] = -w[:5]
        w[5:8, 0] = 0.0
        v = np.concatenate([xy * w, z * w, w], axis=-1).reshape(1, n * 3, 4)
        f = np.arange(n * 3).reshape(n, 3)
        f[10] = f[10][[0, 0, 1]]            # zero-area
        return v.astype(np.float32), f.astype(np.int32), (37, 53)
    if kind == 'large':                      # screen-filling, overlapping, interpenetrating triangles -> the queued path
        v = np.array([[[-1.5, -1.5, 0.2, 1], [1.5, -1.5, 0.2, 1], [1.5, 1.5, 0.2, 1], [-1.5, 1.5, 0.2, 1],
                       [-0.9, -0.8, -0.5, 1], [0.9, -0.7, 0.9, 1], [0.0, 0.95, 0.1, 1],
                       [-3, -3, 0.5, 2], [3, -3, 0.5, 2], [0, 3, -0.5, 2]]], np.float32)
        f = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [7, 8, 9], [0, 1, 2]], np.int32)   # the last duplicates the first: tie -> lower id
        return v, f, (64, 64)
    raise KeyError(kind)


@pytest.mark.parametrize('kind', ['sphere', 'soup', 'large'])
def test_rasterize_matches_oracle_bit_for_bit(kind):
    pos, tri, res = _scene(kind)
    r_o, db_o = ro.rasterize(pos, tri, res)
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), torch.from_numpy(pos), torch.from_numpy(tri), res)
    rast, db = rast.numpy(), db.numpy()
    assert (rast[..., 3] == r_o[..., 3]).all()
    assert (rast[..., 3] > 0).sum() > 50
    np.testing.assert_array_equal(rast.view(np.uint32), r_o.view(np.uint32))
    np.testing.assert_array_equal(db.view(np.uint32), db_o.view(np.uint32))
    fg = rast[..., 3] > 0
    assert (rast[fg][:, :2] >= 0).all() and (rast[fg][:, 0] + rast[fg][:, 1] <= 1 + 1e-6).all()
    assert (np.abs(rast[fg][:, 2]) <= 1).all()
    r2, db2 = dr.rasterize(dr.RasterizeCudaContext(), torch.from_numpy(pos), torch.from_numpy(tri), res, grad_db=False)
    assert (db2.numpy() == db).all() and (r2.numpy() == rast).all()      # rast_db is produced whatever grad_db says (as nvdiffrast's CUDA context)


def test_closed_mesh_is_watertight():
    """Shared edges are exactly negated edge functions: no pixel inside the silhouette of a closed mesh is left empty."""
    v, f = synth_mesh.icosphere(3)
    pos = synth_mesh.project(v * 0.6, synth_mesh.surround_poses(2, 1), fov_deg=30.0).astype(np.float32)
    H = W = 96
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), torch.from_numpy(pos), torch.from_numpy(f.astype(np.int32)), (H, W))
    fg = (rast[..., 3] > 0).numpy()
    for b in range(fg.shape[0]):
        for y in range(H):
            xs = np.nonzero(fg[b, y])[0]
            if len(xs):
                assert fg[b, y, xs[0]:xs[-1] + 1].all()     # a convex silhouette: every row is one run
        assert 0.1 < fg[b].mean() < 0.6
    # nearest surface wins: all winners face the camera
    ids = (rast[..., 3].long() - 1).numpy()
    for b in range(fg.shape[0]):
        t = np.unique(ids[b][fg[b]])
        p = pos[b][f[t]]
        s = p[..., :2] / p[..., 3:]
        area = (s[:, 1, 0] - s[:, 0, 0]) * (s[:, 2, 1] - s[:, 0, 1]) - (s[:, 2, 0] - s[:, 0, 0]) * (s[:, 1, 1] - s[:, 0, 1])
        assert (np.sign(area) == np.sign(area[0])).all()


def test_rasterize_depth_matches_analytic_sphere():
    v, f = synth_mesh.icosphere(4)
    poses = synth_mesh.surround_poses(1, 0)
    pos, vcam = synth_mesh.project(v * 0.5, poses, fov_deg=30.0, return_cam=True)
    H = W = 64
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), torch.from_numpy(pos.astype(np.float32)), torch.from_numpy(f.astype(np.int32)), (H, W))
    depth_attr = torch.from_numpy((-vcam[..., 2:3]).astype(np.float32))
    inv_d, _ = dr.interpolate(depth_attr, rast, torch.from_numpy(f.astype(np.int32)))
    fg = rast[..., 3] > 0
    d = inv_d[..., 0][fg].numpy()
    dist = np.linalg.norm(poses[0, :3, 3])
    assert abs(d.min() - (dist - 0.5)) < 5e-3            # nearest point of the sphere of radius 0.5
    assert d.max() < dist


def _rand_like(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).normal(size=shape).astype(np.float32))


def test_interpolate_forward_backward_vs_oracle():
    pos, tri, res = _scene('sphere')
    pos_t, tri_t = torch.from_numpy(pos), torch.from_numpy(tri)
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), pos_t, tri_t, res)
    B, V = pos.shape[:2]
    for batched in (False, True):
        attr = _rand_like((B if batched else 1, V, 5), 3).requires_grad_(True)
        rast_l = rast.clone().requires_grad_(True)
        out, da = dr.interpolate(attr, rast_l, tri_t, rast_db=db, diff_attrs='all')
        attr_o = attr.detach().double().requires_grad_(True)
        rast_o = rast.double().requires_grad_(True)
        out_o, da_o = ro.interpolate(attr_o, rast_o, tri_t, rast_db=db.double(), diff_attrs='all')
        torch.testing.assert_close(out, out_o.float(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(da, da_o.float(), rtol=1e-4, atol=1e-6)
        g = _rand_like(out.shape, 4)
        out.backward(g)
        out_o.backward(g.double())
        torch.testing.assert_close(attr.grad, attr_o.grad.float(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(rast_l.grad[..., :2], rast_o.grad[..., :2].float(), rtol=1e-4, atol=1e-5)
        assert (rast_l.grad[..., 2:] == 0).all()
    out2, da2 = dr.interpolate(attr, rast, tri_t)
    assert da2.shape[-1] == 0


def test_rasterize_backward_vs_oracle_autograd():
    pos, tri, res = _scene('sphere')
    tri_t = torch.from_numpy(tri)
    pos_t = torch.from_numpy(pos).requires_grad_(True)
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), pos_t, tri_t, res)
    g = _rand_like(rast.shape, 5)
    g[..., 3] = 0
    rast.backward(g)
    pos_o = torch.from_numpy(pos).double().requires_grad_(True)
    ids = rast[..., 3].long() - 1
    uvz = ro.barycentrics(pos_o, tri_t, ids)
    torch.testing.assert_close(rast[..., :3].detach(), uvz.float(), rtol=1e-4, atol=2e-5)
    uvz.backward(g[..., :3].double())
    scale = pos_o.grad.abs().max()
    assert scale > 0
    assert (pos_t.grad - pos_o.grad.float()).abs().max() <= 2e-3 * scale


def test_gradient_flows_from_interpolated_attribute_to_positions():
    """The chain mesh_optim relies on: loss(interpolate(attr, rasterize(pos))) -> d pos (finite-difference check, fp64 oracle chain)."""
    pos, tri, res = _scene('sphere')
    tri_t = torch.from_numpy(tri)
    pos_t = torch.from_numpy(pos).requires_grad_(True)
    attr = _rand_like((1, pos.shape[1], 3), 7)
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), pos_t, tri_t, res)
    out, _ = dr.interpolate(attr, rast, tri_t)
    wgt = _rand_like(out.shape, 8)
    (out * wgt).sum().backward()
    ids = rast[..., 3].long() - 1
    pos_o = torch.from_numpy(pos).double().requires_grad_(True)
    uvz = ro.barycentrics(pos_o, tri_t, ids)
    rast_o = torch.cat([uvz, rast.detach()[..., 3:].double()], dim=-1)
    out_o, _ = ro.interpolate(attr.double(), rast_o, tri_t)
    (out_o * wgt.double()).sum().backward()
    scale = pos_o.grad.abs().max()
    assert (pos_t.grad - pos_o.grad.float()).abs().max() <= 2e-3 * scale


def test_edge_opposites_matches_oracle():
    v, f = synth_mesh.icosphere(1)
    f = np.concatenate([f[:-3], [[0, 1, 2]]]).astype(np.int32)      # open edges and an extra user of some edges
    opp = dr.edge_opposites(torch.from_numpy(f)).numpy()
    np.testing.assert_array_equal(opp, ro.edge_opposites(f))
    closed = dr.edge_opposites(torch.from_numpy(synth_mesh.icosphere(1)[1].astype(np.int32))).numpy()
    assert (closed >= 0).all()


@pytest.mark.parametrize('kind', ['sphere', 'large'])
def test_antialias_forward_backward_vs_oracle(kind):
    pos, tri, res = _scene(kind)
    tri_t = torch.from_numpy(tri)
    pos_t = torch.from_numpy(pos).requires_grad_(True)
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), pos_t.detach(), tri_t, res)
    B, H, W, _ = rast.shape
    fg = (rast[..., 3:] > 0).float()
    color = (torch.cat([_rand_like((B, H, W, 3), 9).abs() * fg, fg, _rand_like((B, H, W, 4), 10)], dim=-1)).requires_grad_(True)
    out = dr.antialias(color, rast, pos_t, tri_t)
    color_o = color.detach().double().requires_grad_(True)
    pos_o = pos_t.detach().double().requires_grad_(True)
    out_o = ro.antialias(color_o, rast.double(), pos_o, tri)
    changed = ((out.detach() - color.detach()).abs().sum(-1) > 1e-6)
    assert changed.float().mean() > 0.004                   # silhouettes were found
    if kind == 'sphere':                                    # fractional coverage on the alpha channel along the outline
        assert ((out.detach()[..., 3] > 0.02) & (out.detach()[..., 3] < 0.98)).sum() > 10
    torch.testing.assert_close(out, out_o.float(), rtol=1e-4, atol=1e-5)
    interior = ~changed
    torch.testing.assert_close(out[interior], color[interior])
    g = _rand_like(out.shape, 11)
    out.backward(g)
    out_o.backward(g.double())
    torch.testing.assert_close(color.grad, color_o.grad.float(), rtol=1e-4, atol=1e-5)
    scale = pos_o.grad.abs().max()
    assert scale > 0
    assert (pos_t.grad - pos_o.grad.float()).abs().max() <= 2e-3 * scale


def test_antialias_alpha_tracks_subpixel_motion():
    """Moving a silhouette by a fraction of a pixel changes the antialiased coverage by about that fraction of the edge length --
    the property that makes the alpha loss differentiable w.r.t. geometry."""
    H = W = 32
    tri = torch.tensor([[0, 1, 2], [0, 2, 3]], dtype=torch.int32)

    def cover(x_edge):
        v = torch.tensor([[[-0.5, -0.5, 0, 1], [x_edge, -0.5, 0, 1], [x_edge, 0.5, 0, 1], [-0.5, 0.5, 0, 1]]], dtype=torch.float32)
        rast, _ = dr.rasterize(dr.RasterizeCudaContext(), v, tri, (H, W))
        a = (rast[..., 3:] > 0).float()
        return dr.antialias(a, rast, v, tri).sum().item()

    px = 2.0 / W
    c0, c1 = cover(0.25 + 0.1 * px), cover(0.25 + 0.6 * px)
    assert abs((c1 - c0) - 0.5 * 16) < 0.3             # the edge is 16 pixels long; it moved half a pixel


def _uv_scene(seed=0):
    """uv / uv_da of a textured quad seen at a grazing angle: footprints from sub-texel to many texels."""
    v = np.array([[[-0.9, -0.7, 0.3, 1.0], [0.9, -0.8, 0.6, 1.6], [0.8, 0.9, 0.9, 2.6], [-0.8, 0.8, 0.2, 1.2]]], np.float32)
    v[..., :3] *= v[..., 3:]
    tri = torch.tensor([[0, 1, 2], [0, 2, 3]], dtype=torch.int32)
    vt = torch.tensor([[[0.05, 0.1], [2.3, 0.0], [2.1, 1.7], [-0.4, 1.2]]])         # beyond [0,1]: wrap addressing
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), torch.from_numpy(v), tri, (40, 48))
    uv, uv_da = dr.interpolate(vt, rast, tri, rast_db=db, diff_attrs='all')
    return uv, uv_da


@pytest.mark.parametrize('mode', ['linear', 'linear-mipmap-linear'])
@pytest.mark.parametrize('batched', [False, True])
def test_texture_forward_backward_vs_oracle(mode, batched):
    uv, uv_da = _uv_scene()
    uv, uv_da = uv.expand(2, -1, -1, -1).contiguous(), uv_da.expand(2, -1, -1, -1).contiguous()
    tex = _rand_like((2 if batched else 1, 32, 64, 3), 21).requires_grad_(True)
    out = dr.texture(tex, uv, uv_da=uv_da if mode != 'linear' else None, filter_mode=mode)
    tex_o = tex.detach().double().requires_grad_(True)
    out_o = ro.texture(tex_o, uv.double(), uv_da.double() if mode != 'linear' else None, filter_mode=mode)
    torch.testing.assert_close(out, out_o.float(), rtol=1e-4, atol=2e-5)
    g = _rand_like(out.shape, 22)
    out.backward(g)
    out_o.backward(g.double())
    torch.testing.assert_close(tex.grad, tex_o.grad.float(), rtol=1e-4, atol=1e-4)
    if mode != 'linear':
        lv = 0.5 * torch.log2((uv_da[..., [0, 2]] * torch.tensor([64.0, 32.0])).square().sum(-1).clamp(min=1e-12))
        assert (lv > 1).any() and (lv < 0).any()             # the scene exercises several pyramid levels and the clamp at level 0


def test_texture_of_ones_map_gradient_is_a_coverage_map():
    """The reference's baking trick (base_mesh_renderer.py:470-475): d(sum of fetches) / d(texture of ones) = how much every texel is
    seen; the weights of every fetch sum to one on every level, so the total equals the pixel count."""
    uv, uv_da = _uv_scene()
    ones = torch.ones(1, 64, 64, 1, requires_grad=True)
    out = dr.texture(ones, uv, uv_da=uv_da, filter_mode='linear-mipmap-linear')
    torch.testing.assert_close(out, torch.ones_like(out), rtol=0, atol=1e-6)
    (g,) = torch.autograd.grad(out.sum(), ones)
    assert abs(g.sum().item() - out.numel()) < 1e-2 * out.numel() and (g >= 0).all()


def test_texture_rejects_what_is_not_built():
    uv, uv_da = _uv_scene()
    tex = torch.ones(1, 8, 8, 1)
    with pytest.raises(NotImplementedError):
        dr.texture(tex, uv, filter_mode='nearest')
    with pytest.raises(NotImplementedError):
        dr.texture(tex, uv, boundary_mode='clamp')
    with pytest.raises(NotImplementedError):
        dr.texture(tex, uv.clone().requires_grad_(True))
    assert dr.texture(torch.ones(1, 6, 10, 1), uv, uv_da=uv_da).shape[-1] == 1       # odd halves: the pyramid stops at 3 x 5


def test_empty_face_list_is_rendered_as_an_empty_image():
    """A DMTet field that lost its zero crossing yields no triangles: every op must pass that through (empty raster, zero attributes,
    colours untouched, zero gradients) instead of reading a face that does not exist."""
    pos = torch.randn(2, 5, 4).abs().requires_grad_(True)
    tri = torch.zeros(0, 3, dtype=torch.int32)
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), pos, tri, (6, 7))
    assert rast.shape == (2, 6, 7, 4) and (rast == 0).all() and (db == 0).all()
    attr = torch.randn(1, 5, 3, requires_grad=True)
    out, _ = dr.interpolate(attr, rast, tri)
    col = torch.rand(2, 6, 7, 4, requires_grad=True)
    aa = dr.antialias(col, rast, pos, tri)
    assert (out == 0).all() and torch.equal(aa, col)
    (out.sum() + aa.sum() + rast.sum()).backward()
    assert (attr.grad == 0).all() and (pos.grad == 0).all() and (col.grad == 1).all()


def test_hash_edge_topology_equals_the_sorted_one():
    rng = np.random.default_rng(5)
    cases = [synth_mesh.icosphere(2)[1], synth_mesh.icosphere(3)[1][:-7]]                          # closed; with open edges
    cases.append(np.concatenate([synth_mesh.icosphere(1)[1], synth_mesh.icosphere(1)[1][:9]]))      # edges with three and four users
    cases += [rng.integers(0, 40, (int(rng.integers(1, 300)), 3)) for _ in range(6)]                # soups incl. degenerate edges (a == b)
    for f in cases:
        t = torch.from_numpy(np.asarray(f).astype(np.int32))
        a, b = dr.edge_opposites(t), dr.edge_opposites(t, method='hash')
        assert a.dtype == b.dtype == torch.int32 and torch.equal(a, b)
        np.testing.assert_array_equal(a.numpy(), ro.edge_opposites(np.asarray(f)))
    assert dr.edge_opposites(torch.zeros(0, 3, dtype=torch.int32), method='hash').shape == (0, 3)

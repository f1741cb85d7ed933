"""GPU parity: libmvedit_b200 ray-marching kernels (through the C ABI / reference-shaped Python mirror)
vs (1) the CPU oracle, (2) the reference's own CUDA kernels (oracle/_ref) when the prebuilt .so travelled.

Tolerances (fp32 path): integer/index outputs exact; march samples bit-exact vs the reference kernels
(same arithmetic), and vs the C oracle equal counts on >= 99.5 % of rays (FMA contraction differs between nvcc
and gcc, so a t that lands within an ulp of a voxel face may flip a cell); composite sums rtol 1e-4 / atol 1e-6
(warp-scan order differs from the sequential product).
"""
import numpy as np
import pytest
import torch

from oracle import raymarching_oracle as orc
from oracle import build_ref
from tests import synth

pytestmark = pytest.mark.gpu
AABB = np.array([-1, -1, -1, 1, 1, 1], np.float32)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope='module')
def rm():
    from mvedit_b200 import raymarching
    return raymarching


@pytest.fixture(scope='module')
def ref():
    m = build_ref.load_ref()
    if m is None:
        pytest.skip('oracle/_ref not built')
    return m


@pytest.fixture(scope='module')
def scene():
    H = 64
    grid = synth.sphere_density_grid(H=H, radius=0.5)
    bitfield = orc.packbits(grid, 0.5)
    poses = synth.surround_poses(4, seed=1)
    ro, rd, f = synth.camera_rays(poses, 48)
    rng = np.random.default_rng(7)
    noises = rng.random(ro.shape[0]).astype(np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, AABB, 0.2)
    return dict(H=H, grid=grid, bitfield=bitfield, ro=ro, rd=rd, f=f, noises=noises, nears=nears, fars=fars, max_steps=256)


def test_utils_exact(rm):
    rng = np.random.default_rng(0)
    c = rng.integers(0, 128, (5001, 3)).astype(np.int32)
    idx = rm.morton3D(cu(c)).cpu().numpy()
    assert np.array_equal(idx, orc.morton3D(c))
    assert np.array_equal(rm.morton3D_invert(cu(idx)).cpu().numpy(), c)
    g = rng.random(8 * 4099).astype(np.float32)
    assert np.array_equal(rm.packbits(cu(g), 0.5).cpu().numpy(), orc.packbits(g, 0.5))
    gh = g.astype(np.float16)
    assert np.array_equal(rm.packbits(cu(gh), 0.5).cpu().numpy(), orc.packbits(gh.astype(np.float32), 0.5))
    assert rm.packbits(cu(np.zeros(0, np.float32)), 0.5).numel() == 0  # empty input


def test_near_far_exact(rm, scene):
    rng = np.random.default_rng(2)
    o = np.concatenate([scene['ro'], rng.normal(size=(777, 3)).astype(np.float32) * 3])
    d = rng.normal(size=(777, 3)).astype(np.float32)
    d = np.concatenate([scene['rd'], d / np.linalg.norm(d, axis=-1, keepdims=True)])
    n_o, f_o = orc.near_far_from_aabb(o, d, AABB, 0.2)
    n_g, f_g = rm.near_far_from_aabb(cu(o), cu(d), cu(AABB), 0.2)
    assert np.array_equal(n_g.cpu().numpy(), n_o) and np.array_equal(f_g.cpu().numpy(), f_o)


def _march_gpu(rm, s, **kw):
    x, d, t, rays = rm.march_rays_train(cu(s['ro']), cu(s['rd']), 1.0, cu(s['bitfield']), 1, s['H'], cu(s['nears']), cu(s['fars']),
                                        perturb=True, dt_gamma=1 / s['f'], max_steps=s['max_steps'], noises=cu(s['noises']), **kw)
    return x.cpu().numpy(), d.cpu().numpy(), t.cpu().numpy(), rays.cpu().numpy()


def _check_partition(rays, M):
    order = np.argsort(rays[:, 0], kind='stable')
    o, c = rays[order, 0].astype(np.int64), rays[order, 1].astype(np.int64)
    nz = c > 0
    assert c.sum() == M
    assert np.array_equal(o[nz], np.concatenate([[0], np.cumsum(c[nz])[:-1]]))


def test_march_train_vs_oracle(rm, scene):
    xo, do, to, ro_ = orc.march_rays_train(scene['ro'], scene['rd'], 1.0, scene['bitfield'], 1, scene['H'], scene['nears'], scene['fars'],
                                           scene['noises'], dt_gamma=1 / scene['f'], max_steps=scene['max_steps'])
    xg, dg, tg, rg = _march_gpu(rm, scene)
    _check_partition(rg, xg.shape[0])
    same = rg[:, 1] == ro_[:, 1]
    assert same.mean() >= 0.995, same.mean()
    assert abs(int(xg.shape[0]) - int(xo.shape[0])) <= 0.002 * xo.shape[0]
    for n in np.nonzero(same & (rg[:, 1] > 0))[0][::7]:
        a, b, c = rg[n, 0], ro_[n, 0], rg[n, 1]
        np.testing.assert_allclose(xg[a:a + c], xo[b:b + c], atol=2e-6)
        np.testing.assert_allclose(tg[a:a + c], to[b:b + c], rtol=1e-6)
        assert np.array_equal(dg[a:a + c], do[b:b + c])


def test_march_train_capacity_protocol_matches_two_pass(rm, scene):
    xg, dg, tg, rg = _march_gpu(rm, scene)
    M = xg.shape[0]
    x2, d2, t2, r2, counter = rm.march_rays_train(cu(scene['ro']), cu(scene['rd']), 1.0, cu(scene['bitfield']), 1, scene['H'],
                                                  cu(scene['nears']), cu(scene['fars']), perturb=True, dt_gamma=1 / scene['f'],
                                                  max_steps=scene['max_steps'], noises=cu(scene['noises']), max_points=M + 1000)
    assert int(counter.item()) == M
    r2 = r2.cpu().numpy()
    assert np.array_equal(r2[:, 1], rg[:, 1])
    x2, t2 = x2.cpu().numpy(), t2.cpu().numpy()
    for n in np.nonzero(rg[:, 1] > 0)[0][::11]:
        a, b, c = rg[n, 0], r2[n, 0], rg[n, 1]
        assert np.array_equal(xg[a:a + c], x2[b:b + c]) and np.array_equal(tg[a:a + c], t2[b:b + c])
    # capacity too small: rays that do not fit write nothing but keep their counts
    x3, d3, t3, r3, counter3 = rm.march_rays_train(cu(scene['ro']), cu(scene['rd']), 1.0, cu(scene['bitfield']), 1, scene['H'],
                                                   cu(scene['nears']), cu(scene['fars']), perturb=True, dt_gamma=1 / scene['f'],
                                                   max_steps=scene['max_steps'], noises=cu(scene['noises']), max_points=M // 2)
    assert int(counter3.item()) == M


def test_march_train_vs_reference_kernels_bit_exact(rm, ref, scene):
    s = scene
    N = s['ro'].shape[0]
    ro, rd, bf, nears, fars, noises = (cu(s[k]) for k in ('ro', 'rd', 'bitfield', 'nears', 'fars', 'noises'))
    counter = torch.zeros(1, dtype=torch.int32, device='cuda')
    rays = torch.empty(N, 2, dtype=torch.int32, device='cuda')
    ref.march_rays_train(ro, rd, bf, 1.0, False, 1 / s['f'], s['max_steps'], N, 1, s['H'], nears, fars, None, None, None, rays, counter, noises)
    M = int(counter.item())
    x, d, t = (torch.zeros(M, k, device='cuda') for k in (3, 3, 2))
    ref.march_rays_train(ro, rd, bf, 1.0, False, 1 / s['f'], s['max_steps'], N, 1, s['H'], nears, fars, x, d, t, rays, counter, noises)
    xr, tr, rr = x.cpu().numpy(), t.cpu().numpy(), rays.cpu().numpy()
    xg, dg, tg, rg = _march_gpu(rm, s)
    assert xg.shape[0] == M
    assert np.array_equal(rg[:, 1], rr[:, 1])
    for n in np.nonzero(rg[:, 1] > 0)[0]:
        a, b, c = rg[n, 0], rr[n, 0], rg[n, 1]
        assert np.array_equal(xg[a:a + c], xr[b:b + c]) and np.array_equal(tg[a:a + c], tr[b:b + c])
    # nears/fars of the reference kernel
    n_r, f_r = torch.empty(N, device='cuda'), torch.empty(N, device='cuda')
    ref.near_far_from_aabb(ro, rd, cu(AABB), N, 0.2, n_r, f_r)
    assert np.array_equal(n_r.cpu().numpy(), s['nears']) and np.array_equal(f_r.cpu().numpy(), s['fars'])


def _composite_inputs(scene, rm, seed=0):
    xg, dg, tg, rg = _march_gpu(rm, scene)
    rng = np.random.default_rng(seed)
    M = xg.shape[0]
    sig = np.exp(rng.normal(size=M) * 1.5 + 1.5).astype(np.float32)
    rgb = rng.random((M, 3)).astype(np.float32)
    return sig, rgb, tg, rg


@pytest.mark.parametrize('T_thresh,binarize', [(1e-4, False), (0.2, False), (1e-4, True)])
def test_composite_train_fwd_bwd_vs_oracle(rm, scene, T_thresh, binarize):
    sig, rgb, ts, rays = _composite_inputs(scene, rm)
    N = rays.shape[0]
    w_o, ws_o, d_o, img_o = orc.composite_rays_train_forward(sig, rgb, ts, rays, T_thresh, binarize)
    s_t, c_t = cu(sig).requires_grad_(True), cu(rgb).requires_grad_(True)
    w, ws, d, img = rm.composite_rays_train(s_t, c_t, cu(ts), cu(rays), T_thresh, binarize)
    tol = dict(rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(w.detach().cpu().numpy(), w_o, **tol)
    np.testing.assert_allclose(ws.detach().cpu().numpy(), ws_o, **tol)
    np.testing.assert_allclose(d.detach().cpu().numpy(), d_o, **tol)
    np.testing.assert_allclose(img.detach().cpu().numpy(), img_o, **tol)
    rng = np.random.default_rng(3)
    gw, gws, gd, gi = (rng.normal(size=s).astype(np.float32) for s in [(sig.shape[0],), (N,), (N,), (N, 3)])
    torch.autograd.backward([w, ws, d, img], [cu(gw), cu(gws), cu(gd), cu(gi)])
    gs_o, gc_o = orc.composite_rays_train_backward(gw, gws, gd, gi, sig, rgb, ts, rays, ws_o, d_o, img_o, T_thresh, binarize)
    np.testing.assert_allclose(c_t.grad.cpu().numpy(), gc_o, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(s_t.grad.cpu().numpy(), gs_o, rtol=2e-3, atol=2e-5)


def test_composite_train_vs_reference_kernels(rm, ref, scene):
    sig, rgb, ts, rays = _composite_inputs(scene, rm, seed=5)
    M, N = sig.shape[0], rays.shape[0]
    s_t, c_t, t_t, r_t = cu(sig), cu(rgb), cu(ts), cu(rays)
    w_r, ws_r, d_r, img_r = torch.zeros(M, device='cuda'), torch.empty(N, device='cuda'), torch.empty(N, device='cuda'), torch.empty(N, 3, device='cuda')
    ref.composite_rays_train_forward(s_t, c_t, t_t, r_t, M, N, 1e-4, False, w_r, ws_r, d_r, img_r)
    s_g, c_g = s_t.clone().requires_grad_(True), c_t.clone().requires_grad_(True)
    w, ws, d, img = rm.composite_rays_train(s_g, c_g, t_t, r_t, 1e-4, False)
    tol = dict(rtol=1e-4, atol=2e-6)
    for a, b in ((w, w_r), (ws, ws_r), (d, d_r), (img, img_r)):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.cpu().numpy(), **tol)
    rng = np.random.default_rng(9)
    gw, gws, gd, gi = (cu(rng.normal(size=s).astype(np.float32)) for s in [(M,), (N,), (N,), (N, 3)])
    gs_r, gc_r = torch.zeros(M, device='cuda'), torch.zeros(M, 3, device='cuda')
    ref.composite_rays_train_backward(gw, gws, gd, gi, s_t, c_t, t_t, r_t, ws_r, d_r, img_r, M, N, 1e-4, False, gs_r, gc_r)
    torch.autograd.backward([w, ws, d, img], [gw, gws, gd, gi])
    np.testing.assert_allclose(c_g.grad.cpu().numpy(), gc_r.cpu().numpy(), rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(s_g.grad.cpu().numpy(), gs_r.cpu().numpy(), rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize('counts', [[0, 0, 0], [1], [3, 0, 40, 7, 1, 0, 9], [1024, 5]])
def test_composite_train_ragged_and_empty(rm, counts):
    counts = np.array(counts, np.int32)
    rays = np.stack([np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32), counts], -1)
    M = int(counts.sum())
    rng = np.random.default_rng(1)
    sig = np.exp(rng.normal(size=M)).astype(np.float32)
    rgb = rng.random((M, 3)).astype(np.float32)
    ts = np.stack([2 + np.sort(rng.random(M)), 0.003 + 0.01 * rng.random(M)], -1).astype(np.float32)
    w_o, ws_o, d_o, img_o = orc.composite_rays_train_forward(sig, rgb, ts, rays)
    if M == 0:
        sig_t, rgb_t, ts_t = torch.zeros(0, device='cuda'), torch.zeros(0, 3, device='cuda'), torch.zeros(0, 2, device='cuda')
    else:
        sig_t, rgb_t, ts_t = cu(sig), cu(rgb), cu(ts)
    w, ws, d, img = rm.composite_rays_train(sig_t, rgb_t, ts_t, cu(rays))
    np.testing.assert_allclose(ws.cpu().numpy(), ws_o, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(img.cpu().numpy(), img_o, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(w.cpu().numpy(), w_o, rtol=1e-4, atol=2e-6)


def test_inference_protocol_vs_oracle_and_reference(rm, scene):
    s = scene
    ref = build_ref.load_ref()
    N = s['ro'].shape[0]
    field = lambda x: (np.exp(3 * np.sin(x * 7).sum(-1)).astype(np.float32), (0.5 + 0.5 * np.cos(x * 5)).astype(np.float32))
    # oracle loop
    ws_o, d_o, img_o = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    alive_o, t_o = np.arange(N, dtype=np.int32), s['nears'].copy()
    # gpu loop
    ro, rd, bf, nears, fars = (cu(s[k]) for k in ('ro', 'rd', 'bitfield', 'nears', 'fars'))
    ws, d, img = torch.zeros(N, device='cuda'), torch.zeros(N, device='cuda'), torch.zeros(N, 3, device='cuda')
    alive, rt = torch.arange(N, dtype=torch.int32, device='cuda'), nears.clone()
    if ref is not None:
        ws_r, d_r, img_r = torch.zeros(N, device='cuda'), torch.zeros(N, device='cuda'), torch.zeros(N, 3, device='cuda')
        alive_r, rt_r = alive.clone(), nears.clone()
    step = 0
    while step < s['max_steps'] and alive.numel():
        n_alive = alive.numel()
        assert alive_o.size == n_alive
        n_step = min(max(N // n_alive, 1), 8)
        xo, _, to = orc.march_rays(n_alive, n_step, alive_o, t_o, s['ro'], s['rd'], 1.0, s['bitfield'], 1, s['H'], s['nears'], s['fars'], None,
                                   dt_gamma=1 / s['f'], max_steps=s['max_steps'])
        x, dd, t = rm.march_rays(n_alive, n_step, alive, rt, ro, rd, 1.0, bf, 1, s['H'], nears, fars, dt_gamma=1 / s['f'], max_steps=s['max_steps'])
        mism = np.any(t.cpu().numpy() != to, axis=-1).reshape(n_alive, n_step).any(-1)
        assert mism.mean() <= 0.005
        sg, cg = field(x.cpu().numpy())
        if ref is not None:
            xr, dr, tr = torch.zeros_like(x), torch.zeros_like(dd), torch.zeros_like(t)
            ref.march_rays(n_alive, n_step, alive_r, rt_r, ro, rd, 1.0, False, 1 / s['f'], s['max_steps'], 1, s['H'], bf, nears, fars, xr, dr, tr,
                           torch.zeros(n_alive, device='cuda'))
            assert torch.equal(xr, x) and torch.equal(tr, t) and torch.equal(dr, dd)
            ref.composite_rays(n_alive, n_step, 1e-2, False, alive_r, rt_r, cu(sg), cu(cg), tr, ws_r, d_r, img_r)
        orc.composite_rays(n_alive, n_step, alive_o, t_o, *field(xo), to, ws_o, d_o, img_o, T_thresh=1e-2)
        rm.composite_rays(n_alive, n_step, alive, rt, cu(sg), cu(cg), t, ws, d, img, 1e-2)
        if ref is not None:
            assert torch.equal(alive_r, alive)
            alive_r = alive_r[alive_r >= 0]
        # use the GPU's alive set for both so a 1-ulp march flip cannot desynchronise the loops
        alive = alive[alive >= 0]
        alive_o = np.ascontiguousarray(alive.cpu().numpy())
        t_o = rt.cpu().numpy().copy()
        ws_o, d_o, img_o = ws.cpu().numpy().copy(), d.cpu().numpy().copy(), img.cpu().numpy().copy()
        step += n_step
    if ref is not None:
        np.testing.assert_allclose(ws.cpu().numpy(), ws_r.cpu().numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(img.cpu().numpy(), img_r.cpu().numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(d.cpu().numpy(), d_r.cpu().numpy(), rtol=1e-6, atol=1e-7)
    assert float(ws.max()) <= 1.0 + 1e-5 and float(ws.max()) > 0.9


def test_full_size_properties_config5(rm):
    """BASELINE config 5 size (64 views x 256^2 rays, 128^3 grid): size-independent properties."""
    H = 128
    grid = synth.sphere_density_grid(H=H, radius=0.5)
    bf = rm.packbits(cu(grid), 0.5)
    assert np.array_equal(bf.cpu().numpy(), synth.pack_bitfield_np(grid))
    poses = synth.surround_poses(64, seed=0)
    ro, rd, f = synth.camera_rays(poses, 256)
    ro, rd = cu(ro), cu(rd)
    N = ro.shape[0]
    nears, fars = rm.near_far_from_aabb(ro, rd, cu(AABB), 0.2)
    g = torch.Generator(device='cuda').manual_seed(0)
    noises = torch.rand(N, device='cuda', generator=g)
    x, d, t, rays = rm.march_rays_train(ro, rd, 1.0, bf, 1, H, nears, fars, perturb=True, dt_gamma=1 / f, max_steps=1024, noises=noises)
    M = x.shape[0]
    assert M > 50 * N // 4
    _check_partition(rays.cpu().numpy(), M)
    assert float(x.norm(dim=-1).max()) < 0.5 + 3 * 2 / H  # all samples inside occupied (sphere) voxels
    sig = torch.exp(torch.randn(M, device='cuda', generator=g))
    c1, c2 = torch.rand(M, 3, device='cuda', generator=g), torch.rand(M, 3, device='cuda', generator=g)
    w1, ws1, d1, i1 = rm.composite_rays_train(sig, c1, t, rays)
    w2, ws2, d2, i2 = rm.composite_rays_train(sig, c2, t, rays)
    w3, ws3, d3, i3 = rm.composite_rays_train(sig, c1 + c2, t, rays)
    assert torch.equal(w1, w2) and torch.equal(ws1, ws3)
    torch.testing.assert_close(i3, i1 + i2, rtol=1e-4, atol=1e-5)              # linear in rgb
    assert float(ws1.max()) <= 1 + 1e-5 and float(ws1.min()) >= 0
    # weights_sum == segment sum of weights
    seg = torch.zeros(N, device='cuda', dtype=torch.float64)
    ray_id = torch.repeat_interleave(torch.arange(N, device='cuda'), rays[:, 1].long())
    order = torch.argsort(rays[:, 0].long() + (rays[:, 1] == 0).long() * (2 ** 40), stable=True)
    ray_id = torch.repeat_interleave(order, rays[order, 1].long())
    seg.index_add_(0, ray_id, w1.double())
    torch.testing.assert_close(seg.float(), ws1, rtol=1e-4, atol=1e-5)
    # backward: grad_rgbs = grad_image * weight exactly
    c1g = c1.clone().requires_grad_(True)
    _, _, _, ig = rm.composite_rays_train(sig, c1g, t, rays)
    ig.sum().backward()
    torch.testing.assert_close(c1g.grad, w1[:, None].expand(-1, 3), rtol=1e-6, atol=1e-7)

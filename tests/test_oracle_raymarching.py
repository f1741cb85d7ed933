"""CPU tests of the ray-marching oracle (oracle/raymarching_oracle.c).

The oracle restates /root/reference/lib/ops/raymarching/src/raymarching.cu; upstream has no golden
vectors, so here it is checked (a) against independent numpy / torch-fp64-autograd formulations of
the same maths and (b) -- in tests/test_golden_raymarching.py -- against outputs of the reference's
own CUDA kernels captured on a B200.
"""
import numpy as np
import pytest
import torch

from oracle import raymarching_oracle as orc
from tests import synth


def make_scene(n_views=2, size=16, H=32, radius=0.5, seed=0):
    grid = synth.sphere_density_grid(H=H, radius=radius)
    bitfield = orc.packbits(grid, 0.5)
    poses = synth.surround_poses(n_views, seed=seed)
    ro, rd, f = synth.camera_rays(poses, size)
    return grid, bitfield, ro, rd, f


def test_morton_roundtrip_and_numpy():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 128, (1000, 3)).astype(np.int32)
    idx = orc.morton3D(c)
    assert np.array_equal(idx, synth.morton3d_np(c[:, 0], c[:, 1], c[:, 2]).astype(np.int32))
    assert np.array_equal(orc.morton3D_invert(idx), c)
    assert idx.max() < 128 ** 3 and idx.min() >= 0


def test_packbits_matches_numpy_and_threshold_is_inclusive():
    rng = np.random.default_rng(1)
    g = rng.random(8 * 513).astype(np.float32)
    g[3] = 0.5  # >= is inclusive (raymarching.cu:283)
    assert np.array_equal(orc.packbits(g, 0.5), synth.pack_bitfield_np(g, 0.5))
    assert orc.packbits(g, 0.5)[0] & 8


def test_near_far_against_slab_formula():
    rng = np.random.default_rng(2)
    o = rng.normal(size=(500, 3)).astype(np.float32) * 3
    d = rng.normal(size=(500, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(o, d, aabb, 0.2)
    t0, t1 = (aabb[:3] - o) / d, (aabb[3:] - o) / d
    tn, tf = np.minimum(t0, t1).max(-1), np.maximum(t0, t1).min(-1)
    hit = tn <= tf
    miss = nears == np.finfo(np.float32).max
    assert np.array_equal(hit, ~miss)
    np.testing.assert_allclose(nears[hit], np.maximum(tn[hit], 0.2), rtol=1e-6)
    np.testing.assert_allclose(fars[hit], tf[hit], rtol=1e-6)
    assert np.all(fars[miss] == np.finfo(np.float32).max)


def test_march_train_samples_lie_in_occupied_cells_and_are_ordered():
    H = 32
    grid, bitfield, ro, rd, f = make_scene(H=H)
    nears, fars = orc.near_far_from_aabb(ro, rd, np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.2)
    rng = np.random.default_rng(3)
    noises = rng.random(ro.shape[0]).astype(np.float32)
    xyzs, dirs, ts, rays = orc.march_rays_train(ro, rd, 1.0, bitfield, 1, H, nears, fars, noises, dt_gamma=1 / f, max_steps=256)
    M = xyzs.shape[0]
    assert M > 0 and rays[:, 1].sum() == M
    assert np.array_equal(rays[:, 0], np.concatenate([[0], np.cumsum(rays[:, 1])[:-1]]))
    # every sample sits in an occupied voxel
    cell = np.clip((0.5 * (xyzs + 1) * H).astype(np.int64), 0, H - 1)
    assert np.all(grid[synth.morton3d_np(cell[:, 0], cell[:, 1], cell[:, 2])] > 0.5)
    # ts[:,0] strictly increasing inside a ray, dt within [dt_min, dt_max]
    dt_min, dt_max = 2 * np.sqrt(3) / 256, 2 * np.sqrt(3) / H
    assert np.all(ts[:, 1] >= np.float32(dt_min) * (1 - 1e-6)) and np.all(ts[:, 1] <= np.float32(dt_max) * (1 + 1e-6))
    for n in np.nonzero(rays[:, 1] > 1)[0][:200]:
        o_, c_ = rays[n]
        assert np.all(np.diff(ts[o_:o_ + c_, 0]) > 0)
        np.testing.assert_allclose(dirs[o_:o_ + c_], np.broadcast_to(rd[n], (c_, 3)))
        # sample position = o + (t_after - dt) d up to float rounding
        np.testing.assert_allclose(xyzs[o_:o_ + c_], ro[n] + (ts[o_:o_ + c_, :1] - ts[o_:o_ + c_, 1:]) * rd[n], atol=2e-6)
    # rays that miss the box have no samples
    assert np.all(rays[nears == np.finfo(np.float32).max, 1] == 0)


def test_march_train_max_steps_cap_and_empty_grid():
    H = 32
    grid, bitfield, ro, rd, f = make_scene(H=H)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    _, _, _, rays = orc.march_rays_train(ro, rd, 1.0, bitfield, 1, H, nears, fars, None, dt_gamma=0.0, max_steps=4)
    assert rays[:, 1].max() == 4
    xyzs, _, _, rays = orc.march_rays_train(ro, rd, 1.0, np.zeros_like(bitfield), 1, H, nears, fars, None, max_steps=64)
    assert xyzs.shape[0] == 0 and np.all(rays[:, 1] == 0)


def _torch_composite(sigmas, rgbs, ts, rays, T_thresh):
    """Independent fp64 formulation with autograd (no early-out masking other than the reference's break rule)."""
    N = rays.shape[0]
    ws_all, d_all, img_all = [], [], []
    w_full = torch.zeros_like(sigmas)
    for n in range(N):
        o, c = int(rays[n, 0]), int(rays[n, 1])
        if c == 0:
            ws_all.append(sigmas.new_zeros(())); d_all.append(sigmas.new_zeros(())); img_all.append(sigmas.new_zeros(3))
            continue
        s, col, t = sigmas[o:o + c], rgbs[o:o + c], ts[o:o + c]
        alpha = 1 - torch.exp(-s * t[:, 1])
        T_after = torch.cumprod(1 - alpha, 0)
        T_before = torch.cat([T_after.new_ones(1), T_after[:-1]])
        stop = torch.nonzero(T_after < T_thresh)
        last = int(stop[0]) if len(stop) else c - 1
        keep = (torch.arange(c) <= last).to(s.dtype)
        w = alpha * T_before * keep
        w_full = w_full + torch.nn.functional.pad(w, (o, sigmas.shape[0] - o - c))
        ws_all.append(w.sum()); d_all.append((w / t[:, 0]).sum()); img_all.append((w[:, None] * col).sum(0))
    return w_full, torch.stack(ws_all), torch.stack(d_all), torch.stack(img_all)


@pytest.mark.parametrize('T_thresh', [1e-4, 0.3])
def test_composite_train_forward_backward_vs_fp64_autograd(T_thresh):
    rng = np.random.default_rng(4)
    counts = np.array([0, 1, 5, 33, 64, 70, 2, 0, 17], np.int32)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays = np.stack([offs, counts], -1)
    M = int(counts.sum())
    sigmas = np.exp(rng.normal(size=M) * 1.5 + 1).astype(np.float32)
    rgbs = rng.random((M, 3)).astype(np.float32)
    ts = np.stack([2 + np.sort(rng.random(M)), 0.01 + 0.02 * rng.random(M)], -1).astype(np.float32)
    w, ws, d, img = orc.composite_rays_train_forward(sigmas, rgbs, ts, rays, T_thresh)
    s64 = torch.tensor(sigmas, dtype=torch.float64, requires_grad=True)
    c64 = torch.tensor(rgbs, dtype=torch.float64, requires_grad=True)
    tw, tws, td, timg = _torch_composite(s64, c64, torch.tensor(ts, dtype=torch.float64), rays, T_thresh)
    np.testing.assert_allclose(w, tw.detach().numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(ws, tws.detach().numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(d, td.detach().numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(img, timg.detach().numpy(), rtol=2e-5, atol=1e-7)
    if T_thresh > 1e-3:
        assert (w == 0).sum() > counts.size  # early termination really happened
    gw, gws, gd, gi = (rng.normal(size=s).astype(np.float32) for s in [(M,), (len(counts),), (len(counts),), (len(counts), 3)])
    # Reference quirk kept by the oracle: the grad_weights term is (gws + gw_i) * (T_i - sum_{j>i} w_j)
    # (raymarching.cu:676), i.e. it applies gw_i -- not gw_j -- to the later samples.  That equals the true
    # gradient only when grad_weights is constant along a ray, so the autograd comparison uses such a gw.
    gw = np.repeat(rng.normal(size=len(counts)).astype(np.float32), counts)
    gs, gc = orc.composite_rays_train_backward(gw, gws, gd, gi, sigmas, rgbs, ts, rays, ws, d, img, T_thresh)
    loss = (tw * torch.tensor(gw, dtype=torch.float64)).sum() + (tws * torch.tensor(gws, dtype=torch.float64)).sum() + \
        (td * torch.tensor(gd, dtype=torch.float64)).sum() + (timg * torch.tensor(gi, dtype=torch.float64)).sum()
    loss.backward()
    # grad_rgbs = grad_image * weight (raymarching.cu:668-670) -- identical to autograd.
    np.testing.assert_allclose(gc, c64.grad.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gs, s64.grad.numpy(), rtol=2e-3, atol=2e-5)


def test_composite_train_invalid_rays_are_zeroed():
    rays = np.array([[0, 4], [4, 10]], np.int32)  # second ray overflows M=8
    M = 8
    sig, rgb = np.ones(M, np.float32), np.ones((M, 3), np.float32)
    ts = np.stack([np.arange(1, M + 1), np.full(M, 0.1)], -1).astype(np.float32)
    w, ws, d, img = orc.composite_rays_train_forward(sig, rgb, ts, rays)
    assert ws[0] > 0 and ws[1] == 0 and d[1] == 0 and np.all(img[1] == 0) and np.all(w[4:] == 0)


def test_inference_loop_equals_train_composite_when_not_truncated():
    """march_rays/composite_rays (chunked, in place) must reproduce the train path on the same ray when no
    early termination happens (T_thresh=0): same sample sequence, same sums."""
    H = 32
    grid, bitfield, ro, rd, f = make_scene(n_views=1, size=12, H=H)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    N = ro.shape[0]
    xyzs, dirs, ts, rays = orc.march_rays_train(ro, rd, 1.0, bitfield, 1, H, nears, fars, None, dt_gamma=1 / f, max_steps=128)
    rng = np.random.default_rng(5)
    field = lambda x: (np.exp(np.sin(x * 7).sum(-1)).astype(np.float32), (0.5 + 0.5 * np.cos(x * 5)).astype(np.float32))
    sig, rgb = field(xyzs)
    _, ws_t, d_t, img_t = orc.composite_rays_train_forward(sig, rgb, ts, rays, T_thresh=0.0)
    ws, d, img = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    alive = np.arange(N, dtype=np.int32)
    rays_t = nears.copy()
    step = 0
    while step < 128 and alive.size:
        n_alive = alive.size
        n_step = min(max(N // n_alive, 1), 8)
        x, dd, t = orc.march_rays(n_alive, n_step, alive, rays_t, ro, rd, 1.0, bitfield, 1, H, nears, fars, None, dt_gamma=1 / f, max_steps=128)
        s_, c_ = field(x)
        orc.composite_rays(n_alive, n_step, alive, rays_t, s_, c_, t, ws, d, img, T_thresh=-1.0)
        alive = np.ascontiguousarray(alive[alive >= 0])
        step += n_step
    full = rays[:, 1] < 128
    np.testing.assert_allclose(ws[full], ws_t[full], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(d[full], d_t[full], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(img[full], img_t[full], rtol=1e-4, atol=1e-6)

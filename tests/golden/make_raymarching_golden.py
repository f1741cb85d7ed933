"""Generate tests/golden/raymarching_ref_small.npz FROM THE REFERENCE'S OWN CUDA KERNELS.

Run on a GPU box (gpurun) where oracle/_ref/_raymarching_ref.so (built by oracle/build_ref.py from
/root/reference/lib/ops/raymarching/src) is present:

    python tests/golden/make_raymarching_golden.py gpurun_out/raymarching_ref_small.npz

then copy the file into tests/golden/.  Inputs are seeded (tests/synth.py); outputs are whatever the reference
kernels produce on sm_100a.  tests/test_golden_raymarching.py pins the CPU oracle to these vectors.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from tests import synth  # noqa: E402


def main(out):
    ref = build_ref.load_ref()
    assert ref is not None, 'oracle/_ref not built'
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    H, max_steps = 32, 128
    grid = synth.sphere_density_grid(H=H, radius=0.55)
    poses = synth.surround_poses(3, seed=11)
    ro, rd, f = synth.camera_rays(poses, 20)
    N = ro.shape[0]
    rng = np.random.default_rng(11)
    noises = rng.random(N).astype(np.float32)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    g = {}
    # packbits / morton
    gt = cu(grid)
    bf = torch.empty(H ** 3 // 8, dtype=torch.uint8, device='cuda')
    ref.packbits(gt, H ** 3 // 8, 0.5, bf)
    coords = rng.integers(0, 128, (257, 3)).astype(np.int32)
    idx = torch.empty(257, dtype=torch.int32, device='cuda')
    ref.morton3D(cu(coords), 257, idx)
    inv = torch.empty(257, 3, dtype=torch.int32, device='cuda')
    ref.morton3D_invert(idx, 257, inv)
    # near far
    rot, rdt = cu(ro), cu(rd)
    nears, fars = torch.empty(N, device='cuda'), torch.empty(N, device='cuda')
    ref.near_far_from_aabb(rot, rdt, cu(aabb), N, 0.2, nears, fars)
    # march train
    counter = torch.zeros(1, dtype=torch.int32, device='cuda')
    rays = torch.empty(N, 2, dtype=torch.int32, device='cuda')
    nz = cu(noises)
    ref.march_rays_train(rot, rdt, bf, 1.0, False, 1 / f, max_steps, N, 1, H, nears, fars, None, None, None, rays, counter, nz)
    M = int(counter.item())
    x, d, t = (torch.zeros(M, k, device='cuda') for k in (3, 3, 2))
    ref.march_rays_train(rot, rdt, bf, 1.0, False, 1 / f, max_steps, N, 1, H, nears, fars, x, d, t, rays, counter, nz)
    # composite
    sig = np.exp(rng.normal(size=M) * 1.5 + 1.5).astype(np.float32)
    rgb = rng.random((M, 3)).astype(np.float32)
    w, ws, dep, img = torch.zeros(M, device='cuda'), torch.empty(N, device='cuda'), torch.empty(N, device='cuda'), torch.empty(N, 3, device='cuda')
    ref.composite_rays_train_forward(cu(sig), cu(rgb), t, rays, M, N, 1e-4, False, w, ws, dep, img)
    gw, gws, gd, gi = (rng.normal(size=s).astype(np.float32) for s in [(M,), (N,), (N,), (N, 3)])
    gs, gc = torch.zeros(M, device='cuda'), torch.zeros(M, 3, device='cuda')
    ref.composite_rays_train_backward(cu(gw), cu(gws), cu(gd), cu(gi), cu(sig), cu(rgb), t, rays, ws, dep, img, M, N, 1e-4, False, gs, gc)
    # one inference round: 4 steps for all rays, then composite
    n_step = 4
    alive = torch.arange(N, dtype=torch.int32, device='cuda')
    rt = nears.clone()
    xi, di, ti = (torch.zeros(N * n_step, k, device='cuda') for k in (3, 3, 2))
    ref.march_rays(N, n_step, alive, rt, rot, rdt, 1.0, False, 1 / f, max_steps, 1, H, bf, nears, fars, xi, di, ti, torch.zeros(N, device='cuda'))
    sig_i = np.exp(rng.normal(size=N * n_step) * 1.5 + 2.5).astype(np.float32)
    rgb_i = rng.random((N * n_step, 3)).astype(np.float32)
    ws_i, d_i, img_i = torch.zeros(N, device='cuda'), torch.zeros(N, device='cuda'), torch.zeros(N, 3, device='cuda')
    ref.composite_rays(N, n_step, 1e-2, False, alive, rt, cu(sig_i), cu(rgb_i), ti, ws_i, d_i, img_i)
    c = lambda v: v.cpu().numpy()
    np.savez_compressed(
        out, H=H, max_steps=max_steps, f=f, grid=grid, ro=ro, rd=rd, noises=noises, aabb=aabb, bitfield=c(bf), coords=coords, morton=c(idx),
        morton_inv=c(inv), nears=c(nears), fars=c(fars), rays=c(rays), xyzs=c(x), dirs=c(d), ts=c(t), sigmas=sig, rgbs=rgb, weights=c(w),
        weights_sum=c(ws), depth=c(dep), image=c(img), gw=gw, gws=gws, gd=gd, gi=gi, grad_sigmas=c(gs), grad_rgbs=c(gc),
        inf_xyzs=c(xi), inf_ts=c(ti), inf_sig=sig_i, inf_rgb=rgb_i, inf_alive=c(alive), inf_rays_t=c(rt), inf_ws=c(ws_i), inf_depth=c(d_i),
        inf_image=c(img_i), device=torch.cuda.get_device_name(0))
    print('wrote', out, 'M =', M)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/raymarching_ref_small.npz')

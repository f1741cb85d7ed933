"""Data-parallel ``mesh_optim`` (views of every iteration split across ranks, one all-reduce of the gradients) on CPU: two gloo ranks
must end with the same SDF / deformation / field as one process running all the views (the reference is single-GPU).
The rasteriser runs through tests/host_harness.py, the field is an analytic stand-in."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from types import SimpleNamespace
    from tests import host_harness, synth_mesh
    from tests.test_mesh_stage_host import ToyField
    from mvedit_b200 import mesh_raster as dr
    from mvedit_b200 import mesh_optim as mopt
    from mvedit_b200.mesh_renderer import DMTet, Mesh, MeshRenderer, make_tet_grid
    from mvedit_b200.nerf import L1LossMod
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(1234 + rank)                       # ranks draw differently: the shared draws must come from rank 0
    n, size, steps = 4, 32, 3
    poses = torch.from_numpy(synth_mesh.surround_poses(n, 2)).float()
    intr = torch.from_numpy(synth_mesh.intrinsics(size)).float()[None].expand(n, -1).contiguous()
    lights = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(4)), dim=-1)
    cam_weights = torch.tensor([1.0, 0.5, 1.0, 2.0])
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    tgt_masks = disc[None, None, :, :, None].expand(1, n, -1, -1, -1).contiguous()
    tgt_images = (torch.rand(1, n, size, size, 3, generator=torch.Generator().manual_seed(5)) * 0.5 + 0.25) * tgt_masks + (1 - tgt_masks)
    # the single-process run is handed the draws; the 2-rank run draws them itself on rank 0 with the same generator state
    noise = dict(camera_perm=torch.tensor([2, 0, 3, 1]), jitter=torch.rand(steps, 4, 2, generator=torch.Generator().manual_seed(6)))
    field = ToyField()
    nerf = SimpleNamespace(decoder=field, pixel_loss=L1LossMod(loss_weight=1.2), patch_loss=None, data_parallel=world > 1)
    with host_harness.routed(dr):
        tet_verts, tet_indices, tet_sdf = mopt.init_tet(nerf, None, density_thresh=5.0, tets=make_tet_grid(12))
        deform = torch.zeros_like(tet_verts).requires_grad_(True)
        tet_sdf.requires_grad_(True)
        opt = torch.optim.Adam([{'params': list(field.parameters())}, {'params': [tet_sdf, deform], 'lr': 1e-3}], lr=0.01)
        dm = DMTet('cpu')
        with torch.enable_grad():
            mv, mf = dm(tet_verts + deform, tet_sdf, tet_indices)
            mesh = Mesh(v=mv, f=mf.int())
            mesh.auto_normal()
        pipe = SimpleNamespace(nerf=nerf, mesh_renderer=MeshRenderer(near=0.01, far=100), normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
        mesh = mopt.mesh_optim(pipe, tgt_images, tgt_masks, None, opt, 0.01, 0.8, steps, 4, 8, 24, 0.0, 0.0, 0.02, 0.1, 5.0, None,
                               tet_verts, deform, tet_sdf, tet_indices, dm, mesh, size, intr, size, poses, cam_weights, lights, 16,
                               False, 0.2, 1.0, noise=noise)
    q.put((rank, tet_sdf.detach().numpy().copy(), deform.detach().numpy().copy(), field.w.detach().numpy().copy(), mesh.f.numpy().copy()))   # by value
    if world > 1:
        dist.destroy_process_group()


def _launch(world, port):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_run, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_two_ranks_match_one_process():
    single = _launch(1, 29611)[0]
    r0, r1 = _launch(2, 29613)
    import numpy as np
    for a, b in zip(r0[1:], r1[1:]):
        assert np.array_equal(a, b)                                # replicas stay bit-identical (same all-reduced gradient, same Adam step)
    assert all(np.abs(r0[k] - single[k]).max() < 2e-5 for k in (1, 2, 3))
    assert np.array_equal(r0[4], single[4])
    assert np.abs(single[2]).max() > 1e-4                           # and the geometry did move


def test_shared_draws_come_from_rank0():
    """Without supplied noise the camera permutation / jitter are rank 0's on every rank (ranks are seeded differently above)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_shared_worker, args=(r, 2, 29617, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert out[0][1] == out[1][1] and out[0][2] == out[1][2] and sorted(out[0][1]) == list(range(16))


def _shared_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mvedit_b200 import mesh_optim as mopt
    torch.manual_seed(rank)
    q.put((rank, mopt._from_rank0(torch.randperm(16)).tolist(), mopt._from_rank0(torch.rand(3, 2)).tolist()))
    dist.destroy_process_group()

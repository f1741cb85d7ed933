"""N>1 path on CPU: world_size-2 gloo processes exercise the view partition, the single packed all_gather of decoded targets, the
collectives of the data-parallel reconstruction (per-ray output gather, flat gradient all-reduce, patch-order broadcast) and the
flat field broadcast of the replicated mode."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_views, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mvedit_b200 import view_shard
    lo, hi = view_shard.local_range(n_views)
    # images + masks, bf16-exact values (the pack is bf16): small integers
    full = (torch.arange(n_views * 2 * 3 * 3, dtype=torch.float32) % 251).reshape(n_views, 2, 3, 3)
    msk = (torch.arange(n_views * 2 * 3, dtype=torch.float32) % 2).reshape(n_views, 2, 3, 1)
    got, gotm = view_shard.gather_views(full[lo:hi].clone(), msk[lo:hi].clone(), n_views)
    ok = torch.equal(got, full) and torch.equal(gotm, msk)
    # data-parallel reconstruction: row strips of P patches -> full patches in patch-major order
    P, ps, C = 3, 4, 5
    rows = ps // world
    patches = torch.arange(P * ps * ps * C, dtype=torch.float32).reshape(P, ps, ps, C)
    mine = patches[:, rank * rows:(rank + 1) * rows].reshape(-1, C)
    ok = ok and torch.equal(view_shard.gather_rays(mine, P), patches.reshape(-1, C))
    ok = ok and torch.equal(view_shard.gather_rays(patches[0, rank * rows:(rank + 1) * rows].reshape(-1, C), 1), patches[0].reshape(-1, C))
    flat = torch.full((7,), float(rank + 1))
    view_shard.allreduce_flat(flat)
    ok = ok and bool((flat == sum(range(1, world + 1))).all())
    torch.manual_seed(100 + rank)
    batches = torch.randperm(12)[None].split(4, dim=1)
    b2 = view_shard.broadcast_patch_order(batches)
    allb = [torch.zeros(1, 12, dtype=torch.long) for _ in range(world)]
    dist.all_gather(allb, torch.cat(list(b2), dim=1))
    ok = ok and all(torch.equal(a, allb[0]) for a in allb) and len(b2) == 3 and b2[0].shape == (1, 4)
    lin = torch.nn.Linear(4, 4)
    torch.manual_seed(rank)
    with torch.no_grad():
        lin.weight.normal_()
    grid = torch.full((8,), float(rank))
    bits = torch.full((8,), rank, dtype=torch.uint8)
    view_shard.broadcast_field(lin, grid, bits)
    ref = [torch.zeros_like(lin.weight) for _ in range(world)]
    dist.all_gather(ref, lin.weight.data)
    ok = ok and all(torch.equal(r, ref[0]) for r in ref) and float(grid[0]) == 0.0 and int(bits[0]) == 0
    q.put((rank, lo, hi, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize('n_views', [32, 9, 3])
def test_two_rank_gloo(n_views):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n_views
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[3] for r in res), res
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_views      # contiguous cover of all views
    assert abs((res[0][2] - res[0][1]) - (res[1][2] - res[1][1])) <= 1


def test_partition_properties():
    from mvedit_b200 import view_shard
    for n in (1, 7, 32, 37):
        for g in (1, 2, 4, 8):
            rs = [view_shard.local_range(n, r, g) for r in range(g)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_scheduler_matches_closed_form():
    from mvedit_b200.pipeline import EulerAncestralScheduler
    s = EulerAncestralScheduler()
    s.set_timesteps(24)
    assert s.timesteps[0] == 999 and len(s.timesteps) == 24 and s.sigmas[-1] == 0
    assert abs(s.init_noise_sigma - 14.6146) < 1e-2                      # SD1.5 sigma_max
    a, b = s.noise_scales(s.timesteps[3])
    assert abs(float(a) ** 2 + float(b) ** 2 - 1) < 1e-5
    x = torch.randn(2, 4, 8, 8)
    eps = torch.randn(2, 4, 8, 8)
    # with zero ancestral noise on the last step (sigma_to = 0) the update lands on pred_x0
    out = s.step(eps, 23, x, torch.zeros_like(x))
    torch.testing.assert_close(out, x - s.sigmas[23] * eps)

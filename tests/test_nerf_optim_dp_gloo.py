"""Data-parallel ``nerf_optim`` (the rays of every drawn patch split into row strips across ranks, per-ray outputs gathered, objective
replicated on the full patches, ONE gradient all-reduce per iteration) on the CPU: two gloo ranks must follow the
trajectory of one process -- same losses per iteration, same field afterwards.  The kernels are the product's sources compiled unchanged
as C++ (tests/host_shim: ``mve_patch_rays`` with its row-strip arguments, the objective kernels); the field is the analytic ToyField of
tests/test_nerf_optim_host.py.  Cases: the text-to-3D terms with a patch loss, and the image-to-3D targets (normals, depths, normal
patch term)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    from tests import host_harness
    from tests.test_nerf_optim_host import ToyField, ToyNeRF, scene, CASES, ITERS, N_RAYS, RS, PS
    from mvedit_b200 import nerf as pnerf
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(77 + rank)                         # ranks draw differently: nothing may depend on a rank's own stream
    c = CASES[case]
    poses, intr, images, masks, normals, depths, cam_w, cam_lights, batches = scene()
    dec = ToyField()
    nerf = ToyNeRF(dec, batches if rank == 0 else tuple(b.flip(1) for b in batches))      # the patch order must come from rank 0
    nerf.data_parallel = world > 1
    grid, bits = torch.zeros(1, 32 ** 3, dtype=torch.float16), torch.full((1, 32 ** 3 // 8), 255, dtype=torch.uint8)
    libs = host_harness.Libraries(host_harness.shimmed('recon.cu'), host_harness.shimmed('nerf_loss.cu'))
    with host_harness.routed(pnerf, libs):
        log = pnerf.nerf_optim(nerf, images, masks, normals if c.get('normals') else None, torch.optim.Adam(dec.parameters(), lr=0.01), 0.02, ITERS,
                               N_RAYS, c['patch_rgb'], c.get('patch_normal', 0.0), 0.02, 0.1, 0.01, [None], grid, bits, RS, intr, RS, poses, cam_w,
                               cam_lights, PS, c['is_init'], 0.015, 0.2, 1.0, c['init_shaded'], debug=True,
                               tgt_depths=depths if c.get('depths') else None, depth_weight=c.get('depth_weight', 0.0))
    q.put((rank, [l['loss'] for l in log], {k: v.detach().numpy().copy() for k, v in dec.named_parameters()}))
    if world > 1:
        dist.destroy_process_group()


def _launch(world, case, port):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_run, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    return res


@pytest.mark.parametrize('case', ['shaded', 'all'])
def test_two_ranks_follow_the_single_process_trajectory(case):
    import numpy as np
    port = 29700 + (os.getpid() % 500) + (7 if case == 'all' else 0)
    single = _launch(1, case, port)[0]
    two = _launch(2, case, port + 1)
    for r in two:
        np.testing.assert_allclose(r[1], single[1], rtol=2e-5, atol=1e-7)                 # the losses of every iteration
        for k, v in single[2].items():
            np.testing.assert_allclose(r[2][k], v, rtol=0, atol=2e-5, err_msg=k)          # the field after the Adam steps
    for k in single[2]:
        np.testing.assert_array_equal(two[0][2][k], two[1][2][k])                         # the replicas stay bit-identical

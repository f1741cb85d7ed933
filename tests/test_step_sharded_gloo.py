"""``MVEdit3DStep.step`` under view sharding (every rank denoises / decodes / renders its slice of the views, the decoded
targets are exchanged by ONE packed all_gather, the reconstruction sees all views) on two gloo ranks, CPU: each rank's new latents and
conditions are the single-process ones of its views, and every rank hands the reconstruction the full target set (bf16-packed).
Toy components as in tests/test_pipeline_loop_pins.py (where ``step`` is held to the reference-pinned loop)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import importlib.util
    from oracle import nerf_oracle as no
    from mvedit_b200 import mvedit_3d_pipeline as P, pipeline as S, view_shard
    from mvedit_b200.schedulers import EulerAncestralScheduler
    spec = importlib.util.spec_from_file_location('make_pipeline_loop_pins', os.path.join(ROOT, 'tests', 'golden', 'make_pipeline_loop_pins.py'))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    field, log = gen.ToyField(), []
    field.patch_size = 64
    S.nerf_optim = lambda nerf, *a, **k: gen.record_call(log, field, *a, **k)
    sch = EulerAncestralScheduler()
    sch.set_timesteps(6)
    pipe = P.MVEdit3DPipeline(gen.ToyVAE(), None, None, gen.ToyUNet(), gen.mixin_gen.toy_nets(2), sch, field, segmentation=gen.toy_segmentation)

    class W:
        bg_color = field.bg_color

        def render(self, bitfield, h, w, intrinsics, poses, cfg=None, normal_bg=(0.5, 0.5, 1.0)):
            return field.render(None, None, bitfield, h, w, intrinsics, poses, cfg=cfg, normal_bg=normal_bg)
    pipe.render_views = lambda bitfield, poses, intr, intr_size, rs, cam_lights, ambient, tdg, render_bs=None, **kw: no.render_views(
        W(), bitfield, poses, intr, intr_size, rs, cam_lights, ambient, tdg, render_bs=render_bs, out_dtype=torch.float32)
    poses, intr, _, embeds = gen.inputs()
    n = gen.N
    g = torch.Generator().manual_seed(5)
    latents, noise = torch.randn(n, 4, 64, 64, generator=g) * 3, torch.randn(n, 4, 64, 64, generator=g)
    lights = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    lo, hi = view_shard.local_range(n)
    pe = torch.cat([embeds[:n][lo:hi], embeds[n:][lo:hi]], dim=0)
    grid, bits = torch.zeros(1, 8 ** 3, dtype=torch.float16), torch.zeros(1, 8 ** 3 // 8, dtype=torch.uint8)
    new, ci, cd = pipe.step(2, latents[lo:hi].clone(), pe, None, grid, bits, None, poses, intr[None].expand(n, -1).contiguous(), gen.IMG,
                            torch.linspace(1, 2, n), lights, noise[lo:hi].clone(), guidance_scale=5.0, render_size=128, n_inverse_steps=3,
                            n_inverse_rays=4096, render_bs=2)
    q.put((rank, lo, hi, new.numpy().copy(), ci.float().numpy().copy(), log[0]['tgt_images'].numpy().copy(), log[0]['tgt_masks'].numpy().copy()))
    if world > 1:
        dist.destroy_process_group()


def _launch(world, port):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_run, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    return res


def test_two_ranks_step_their_views_like_one_process():
    import numpy as np
    port = 29900 + (os.getpid() % 90)
    single = _launch(1, port)[0]
    two = _launch(2, port + 1)
    assert [(r[1], r[2]) for r in two] == [(0, 3), (3, 5)] or two[0][2] == two[1][1]
    for r in two:
        lo, hi = r[1], r[2]
        np.testing.assert_allclose(r[3], single[3][lo:hi], rtol=1e-4, atol=1e-4)          # new latents of this rank's views
        np.testing.assert_allclose(r[4], single[4][lo:hi], rtol=0, atol=1e-2)             # rendered conditions (bf16)
        assert r[5].shape == single[5].shape                                              # the reconstruction sees ALL views ...
        np.testing.assert_allclose(r[5], single[5], rtol=0, atol=8e-3)                    # ... through the bf16-packed exchange
        np.testing.assert_allclose(r[6], single[6], rtol=0, atol=8e-3)
    np.testing.assert_array_equal(two[0][5], two[1][5])                                   # every rank holds the same targets

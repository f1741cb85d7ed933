"""Generates tests/golden/nerf_loop_pins.npz by RUNNING the reference's own ``MVEdit3DPipeline.nerf_optim``
(lib/pipelines/mvedit_3d_pipeline.py:452-656), cut out by AST and executed unmodified in the build container, on the CPU.  What it
drives is this repo's oracle stack -- ``OracleNeRF`` / ``OracleDecoder`` (oracle/nerf_oracle.py) over the C restatement of the
ray-marching kernels (``CpuOps``; pinned bit-exactly against the reference's compiled kernels elsewhere) and the plain-torch hash grid --
because the reference's own BaseNeRF / iNGPDecoder need mmcv / mmgen / tinycudann.  The loop itself (targets, ray batches, per-iteration
occupancy refresh, every loss term, Adam) is the reference's code: the fixture pins ``oracle/nerf_oracle.nerf_optim`` -- the restatement
the GPU parity tests compare the product against -- to it.

Random draws (patch order, marching perturbation, occupancy jitter) are drawn once and fed to both sides through thin subclasses.
Stand-ins: ``TVLoss`` = the reference's ``tv_loss`` body under mmgen's mean reduction, ``L1LossMod`` / geometry helpers as pinned by
tests/golden/make_reference_pins.py, a weighted-MSE patch metric with the LPIPSLoss call shape.

Run:  python tests/golden/make_nerf_loop_pins.py      (CPU, ~1 min)
"""
import ast
import importlib.util
import math
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
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'nerf_loop_pins.npz')

V, RS, PS, GRID, ITERS, N_RAYS = 3, 32, 16, 32, 4, 2 * 16 * 16
DEC = dict(max_steps=64, weight_culling_th=0.001, base_resolution=4, max_resolution=24, n_levels=12)


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
    assert not (set(names) - set(found))
    return found


def scene(seed=0):
    """Cameras, targets (a shaded disc), draws and an initial field -- shared by the generator and the test."""
    from tests import synth
    g = torch.Generator().manual_seed(seed)
    poses = torch.from_numpy(synth.surround_poses(V, seed=0)).float()
    f = 0.5 * RS / math.tan(math.radians(15))
    intr = torch.tensor([[f, f, RS / 2, RS / 2]] * V)
    yy, xx = torch.meshgrid(torch.arange(RS), torch.arange(RS), indexing='ij')
    disc = (((xx - 15.5) ** 2 + (yy - 15.5) ** 2).float().sqrt() < 9).float()
    masks = disc[None, None, :, :, None].expand(1, V, -1, -1, -1).contiguous()
    images = (torch.rand(1, V, RS, RS, 3, generator=g) * 0.5 + 0.25) * masks + (1 - masks)
    n_patches = V * (RS // PS) ** 2
    draws = dict(raybatch=torch.randperm(n_patches, generator=g)[None], march=[torch.rand(N_RAYS, generator=g) for _ in range(ITERS)],
                 grid=[torch.rand(GRID ** 3, 3, generator=g) for _ in range(ITERS)])
    return poses, intr, images, masks, torch.tensor([1.0, 0.5, 2.0]), F.normalize(torch.randn(V, 3, generator=g), dim=-1), draws


def targets(seed=5):
    """Target normals (opengl, [0,1]; a dome over the disc) and target 1/z for the image-to-3D case."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(RS), torch.arange(RS), indexing='ij')
    nx, ny = (xx - 15.5) / 12.0, -(yy - 15.5) / 12.0
    n = F.normalize(torch.stack([nx, ny, (1 - nx ** 2 - ny ** 2).clamp(min=0.05).sqrt()], -1) + 0.05 * torch.randn(RS, RS, 3, generator=g), dim=-1)
    normals = (n / 2 + 0.5)[None, None].expand(1, V, -1, -1, -1).contiguous()
    depths = (0.25 + 0.1 * torch.rand(1, V, RS, RS, 1, generator=g))
    return normals, depths


def make_field(seed=1):
    from oracle import nerf_oracle as no
    dec = no.OracleDecoder(no.CpuOps(), **DEC)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        dec.encoder.params.copy_((torch.rand(dec.encoder.params.numel(), generator=g) - 0.5) * 2.0)
        for m in dec.mlp.net:
            nn.init.xavier_uniform_(m.weight, generator=g)
            m.bias.zero_()
    return dec


class WeightedMSE:
    """A patch metric with ``LPIPSLoss``'s call shape (NCHW in, per-patch weight)."""

    def __call__(self, pred, target, weight=None, avg_factor=None):
        per = (pred - target).square().flatten(1).mean(1)
        return (per if weight is None else per * weight).mean() * 1.2


def main():
    from oracle import nerf_oracle as no
    for stub in ('mcubes', 'skimage'):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    sys.modules['skimage'].morphology = types.ModuleType('morphology')
    spec = importlib.util.spec_from_file_location('ref_geometry_utils', os.path.join(REF, 'lib/core/utils/geometry_utils.py'))
    gu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gu)
    tv_body = extract('lib/models/losses/tv_loss.py', ['tv_loss'], dict(torch=torch))['tv_loss']

    class TVLoss(nn.Module):
        def __init__(self, dims=[-2, -1], power=1, loss_weight=1.0):
            super().__init__()
            self.dims, self.power, self.loss_weight = dims, power, loss_weight

        def forward(self, pred, target=None, weight=None, avg_factor=None):
            return tv_body(pred, target, self.dims, power=self.power, dense_weight=weight).mean() * self.loss_weight

    import torchvision.transforms.v2.functional as F_t
    env = dict(torch=torch, F=F, F_t=F_t, np=np, math=math, TVLoss=TVLoss, get_module_device=lambda m: 'cpu', get_ray_directions=gu.get_ray_directions,
               get_rays=gu.get_rays, depth_to_normal=gu.depth_to_normal)
    env['highpass'] = extract('lib/pipelines/utils.py', ['highpass'], dict(F_t=F_t))['highpass']
    ref_nerf_optim = extract('lib/pipelines/mvedit_3d_pipeline.py', ['nerf_optim'], env)['nerf_optim']

    out = {}
    for case in ('p1', 't1'):                    # p1: text-to-3D terms;  t1: + target normals (TV target, high-passed patch term) and depths
        poses, intr, images, masks, cam_w, cam_lights, draws = scene()
        dec = make_field()

        class FedDecoder(type(dec)):             # the reference loop calls decoder(...) / update_extra_state(...) without noise arguments
            def forward(self, *a, **k):
                return super().forward(*a, noises=self._march.pop(0), **k)

            def update_extra_state(self, code, dg, db, it, **kw):
                return super().update_extra_state(code, dg, db, it, noise=self._grid.pop(0), **kw)
        dec.__class__ = FedDecoder
        dec._march, dec._grid = list(draws['march']), list(draws['grid'])

        class FedNeRF(no.OracleNeRF):
            def get_raybatch_inds(self, cond_imgs, n_inverse_rays):
                b = draws['raybatch'].split(n_inverse_rays // (self.patch_size ** 2), dim=1)
                return b, len(b)
        nerf = FedNeRF(dec, grid_size=GRID, patch_size=PS, update_extra_interval=2)
        nerf.patch_loss = WeightedMSE()
        density_grid = torch.zeros(1, GRID ** 3, dtype=torch.float16)
        density_bitfield = torch.full((1, GRID ** 3 // 8), 255, dtype=torch.uint8)
        opt = torch.optim.Adam(dec.parameters(), lr=0.01)
        p0 = {k: v.detach().clone() for k, v in dec.state_dict().items()}
        self_ = types.SimpleNamespace(nerf=nerf, normal_bg=[0.5, 0.5, 1.0], tonemapping=None)
        if case == 'p1':
            ref_nerf_optim(self_, images, masks, None, opt, 0.01, ITERS, N_RAYS, 0.4, 0.0, 0.02, 0.1, 0.01, [None], density_grid, density_bitfield,
                           RS, intr, RS, poses, cam_w, cam_lights, PS, False, 0.015, 0.2, 1.0, False)
        else:
            normals, depths = targets()
            ref_nerf_optim(self_, images, masks, normals, opt, 0.01, ITERS, N_RAYS, 0.4, 0.7, 0.02, 0.1, 0.01, [None], density_grid,
                           density_bitfield, RS, intr, RS, poses, cam_w, cam_lights, PS, False, 0.015, 0.2, 1.0, False, tgt_depths=depths,
                           depth_weight=0.3)
        out.update({case + '_' + k: v.detach().numpy() for k, v in dec.state_dict().items()})   # (the initial field is make_field(): not stored)
        out.update({'grid1' if case == 'p1' else 'grid_t1': density_grid.numpy(), 'bits1' if case == 'p1' else 'bits_t1': density_bitfield.numpy()})
        moved = max(float((dec.state_dict()[k] - p0[k]).abs().max()) for k in p0)
        print(case, 'max parameter change', moved, 'occupied bytes', int((density_bitfield != 0).sum()))
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main()

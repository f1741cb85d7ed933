"""Generates tests/golden/volume_renderer_pins.npz by RUNNING the reference's own ``VolumeRenderer``
(lib/models/decoders/base_volume_renderer.py:17-343: ``update_extra_state``, the training branch of ``forward`` with weight culling, the
inference while-loop), cut out by AST and executed unmodified in the build container on the CPU.  Its native ops (``lib.ops``:
``march_rays_train``, ``batch_composite_rays_train``, ``march_rays``, ``composite_rays``, ``morton3D``, ``packbits``,
``batch_near_far_from_aabb``) are served, under the reference's own wrapper signatures (lib/ops/raymarching/raymarching.py:70-524), by
the C restatement of the kernels (oracle ``CpuOps``; itself pinned bit-exactly against the reference's compiled kernels), and the field
(``point_decode``; tinycudann in the reference) by the plain-torch hash grid.  The fixture pins ``oracle/nerf_oracle.OracleDecoder`` --
what the GPU parity tests compare ``iNGPDecoder.forward`` / ``update_extra_state`` / the fused renderer against -- to the reference's
control flow: culling, re-indexing of rays, compaction steps, EMA / threshold / packbits of the occupancy refresh.

Run:  python tests/golden/make_volume_renderer_pins.py      (CPU, seconds)
"""
import ast
import importlib.util
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
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'volume_renderer_pins.npz')
GRID, SIZE, VIEWS = 32, 16, 2
DEC = dict(max_steps=64, weight_culling_th=0.001, base_resolution=4, max_resolution=24, n_levels=12)


def inputs():
    """Rays of two views, a sphere-shaped occupancy bitfield, noises, a random field -- shared by the generator and the test."""
    import math
    from tests import synth
    from oracle import nerf_oracle as no, raymarching_oracle as orc
    g = torch.Generator().manual_seed(3)
    poses = torch.from_numpy(synth.surround_poses(VIEWS, seed=1)).float()
    f = 0.5 * SIZE / math.tan(math.radians(15))
    intr = torch.tensor([[f, f, SIZE / 2, SIZE / 2]] * VIEWS)
    d = no.get_ray_directions(SIZE, SIZE, intr[None])
    ro, rd = no.get_rays(d, poses[None], norm=True)
    bitfield = torch.from_numpy(orc.packbits(synth.sphere_density_grid(H=GRID, radius=0.6), 0.5))[None].clone()
    dec = no.OracleDecoder(no.CpuOps(), **DEC)
    with torch.no_grad():
        dec.encoder.params.copy_((torch.rand(dec.encoder.params.numel(), generator=g) - 0.5) * 2.0)
        for m in dec.mlp.net:
            nn.init.xavier_uniform_(m.weight, generator=g)
            m.bias.zero_()
    n = VIEWS * SIZE * SIZE
    return (ro.reshape(1, n, 3).contiguous(), rd.reshape(1, n, 3).contiguous(), bitfield, dec,
            dict(march=torch.rand(n, generator=g), grid=torch.rand(GRID ** 3, 3, generator=g)), float(1.0 / f))


def render_cameras():
    import math
    from tests import synth
    poses = torch.from_numpy(synth.surround_poses(VIEWS, seed=5)).float()
    f = 0.5 * SIZE / math.tan(math.radians(15))
    return poses, torch.tensor([[f, 1.1 * f, SIZE / 2 - 0.5, SIZE / 2 + 0.25]] * VIEWS)


def main():
    from oracle import nerf_oracle as no
    tree = ast.parse(open(os.path.join(REF, 'lib/models/decoders/base_volume_renderer.py')).read())
    node = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'VolumeRenderer'][0]
    ro, rd, bitfield, od, noise, dt_gamma = inputs()
    ops = od.ops
    queue = dict(march=[noise['march']], grid=[noise['grid']])

    class TorchFed:                              # ``torch`` inside the class: rand_like returns the supplied occupancy jitter
        def __getattr__(self, k):
            return getattr(torch, k)

        def rand_like(self, t, **kw):
            return queue['grid'].pop(0).to(t)

    def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1,
                         force_all_rays=False, dt_gamma=0, max_steps=1024):
        return ops.march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, perturb=perturb, dt_gamma=dt_gamma, max_steps=max_steps,
                                    noises=queue['march'].pop(0) if perturb else None)

    def batch_composite_rays_train(sigmas, rgbs, ts, rays, num_points, T_thresh=1e-4, binarize=False):
        assert len(ts) == 1
        w, ws, d, img = ops.composite_rays_train(sigmas, rgbs, ts[0], rays[0], T_thresh, binarize)
        return w, ws[None], d[None], img[None]

    def batch_near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        n, f = ops.near_far_from_aabb(rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), aabb, min_near)
        return n.reshape(rays_o.shape[:2]), f.reshape(rays_o.shape[:2])

    env = dict(torch=TorchFed(), nn=nn, F=F, build_module=None, get_module_device=lambda m: torch.device('cpu'),
               custom_meshgrid=lambda *a: torch.meshgrid(*a, indexing='ij'), march_rays_train=march_rays_train,
               batch_composite_rays_train=batch_composite_rays_train, batch_near_far_from_aabb=batch_near_far_from_aabb,
               march_rays=lambda *a, **k: ops.march_rays(*a, **k), composite_rays=lambda *a, **k: ops.composite_rays(*a, **k),
               morton3D=ops.morton3D, morton3D_invert=None,
               packbits=lambda grid, thresh, bits: ops.packbits(grid[0], float(thresh), bits[0]))
    mod = ast.Module(body=[node], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, 'base_volume_renderer.py', 'exec'), env)
    VolumeRenderer = env['VolumeRenderer']

    class RefDecoder(VolumeRenderer):            # the subclass role of iNGPDecoder: only the field evaluation is supplied
        def __init__(self):
            super().__init__(bound=1, min_near=0.2, max_steps=DEC['max_steps'], weight_culling_th=DEC['weight_culling_th'])
            self.dummy = nn.Parameter(torch.zeros(1))

        def point_decode(self, xyzs, dirs, code, use_2nd_order=False):
            return od.point_decode(xyzs, dirs, code)

        def point_density_decode(self, xyzs, code):
            return od.point_density_decode(xyzs, code)

    ref = RefDecoder()
    out = {}
    # occupancy refresh from a clean grid (base_volume_renderer.py:105-177)
    grid = torch.zeros(1, GRID ** 3, dtype=torch.float16)
    bits = torch.zeros(1, GRID ** 3 // 8, dtype=torch.uint8)
    ref.update_extra_state(None, grid, bits, 0, density_thresh=0.1)
    out.update(ue_grid=grid.numpy().copy(), ue_bits=bits.numpy().copy())
    # training forward with culling (:179-262) and inference forward (:264-329)
    ref.train(True)
    with torch.no_grad():
        r = ref(ro, rd, None, bitfield, GRID, dt_gamma=dt_gamma, perturb=True)
    out.update(tr_weights=r['weights'].numpy(), tr_weights_sum=r['weights_sum'][0].numpy(), tr_depth=r['depth'][0].numpy(), tr_image=r['image'][0].numpy(),
               tr_rays=r['rays'][0].numpy(), tr_ts=r['ts'][0].numpy())
    ref.train(False)
    with torch.no_grad():
        e = ref(ro, rd, [None], bitfield, GRID, dt_gamma=torch.tensor([dt_gamma]), perturb=False)
    out.update(ev_weights_sum=e['weights_sum'][0].numpy(), ev_depth=e['depth'][0].numpy(), ev_image=e['image'][0].numpy())
    # BaseNeRF.render (lib/models/autoencoders/base_nerf.py:489-556) on the oracle decoder: rays, dt_gamma, rgba / 1/z depth, normals from depth
    for stub in ('mcubes', 'skimage'):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    sys.modules['skimage'].morphology = types.ModuleType('morphology')
    spec = importlib.util.spec_from_file_location('ref_geometry_utils', os.path.join(REF, 'lib/core/utils/geometry_utils.py'))
    gu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gu)
    btree = ast.parse(open(os.path.join(REF, 'lib/models/autoencoders/base_nerf.py')).read())
    rnode = [n for n in ast.walk(btree) if isinstance(n, ast.FunctionDef) and n.name == 'render'][0]
    rnode.decorator_list = []
    benv = dict(torch=torch, get_ray_directions=gu.get_ray_directions, get_rays=gu.get_rays, depth_to_normal=gu.depth_to_normal)
    rmod = ast.Module(body=[rnode], type_ignores=[])
    ast.fix_missing_locations(rmod)
    exec(compile(rmod, 'base_nerf.py', 'exec'), benv)
    poses, intr = render_cameras()
    with torch.no_grad():
        rgba, depth, normal, normal_fg = benv['render'](types.SimpleNamespace(bg_color=1.0, grid_size=GRID), od, None, bitfield, SIZE, SIZE, intr[None],
                                                        poses[None], cfg=dict(return_rgba=True, compute_normal=True, dt_gamma_scale=0.5))
    out.update(rn_rgba=rgba.numpy(), rn_depth=depth.numpy(), rn_normal=normal.numpy(), rn_normal_fg=normal_fg.numpy())
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, 'samples after culling', r['weights'].shape[0], 'occupied cells', int((grid > 0.01).sum()), 'alpha mean', float(e['weights_sum'][0].mean()))


if __name__ == '__main__':
    main()

"""CUDA-event timing of the mesh rasteriser ops at the mesh stage's size (render_bs = 8 views x 512^2, a DMTet-sized closed mesh):
python tests/gpu_profile_mesh.py [out.json].  Algorithmic bytes per DESIGN.md §3 (z-buffer 8 B + rast 16 B + rast_db 16 B per pixel, ...)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import synth_mesh                      # noqa: E402
from mvedit_b200 import mesh_raster as dr         # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B, S = 8, 512
    out = {}
    for sub in (6, 7):
        v, f = synth_mesh.icosphere(sub)
        pos = torch.from_numpy(synth_mesh.project(v * 0.75, synth_mesh.surround_poses(B, 0), fov_deg=30.0).astype(np.float32)).cuda().requires_grad_(True)
        tri = torch.from_numpy(f.astype(np.int32)).cuda()
        ctx = dr.RasterizeCudaContext()
        opp = dr.edge_opposites(tri)
        attr = torch.randn(1, v.shape[0], 3, device='cuda', requires_grad=True)
        rast, db = dr.rasterize(ctx, pos, tri, (S, S))
        cover = float((rast[..., 3] > 0).float().mean())
        color = torch.rand(B, S, S, 8, device='cuda', requires_grad=True)
        g4, g3, g8 = torch.randn(B, S, S, 4, device='cuda'), torch.randn(B, S, S, 3, device='cuda'), torch.randn(B, S, S, 8, device='cuda')
        r = dict(triangles=int(f.shape[0]), views=B, size=S, coverage=cover)
        r['rasterize_ms'] = timed(lambda: dr.rasterize(ctx, pos.detach(), tri, (S, S)))
        r['rasterize_bwd_ms'] = timed(lambda: torch.autograd.grad(dr.rasterize(ctx, pos, tri, (S, S))[0], pos, g4)) - r['rasterize_ms']
        r['interpolate_ms'] = timed(lambda: dr.interpolate(attr.detach(), rast.detach(), tri))
        out_i = dr.interpolate(attr, rast.detach(), tri)[0]
        r['interpolate_bwd_ms'] = timed(lambda: torch.autograd.grad(out_i, attr, g3, retain_graph=True))
        r['edge_opposites_ms'] = timed(lambda: dr.edge_opposites(tri), reps=5)
        r['antialias_ms'] = timed(lambda: dr.antialias(color.detach(), rast.detach(), pos.detach(), tri, topology_hash=opp))
        out_a = dr.antialias(color, rast.detach(), pos, tri, topology_hash=opp)
        r['antialias_bwd_ms'] = timed(lambda: torch.autograd.grad(out_a, (color, pos), g8, retain_graph=True))
        npx = B * S * S
        r['rasterize_gbs'] = npx * (8 + 8 + 16 + 16) / r['rasterize_ms'] / 1e6          # memset + read of the z-buffer, rast, rast_db
        r['interpolate_gbs'] = npx * (16 + 12) / r['interpolate_ms'] / 1e6
        r['antialias_gbs'] = npx * (32 * 3 + 16) / r['antialias_ms'] / 1e6               # copy (r + w) + colour read, rast read
        out['icosphere_%d' % sub] = r
    print(json.dumps(out))
    if len(sys.argv) > 1:
        os.makedirs(os.path.dirname(os.path.abspath(sys.argv[1])), exist_ok=True)
        json.dump(out, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()

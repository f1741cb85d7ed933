"""Generates tests/golden/pipeline_input_pins.npz by RUNNING the reference's own ``MVEdit3DPipeline.enable_normals`` and ``load_depths``
(lib/pipelines/mvedit_3d_pipeline.py:232-306), cut out by AST and executed unmodified on the CPU, with the reference's own ``Tonemapping``
module (lib/models/decoders/tonemapping.py, loaded by path) and toy stand-ins for the two networks they call (``normal_model``: the
omnidata DPT, ``image_enhancer``: SRVGG -- deterministic functions defined here and shared with the test).

Run:  python tests/golden/make_pipeline_input_pins.py      (CPU, seconds)
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'pipeline_input_pins.npz')
N, S = 4, 64


def toy_normal_model(x):
    """images (n,3,384,384) -> 'predicted' opencv normals in [0,1]."""
    gx = x.mean(dim=1, keepdim=True)
    return torch.cat([torch.sigmoid(4 * (gx - 0.5)), x[:, 1:2] * 0.8 + 0.1, 1 - 0.3 * x[:, 2:3]], dim=1)


def toy_enhancer(x):
    """(n,3,h,w) -> (n,3,2h,2w), slightly out of [0,1] like a real super-resolution net."""
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) * 1.04 - 0.02


def inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing='ij')
    disc = (((xx - 31.5) ** 2 + (yy - 31.5) ** 2).float().sqrt() < 22).float()
    masks = (disc[None, None] * (0.6 + 0.4 * torch.rand(N, 1, S, S, generator=g))).contiguous()
    masks[:, :, 20:30, 20:30] = 1.0
    images = torch.rand(N, 3, S, S, generator=g) * 0.9 + 0.05
    lights = F.normalize(torch.randn(N, 3, generator=g), dim=-1)
    normals = [None,
               (torch.rand(S, S, 3, generator=g) * 255).to(torch.uint8).numpy(),           # at the working size
               (torch.rand(S // 2, S // 2, 3, generator=g) * 255).to(torch.uint8).numpy(),  # smaller: enhanced, then resized
               None]
    depths = [torch.rand(40, 40, generator=g).numpy(), torch.rand(S, S, 1, generator=g), torch.rand(96, 96, generator=g), torch.rand(S, S, generator=g).numpy()]
    return images, masks, lights, normals, depths


def main():
    tree = ast.parse(open(os.path.join(REF, 'lib/pipelines/mvedit_3d_pipeline.py')).read())
    env = dict(torch=torch, F=F, np=np, get_module_device=lambda m: 'cpu')
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ('enable_normals', 'load_depths'):
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, 'mvedit_3d_pipeline.py', 'exec'), env)
    spec = importlib.util.spec_from_file_location('ref_tonemapping', os.path.join(REF, 'lib/models/decoders/tonemapping.py'))
    tm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tm)
    images, masks, lights, normals, depths = inputs()
    out = {}
    for name, tone in (('plain', None), ('tone', tm.Tonemapping())):
        self_ = types.SimpleNamespace(unet=types.SimpleNamespace(dtype=torch.float32), nerf=None, normal_model=toy_normal_model,
                                      image_enhancer=toy_enhancer, tonemapping=tone, normal_bg=[0.5, 0.5, 1.0])
        im, nm = env['enable_normals'](self_, images.clone(), masks, lights, 0.2, normals=normals)
        out[name + '_images'], out[name + '_normals'] = im.numpy(), nm.numpy()
    self_ = types.SimpleNamespace(unet=types.SimpleNamespace(dtype=torch.float32), nerf=None)
    out['depths'] = env['load_depths'](self_, depths, diff_size=S).numpy()
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()

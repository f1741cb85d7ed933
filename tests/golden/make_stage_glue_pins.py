"""Generates tests/golden/stage_glue_pins.npz by RUNNING two more of the reference's own functions unmodified on the CPU:
``init_tet`` (lib/pipelines/utils.py:156-184; its tet grid is read from a temporary ``demo/tets/12_tets.npz`` written from
``make_tet_grid(12)``, and its hard-wired ``device='cuda'`` is mapped to the CPU) and ``Adapter3DMixin.load_init_mesh``
(lib/pipelines/adapter3d_mixin.py:21-66) around a recording toy renderer.

Run:  python tests/golden/make_stage_glue_pins.py      (CPU, seconds)
"""
import ast
import os
import sys
import tempfile
import types
from copy import copy

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stage_glue_pins.npz')


class BlobDensity(nn.Module):
    """decoder.point_density_decode of an off-centre ellipsoid (so that the fitted box is neither centred nor cubic)."""

    def __init__(self):
        super().__init__()
        self.p = nn.Parameter(torch.zeros(1))

    def point_density_decode(self, xyzs, code, **kw):
        x = xyzs[0]
        r = ((x - x.new_tensor([0.1, -0.05, 0.15])) / x.new_tensor([0.45, 0.3, 0.35])).norm(dim=-1)
        return 30 * (1 - r), [len(x)]


class RecordingRenderer:
    def __init__(self):
        self.ssaa, self.calls = 1, []

    def __call__(self, meshes, poses, intrinsics, h, w, shading_fun=None, **kw):
        self.calls.append(dict(ssaa=self.ssaa, poses=poses.clone(), intrinsics=intrinsics.clone(), h=h, w=w, fun=shading_fun))
        n = poses.shape[1]
        g = torch.Generator().manual_seed(len(self.calls))
        rgba = torch.rand(1, n, h, w, 4, generator=g) * 1.2 - 0.1
        return dict(rgba=rgba, depth=torch.rand(1, n, h, w, generator=g))


class ToyMesh:
    def detach(self):
        return self

    def to(self, device):
        return self


def glue_inputs():
    from tests import synth
    poses = torch.from_numpy(synth.surround_poses(5, seed=4)).float()
    return poses, torch.tensor([[70.0, 72.0, 16.0, 15.0]] * 5) * torch.linspace(1, 1.2, 5)[:, None]


def decode_inputs():
    from oracle import field_oracle as fo
    levels, n = fo.level_table(12, 16, 320)
    g = torch.Generator().manual_seed(5)
    xyz = torch.rand(4096, 3, generator=g) * 2 - 1
    xyz[:64] *= 0.2                                          # inside the density blob's clamp radius
    params = fo.init_params(levels, n, seed=6, table_scale=0.5)
    return xyz, (params[0], params[1], params[2] + 0.1, params[3], params[4] - 0.2), levels


def triplane_inputs():
    from oracle import field_oracle as fo
    levels, n = fo.level_table(12, 16, 320)
    g = torch.Generator().manual_seed(7)
    r = lambda *sh: torch.randn(*sh, generator=g)
    C, H = 6, 32
    sd = {'encoder.params': (torch.rand(n * 2, generator=g) - 0.5), 'base_net.0.weight': r(H, 3 * C) * 0.3, 'base_net.0.bias': r(H) * 0.1,
          'ingp_base_net.0.weight': r(H, 24) * 0.3, 'ingp_base_net.0.bias': r(H) * 0.1, 'density_net.0.weight': r(1, H) * 0.3,
          'density_net.0.bias': r(1) * 0.1, 'color_net.0.weight': r(3, H) * 0.3, 'color_net.0.bias': r(3) * 0.1}
    return torch.rand(2048, 3, generator=g) * 2.2 - 1.1, r(1, 3, C, 20, 20), sd, levels      # some points beyond the planes: border padding


def shading_inputs():
    g = torch.Generator().manual_seed(11)
    fg = torch.rand(1, 3, 16, 16, generator=g) > 0.4
    n = int(fg.sum())
    lights = torch.nn.functional.normalize(torch.randn(3, 16, 16, 3, generator=g), dim=-1)
    return lights, torch.rand(n, 3, generator=g) * 0.9 + 0.05, torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1), fg


def init_nerf_inputs():
    import importlib.util as iu
    sp = iu.spec_from_file_location('make_pipeline_loop_pins', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'make_pipeline_loop_pins.py'))
    L = iu.module_from_spec(sp)
    sp.loader.exec_module(L)
    poses, intr = glue_inputs()
    g = torch.Generator().manual_seed(13)
    lw = torch.nn.functional.normalize(torch.randn(5, 3, generator=g), dim=-1)
    return L.ToyField(), poses, intr * 2, lw, (lw[:, None, :] @ poses[:, :3, :3]).squeeze(-2)


def extract(rel, name, env):
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    node = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name)
    node.decorator_list = []
    mod = ast.Module(body=[node], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, rel, 'exec'), env)
    return env[name]


def main():
    from mvedit_b200.mesh_renderer import make_tet_grid
    out = {}
    # ---- init_tet
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, 'demo', 'tets'))
        os.makedirs(os.path.join(tmp, 'lib', 'pipelines'))
        grid = make_tet_grid(12)
        np.savez(os.path.join(tmp, 'demo', 'tets', '12_tets.npz'), vertices=grid['vertices'].numpy(), indices=grid['indices'].numpy())

        class TorchOnCpu:
            def __getattr__(self, k):
                return getattr(torch, k)

            def tensor(self, *a, **k):
                k.pop('device', None)
                return torch.tensor(*a, **k)
        fn = extract('lib/pipelines/utils.py', 'init_tet', dict(torch=TorchOnCpu(), np=np, os=os, hf_hub_download=None,
                                                              __file__=os.path.join(tmp, 'lib', 'pipelines', 'utils.py')))
        verts, indices, sdf = fn(types.SimpleNamespace(decoder=BlobDensity()), None, density_thresh=5.0, resolution=12)
    out.update(tet_verts=verts.numpy(), tet_indices=indices.numpy(), tet_sdf=sdf.numpy())
    # ---- load_init_mesh
    poses, intr = glue_inputs()
    rend = RecordingRenderer()
    self_ = types.SimpleNamespace(unet=types.SimpleNamespace(device='cpu'), mesh_renderer=rend, bg_color=0.7)
    fn = extract('lib/pipelines/adapter3d_mixin.py', 'load_init_mesh', dict(torch=torch, copy=copy, Mesh=None))
    funs = ['f0', 'f1', 'f2']
    mesh = ToyMesh()
    m, images, alphas, depths = fn(self_, mesh, poses, intr, 32, 2, funs, diff_size=48)
    assert m is mesh and rend.ssaa == 1                     # the 2x supersampling is set on a COPY of the renderer
    out.update(lim_images=images.numpy(), lim_alphas=alphas.numpy(), lim_depths=depths.numpy(), lim_ssaa=np.array([c['ssaa'] for c in rend.calls]),
               lim_intr=torch.cat([c['intrinsics'][0] for c in rend.calls]).numpy(), lim_sizes=np.array([[c['h'], c['w']] for c in rend.calls]),
               lim_funs=np.array([funs.index(c['fun']) for c in rend.calls]))
    rend2 = RecordingRenderer()
    self_.mesh_renderer = rend2
    fn(self_, mesh, poses, intr, 32, 4, None)
    out.update(lim_default_sizes=np.array([[c['h'], c['w']] for c in rend2.calls]), lim_default_fun_none=np.array([c['fun'] is None for c in rend2.calls]))
    # ---- iNGPDecoder.point_decode / density_blob / MLP (ingp_decoder.py:20-40,101-120) and TruncExp (lib/ops/activation.py) around the
    # plain-torch hash grid of oracle/field_oracle.py in place of tinycudann's encoder (absent; "parity unpinned" for the grid itself)
    import importlib.util
    import torch.nn.functional as F
    from oracle import field_oracle as fo
    sp = importlib.util.spec_from_file_location('ref_activation', os.path.join(REF, 'lib/ops/activation.py'))
    act = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(act)
    tree = ast.parse(open(os.path.join(REF, 'lib/models/decoders/ingp_decoder.py')).read())
    env = dict(torch=torch, nn=nn, F=F)
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == 'MLP' or isinstance(node, ast.FunctionDef) and node.name in ('density_blob', 'point_decode', 'point_density_decode'):
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, 'ingp_decoder.py', 'exec'), env)
    xyz, (table, w1, b1, w2, b2), levels = decode_inputs()
    mlp = env['MLP'](w1.shape[1], 4, 64, 2)
    with torch.no_grad():
        mlp.net[0].weight.copy_(w1); mlp.net[0].bias.copy_(b1); mlp.net[1].weight.copy_(w2); mlp.net[1].bias.copy_(b2)
    dec = types.SimpleNamespace(encoder=lambda x01: fo.hash_encode(x01, table, levels), mlp=mlp, bound=1, sigma_activation=act.TruncExp(),
                                blob_density=1.0, blob_radius=0.2, sigmoid_saturation=0.001)
    for n in ('density_blob', 'point_decode', 'point_density_decode'):
        setattr(dec, n, types.MethodType(env[n], dec))
    with torch.no_grad():
        sig, rgb, num = dec.point_decode([xyz], None, [None])
        sig2, num2 = dec.point_density_decode([xyz], [None])
    assert num == [len(xyz)] and torch.equal(sig, sig2)
    out.update(dec_sigma=sig.numpy(), dec_rgb=rgb.numpy())
    # ---- TriPlaneiNGPDecoder.point_decode (triplane_ingp_decoder.py:142-212) + TriPlaneDecoder.xyz_transform (triplane_decoder.py:106-130):
    # the reference's code with Sequential(Linear[, activation]) nets as its constructor builds them (:59-92) around the oracle's hash grid
    t1 = ast.parse(open(os.path.join(REF, 'lib/models/decoders/triplane_ingp_decoder.py')).read())
    t2 = ast.parse(open(os.path.join(REF, 'lib/models/decoders/triplane_decoder.py')).read())
    tenv = dict(torch=torch, nn=nn, F=F)
    pd = next(n for n in ast.walk(t1) if isinstance(n, ast.FunctionDef) and n.name == 'point_decode')
    xt = next(n for n in ast.walk(t2) if isinstance(n, ast.FunctionDef) and n.name == 'xyz_transform')
    for node in (pd, xt):
        mod = ast.Module(body=[node], type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, 'triplane', 'exec'), tenv)
    xyz_t, code, sd, tl = triplane_inputs()
    lin = lambda k: nn.Linear(sd[k + '.0.weight'].shape[1], sd[k + '.0.weight'].shape[0])
    nets = {k: lin(k) for k in ('base_net', 'ingp_base_net', 'density_net', 'color_net')}
    with torch.no_grad():
        for k, m in nets.items():
            m.weight.copy_(sd[k + '.0.weight']); m.bias.copy_(sd[k + '.0.bias'])
    for cfg_name, plane_cfg, flip in (('a', ('xy', 'xz', 'yz'), False), ('b', ['yx', 'yz', 'xz'], True)):
        tp = types.SimpleNamespace(encoder=lambda x01: fo.hash_encode(x01, sd['encoder.params'].view(-1, 2), tl), bound=1, code_dropout=None,
                                   scene_base=None, interp_mode='bilinear', plane_cfg=plane_cfg, flip_z=flip, use_dir_enc=False,
                                   base_net=nn.Sequential(nets['base_net']), ingp_base_net=nn.Sequential(nets['ingp_base_net']),
                                   base_activation=nn.SiLU(), density_net=nn.Sequential(nets['density_net'], act.TruncExp()),
                                   color_net=nn.Sequential(nets['color_net'], nn.Sigmoid()), sigmoid_saturation=0.001)
        tp.xyz_transform = types.MethodType(tenv['xyz_transform'], tp)
        with torch.no_grad():
            s_, c_, _ = tenv['point_decode'](tp, [xyz_t], None, code)
        out['tri_sigma_' + cfg_name], out['tri_rgb_' + cfg_name] = s_.numpy(), c_.numpy()
    # ---- MVEdit3DPipeline.make_shading_fun (mvedit_3d_pipeline.py:410-423): Lambert shading of a mesh's own albedo, plain and tone-mapped
    sp2 = importlib.util.spec_from_file_location('ref_tonemapping', os.path.join(REF, 'lib/models/decoders/tonemapping.py'))
    tmm = importlib.util.module_from_spec(sp2)
    sp2.loader.exec_module(tmm)
    msf = extract('lib/pipelines/mvedit_3d_pipeline.py', 'make_shading_fun', dict(torch=torch))
    lights, albedo, normal, fg = shading_inputs()
    for name, tone in (('plain', None), ('tone', tmm.Tonemapping())):
        fun = msf(types.SimpleNamespace(tonemapping=tone), lights, 0.2)
        out['shade_' + name] = fun(world_pos=None, albedo=albedo, world_normal=normal, fg_mask=fg).numpy()
    # ---- MVEdit3DPipeline.load_init_nerf (mvedit_3d_pipeline.py:138-173): initial targets rendered from the field, shaded per view
    lin = extract('lib/pipelines/mvedit_3d_pipeline.py', 'load_init_nerf', dict(torch=torch))
    field, poses5, intr5, l_world, l_cam = init_nerf_inputs()
    for name, tone in (('plain', None), ('tone', tmm.Tonemapping())):
        im, al = lin(types.SimpleNamespace(nerf=field, tonemapping=tone, normal_bg=[0.5, 0.5, 1.0]), [None], None, poses5, intr5, 64, l_world, l_cam, 0.2,
                     2, 0.25, diff_size=48)
        out['lin_images_' + name], out['lin_alphas_' + name] = im.numpy(), al.numpy()
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()

"""Generates tests/golden/texture_loop_pins.npz by RUNNING the reference's own ``MVEditTexturePipeline.__call__``
(lib/pipelines/mvedit_texture_pipeline.py:173-544) and ``MVEditTextureSuperResPipeline.__call__``
(lib/pipelines/mvedit_texture_superres_pipeline.py) -- cut out by AST, executed unmodified on the CPU -- around the toy components of
make_pipeline_loop_pins.py, with the reference's own ``camera_dense_weighting``, input loaders, denoiser mixin, noise scales, pruning and
schedules.  Recorded: every ``bake_multiview`` / ``texture_optim`` / ``bake_xyz_shading_fun`` hand-over (targets, dense camera weights,
cameras, weights, sizes) -- the loop bodies of SURVEY §8 a-11's two pipelines.

Run:  python tests/golden/make_texture_loop_pins.py      (CPU, ~1 min)
"""
import importlib.util
import math
import os
import sys
import types
from copy import deepcopy

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, 'texture_loop_pins.npz')
spec = importlib.util.spec_from_file_location('make_pipeline_loop_pins', os.path.join(HERE, 'make_pipeline_loop_pins.py'))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)
N, S = 5, 256


class ToyMesh:
    albedo, textureless = None, True

    def detach(self):
        return self


class ToyTexRenderer:
    """MeshRenderer stand-in of the texture pipelines: renders depend on how many bakes the "texture" has seen; ``bake_multiview`` and
    ``bake_xyz_shading_fun`` record what they are handed."""
    ssaa = 1

    def __init__(self, log):
        self.log, self.field = log, L.ToyField()

    def __call__(self, meshes, poses, intrinsics, h, w, shading_fun=None, **kw):
        rgba, depth, normal, _ = self.field.render(None, None, None, h, w, intrinsics, poses)
        return dict(rgba=rgba, depth=depth, normal=normal)

    def bake_multiview(self, meshes, images, cam_weights, cam_poses, intrinsics, cos_weight_pow=4.0, render_bs=8, **kw):
        self.log.append(dict(kind=0.0, maps=images.detach().float().clone(), weights=cam_weights[0].detach().float().clone(),
                             camera_poses=cam_poses[0].clone(), intrinsics=intrinsics[0].clone(), cos_weight_pow=cos_weight_pow, render_bs=render_bs))
        self.field.fits += 1
        return meshes

    def bake_xyz_shading_fun(self, meshes, fun, **kw):
        self.log.append(dict(kind=2.0, **{k: float(v) for k, v in kw.items()}))
        out = ToyMesh()
        ms = int(kw.get('map_size', 64))
        t = torch.linspace(0, 1, ms)
        out.albedo = torch.stack([t[:, None].expand(ms, ms), t[None, :].expand(ms, ms), 0.5 + 0.4 * torch.sin(9 * t[:, None] * t[None, :]),
                                  torch.ones(ms, ms)], dim=-1)
        out.textureless = False
        return [out]

    def get_cam_weights_uv(self, meshes, cam_poses, intrinsics, map_size=1024, render_bs=8, cos_weight_pow=4.0, **kw):
        n = cam_poses.shape[1]
        self.log.append(dict(kind=3.0, camera_poses=cam_poses[0].clone(), intrinsics=intrinsics[0].clone(), map_size=map_size, render_bs=render_bs,
                             cos_weight_pow=cos_weight_pow))
        t = torch.linspace(0, 1, map_size)
        w = torch.stack([0.5 + 0.5 * torch.sin(3 * (k + 1) * t[:, None] + 2 * t[None, :]) for k in range(n)], dim=0)[None, ..., None] / n
        return w, (t[:, None] + t[None, :]) > 0.2


def record_texture_optim(log, tgt_images, optimizer, lr, inverse_steps, render_bs, patch_bs, patch_rgb_weight, nerf_code, in_mesh, render_size,
                         intrinsics, intrinsics_size, camera_poses, cam_weights_dense, patch_size, **kw):
    log.append(dict(kind=1.0, maps=tgt_images.detach().float().clone(), weights=cam_weights_dense.detach().float().clone(), lr=lr,
                    inverse_steps=inverse_steps, render_bs=render_bs, patch_bs=patch_bs, patch_rgb_weight=patch_rgb_weight, render_size=render_size,
                    intrinsics=intrinsics.clone(), intrinsics_size=intrinsics_size, camera_poses=camera_poses.clone(), patch_size=patch_size,
                    **{k: float(v) for k, v in kw.items() if isinstance(v, (int, float)) and not isinstance(v, bool)}))


def toy_load_init_mesh(renderer):
    def load_init_mesh(in_model, camera_poses, intrinsics, intrinsics_size, render_bs, shading_fun=None, diff_size=512):
        out = renderer([in_model], camera_poses[None], intrinsics[None] * (diff_size / intrinsics_size), diff_size, diff_size)
        rgba = out['rgba'][0]
        return in_model, (rgba[..., :3] + (1 - rgba[..., 3:])).clamp(0, 1), rgba[..., 3:], out['depth'][0]
    return load_init_mesh


def inputs():
    from tests import synth
    g = torch.Generator().manual_seed(3)
    poses = torch.from_numpy(synth.surround_poses(N + 2, seed=2)).float()
    f = 0.5 * S / math.tan(math.radians(15))
    return poses, torch.tensor([f, f, S / 2, S / 2]), torch.randn(2 * N, 77, L.D, generator=g)


CASES = dict(
    tex_optim_only=dict(pipe='texture', optim_only=True, num_inference_steps=3),
    tex_one_pass=dict(pipe='texture', mode='1-pass', use_reference=False),
    tex_two_pass_reference=dict(pipe='texture', mode='2-pass', use_reference=True, weighted_cam_pruning=True, cam_weights=[1.0, 2.0, 0.5, 1.0, 1.5]),
    tex_from_noise_reference=dict(pipe='texture', mode='1-pass', use_reference=True, denoising_strength=None, num_inference_steps=3),
    sr_plain=dict(pipe='superres', use_reference=False),
    sr_reference_reg=dict(pipe='superres', use_reference=True, reg=True),
    tex_ip_adapter=dict(pipe='texture', mode='1-pass', use_reference=True, ip=True, cond=True),
    sr_ip_adapter_some_cond=dict(pipe='superres', use_reference=False, ip=True, cond=True, ip_adapter_use_cond_idx=[0, 3]))


def call_kwargs(case, poses, intr, embeds=None):
    c = dict(CASES[case])
    kind = c.pop('pipe')
    kw = dict(prompt='a toy', in_model=ToyMesh(), camera_poses=poses[:N], intrinsics=intr, intrinsics_size=S, guidance_scale=5.0,
              num_inference_steps=6, denoising_strength=0.5, diff_size=S, patch_size=128, patch_bs=2, diff_bs=4, render_bs=2, n_inverse_steps=9,
              lr=0.02, bake_texture_kwargs=dict(map_size=64))
    if kind == 'texture':
        kw.update(keep_views=[1], max_num_views=lambda p: 5 if p < 0.5 else 3)
    else:
        kw['camera_poses'] = poses[:N, :3]                   # the reference's empty regulariser set is (0, 3, 4): 3 x 4 poses throughout
        if c.pop('reg', False):
            kw.update(reg_camera_poses=poses[N:, :3], reg_cam_weights=[0.5, 0.25])
    if c.pop('ip', False):
        kw['ip_adapter'] = L.ToyIPAdapter(embeds)
    if c.pop('cond', False):
        g = torch.Generator().manual_seed(12)
        kw['cond_images'] = [(torch.rand(96, 96, 3, generator=g) * 255).to(torch.uint8).numpy() for _ in range(N)]
    kw.update(c)
    if case == 'sr_reference_reg':                   # an input mesh that brings its own texture: the baked result is blended with it (:468-487)
        kw['in_model'].albedo, kw['in_model'].textureless = torch.rand(32, 32, 4, generator=torch.Generator().manual_seed(8)), False
    return kind, kw


def flatten(log, prefix):
    out = {prefix + 'calls': np.array(len(log))}
    for i, rec in enumerate(log):
        for k, v in rec.items():
            if k in ('maps', 'weights'):
                x = v.reshape(-1, *v.shape[-3:]).permute(0, 3, 1, 2)
                out['%s%d_%s_pooled' % (prefix, i, k)] = F.avg_pool2d(x, 16 if x.shape[-1] >= 128 else 4).numpy()
                out['%s%d_%s_std' % (prefix, i, k)] = x.flatten(2).std(dim=2).numpy()
                out['%s%d_%s_shape' % (prefix, i, k)] = np.array(v.shape)
            else:
                out['%s%d_%s' % (prefix, i, k)] = v.numpy() if torch.is_tensor(v) else np.array(float(v))
    return out


def main():
    import PIL
    import PIL.Image
    for stub in ('mcubes', 'skimage'):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    sys.modules['skimage'].morphology = types.ModuleType('morphology')
    gu, diff, rot = L.load('ref_gu', 'lib/core/utils/geometry_utils.py'), L.load('ref_diffusion', 'lib/core/diffusion.py'), L.load('ref_rot', 'lib/ops/rotation_conversions.py')
    U = L.extract('lib/pipelines/utils.py', ['get_camera_dists', 'prune_cameras', 'join_prompts'], dict(torch=torch, F=F, np=np, matrix_to_quaternion=rot.matrix_to_quaternion))

    class MultiControlNetModel:
        def __init__(self, nets):
            self.nets = list(nets)

        def __call__(self, sample, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=None, guess_mode=False,
                     added_cond_kwargs=None, return_dict=True):
            acc = None
            for net, cond, sc in zip(self.nets, controlnet_cond, conditioning_scale):
                acc = net(sample, t, encoder_hidden_states=encoder_hidden_states, controlnet_cond=cond, conditioning_scale=sc, accumulate=acc)
            return acc
    M = L.extract('lib/pipelines/adapter3d_mixin.py', ['get_noise_pred', 'get_noise_pred_p1', 'get_noise_pred_p2'],
                  dict(torch=torch, copy=__import__('copy').copy, MultiControlNetModel=MultiControlNetModel,
                       unet_enc=lambda unet, *a, **k: unet.enc(*a, **k), unet_dec=lambda unet, *a, **k: unet.dec(*a, **k)))

    class _Never:
        pass

    class _NumpyCompat:
        cumproduct = staticmethod(np.cumprod)

        def __getattr__(self, k):
            return getattr(np, k)
    tb = []

    def env():
        return dict(torch=torch, F=F, np=_NumpyCompat(), PIL=PIL, math=math, deepcopy=deepcopy, get_module_device=lambda m: 'cpu',
                    join_prompts=U['join_prompts'], get_camera_dists=U['get_camera_dists'], prune_cameras=U['prune_cameras'],
                    DPMSolverSDEScheduler=_Never, DPMSolverMultistepScheduler=_Never, get_noise_scales=diff.get_noise_scales,
                    normalize_depth=gu.normalize_depth, get_ray_directions=gu.get_ray_directions, depth_to_normal=gu.depth_to_normal, apply_cross_image_attn_proc=lambda u: None,
                    remove_cross_image_attn_proc=lambda u: None, tqdm=lambda x: x, edge_dilation=None,
                    traceback=types.SimpleNamespace(format_exc=lambda: tb.append(__import__('traceback').format_exc()) or tb[-1]))
    tenv, senv = env(), env()
    from mvedit_b200.mesh_renderer import edge_dilation
    senv['edge_dilation'] = edge_dilation                       # (pinned against lib/ops/edge_dilation.py by test_mesh_pins.py)
    T = L.extract('lib/pipelines/mvedit_texture_pipeline.py', ['default_patch_rgb_weight', 'default_max_num_views', 'camera_dense_weighting'], tenv)
    T.update(L.extract('lib/pipelines/mvedit_texture_pipeline.py', ['__call__'], tenv))
    senv['camera_dense_weighting'], senv['default_patch_rgb_weight'] = T['camera_dense_weighting'], T['default_patch_rgb_weight']      # (:25-26)
    Sx = L.extract('lib/pipelines/mvedit_texture_superres_pipeline.py', ['__call__'], senv)
    P3 = L.extract('lib/pipelines/mvedit_3d_pipeline.py', ['load_init_images', 'load_cond_images', 'get_prompt_embeds'], env())
    SP = L.extract('lib/pipelines/mvedit_texture_superres_pipeline.py', ['get_prompt_embeds'], env())
    poses, intr, embeds = inputs()
    out = {}
    for case in CASES:
        log = []
        renderer = ToyTexRenderer(log)
        self_ = types.SimpleNamespace(
            nerf=renderer.field, unet=L.ToyUNet(), controlnet=MultiControlNetModel(L.mixin_gen.toy_nets(2)), vae=L.ToyVAE(),
            scheduler=L.DiffusersShapedScheduler(), image_enhancer=L.ToyEnhancer(), segmentation=None, tonemapping=None, bg_color=0.5,
            normal_bg=[0.5, 0.5, 1.0], mesh_renderer=renderer, clip_img_size=224, clip_img_mean=[0.48145466, 0.4578275, 0.40821073],
            clip_img_std=[0.26862954, 0.26130258, 0.27577711])
        for n, fn in list(P3.items()) + list(M.items()):
            setattr(self_, n, types.MethodType(fn, self_))
        self_.load_init_mesh = toy_load_init_mesh(renderer)
        self_._encode_prompt = lambda *a, **k: embeds.clone()
        self_.make_nerf_albedo_shading_fun = lambda code: None
        kind, kw = call_kwargs(case, poses, intr, embeds)
        if kind == 'superres':                   # the super-resolution class overrides get_prompt_embeds (:62-87)
            self_.get_prompt_embeds = types.MethodType(SP['get_prompt_embeds'], self_)
        if kind == 'texture':
            self_.texture_optim = lambda *a, **k: record_texture_optim(log, *a, **k)
        else:       # the super-resolution variant's own texture_optim takes num_cameras second (:89-91); the product passes it as patch_views=
            self_.texture_optim = lambda tgt, num_cameras, *a, **k: record_texture_optim(log, tgt, *a, patch_views=num_cameras, **k)
        del tb[:]
        torch.manual_seed(4321)
        res = (T if kind == 'texture' else Sx)['__call__'](self_, prog_bar=lambda x: x, **kw)
        assert not tb, tb
        if kind == 'superres':
            log.append(dict(kind=4.0, maps=res.albedo[..., :3][None].clone()))       # the returned texture
        out.update(flatten(log, case + '_'))
        if kw.get('ip_adapter') is not None:
            assert len(kw['ip_adapter'].seen) == 1
            out[case + '_ipa_images'] = F.avg_pool2d(kw['ip_adapter'].seen[0], 16).numpy()
        print(case, [int(r['kind']) for r in log], [tuple(r['maps'].shape) for r in log if 'maps' in r][-1])
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, sum(v.nbytes for v in out.values()) // 1024, 'KiB')


if __name__ == '__main__':
    main()

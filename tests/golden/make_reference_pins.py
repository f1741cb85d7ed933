"""Generates tests/golden/reference_pins.npz by RUNNING the reference's own torch-only modules in this container
(they cannot travel to the GPU box: /root/reference does not exist there, so their outputs are committed as fixtures).

Loaded by path (``import lib`` itself is impossible here: mmcv / mmgen / diffusers / trimesh are absent):
  lib/ops/activation.py                    trunc_exp fwd / bwd                       -> pins oracle/field_oracle.TruncExpFn
  lib/models/architecture/joint_attn.py    CrossImageAttnProcWrapper                 -> pins oracle/unet_oracle.attention(num_cross_attn_imgs=2)
  lib/models/decoders/tonemapping.py       Tonemapping (lut / inverse_lut / smooth_forward) -> pins oracle/nerf_oracle.Tonemapping and
                                           mvedit_b200.tonemapping.Tonemapping
  lib/models/decoders/image_space_ss.py    SRVGGNetCompact (class cut out by AST)     -> pins oracle/enhancer_oracle.srvgg_forward
  lib/core/diffusion.py                    get_noise_scales                          -> pins oracle/nerf_oracle.get_noise_scales,
                                                                                        mvedit_b200.pipeline scheduler.noise_scales
  lib/core/utils/geometry_utils.py         get_ray_directions / get_rays / depth_to_normal / normalize_depth   (mcubes, skimage stubbed)
                                                                                     -> pins oracle/nerf_oracle geometry helpers
  lib/core/utils/camera_utils.py           random_surround_views(use_linspace) / light_sampling               -> pins tests/synth rig helpers
  lib/ops/rotation_conversions.py          matrix_to_quaternion (for get_camera_dists)
Functions cut out of files that import absent packages (AST extraction of the pure-torch function bodies, executed unmodified):
  lib/pipelines/utils.py                   get_camera_dists, prune_cameras, highpass
  lib/pipelines/mvedit_3d_pipeline.py      default_* schedules (:41-78)
  lib/models/autoencoders/base_nerf.py     BaseNeRF.ray_sample, BaseNeRF.get_raybatch_inds (:245-322)
  lib/models/losses/tv_loss.py             tv_loss (:8-42, the un-decorated function body; mmgen's @weighted_loss stripped)

Run:  python tests/golden/make_reference_pins.py      (CPU, seconds)
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_pins.npz')


def load_by_path(name, rel, stubs=()):
    for s in stubs:
        mod = sys.modules.setdefault(s.split(':')[0], types.ModuleType(s.split(':')[0]))
        for attr in s.split(':')[1:]:
            setattr(mod, attr, types.ModuleType(attr))
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def extract_functions(rel, names, env, strip_decorators=True):
    """exec the named top-level functions (or methods of top-level classes, as plain functions) of a reference file, unmodified."""
    src = open(os.path.join(REF, rel)).read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            if strip_decorators:
                node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, rel, 'exec'), env)
            found[node.name] = env[node.name]
    missing = set(names) - set(found)
    assert not missing, missing
    return found


def main():
    import torch.nn.functional as F
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)
    out = {}

    # ---- trunc_exp (activation.py:8-23)
    act = load_by_path('ref_activation', 'lib/ops/activation.py')
    x = torch.tensor([-20.0, -3.0, -0.5, 0.0, 0.7, 5.0, 12.0, 15.0, 30.0], requires_grad=True)
    y = act.trunc_exp(x)
    gy = torch.linspace(0.5, 1.5, x.numel())
    y.backward(gy)
    out.update(truncexp_x=x.detach().numpy(), truncexp_y=y.detach().numpy(), truncexp_gy=gy.numpy(), truncexp_gx=x.grad.numpy())

    # ---- CrossImageAttnProcWrapper (joint_attn.py:11-37) around a plain SDPA processor with fixed weights
    ja = load_by_path('ref_joint_attn', 'lib/models/architecture/joint_attn.py')
    B, S, C, T, Dc, heads = 4, 6, 16, 5, 8, 2
    W = {k: torch.randn(C, C if k != 'k2' and k != 'v2' else Dc, generator=g) / 4 for k in ('q', 'k', 'v', 'o', 'k2', 'v2')}
    bo = torch.randn(C, generator=g) * 0.1

    def base_proc(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        kv = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        wk, wv = (W['k'], W['v']) if encoder_hidden_states is None else (W['k2'], W['v2'])
        q, k, v = hidden_states @ W['q'].t(), kv @ wk.t(), kv @ wv.t()
        sh = lambda t: t.reshape(t.shape[0], t.shape[1], heads, C // heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(sh(q), sh(k), sh(v)).transpose(1, 2).reshape(hidden_states.shape)
        return o @ W['o'].t() + bo

    wrap = ja.CrossImageAttnProcWrapper(base_proc)
    hs = torch.randn(B, S, C, generator=g)
    ctx = torch.randn(B, T, Dc, generator=g)
    out.update(ja_hs=hs.numpy(), ja_ctx=ctx.numpy(), ja_bo=bo.numpy(), **{'ja_w_' + k: v.numpy() for k, v in W.items()},
               ja_self=wrap(None, hs, num_cross_attn_imgs=2).numpy(), ja_cross=wrap(None, hs, encoder_hidden_states=ctx, num_cross_attn_imgs=2).numpy(),
               ja_self_1=wrap(None, hs, num_cross_attn_imgs=1).numpy())

    # ---- IPAttnProcessor2_0 / CNAttnProcessor2_0 (ip_adapter/attention_processor.py:283-396, :466-556) on a stand-in Attention module
    ap = load_by_path('ref_ip_attn', 'lib/models/architecture/ip_adapter/attention_processor.py')
    torch.manual_seed(3)
    Ci, Di, Hh, Ti, Ni = 16, 8, 2, 9, 4            # hidden, cross dim, heads, context tokens (5 text + 4 image), image tokens
    lin = lambda i, o, b: torch.nn.Linear(i, o, bias=b)
    attn = types.SimpleNamespace(spatial_norm=None, group_norm=None, norm_cross=False, heads=Hh, residual_connection=False, rescale_output_factor=1.0,
                                 to_q=lin(Ci, Ci, False), to_k=lin(Di, Ci, False), to_v=lin(Di, Ci, False),
                                 to_out=[lin(Ci, Ci, True), torch.nn.Identity()])
    ipp = ap.IPAttnProcessor2_0(hidden_size=Ci, cross_attention_dim=Di, scale=0.7, num_tokens=Ni)
    cnp = ap.CNAttnProcessor2_0(num_tokens=2)
    hs_i, ctx_i = torch.randn(3, 6, Ci, generator=g), torch.randn(3, Ti, Di, generator=g)
    with torch.no_grad():
        out.update(ip_hs=hs_i.numpy(), ip_ctx=ctx_i.numpy(), ip_out=ipp(attn, hs_i, encoder_hidden_states=ctx_i).numpy(),
                   cn_out=cnp(attn, hs_i, encoder_hidden_states=ctx_i).numpy(),
                   ip_wq=attn.to_q.weight.numpy(), ip_wk=attn.to_k.weight.numpy(), ip_wv=attn.to_v.weight.numpy(),
                   ip_wo=attn.to_out[0].weight.numpy(), ip_bo=attn.to_out[0].bias.numpy(), ip_wk_ip=ipp.to_k_ip.weight.numpy(),
                   ip_wv_ip=ipp.to_v_ip.weight.numpy())

    # ---- get_noise_scales (diffusion.py:4-21) on SD1.5's scaled_linear schedule
    dif = load_by_path('ref_diffusion', 'lib/core/diffusion.py')
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
    ab = np.cumprod(1 - betas)
    t_int = torch.tensor([0, 1, 250, 500, 998, 999])
    t_flt = torch.tensor([0.0, 0.25, 41.625, 499.5, 957.375, 998.999, 999.0])
    out.update(ns_alphas_bar=ab, ns_t_int=t_int.numpy(), ns_t_flt=t_flt.numpy(),
               ns_int=np.stack([v.numpy() for v in dif.get_noise_scales(ab, t_int, 1000)]),
               ns_flt=np.stack([v.numpy() for v in dif.get_noise_scales(ab, t_flt, 1000)]))

    # ---- geometry (geometry_utils.py:18-55,119-168)
    geo = load_by_path('ref_geometry', 'lib/core/utils/geometry_utils.py', stubs=('mcubes', 'skimage:morphology'))
    cam = load_by_path('ref_camera', 'lib/core/utils/camera_utils.py')
    poses = cam.random_surround_views(3.7, 6, 0, 0.6, use_linspace=True)          # BASELINE configs[0] rig (draws torch.rand for elevation)
    h = w = 12
    K = torch.tensor([[40.0, 41.0, 6.2, 5.7]] * 6) * torch.linspace(1.0, 1.2, 6)[:, None]
    dirs = geo.get_ray_directions(h, w, K[None], norm=False)
    dirs_n = geo.get_ray_directions(h, w, K[None], norm=True)
    ro, rd = geo.get_rays(dirs, poses[None], norm=True)
    depth = (0.25 + 0.1 * torch.rand(1, 6, h, w, generator=g)) * (torch.rand(1, 6, h, w, generator=g) > 0.2)
    alphas = torch.rand(6, h, w, 1, generator=g)
    out.update(geo_poses=poses.numpy(), geo_K=K.numpy(), geo_dirs=dirs.numpy(), geo_dirs_n=dirs_n.numpy(), geo_ro=ro.numpy(), geo_rd=rd.numpy(),
               geo_depth=depth.numpy(), geo_normal_gl=geo.depth_to_normal(depth, dirs).numpy(),
               geo_normal_cv=geo.depth_to_normal(depth, dirs, format='opencv').numpy(), geo_alphas=alphas.numpy(),
               geo_depth_norm=geo.normalize_depth(depth[0] * alphas.squeeze(-1), alphas).numpy())
    torch.manual_seed(5)
    wl, cl = cam.light_sampling(poses)
    torch.manual_seed(5)
    out.update(light_world=wl.numpy(), light_cam=cl.numpy(), light_seed=np.array(5))

    # ---- camera pruning (pipelines/utils.py:350-379) and schedules (mvedit_3d_pipeline.py:41-78)
    rot = load_by_path('ref_rotconv', 'lib/ops/rotation_conversions.py')
    env = dict(torch=torch, matrix_to_quaternion=rot.matrix_to_quaternion, F=F, np=np)
    try:
        import torchvision.transforms.functional as F_t
        env['F_t'] = F_t
    except Exception:
        pass
    fn = extract_functions('lib/pipelines/utils.py', ['get_camera_dists', 'prune_cameras'], env)
    poses9 = cam.random_surround_views(3.0, 9, 0.1, 0.5, use_linspace=True)
    cw = torch.linspace(1.0, 2.0, 9)
    dists = fn['get_camera_dists'](poses9, cw, 'cpu')
    pix = torch.rand(9, generator=g)
    keep, d2 = fn['prune_cameras'](dists.clone(), 2, 5, 'cpu', pixel_dist=pix.clone())
    keep0, _ = fn['prune_cameras'](dists.clone(), 0, 4, 'cpu')
    out.update(prune_poses=poses9.numpy(), prune_cw=cw.numpy(), prune_dists=dists.numpy(), prune_pix=pix.numpy(), prune_keep=keep.numpy(),
               prune_dists_after=d2.numpy(), prune_keep_nopix=keep0.numpy())
    names = ['default_lr_multiplier', 'default_max_num_views', 'default_render_size_p', 'default_lr_schedule', 'default_patch_rgb_weight',
             'default_patch_normal_weight', 'default_entropy_weight', 'default_normal_reg_weight']
    sch = extract_functions('lib/pipelines/mvedit_3d_pipeline.py', names, {})
    ps = np.linspace(0, 1, 21)
    out['sched_p'] = ps
    for n in names:
        f = sch[n]
        out['sched_' + n] = np.array([f(float(p), 0.6) if n in ('default_lr_multiplier', 'default_max_num_views') else f(float(p)) for p in ps], np.float64)

    # ---- BaseNeRF.ray_sample / get_raybatch_inds (base_nerf.py:245-322), patch branch
    meth = extract_functions('lib/models/autoencoders/base_nerf.py', ['ray_sample', 'get_raybatch_inds'], dict(torch=torch))
    holder = types.SimpleNamespace(patch_size=4, patch_loss=object())
    V, hh = 3, 8
    cro, crd, cim = (torch.randn(1, V, hh, hh, 3, generator=g) for _ in range(3))
    ex = [torch.randn(1, V, hh, hh, 1, generator=g), torch.arange(V)[None, :, None, None, None].expand(1, V, hh, hh, 1).float()]
    inds = torch.tensor([[7, 0, 10]])
    rs = meth['ray_sample'](holder, cro, crd, cim, 3 * 16, sample_inds=inds, cond_extras=ex)
    out.update(rs_ro=cro.numpy(), rs_rd=crd.numpy(), rs_img=cim.numpy(), rs_ex0=ex[0].numpy(), rs_ex1=ex[1].numpy(), rs_inds=inds.numpy(),
               **{'rs_out%d' % i: t.numpy() for i, t in enumerate(rs)})
    torch.manual_seed(11)
    rb, nb = meth['get_raybatch_inds'](holder, cim, 2 * 16)
    out.update(rb_seed=np.array(11), rb_num=np.array(nb), rb_cat=torch.cat(list(rb), dim=1).numpy(), rb_len0=np.array(rb[0].shape[1]))

    # ---- tv_loss (tv_loss.py:8-42): the function body under mmgen's @weighted_loss
    tv = extract_functions('lib/models/losses/tv_loss.py', ['tv_loss'], dict(torch=torch))['tv_loss']
    pred, tgt = torch.rand(2, 3, 6, 6, generator=g), torch.rand(2, 3, 6, 6, generator=g)
    dw = torch.rand(2, 1, 6, 6, generator=g)
    out.update(tv_pred=pred.numpy(), tv_tgt=tgt.numpy(), tv_w=dw.numpy(), tv_plain=tv(pred, None, dims=[-2, -1], power=1.5).numpy(),
               tv_full=tv(pred, tgt, dims=[-2, -1], power=1.5, dense_weight=dw).numpy())

    # ---- Tonemapping (lib/models/decoders/tonemapping.py:5-52): the module itself (torch only)
    tmod = load_by_path('ref_tonemapping', 'lib/models/decoders/tonemapping.py')
    tm = tmod.Tonemapping()
    xs = torch.cat([torch.linspace(-11, 5, 97), tm.lut_x, tm.lut_x + 1e-4])            # beyond both ends, and on / next to the knots
    ys = torch.cat([torch.linspace(-0.1, 1.25, 83), tm.lut_y])
    alb, shd = torch.rand(5, 7, 3, generator=g), torch.rand(5, 7, 1, generator=g) * 0.8 + 0.2
    out.update(tm_lut_x=tm.lut_x.numpy(), tm_lut_y=tm.lut_y.numpy(), tm_xs=xs.numpy(), tm_ys=ys.numpy(),
               tm_lut=tm.lut(xs).numpy(), tm_lut_lin=tm.lut(torch.exp2(xs), input_mode='linear').numpy(),
               tm_inv=tm.inverse_lut(ys).numpy(), tm_inv_lin=tm.inverse_lut(ys, output_mode='linear').numpy(),
               tm_smooth=tm.smooth_forward(xs).numpy(), tm_alb=alb.numpy(), tm_shd=shd.numpy(),
               tm_shaded=tm.lut(tm.inverse_lut(alb) + shd.clamp(min=1e-6).log2()).numpy())       # mvedit_3d_pipeline.py:418-422

    # ---- SRVGGNetCompact (image_space_ss.py:8-75): the class cut out of its file (mmgen / mmcv imports are not needed by it)
    src = open(os.path.join(REF, 'lib/models/decoders/image_space_ss.py')).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'SRVGGNetCompact'][0]
    cls.decorator_list = []
    env = dict(nn=torch.nn, F=torch.nn.functional, torch=torch)
    exec(compile(ast.Module(body=[cls], type_ignores=[]), 'image_space_ss.py', 'exec'), env)
    torch.manual_seed(21)
    net = env['SRVGGNetCompact'](num_in_ch=3, num_out_ch=3, num_feat=8, num_conv=3, upscale=4, act_type='prelu').eval()
    with torch.no_grad():
        for m in net.body:
            if isinstance(m, torch.nn.PReLU):
                m.weight.uniform_(0.05, 0.4)
        xin = torch.rand(2, 3, 6, 5, generator=g)
        yout = net(xin)
    out.update(sr_x=xin.numpy(), sr_y=yout.numpy(), **{'sr_sd.' + k: v.numpy() for k, v in net.state_dict().items()})

    np.savez_compressed(OUT, **out)
    print('wrote', OUT, len(out), 'arrays', os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()

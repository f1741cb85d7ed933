"""Host-side mirror of the NeRF-adapter slice of the reference that sits on the hot path:

  * ``BaseNeRF``            /root/reference/lib/models/autoencoders/base_nerf.py:78-322,489-556  (ray_sample, get_raybatch_inds, render,
                            density grid / bitfield init) -- the mmgen registry / training-harness parts are out of scope
  * ``nerf_optim``          /root/reference/lib/pipelines/mvedit_3d_pipeline.py:452-656, the reconstruction inner loop

Same names / arguments / return values; torch is the plumbing, the arithmetic of march / field / composite / objective / render /
shading is libmvedit_b200.  The op-by-op torch restatements of the reference (geometry helpers, loss modules, the eager objective)
live in ``oracle/nerf_oracle.py`` as test infrastructure; nothing here falls back to them.

The objective covers L1 rgb / alpha, TV-normal, entropy, the LPIPS patch term (``nerf.patch_loss`` = mvedit_b200.lpips.LPIPSLoss) and
shading in tone-mapped space (``tonemapping``), and the optional targets of image-to-3D runs: target normals inside the TV term, the
L1 term on 1/z against target depths (both inside ``mve_nerf_patch_loss_targets``) and the high-passed normal patch term (LPIPS on
``mve_nerf_patch_out_normal``'s output; the 31-tap Gaussian of the high-pass is two depthwise torch convolutions).  There is no eager
fallback path.
"""
import collections
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ingp_decoder import iNGPDecoder
from . import view_shard
from .optim import FusedAdam
from .tonemapping import tone_args
from ._lib import call, ptr, stream, c_int, c_u32, c_f32


# ------------------------------------------------------------------------------------------------ small host helpers
def pixel_directions(intrinsics, h, w):
    """Camera-space direction (z = 1) through every pixel centre: [..., 4] (fx, fy, cx, cy) -> [..., h, w, 3]
    (what geometry_utils.get_ray_directions produces with norm=False; used for the static ``dirs`` slot of the recon program)."""
    dev = intrinsics.device
    fx, fy, cx, cy = (intrinsics[..., i, None, None] for i in range(4))
    u = (torch.arange(w, device=dev, dtype=torch.float32) + 0.5)[None, :]
    v = (torch.arange(h, device=dev, dtype=torch.float32) + 0.5)[:, None]
    shape = intrinsics.shape[:-1] + (h, w)
    return torch.stack([((u - cx) / fx).expand(shape), ((v - cy) / fy).expand(shape), torch.ones(shape, device=dev)], dim=-1)


def blur_masks(x, kernel_size, sigma):
    """Separable Gaussian blur with reflect padding on [V,1,h,w] (torchvision ``gaussian_blur`` semantics, which is what
    mvedit_3d_pipeline.py:473-476 calls): two 1-D depthwise passes."""
    r = kernel_size // 2
    t = torch.arange(kernel_size, device=x.device, dtype=x.dtype) - (kernel_size - 1) * 0.5
    k = torch.exp(-0.5 * (t / sigma) ** 2)
    k = (k / k.sum())
    c = x.shape[1]
    x = F.conv2d(F.pad(x, (r, r, 0, 0), mode='reflect'), k.view(1, 1, 1, -1).expand(c, 1, 1, -1), groups=c)
    return F.conv2d(F.pad(x, (0, 0, r, r), mode='reflect'), k.view(1, 1, -1, 1).expand(c, 1, -1, 1), groups=c)


def highpass(x, std=5, offset=0.5):
    """lib/pipelines/utils.py:187-188 on NCHW: offset + x - gaussian_blur(x, 6 round(std) + 1, std)."""
    return offset + x - blur_masks(x, int(round(std)) * 6 + 1, std)


_default_highpass = highpass          # (``nerf_optim`` has an argument of the same name)


class L1LossMod(nn.Module):
    """Holder of the pixel-loss weight (lib/models/losses/pixelwise_loss.py:9-35; the pipelines build it with loss_weight=1.2,
    lib/pipelines/utils.py:231).  The fused objective kernel reads ``loss_weight``; ``forward`` is the plain weighted-mean L1."""

    def __init__(self, loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None):
        e = (pred - target).abs()
        e = e if weight is None else e * weight
        return (e.mean() if avg_factor is None else e.sum() / avg_factor) * self.loss_weight


# ------------------------------------------------------------------------------------------------ BaseNeRF
class BaseNeRF(nn.Module):
    """The slice of base_nerf.py:78-322,489-556 that MVEdit's pipelines call (init_mvedit, lib/pipelines/utils.py:216-233)."""

    def __init__(self, grid_size=128, decoder=None, bg_color=1.0, pixel_loss=None, patch_loss=None, patch_size=128,
                 update_extra_interval=16, update_extra_iters=1):
        super().__init__()
        self.grid_size = grid_size
        self.decoder = decoder if decoder is not None else iNGPDecoder(max_resolution=320, n_levels=12, max_steps=1024,
                                                                       weight_culling_th=0.001)
        self.bg_color = bg_color
        self.pixel_loss = pixel_loss if pixel_loss is not None else L1LossMod(loss_weight=1.2)
        self.patch_loss = patch_loss
        self.patch_size = patch_size
        self.update_extra_interval = update_extra_interval
        self.update_extra_iters = update_extra_iters

    def get_init_density_grid(self, num_scenes, device=None):
        return torch.zeros(self.grid_size ** 3 if num_scenes is None else (num_scenes, self.grid_size ** 3), device=device,
                           dtype=torch.float16)

    def get_init_density_bitfield(self, num_scenes, device=None):
        return torch.zeros(self.grid_size ** 3 // 8 if num_scenes is None else (num_scenes, self.grid_size ** 3 // 8), device=device,
                           dtype=torch.uint8)

    def ray_sample(self, cond_rays_o, cond_rays_d, cond_imgs, n_samples, sample_inds=None, cond_extras=None):
        """base_nerf.py:245-303, patch-wise branch: (1,V,h,w,C) tensors -> the rays / targets of the drawn ps x ps patches,
        ``(rays_o (1,n,3), rays_d (1,n,3), target_rgbs (P,ps,ps,3), *extras (P,ps,ps,C))``.  Patches are numbered (view, row, col)
        as in the reference; only the drawn patches are gathered (the reference reshuffles all V*h*w pixels first)."""
        ps = self.patch_size
        assert cond_rays_o.shape[0] == 1 and n_samples % (ps * ps) == 0
        V, h, w = cond_rays_o.shape[1:4]
        n_total, n_sel = V * (h // ps) * (w // ps), n_samples // (ps * ps)
        if V * h * w > n_samples:
            if sample_inds is None:
                sample_inds = torch.randperm(n_total, device=cond_rays_o.device)[None, :n_sel]
            inds = sample_inds[0]
        else:
            inds = torch.arange(n_total, device=cond_rays_o.device)
        pick = lambda t: _gather_patches(_patch_view(t, ps), inds)[0]
        extras = [] if cond_extras is None else [pick(e) for e in cond_extras]
        return (pick(cond_rays_o).reshape(1, -1, 3), pick(cond_rays_d).reshape(1, -1, 3), pick(cond_imgs), *extras)

    def get_raybatch_inds(self, cond_imgs, n_inverse_rays):
        """base_nerf.py:305-322 (patch branch): one random permutation of all patches, cut into per-iteration batches."""
        _, V, h, w, _ = cond_imgs.shape
        pp = self.patch_size ** 2
        if V * h * w <= n_inverse_rays:
            return None, None
        perm = torch.randperm(V * h * w // pp, device=cond_imgs.device)[None]
        batches = perm.split(n_inverse_rays // pp, dim=1)
        return batches, len(batches)

    def render(self, decoder, code, density_bitfield, h, w, intrinsics, poses, cfg=dict(), bg_color=None, perturb=False,
               normal_bg=(0.5, 0.5, 1.0)):
        """base_nerf.py:489-556.  intrinsics (1,V,4), poses (1,V,3|4,4).  One fused launch renders all V*h*w rays (ray origins /
        directions are generated inside the kernel; the reference materialises two (1,V,h,w,3) tensors per call); with
        ``compute_normal`` the depth -> normal stencil runs in mve_shade_views."""
        assert not perturb
        if bg_color is None:
            bg_color = self.bg_color
        assert intrinsics.dim() == 3 and intrinsics.size(0) == 1, 'one scene'
        K = intrinsics[0].float().contiguous()
        dt_gamma = float(cfg.get('dt_gamma_scale', 0.0) * 2 / (K[:, 0] + K[:, 1]).mean())
        extra = dict(code=code) if not getattr(decoder, 'supports_capacity', True) else {}      # tri-plane decoders render from `code`
        ws, depth, image = decoder.render_cameras(poses[0], K, h, w, density_bitfield, self.grid_size, dt_gamma=dt_gamma, **extra)
        return_rgba = cfg.get('return_rgba', False)
        if return_rgba:
            out_image = torch.cat([image, ws.unsqueeze(-1)], dim=-1)[None]
        else:
            out_image = (image + bg_color * (1 - ws.unsqueeze(-1)))[None]
        out_depth = depth[None]
        if cfg.get('inverse_z_depth', True):
            out_depth = out_depth * pixel_directions(intrinsics, h, w).norm(dim=-1)          # 1/r -> 1/z
        if cfg.get('compute_normal', False):
            assert cfg.get('inverse_z_depth', True) and return_rgba
            V = K.shape[0]
            nfg = torch.empty(V, h, w, 3, dtype=torch.float32, device=K.device)
            call('mve_shade_views', ptr(ws), ptr(depth), ptr(image), ptr(K), ptr(None), c_u32(V), c_u32(h), c_u32(w), c_f32(0.0),
                 c_f32(float(bg_color)), c_f32(0.25), c_f32(0.5), c_f32(1e-5), ptr(None), ptr(None), ptr(None), ptr(nfg), None, c_u32(0),
                 stream())
            out_normal_fg = nfg[None]
            a = out_image[..., 3:]
            out_normal = out_normal_fg * a + out_normal_fg.new_tensor(normal_bg) * (1 - a)
            return out_image, out_depth, out_normal, out_normal_fg
        return out_image, out_depth


def patch_loss(image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, shaded, ambient, bg_color, bg_width,
               pixel_loss_weight, w_alpha_mul, w_normal_reg, w_entropy, tonemapping=None, g_out_extra=None,
               tgt_normal=None, g_normal_extra=None, normal_bg=(0.5, 0.5, 1.0), tgt_depth=None, w_depth=None):
    """Fused objective of one nerf_optim iteration on P patches of ps x ps rays (mve_nerf_patch_loss, four launches):
    -> (loss5 = [total, rgb, alpha, normal_reg, background-entropy], d total / d image [N,3], / d alpha [N], / d depth [N]).
    w_* are device scalars (schedule dependent; they must stay valid inside a captured graph).
    Optional targets (mve_nerf_patch_loss_targets): ``tgt_normal`` [N,3] inside the TV term, ``g_normal_extra`` [N,3] the gradient of the
    normal patch term w.r.t. the composited normals, ``tgt_depth`` [N] with the device scalar ``w_depth`` -- then a 6th value, the depth
    term, is appended to the returned losses."""
    N = alpha.numel()
    P = N // (ps * ps)
    dev = alpha.device
    f = lambda t: None if t is None else t.detach().float().contiguous()
    scratch = torch.empty(N * 10, dtype=torch.float32, device=dev)
    g_image, g_alpha, g_depth = torch.empty(N, 3, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev)
    loss5 = torch.empty(6 if tgt_depth is not None else 5, dtype=torch.float32, device=dev)
    # named references: a converted copy must stay alive until the launches are enqueued
    image_, alpha_, depth_, trgb, tmsk, dirs_, pw, lt, ge_rgb = (f(t) for t in (image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights,
                                                                                g_out_extra))
    base = (ptr(image_.view(N, 3)), ptr(alpha_.view(N)), ptr(depth_.view(N)), ptr(trgb), ptr(tmsk), ptr(dirs_), ptr(pw), ptr(lt),
            c_u32(P), c_u32(ps), c_int(int(shaded)), c_f32(ambient), c_f32(bg_color), c_f32(bg_width), c_f32(pixel_loss_weight),
            ptr(w_alpha_mul), ptr(w_normal_reg), ptr(w_entropy), ptr(scratch), ptr(g_image), ptr(g_alpha), ptr(g_depth), ptr(loss5),
            ptr(ge_rgb), *tone_args(tonemapping))
    if tgt_normal is None and g_normal_extra is None and tgt_depth is None:
        call('mve_nerf_patch_loss', *base, stream())
    else:
        tn, ge, td = f(tgt_normal), f(g_normal_extra), f(tgt_depth)
        loss_d = loss5[5:] if td is not None else None
        call('mve_nerf_patch_loss_targets', *base, ptr(tn), ptr(ge), c_f32(normal_bg[0]), c_f32(normal_bg[1]), c_f32(normal_bg[2]),
             ptr(td), ptr(w_depth if td is not None else None), ptr(loss_d), stream())
    return loss5, g_image, g_alpha, g_depth


def patch_out_normal(alpha, depth, dirs, ps, normal_bg=(0.5, 0.5, 1.0)):
    """-> [N,3] the alpha-composited normal map of P patches (mvedit_3d_pipeline.py:549-554), the input of the normal patch term."""
    N = alpha.numel()
    alpha_, depth_, dirs_ = (t.detach().float().contiguous() for t in (alpha, depth, dirs))
    scratch = torch.empty(N * 10, dtype=torch.float32, device=alpha.device)
    out = torch.empty(N, 3, dtype=torch.float32, device=alpha.device)
    call('mve_nerf_patch_out_normal', ptr(alpha_.view(N)), ptr(depth_.view(N)), ptr(dirs_), c_u32(N // (ps * ps)), c_u32(ps),
         c_f32(normal_bg[0]), c_f32(normal_bg[1]), c_f32(normal_bg[2]), ptr(scratch), ptr(out), stream())
    return out


# ------------------------------------------------------------------------------------------------ nerf_optim
def _patch_view(t, ps):
    """(1, V, h, w, C) -> non-materialised view (V, h/ps, w/ps, ps, ps, C): indexing it gathers only the chosen patches
    (the reference reshapes the permuted tensor, i.e. copies all V*h*w pixels every iteration, base_nerf.py:266-282)."""
    _, V, h, w, C = t.shape
    return t[0].reshape(V, h // ps, ps, w // ps, ps, C).permute(0, 1, 3, 2, 4, 5)


def _gather_patches(view, inds):
    V, nh, nw = view.shape[:3]
    v = torch.div(inds, nh * nw, rounding_mode='floor')
    r = inds - v * (nh * nw)
    ih = torch.div(r, nw, rounding_mode='floor')
    return view[v, ih, r - ih * nw], v


def nerf_optim(nerf, tgt_images, tgt_masks, tgt_normals, optimizer, lr, inverse_steps, n_inverse_rays,
               patch_rgb_weight, patch_normal_weight, alpha_soften, normal_reg_weight, entropy_weight,
               nerf_code, density_grid, density_bitfield,
               render_size, intrinsics, intrinsics_size, camera_poses, cam_weights, cam_lights, patch_size,
               is_init, bg_width, ambient_light, dt_gamma_scale, init_shaded,
               alpha_blur_std=1.5, debug=False, tgt_depths=None, depth_weight=0.0, tonemapping=None, normal_bg=(0.5, 0.5, 1.0),
               highpass=None):
    """MVEdit3DPipeline.nerf_optim (mvedit_3d_pipeline.py:452-656), same arguments (``self`` -> ``nerf`` + ``tonemapping``).

    tgt_images (1,V,rs,rs,3), tgt_masks (1,V,rs,rs,1), intrinsics (V,4), camera_poses (V,3|4,4), cam_weights (V,), cam_lights (V,3).

    Same objective, term by term (checked against oracle/nerf_oracle.py); the execution plan differs:
      * rays of the drawn patches are generated from (pose, directions[patch]) -- the two (1,V,rs,rs,3) origin/direction tensors and
        the per-iteration whole-image patch reshuffle are never materialised;
      * the objective and its gradient w.r.t. the renderer outputs are four kernels (mve_nerf_patch_loss);
      * ``nerf.decoder.sample_capacity > 0`` (set automatically when 0): no host sync inside an iteration, and with
        ``nerf.use_cuda_graph`` (needs a capturable optimizer) forward + objective + backward + optimizer step of one iteration replay
        as ONE CUDA graph; the occupancy refresh every ``update_extra_interval`` iterations runs between replays;
      * under ``torch.distributed`` with ``nerf.data_parallel`` the rays of each patch batch are split across ranks and the gradients
        are all-reduced once per iteration (view_shard.allreduce_grads).
    Returns the per-iteration loss log when ``debug`` else None."""
    device = tgt_images.device
    use_normal = tgt_normals is not None              # :462-463
    use_depth = tgt_depths is not None and depth_weight > 0
    use_pn = use_normal and patch_normal_weight > 0   # :619
    use_prgb = patch_rgb_weight > 0
    tone = tone_args(tonemapping)                     # knots by value into the kernels (static: safe inside the captured graph)
    lpips = nerf.patch_loss if (use_prgb or use_pn) else None
    if lpips is None and (use_prgb or use_pn):
        raise RuntimeError('nerf_optim: a patch term has a positive weight but nerf.patch_loss is None')
    if lpips is not None and not hasattr(lpips, 'loss_and_grad'):
        raise NotImplementedError('nerf_optim: nerf.patch_loss must be a mvedit_b200.lpips.LPIPSLoss (kernels, no autograd graph); got %r'
                                  % type(lpips).__name__)
    hp = _default_highpass if highpass is None else highpass
    normal_bg = tuple(float(v) for v in normal_bg)
    ps = nerf.patch_size
    assert patch_size == ps
    dec = nerf.decoder
    V = camera_poses.shape[0]
    n_sel = max(n_inverse_rays // (ps * ps), 1)
    n_patches_total = V * (render_size // ps) ** 2
    dec.check_sample_overflow()                       # result of the PREVIOUS call's iterations (no sync on the fast path)
    if getattr(dec, 'supports_capacity', True) and getattr(dec, 'auto_capacity', dec.sample_capacity <= 0):
        # post-cull sample buffers: the initial fit starts from fog (hundreds of surviving samples per ray), later calls see a
        # fitted field.  N * max_steps can never overflow; N * 256 is checked (check_sample_overflow raises if rays were dropped).
        dec.auto_capacity = True
        dec.sample_capacity = n_sel * ps * ps * (int(dec.max_steps) if is_init else min(256, int(dec.max_steps)))
    use_graph = bool(getattr(nerf, 'use_cuda_graph', False)) and bool(optimizer.defaults.get('capturable', False)) and not debug \
        and getattr(dec, 'supports_capacity', True)
    rank, world = view_shard.world() if getattr(nerf, 'data_parallel', False) else (0, 1)
    if world > 1 and os.environ.get('MVE_DP_GRAPH', '1') == '0':
        use_graph = False                             # escape hatch: eager iterations around the collectives
    assert ps % world == 0, 'data-parallel reconstruction: the patch rows must divide by the world size'

    # ---- per-configuration "program": static input slots + the iteration closure (+ its captured CUDA graph).  It persists on the
    # nerf object (LRU of 2) so that a graph is captured ONCE per configuration and replayed by every later nerf_optim call; inputs are
    # copied into the static slots at the start of each call.  Everything the captured graph bakes in is part of the key.
    key = (V, render_size, ps, n_sel, bool(is_init), bool(init_shaded), use_graph, float(dt_gamma_scale),
           float(ambient_light), float(bg_width), float(intrinsics_size), id(optimizer), density_bitfield.data_ptr(), id(nerf_code),
           float(nerf.bg_color), float(nerf.pixel_loss.loss_weight), int(dec.sample_capacity), int(dec.max_steps),
           float(dec.weight_culling_th), bool(dec.mlp_tf32), int(nerf.grid_size), rank, world, id(lpips),
           None if tonemapping is None else tuple(tonemapping.knots()[0]), use_normal, use_depth, use_pn, use_prgb, normal_bg)
    cache = nerf.__dict__.setdefault('_recon_programs', collections.OrderedDict())
    prog = cache.get(key) if use_graph else None
    if prog is None:
        f32 = dict(dtype=torch.float32, device=device)
        prog = dict(
            img=torch.empty(1, V, render_size, render_size, 3, **f32), msk=torch.empty(1, V, render_size, render_size, 1, **f32),
            R=torch.empty(V, 3, 3, **f32), Tr=torch.empty(V, 3, **f32),
            camw=torch.empty(V, **f32), lights=torch.empty(V, 3, **f32), intr=torch.empty(V, 4, **f32),
            sc=dict(normal_reg=torch.zeros((), **f32), entropy=torch.zeros((), **f32), alpha_mul=torch.zeros((), **f32),
                    patch_rgb=torch.zeros((), **f32), patch_normal=torch.zeros((), **f32), depth=torch.zeros((), **f32)),
            nrm=torch.empty(1, V, render_size, render_size, 3, **f32) if use_normal else None,
            dpt=torch.empty(1, V, render_size, render_size, 1, **f32) if use_depth else None,
            inds=torch.zeros(min(n_sel, n_patches_total), dtype=torch.long, device=device), graph=None, vals=None)
        if use_graph:
            cache[key] = prog
            while len(cache) > 2:                      # each entry pins ~0.3 GB of slots + a graph pool: keep the two most recent
                cache.popitem(last=False)
    elif use_graph:
        cache.move_to_end(key)
    with torch.no_grad():
        if alpha_blur_std > 0:
            kernel_size = int((alpha_blur_std * 6) // 2 * 2 + 1)
            m = blur_masks(tgt_masks[0].permute(0, 3, 1, 2).square(), kernel_size, alpha_blur_std).permute(0, 2, 3, 1)[None]
        else:
            m = tgt_masks
        prog['msk'].copy_(m.clamp(min=alpha_soften ** 2, max=(1 - alpha_soften) ** 2).sqrt())
        prog['img'].copy_(tgt_images)
        prog['R'].copy_(camera_poses[:, :3, :3]); prog['Tr'].copy_(camera_poses[:, :3, 3])
        prog['camw'].copy_(cam_weights); prog['lights'].copy_(cam_lights); prog['intr'].copy_(intrinsics)
        prog['sc']['normal_reg'].fill_(float(normal_reg_weight) * 10)
        prog['sc']['entropy'].fill_(float(entropy_weight))
        prog['sc']['alpha_mul'].fill_(5.0 if is_init else 1.0)
        prog['sc']['patch_rgb'].fill_(float(patch_rgb_weight))
        prog['sc']['patch_normal'].fill_(float(patch_normal_weight))
        prog['sc']['depth'].fill_(float(depth_weight))
        if use_normal:
            prog['nrm'].copy_(tgt_normals)
        if use_depth:
            prog['dpt'].copy_(tgt_depths)
    R, Tr = prog['R'], prog['Tr']
    sc, inds_static = prog['sc'], prog['inds']
    decoder_training_prev = dec.training
    dec.train(True)
    log = [] if debug else None
    fused_opt = isinstance(optimizer, FusedAdam)
    params = [p_ for p_ in dec.parameters() if p_.requires_grad]
    P = inds_static.numel()
    rows = ps // world                                   # data-parallel: this rank's row strip of every drawn patch
    row_lo, row_hi = rank * rows, (rank + 1) * rows
    n, n_loc = P * ps * ps, P * rows * ps
    if 'rays_o' not in prog:
        f32 = dict(dtype=torch.float32, device=device)
        prog.update(rays_o=torch.empty(1, n_loc, 3, **f32), rays_d=torch.empty(1, n_loc, 3, **f32), pdirs=torch.empty(n, 3, **f32),
                    trgb=torch.empty(n, 3, **f32), tmsk=torch.empty(n, **f32), pw=torch.empty(P, **f32), pl=torch.empty(P, 3, **f32),
                    dtg=torch.empty(1, **f32), scratch=torch.empty(n * 10, **f32), g_img=torch.empty(n, 3, **f32),
                    g_a=torch.empty(n, **f32), g_d=torch.empty(n, **f32), loss5=torch.empty(5, **f32), out_rgb=torch.empty(n, 3, **f32),
                    lp=torch.zeros((), **f32), lpn=torch.zeros((), **f32), ld=torch.zeros(1, **f32),
                    out_nrm=torch.empty(n, 3, **f32) if use_pn else None)
    b = prog
    nrm_view = _patch_view(prog['nrm'], ps) if use_normal else None
    dpt_view = _patch_view(prog['dpt'], ps) if use_depth else None

    def iteration():
        # 1. rays of this rank's strip + targets of the full patches: one launch (replaces ray_sample / get_rays / per-patch scalars)
        call('mve_patch_rays', ptr(inds_static), c_u32(P), c_u32(V), c_u32(render_size), c_u32(ps), ptr(R), ptr(Tr), ptr(prog['intr']),
             c_f32(render_size / intrinsics_size), ptr(prog['img']), ptr(prog['msk']), ptr(prog['camw']), ptr(prog['lights']),
             c_f32(float(dt_gamma_scale)), c_u32(row_lo), c_u32(row_hi), ptr(b['rays_o']), ptr(b['rays_d']), ptr(b['pdirs']), ptr(b['trgb']),
             ptr(b['tmsk']), ptr(b['pw']), ptr(b['pl']), ptr(b['dtg']), stream())
        # 2. march -> cull -> field -> composite of the strip (autograd graph: composite <- field <- parameters).  The perturbation
        # noise is drawn for ALL rays of the iteration on every rank (the draw the reference makes, raymarching.py:279-282) and sliced,
        # so the random stream -- and with it the training trajectory -- does not depend on the number of ranks
        noise = torch.rand(n, device=device)
        if world > 1:
            noise = noise.view(P, ps, ps)[:, row_lo:row_hi].reshape(-1)
        out = dec(b['rays_o'], b['rays_d'], nerf_code, density_bitfield, nerf.grid_size, dt_gamma=b['dtg'], perturb=True, noises=noise,
                  fused_entropy=(sc['entropy'], 1.0 / n))
        image, alpha, depth = out['image'][0], out['weights_sum'][0], out['depth'][0]
        if world > 1:        # 3. exchange the per-ray outputs (5 floats per ray): every rank sees the full patches
            full = view_shard.gather_rays(torch.cat([image.detach(), alpha.detach()[:, None], depth.detach()[:, None]], dim=1), P)
            f_img, f_a, f_d = full[:, :3].contiguous(), full[:, 3].contiguous(), full[:, 4].contiguous()
        else:
            f_img, f_a, f_d = image.detach(), alpha.detach(), depth.detach()
        shaded = int((not is_init) or init_shaded)
        # 4a. LPIPS patch term (:611-617) on the shaded, composited rgb: VGG16 forward over [rendered ; target] patches and the
        # gradient back to the rendered pixels, ~42 launches of the conv / lpips kernels (mvedit_b200.lpips), no autograd graph
        g_extra = g_nextra = tn = td = None
        if use_prgb:
            call('mve_nerf_patch_out_rgb', ptr(f_img), ptr(f_a), ptr(f_d), ptr(b['pdirs']), ptr(b['pl']), c_u32(P), c_u32(ps), c_int(shaded),
                 c_f32(ambient_light), c_f32(float(nerf.bg_color)), ptr(b['scratch']), ptr(b['out_rgb']), *tone, stream())
            lp, g_extra, _ = lpips.loss_and_grad(b['out_rgb'].view(P, ps, ps, 3), b['trgb'].view(P, ps, ps, 3), b['pw'], sc['patch_rgb'])
            b['lp'].copy_(lp)
        # 4a'. image-to-3D targets (:517-529): normals / depths of the drawn patches, gathered from the static slots
        if use_normal:
            tn = _gather_patches(nrm_view, inds_static)[0].reshape(n, 3).contiguous()
        if use_depth:
            td = _gather_patches(dpt_view, inds_static)[0].reshape(n).contiguous()
        if use_pn:
            # high-passed normal patch term (:619-626): LPIPS between highpass(composited normals) and highpass(target normals); the
            # Gaussian of the high-pass is two depthwise torch convolutions, its transpose comes from autograd
            call('mve_nerf_patch_out_normal', ptr(f_a), ptr(f_d), ptr(b['pdirs']), c_u32(P), c_u32(ps), c_f32(normal_bg[0]),
                 c_f32(normal_bg[1]), c_f32(normal_bg[2]), ptr(b['scratch']), ptr(b['out_nrm']), stream())
            x = b['out_nrm'].view(P, ps, ps, 3).permute(0, 3, 1, 2).detach().requires_grad_(True)
            hx = hp(x)
            ht = hp(tn.view(P, ps, ps, 3).permute(0, 3, 1, 2))
            lpn, g_hx, _ = lpips.loss_and_grad(hx.detach().permute(0, 2, 3, 1).contiguous(), ht.permute(0, 2, 3, 1).contiguous(), b['pw'],
                                               sc['patch_normal'])
            g_x, = torch.autograd.grad(hx, x, g_hx.permute(0, 3, 1, 2))
            g_nextra = g_x.permute(0, 2, 3, 1).reshape(n, 3).contiguous()
            b['lpn'].copy_(lpn)
        # 4b. objective on the full patches: loss terms and d/d(image, alpha, depth) in four launches (replicated: 16 384 pixels)
        base = (ptr(f_img), ptr(f_a), ptr(f_d), ptr(b['trgb']), ptr(b['tmsk']), ptr(b['pdirs']), ptr(b['pw']), ptr(b['pl']),
                c_u32(P), c_u32(ps), c_int(shaded), c_f32(ambient_light), c_f32(float(nerf.bg_color)), c_f32(bg_width),
                c_f32(float(nerf.pixel_loss.loss_weight)), ptr(sc['alpha_mul']), ptr(sc['normal_reg']), ptr(sc['entropy']), ptr(b['scratch']),
                ptr(b['g_img']), ptr(b['g_a']), ptr(b['g_d']), ptr(b['loss5']), ptr(g_extra), *tone)
        if use_normal or use_depth:
            call('mve_nerf_patch_loss_targets', *base, ptr(tn), ptr(g_nextra), c_f32(normal_bg[0]), c_f32(normal_bg[1]), c_f32(normal_bg[2]),
                 ptr(td), ptr(sc['depth'] if use_depth else None), ptr(b['ld'] if use_depth else None), stream())
        else:
            call('mve_nerf_patch_loss', *base, stream())
        if world > 1:
            sel = lambda t: t.view(P, ps, ps, -1)[:, row_lo:row_hi].reshape(n_loc, -1)
            g_img, g_a, g_d = sel(b['g_img']), sel(b['g_a']).view(-1), sel(b['g_d']).view(-1)
        else:
            g_img, g_a, g_d = b['g_img'], b['g_a'], b['g_d']
        # 5. backward of the strip, 6. gradient exchange + Adam
        optimizer.zero_grad(set_to_none=False)
        torch.autograd.backward([image, alpha, depth], [g_img, g_a, g_d])
        if world > 1:
            if fused_opt:
                view_shard.allreduce_flat(optimizer.flat_grad)
            else:
                view_shard.allreduce_grads(params)
        optimizer.step()
        return b['loss5']

    if prog.get('iteration') is None:
        prog['iteration'] = iteration       # the closure the graph was (or will be) captured from
    iteration = prog['iteration']
    dec.grad_sink = optimizer if fused_opt else None
    if fused_opt:
        optimizer.hard_zero_grad()
    with torch.enable_grad():
        g0 = optimizer.param_groups[0]                      # the reference sets param_groups[0]['lr'] (:494)
        if isinstance(g0['lr'], torch.Tensor):
            g0['lr'].fill_(float(lr))
        elif use_graph:
            g0['lr'] = torch.tensor(float(lr), device=device)
        else:
            g0['lr'] = lr
        raybatch_inds, num_raybatch = nerf.get_raybatch_inds(tgt_images, n_inverse_rays)
        if world > 1 and raybatch_inds is not None:
            raybatch_inds = view_shard.broadcast_patch_order(raybatch_inds)         # every rank works on the same permutation
        iter_density = 0
        graph, vals_static = prog['graph'], prog['vals']
        for inverse_step_id in range(inverse_steps):
            if inverse_step_id % nerf.update_extra_interval == 0:
                for _ in range(nerf.update_extra_iters):
                    dec.update_extra_state(nerf_code, density_grid, density_bitfield, iter_density, density_thresh=0.1)
            if raybatch_inds is not None:
                inds_static.copy_(raybatch_inds[inverse_step_id % num_raybatch][0][:inds_static.numel()])
            else:
                inds_static.copy_(torch.arange(inds_static.numel(), device=device))
            if not use_graph:
                vals = iteration()
            elif graph is None:
                # iteration 0 runs eagerly on a side stream (allocator warm-up, optimizer state creation) and is then captured;
                # iterations >= 1 are replays of the captured graph
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    vals = iteration()
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                # data-parallel: the captured iteration contains NCCL collectives; the process group's watchdog thread polls CUDA
                # events concurrently, which the default ("global") capture mode turns into a capture error -> thread_local
                with torch.cuda.graph(graph, capture_error_mode='thread_local' if world > 1 else 'global'):
                    vals_static = iteration()
                prog['graph'], prog['vals'] = graph, vals_static
            else:
                graph.replay()
                vals = vals_static
            if debug:
                v = [float(x) for x in vals]
                lpv = float(prog['lp']) if use_prgb else 0.0
                lpn = float(prog['lpn']) if use_pn else 0.0
                log.append(dict(loss=v[0] + lpv + lpn, pixel_rgb=v[1], alpha=v[2], normal_reg=v[3], entropy=v[4], patch_rgb=lpv,
                                patch_normal=lpn, depth=float(prog['ld']) if use_depth else 0.0))
    dec.grad_sink = None
    dec.note_sample_overflow()
    dec.train(decoder_training_prev)
    return log

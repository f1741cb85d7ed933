"""Host-side mirror of the NeRF-adapter slice of the reference that sits on the hot path:

  * geometry helpers        /root/reference/lib/core/utils/geometry_utils.py:18-55,119-168
  * ``BaseNeRF``            /root/reference/lib/models/autoencoders/base_nerf.py:78-322,489-556  (ray_sample, get_raybatch_inds, render,
                            density grid / bitfield init) -- the mmgen registry / training-harness parts are out of scope
  * loss modules            /root/reference/lib/models/losses/{pixelwise_loss,tv_loss}.py (mmgen ``weighted_loss`` semantics restated)
  * ``nerf_optim``          /root/reference/lib/pipelines/mvedit_3d_pipeline.py:452-656, the reconstruction inner loop

Same names / arguments / return values; torch is the plumbing, the arithmetic of march / field / composite / render is libmvedit_b200.
LPIPS (``patch_loss``) is a neighbour of the path: pass any callable ``patch_loss(pred_nchw, tgt_nchw, weight=...)``;
with ``None`` the two patch terms are skipped.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ingp_decoder import iNGPDecoder
from ._lib import call, ptr, stream, get_lib, c_int, c_u32, c_f32


# ------------------------------------------------------------------------------------------------ geometry
def get_ray_directions(h, w, intrinsics, norm=False, device=None):
    """geometry_utils.py:18-40.  intrinsics (*,4) [fx,fy,cx,cy] -> (*,h,w,3) camera-space directions (z=1)."""
    batch_size = intrinsics.shape[:-1]
    x = torch.linspace(0.5, w - 0.5, w, device=device)
    y = torch.linspace(0.5, h - 0.5, h, device=device)
    directions_xy = torch.stack(
        [((x - intrinsics[..., 2:3]) / intrinsics[..., 0:1])[..., None, :].expand(*batch_size, h, w),
         ((y - intrinsics[..., 3:4]) / intrinsics[..., 1:2])[..., :, None].expand(*batch_size, h, w)], dim=-1)
    directions = F.pad(directions_xy, [0, 1], mode='constant', value=1.0)
    if norm:
        directions = F.normalize(directions, dim=-1)
    return directions


def get_rays(directions, c2w, norm=False):
    """geometry_utils.py:43-55."""
    rays_d = directions @ c2w[..., None, :3, :3].transpose(-1, -2)
    rays_o = c2w[..., None, None, :3, 3].expand(rays_d.shape)
    if norm:
        rays_d = F.normalize(rays_d, dim=-1)
    return rays_o, rays_d


def depth_to_normal(depth, directions, format='opengl'):
    """geometry_utils.py:119-148.  depth = inverse depth 1/z (*,h,w); -> normals in [0,1]."""
    out_xyz = directions / depth.unsqueeze(-1).clamp(min=1e-6)
    dx = out_xyz[..., :, 1:, :] - out_xyz[..., :, :-1, :]
    dy = out_xyz[..., 1:, :, :] - out_xyz[..., :-1, :, :]
    right = F.pad(dx, (0, 0, 0, 1, 0, 0), mode='replicate')
    up = F.pad(-dy, (0, 0, 0, 0, 1, 0), mode='replicate')
    left = F.pad(-dx, (0, 0, 1, 0, 0, 0), mode='replicate')
    down = F.pad(dy, (0, 0, 0, 0, 0, 1), mode='replicate')
    out_normal = F.normalize(
        F.normalize(torch.cross(right, up, dim=-1), dim=-1)
        + F.normalize(torch.cross(up, left, dim=-1), dim=-1)
        + F.normalize(torch.cross(left, down, dim=-1), dim=-1)
        + F.normalize(torch.cross(down, right, dim=-1), dim=-1), dim=-1)
    if format == 'opengl':
        out_normal = torch.cat([out_normal[..., :1], -out_normal[..., 1:3]], dim=-1)
    elif format != 'opencv':
        raise ValueError('format should be opengl or opencv')
    return out_normal / 2 + 0.5


def normalize_depth(depths, alphas, far_depth=0.25, alpha_clip=0.5, eps=1e-5):
    """geometry_utils.py:151-168."""
    depths_max = depths.flatten(1).amax(dim=1)[:, None, None]
    depths_fg = depths / alphas.clamp(min=eps).squeeze(-1)
    depths_fg_min = depths_fg.masked_fill(alphas.squeeze(-1) < alpha_clip, 1 / eps).flatten(1).amin(dim=1)[:, None, None]
    depths_fg = (depths_fg - depths_fg_min) / (depths_max - depths_fg_min).clamp(min=eps)
    depths_fg = depths_fg * (1 - far_depth) + far_depth
    return (depths_fg * alphas.squeeze(-1)).clamp(min=0, max=1)


# ------------------------------------------------------------------------------------------------ losses
def _weighted(loss, weight=None, avg_factor=None):
    """mmgen ``weighted_loss`` with reduction='mean': elementwise * weight, then mean (or sum / avg_factor)."""
    if weight is not None:
        loss = loss * weight
    return loss.mean() if avg_factor is None else loss.sum() / avg_factor


class L1LossMod(nn.Module):
    """pixelwise_loss.py:9-35."""

    def __init__(self, loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None):
        return _weighted(torch.abs(pred - target), weight, avg_factor) * self.loss_weight


class TVLoss(nn.Module):
    """tv_loss.py:7-61."""

    def __init__(self, dims=(-2, -1), power=1, loss_weight=1.0):
        super().__init__()
        self.dims, self.power, self.loss_weight = list(dims), power, loss_weight

    def forward(self, pred, target=None, weight=None, avg_factor=None):
        def diffs(t):
            out = []
            for dim in self.dims:
                pad_shape = list(t.size())
                pad_shape[dim] = 1
                out.append(torch.cat([torch.diff(t, dim=dim), t.new_zeros(pad_shape)], dim=dim))
            return torch.stack(out, dim=0)

        diff_loss = diffs(pred) if target is None else diffs(pred) - diffs(target)
        if weight is not None:
            dw = []
            for dim in self.dims:
                pad_shape = list(weight.size())
                pad_shape[dim] = 1
                dw.append(torch.cat([torch.minimum(torch.narrow(weight, dim, 0, weight.size(dim) - 1),
                                                   torch.narrow(weight, dim, 1, weight.size(dim) - 1)),
                                     weight.new_zeros(pad_shape)], dim=dim))
            diff_loss = diff_loss * torch.stack(dw, dim=0)
        loss = diff_loss.norm(dim=0).pow(self.power).mean(dim=self.dims)
        return _weighted(loss, None, avg_factor) * self.loss_weight


def gaussian_blur(x, kernel_size, sigma):
    """torchvision.transforms.functional.gaussian_blur (reflect padding) on NCHW, as used at mvedit_3d_pipeline.py:473-476."""
    ks = kernel_size
    half = (ks - 1) * 0.5
    xs = torch.linspace(-half, half, ks, device=x.device, dtype=x.dtype)
    k1 = torch.exp(-0.5 * (xs / sigma) ** 2)
    k1 = k1 / k1.sum()
    k2 = (k1[:, None] * k1[None, :])[None, None].expand(x.shape[1], 1, ks, ks)
    xp = F.pad(x, [ks // 2] * 4, mode='reflect')
    return F.conv2d(xp, k2, groups=x.shape[1])


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
        """base_nerf.py:245-303 (patch-wise branch: the reference builds BaseNeRF with patch_size=128 and a patch loss, so
        rays are always drawn as whole patches -- kept even when the LPIPS callable itself is absent)."""
        device = cond_rays_o.device
        num_scenes, num_imgs, h, w, _ = cond_rays_o.size()
        num_scene_pixels = num_imgs * h * w
        ps = self.patch_size
        assert n_samples % (ps ** 2) == 0

        def to_patches(t):
            c = t.size(-1)
            return t.reshape(num_scenes, -1, h // ps, ps, w // ps, ps, c).permute(0, 1, 2, 4, 3, 5, 6).reshape(num_scenes, -1, ps, ps, c)

        rays_o, rays_d, target_rgbs = to_patches(cond_rays_o), to_patches(cond_rays_d), to_patches(cond_imgs)
        target_extras = [] if cond_extras is None else [to_patches(e) for e in cond_extras]
        if num_scene_pixels > n_samples:
            if sample_inds is None:
                sample_inds = torch.stack([torch.randperm(target_rgbs.size(1), device=device)[:n_samples // (ps ** 2)]
                                           for _ in range(num_scenes)], dim=0)
            scene_arange = torch.arange(num_scenes, device=device)[:, None]
            rays_o, rays_d, target_rgbs = rays_o[scene_arange, sample_inds], rays_d[scene_arange, sample_inds], target_rgbs[scene_arange, sample_inds]
            target_extras = [e[scene_arange, sample_inds] for e in target_extras]
        rays_o = rays_o.reshape(num_scenes, -1, 3)
        rays_d = rays_d.reshape(num_scenes, -1, 3)
        target_rgbs = target_rgbs.reshape(-1, ps, ps, 3)
        target_extras = [e.reshape(-1, ps, ps, e.size(-1)) for e in target_extras]
        return (rays_o, rays_d, target_rgbs, *target_extras)

    def get_raybatch_inds(self, cond_imgs, n_inverse_rays):
        """base_nerf.py:305-322 (patch branch)."""
        device = cond_imgs.device
        num_scenes, num_imgs, h, w, _ = cond_imgs.size()
        num_scene_pixels = num_imgs * h * w
        if num_scene_pixels > n_inverse_rays:
            raybatch_inds = [torch.randperm(num_scene_pixels // (self.patch_size ** 2), device=device) for _ in range(num_scenes)]
            raybatch_inds = torch.stack(raybatch_inds, dim=0).split(n_inverse_rays // (self.patch_size ** 2), dim=1)
            num_raybatch = len(raybatch_inds)
        else:
            raybatch_inds = num_raybatch = None
        return raybatch_inds, num_raybatch

    def render(self, decoder, code, density_bitfield, h, w, intrinsics, poses, cfg=dict(), bg_color=None, perturb=False,
               normal_bg=(0.5, 0.5, 1.0)):
        """base_nerf.py:489-556.  intrinsics (1,V,4), poses (1,V,3|4,4).  One fused launch renders all V*h*w rays; ray origins /
        directions are generated inside the kernel (the reference materialises two (1,V,h,w,3) tensors per call)."""
        assert not perturb
        if bg_color is None:
            bg_color = self.bg_color
        assert intrinsics.dim() == 3 and intrinsics.size(0) == 1, 'one scene'
        dt_gamma_scale = cfg.get('dt_gamma_scale', 0.0)
        dt_gamma = float((dt_gamma_scale * 2 / (intrinsics[..., 0] + intrinsics[..., 1]).mean(dim=-1))[0])
        ws, depth, image = decoder.render_cameras(poses[0], intrinsics[0], h, w, density_bitfield, self.grid_size, dt_gamma=dt_gamma)
        return_rgba = cfg.get('return_rgba', False)
        if return_rgba:
            out_image = torch.cat([image, ws.unsqueeze(-1)], dim=-1)[None]
        else:
            out_image = (image + bg_color * (1 - ws.unsqueeze(-1)))[None]
        out_depth = depth[None]
        directions = None
        if cfg.get('inverse_z_depth', True):
            directions = get_ray_directions(h, w, intrinsics, norm=False, device=intrinsics.device)
            out_depth = out_depth * torch.linalg.norm(directions, dim=-1)
        if cfg.get('compute_normal', False):
            assert cfg.get('inverse_z_depth', True) and return_rgba
            out_depth_fg = out_depth / out_image[..., 3].clamp(min=1e-6)
            out_normal_fg = depth_to_normal(out_depth_fg, directions)
            out_normal = out_normal_fg * out_image[..., 3:] + out_normal_fg.new_tensor(normal_bg) * (1 - out_image[..., 3:])
            return out_image, out_depth, out_normal, out_normal_fg
        return out_image, out_depth



class _PatchLossFn(torch.autograd.Function):
    """Fused objective of one nerf_optim iteration (libmvedit_b200: mve_nerf_patch_loss).  Returns the 5 loss terms
    [total, rgb, alpha, normal_reg, background-entropy]; only element 0 is meant to be back-propagated."""

    @staticmethod
    def forward(ctx, image, alpha, depth, tgt_rgb, tgt_mask, dirs, patch_w, lights, ps, shaded, ambient, bg_color, bg_width,
                pixel_loss_weight, w_alpha_mul, w_normal_reg, w_entropy):
        N = alpha.numel()
        P = N // (ps * ps)
        dev = alpha.device
        f = lambda t: t.float().contiguous()
        ctx.shapes = (image.shape, alpha.shape, depth.shape)
        image, alpha, depth = f(image).view(N, 3), f(alpha).view(N), f(depth).view(N)
        scratch = torch.empty(N * 10, dtype=torch.float32, device=dev)
        g_image, g_alpha, g_depth = torch.empty(N, 3, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev)
        loss5 = torch.empty(5, dtype=torch.float32, device=dev)
        call('mve_nerf_patch_loss', ptr(image), ptr(alpha), ptr(depth), ptr(f(tgt_rgb)), ptr(f(tgt_mask)), ptr(f(dirs)), ptr(f(patch_w)),
             ptr(f(lights)), c_u32(P), c_u32(ps), c_int(int(shaded)), c_f32(ambient), c_f32(bg_color), c_f32(bg_width), c_f32(pixel_loss_weight),
             ptr(w_alpha_mul), ptr(w_normal_reg), ptr(w_entropy), ptr(scratch), ptr(g_image), ptr(g_alpha), ptr(g_depth), ptr(loss5), stream())
        ctx.save_for_backward(g_image, g_alpha, g_depth)
        return loss5

    @staticmethod
    def backward(ctx, g):
        g_image, g_alpha, g_depth = ctx.saved_tensors
        s = g[0]
        si, sa, sd = ctx.shapes
        return ((g_image * s).view(si), (g_alpha * s).view(sa), (g_depth * s).view(sd)) + (None,) * 14


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

    Same objective, term by term; what differs from the reference is the execution plan:
      * rays of the drawn patches are generated from (pose, directions[patch]) -- the two (1,V,rs,rs,3) origin/direction tensors and
        the per-iteration whole-image patch reshuffle are never materialised;
      * with ``nerf.decoder.sample_capacity > 0`` the iteration has no host sync, and with ``nerf.use_cuda_graph`` (needs
        ``Adam(capturable=True)``) forward + losses + backward + Adam of one iteration replay as ONE CUDA graph; the occupancy
        refresh every ``update_extra_interval`` iterations runs between replays.
    Returns the per-iteration loss log when ``debug`` else None."""
    device = tgt_images.device
    loss_tv = TVLoss(loss_weight=1.0, power=1.5)
    use_normal = tgt_normals is not None
    use_depth = tgt_depths is not None and depth_weight > 0
    ps = nerf.patch_size
    assert patch_size == ps
    V = camera_poses.shape[0]
    n_sel = max(n_inverse_rays // (ps * ps), 1)
    n_patches_total = V * (render_size // ps) ** 2
    # fused objective (4 kernels instead of ~300 eager ops + autograd) for the configuration it covers
    fused = bool(getattr(nerf, 'fused_loss', True)) and tonemapping is None and not use_normal and not use_depth \
        and not (patch_rgb_weight > 0 and nerf.patch_loss is not None) and nerf.decoder.sample_capacity > 0 and tgt_images.is_cuda
    use_graph = bool(getattr(nerf, 'use_cuda_graph', False)) and bool(optimizer.defaults.get('capturable', False)) \
        and nerf.decoder.sample_capacity > 0 and not debug

    # ---- per-configuration "program": static input slots + the iteration closure (+ its captured CUDA graph).  It persists on the
    # nerf object so that a graph is captured ONCE per configuration and replayed by every later nerf_optim call (the pipeline calls
    # nerf_optim once per denoising step); inputs are copied into the static slots at the start of each call.
    key = (V, render_size, ps, n_sel, bool(is_init), bool(init_shaded), use_normal, use_depth, fused, use_graph, float(dt_gamma_scale),
           float(ambient_light), float(bg_width), float(intrinsics_size), float(patch_rgb_weight), float(patch_normal_weight),
           float(depth_weight), id(optimizer), density_bitfield.data_ptr(), id(tonemapping), id(nerf_code), tuple(normal_bg))
    cache = nerf.__dict__.setdefault('_recon_programs', {})
    prog = cache.get(key) if use_graph else None
    if prog is None:
        f32 = dict(dtype=torch.float32, device=device)
        prog = dict(
            img=torch.empty(1, V, render_size, render_size, 3, **f32), msk=torch.empty(1, V, render_size, render_size, 1, **f32),
            dirs=torch.empty(1, V, render_size, render_size, 3, **f32), R=torch.empty(V, 3, 3, **f32), Tr=torch.empty(V, 3, **f32),
            camw=torch.empty(V, **f32), lights=torch.empty(V, 3, **f32), intr=torch.empty(V, 4, **f32),
            nrm=torch.empty(1, V, render_size, render_size, 3, **f32) if use_normal else None,
            dep=torch.empty(1, V, render_size, render_size, 1, **f32) if use_depth else None,
            sc=dict(normal_reg=torch.zeros((), **f32), entropy=torch.zeros((), **f32), alpha_mul=torch.zeros((), **f32)),
            inds=torch.zeros(min(n_sel, n_patches_total), dtype=torch.long, device=device), graph=None, vals=None)
        if use_graph:
            cache[key] = prog
    if alpha_blur_std > 0:
        kernel_size = int((alpha_blur_std * 6) // 2 * 2 + 1)
        tgt_masks_blur = gaussian_blur(tgt_masks.square().squeeze(0).permute(0, 3, 1, 2), kernel_size, alpha_blur_std
                                       ).permute(0, 2, 3, 1)[None].clamp(min=alpha_soften ** 2, max=(1 - alpha_soften) ** 2).sqrt()
    else:
        tgt_masks_blur = tgt_masks.clamp(min=alpha_soften ** 2, max=(1 - alpha_soften) ** 2).sqrt()
    with torch.no_grad():
        prog['img'].copy_(tgt_images); prog['msk'].copy_(tgt_masks_blur)
        prog['dirs'].copy_(get_ray_directions(render_size, render_size, intrinsics[None] * (render_size / intrinsics_size), norm=False,
                                              device=intrinsics.device))
        prog['R'].copy_(camera_poses[:, :3, :3]); prog['Tr'].copy_(camera_poses[:, :3, 3])
        prog['camw'].copy_(cam_weights); prog['lights'].copy_(cam_lights); prog['intr'].copy_(intrinsics)
        if use_normal:
            prog['nrm'].copy_(tgt_normals)
        if use_depth:
            prog['dep'].copy_(tgt_depths)
        prog['sc']['normal_reg'].fill_(float(normal_reg_weight) * 10)
        prog['sc']['entropy'].fill_(float(entropy_weight))
        prog['sc']['alpha_mul'].fill_(5.0 if is_init else 1.0)
    directions, R, Tr = prog['dirs'], prog['R'], prog['Tr']
    cam_weights, cam_lights, intrinsics = prog['camw'], prog['lights'], prog['intr']     # static slots from here on
    sc, inds_static = prog['sc'], prog['inds']
    normal_bg_t = tgt_images.new_tensor(normal_bg)
    pv_img, pv_msk, pv_dir = _patch_view(prog['img'], ps), _patch_view(prog['msk'], ps), _patch_view(directions, ps)
    pv_nrm = _patch_view(prog['nrm'], ps) if use_normal else None
    pv_dep = _patch_view(prog['dep'], ps) if use_depth else None
    decoder_training_prev = nerf.decoder.training
    nerf.decoder.train(True)
    log = [] if debug else None

    def iteration():
        inds = inds_static
        cam_weights_mean = cam_weights.mean()
        target_rgbs, target_cam_ids = _gather_patches(pv_img, inds)
        target_m_blur, _ = _gather_patches(pv_msk, inds)
        target_dir, _ = _gather_patches(pv_dir, inds)
        Rv = R[target_cam_ids]
        rays_d = F.normalize(target_dir @ Rv[:, None].transpose(-1, -2), dim=-1).reshape(1, -1, 3)
        rays_o = Tr[target_cam_ids][:, None, None, :].expand(-1, ps, ps, -1).reshape(1, -1, 3)
        target_w = cam_weights[target_cam_ids][:, None, None, None].expand(-1, ps, ps, 1)
        target_lights = cam_lights[target_cam_ids][:, None, None, :].expand(-1, ps, ps, 3)
        dt_gamma = dt_gamma_scale / (intrinsics[target_cam_ids, :2].mean(dim=-1) * render_size / intrinsics_size)

        if fused:
            n_rays = rays_o.shape[1]
            outputs = nerf.decoder(rays_o, rays_d, nerf_code, density_bitfield, nerf.grid_size, dt_gamma=dt_gamma, perturb=True,
                                   fused_entropy=(sc['entropy'], 1.0 / n_rays))
            terms = _PatchLossFn.apply(outputs['image'], outputs['weights_sum'], outputs['depth'], target_rgbs, target_m_blur, target_dir,
                                       target_w[:, 0, 0, 0] / cam_weights_mean, cam_lights[target_cam_ids], ps, (not is_init) or init_shaded,
                                       ambient_light, float(nerf.bg_color), bg_width, float(nerf.pixel_loss.loss_weight), sc['alpha_mul'],
                                       sc['normal_reg'], sc['entropy'])
            loss = terms[0]
            optimizer.zero_grad(set_to_none=not use_graph)
            loss.backward()
            optimizer.step()
            return terms.detach()
        outputs = nerf.decoder(rays_o, rays_d, nerf_code, density_bitfield, nerf.grid_size, dt_gamma=dt_gamma, perturb=True)
        out_rgbs = outputs['image'].reshape(target_rgbs.size())
        out_alphas = outputs['weights_sum'].reshape(target_m_blur.size())
        out_depth = outputs['depth'].reshape(-1, ps, ps)
        out_depth = out_depth * torch.linalg.norm(target_dir, dim=-1).reshape(out_depth.size())  # 1/r -> 1/z
        out_depth_fg = out_depth / out_alphas.reshape(-1, ps, ps).clamp(min=1e-6)
        out_normals_fg = depth_to_normal(out_depth_fg, target_dir)
        out_normals_fg_mask = out_alphas.reshape(-1, ps, ps, 1)
        out_normals = out_normals_fg * out_normals_fg_mask + normal_bg_t * (1 - out_normals_fg_mask)
        out_normals_fg_weight = -F.max_pool2d(-out_normals_fg_mask.detach().squeeze(-1).unsqueeze(1), 3, stride=1, padding=1
                                              ).squeeze(1).unsqueeze(-1)
        if not is_init or init_shaded:
            out_normals_fg_opencv = torch.cat([out_normals_fg[..., :1] * 2 - 1, -out_normals_fg[..., 1:3] * 2 + 1], dim=-1)
            nerf_shading = ((target_lights[..., None, :] @ out_normals_fg_opencv[..., :, None]).clamp(min=0)
                            * (1 - ambient_light) + ambient_light).squeeze(-1)
            if tonemapping is None:
                out_rgbs = out_rgbs * nerf_shading + nerf.bg_color * (1 - out_alphas)
            else:
                out_rgbs = tonemapping.lut(tonemapping.inverse_lut(out_rgbs / out_alphas.clamp(min=1e-6))
                                           + nerf_shading.clamp(min=1e-6).log2()) * out_alphas + nerf.bg_color * (1 - out_alphas)
        else:
            out_rgbs = out_rgbs + nerf.bg_color * (1 - out_alphas)

        pixel_rgb_loss = nerf.pixel_loss(out_rgbs.reshape(target_rgbs.size()), target_rgbs, weight=target_w / cam_weights_mean) * 4.5
        alphas_loss = nerf.pixel_loss(out_alphas.reshape(target_m_blur.size()), target_m_blur, weight=target_w / cam_weights_mean
                                      ) * sc['alpha_mul']
        target_n = _gather_patches(pv_nrm, inds)[0] if use_normal else None
        normal_reg_loss = loss_tv(out_normals_fg.permute(0, 3, 1, 2), target_n.permute(0, 3, 1, 2) if use_normal else None,
                                  weight=out_normals_fg_weight.permute(0, 3, 1, 2)) * sc['normal_reg']
        loss = pixel_rgb_loss + alphas_loss + normal_reg_loss
        if use_depth:
            target_depth = _gather_patches(pv_dep, inds)[0]
            loss = loss + nerf.pixel_loss(out_depth.reshape(target_depth.size()), target_depth, weight=target_w / cam_weights_mean) * depth_weight
        bin_weights_sum = outputs['weights'].float()
        bin_width = outputs['ts'][0][:, 1].float()
        bg_weights_sum = 1 - outputs['weights_sum'].flatten()
        entropy_loss = -(torch.sum(bin_weights_sum * (torch.log(bin_weights_sum.clamp(min=1e-6)) - torch.log(bin_width.clamp(min=1e-6))))
                         + torch.sum(bg_weights_sum * (torch.log(bg_weights_sum.clamp(min=1e-6)) - math.log(bg_width)))
                         ) * (sc['entropy'] / target_rgbs.shape[:-1].numel())
        loss = loss + entropy_loss
        if patch_rgb_weight > 0 and nerf.patch_loss is not None:
            loss = loss + nerf.patch_loss(out_rgbs.reshape(target_rgbs.size()).permute(0, 3, 1, 2), target_rgbs.permute(0, 3, 1, 2),
                                          weight=target_w[:, 0, 0, 0] / cam_weights_mean) * patch_rgb_weight
        if use_normal and patch_normal_weight > 0 and nerf.patch_loss is not None and highpass is not None:
            loss = loss + nerf.patch_loss(highpass(out_normals.reshape(target_n.size()).permute(0, 3, 1, 2)),
                                          highpass(target_n.permute(0, 3, 1, 2)),
                                          weight=target_w[:, 0, 0, 0] / cam_weights_mean) * patch_normal_weight
        optimizer.zero_grad(set_to_none=not use_graph)
        loss.backward()
        optimizer.step()
        return torch.stack([loss.detach(), pixel_rgb_loss.detach(), alphas_loss.detach(), normal_reg_loss.detach(), entropy_loss.detach()])

    if prog.get('iteration') is None:
        prog['iteration'] = iteration       # the closure the graph was (or will be) captured from
    iteration = prog['iteration']
    with torch.enable_grad():
        if use_graph:
            if not isinstance(optimizer.param_groups[0]['lr'], torch.Tensor):
                optimizer.param_groups[0]['lr'] = torch.tensor(float(lr), device=device)
            optimizer.param_groups[0]['lr'].fill_(float(lr))
        else:
            optimizer.param_groups[0]['lr'] = lr
        raybatch_inds, num_raybatch = nerf.get_raybatch_inds(tgt_images, n_inverse_rays)
        iter_density = 0
        graph, vals_static = prog['graph'], prog['vals']
        for inverse_step_id in range(inverse_steps):
            if inverse_step_id % nerf.update_extra_interval == 0:
                for _ in range(nerf.update_extra_iters):
                    nerf.decoder.update_extra_state(nerf_code, density_grid, density_bitfield, iter_density, density_thresh=0.1)
            if raybatch_inds is not None:
                inds_static.copy_(raybatch_inds[inverse_step_id % num_raybatch][0][:inds_static.numel()])
            else:
                inds_static.copy_(torch.arange(inds_static.numel(), device=device))
            if not use_graph:
                vals = iteration()
            elif graph is None:
                # iteration 0 runs eagerly on a side stream (allocator warm-up, Adam state creation) and is then captured;
                # iterations >= 1 are replays of the captured graph
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    vals = iteration()
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    vals_static = iteration()
                prog['graph'], prog['vals'] = graph, vals_static
            else:
                graph.replay()
                vals = vals_static
            if debug:
                v = [float(x) for x in vals]
                log.append(dict(loss=v[0], pixel_rgb=v[1], alpha=v[2], normal_reg=v[3], entropy=v[4]))
    nerf.decoder.train(decoder_training_prev)
    return log

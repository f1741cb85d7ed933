"""DMTet / mesh stage of the reconstruction loop: ``init_tet``, the NeRF-field shading function of the mesh
renderer and ``mesh_optim`` -- argument for argument the reference's ``MVEdit3DPipeline.mesh_optim``
(``lib/pipelines/mvedit_3d_pipeline.py:658-872``; ``init_tet``: ``lib/pipelines/utils.py:156-184``; shading functions ``:425-450``).

One iteration = render ``render_bs`` full views of the current DMTet mesh through ``MeshRenderer`` (rasterize / interpolate / antialias
kernels, the hash-grid field kernel at the visible surface points), the objective of Appendix F's mesh column, backward through the same
kernels, one fused Adam step over (field parameters | sdf, deform), marching-tets re-extraction.  The objective's elementwise glue is
torch autograd over those kernels (the NeRF stage has it fused in ``nerf_loss.cu``; the mesh stage's version is not fused yet).

Target normals (``tgt_normals``: TV term against the target's differences, geometry lr without the multiplier, high-passed normal
patch term) are covered by the eager composition.  ``mesh_reduction < 1`` decimates the mesh of the last call by quadric-error edge
collapses (``mesh_renderer.simplify_mesh`` = ``mve_mesh_simplify``, host C++ in place of open3d's ``simplify_quadric_decimation``,
``:829-844``) and fits only the texture afterwards.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn.functional as F

from .mesh_renderer import (DMTet, Mesh, laplacian_smooth_loss, make_tet_grid, min_pool, normal_consistency, simplify_mesh,   # noqa: F401  (re-exported)
                            view_cosine)
from .nerf import blur_masks, highpass, pixel_directions
from ._lib import call, ptr, stream, c_u32, c_f32
from . import view_shard


def init_tet(nerf_model, nerf_code=None, density_thresh=5.0, resolution=128, tets=None):
    """Tet grid fitted to the NeRF's occupied box, SDF initialised from its density (``lib/pipelines/utils.py:156-184``).

    ``tets``: a dict / npz path with ``vertices`` [N,3] in [-0.5, 0.5] and ``indices`` [F,4] (the format of the reference's
    ``demo/tets/{resolution}_tets.npz``, which it downloads when missing); default: ``$MVEDIT_TETS_DIR/{resolution}_tets.npz`` if that
    exists, else the regular grid of ``make_tet_grid(resolution)``.  -> (verts [N,3], indices [F,4] int64, sdf [N])."""
    device = next(nerf_model.decoder.parameters()).device
    if tets is None:
        path = os.path.join(os.environ.get('MVEDIT_TETS_DIR', ''), '%d_tets.npz' % resolution)
        tets = path if os.path.exists(path) else make_tet_grid(resolution, device=device)
    if isinstance(tets, str):
        tets = np.load(tets)
    verts = -torch.as_tensor(np.asarray(tets['vertices']) if not torch.is_tensor(tets['vertices']) else tets['vertices'],
                             dtype=torch.float32, device=device) * 2                       # covers [-1, 1]
    indices = torch.as_tensor(np.asarray(tets['indices']) if not torch.is_tensor(tets['indices']) else tets['indices'],
                              dtype=torch.long, device=device)
    if nerf_code is None:
        nerf_code = [None]
    with torch.no_grad():
        sigma = nerf_model.decoder.point_density_decode([verts], nerf_code)[0]
        valid = verts[sigma > density_thresh]
        if valid.shape[0] == 0:
            raise RuntimeError('init_tet: the field has no density above %g -- nothing to extract' % density_thresh)
        vmax, vmin = valid.amax(dim=0) + 0.1, valid.amin(dim=0) - 0.1
        verts = verts * ((vmax - vmin).max() / 2) + (vmax + vmin) / 2
        sigma = nerf_model.decoder.point_density_decode([verts], nerf_code)[0]
        sdf = (sigma - density_thresh).clamp(-1, 1)
        sdf[(verts < -1).any(dim=-1) | (verts > 1).any(dim=-1)] = -1
    return verts.contiguous(), indices, sdf.contiguous()


def make_nerf_shading_fun(decoder, nerf_code, worldspace_point_lights, ambient_light, tonemapping=None):
    """Lambert shading of the field's albedo at the visible surface points by one point light per view (``:425-442``)."""
    def shading_fun(world_pos=None, albedo=None, world_normal=None, fg_mask=None, **kwargs):
        if len(world_pos) == 0:
            return world_pos if albedo is None else albedo
        base_albedo = decoder.point_decode([world_pos], None, nerf_code)[1]
        fg_lights = worldspace_point_lights[fg_mask.squeeze(0)]
        shading = (fg_lights * world_normal).sum(-1, keepdim=True).clamp(min=0) * (1 - ambient_light) + ambient_light
        if tonemapping is None:
            return base_albedo * shading
        return tonemapping.lut(tonemapping.inverse_lut(base_albedo) + shading.clamp(min=1e-6).log2())
    return shading_fun


def make_nerf_albedo_shading_fun(decoder, nerf_code):
    """Unshaded albedo (``:444-450``; used when baking the texture)."""
    def shading_fun(world_pos=None, albedo=None, **kwargs):
        if len(world_pos) == 0:
            return world_pos if albedo is None else albedo
        return decoder.point_decode([world_pos], None, nerf_code)[1]
    return shading_fun


def view_cosine_gate(inv_depth, dirs):
    """``view_cosine`` eroded by a 5x5 window (``mvedit_3d_pipeline.py:750-756``)."""
    return min_pool(view_cosine(inv_depth, dirs))


def tv_normal_loss(pred, weight, power=1.5, target=None):
    """``TVLoss(power=1.5)`` on NCHW maps (``lib/models/losses/tv_loss.py:7-42``): forward differences along h and w (zero at the far
    border), each weighted by the smaller weight of its two pixels, ``||(dh, dw)||_2 ** power`` averaged over everything."""
    def fdiff(t, dim):
        return F.pad(torch.diff(t, dim=dim), (0, 1) if dim == -1 else (0, 0, 0, 1))

    def wmin(t, dim):
        n = t.size(dim) - 1
        return F.pad(torch.minimum(t.narrow(dim, 0, n), t.narrow(dim, 1, n)), (0, 1) if dim == -1 else (0, 0, 0, 1))
    dh, dw = fdiff(pred, -2), fdiff(pred, -1)
    if target is not None:
        dh, dw = dh - fdiff(target, -2), dw - fdiff(target, -1)
    dh, dw = dh * wmin(weight, -2), dw * wmin(weight, -1)
    return torch.stack([dh, dw], dim=0).norm(dim=0).pow(power).mean()      # norm: zero sub-gradient on the flat background


class _LpipsFn(torch.autograd.Function):
    """Bridges ``LPIPSLoss.loss_and_grad`` (forward and input gradient in one call, no autograd graph) into the mesh objective."""

    @staticmethod
    def forward(ctx, pred_nhwc, target_nhwc, weight, patch_loss):
        loss, g, _ = patch_loss.loss_and_grad(pred_nhwc.detach().float().contiguous(), target_nhwc.detach().float().contiguous(), weight, 1.0)
        ctx.save_for_backward(g)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        (g,) = ctx.saved_tensors
        return g * g_loss, None, None, None


def lpips_patch_loss(patch_loss, pred_nchw, target_nchw, weight):
    """``self.nerf.patch_loss(pred, target, weight=...)`` with gradient to ``pred`` (NCHW in [0,1] as the reference passes them)."""
    if hasattr(patch_loss, 'loss_and_grad'):
        return _LpipsFn.apply(pred_nchw.permute(0, 2, 3, 1), target_nchw.permute(0, 2, 3, 1), weight, patch_loss)
    return patch_loss(pred_nchw, target_nchw, weight=weight)


class _MeshObjectiveFn(torch.autograd.Function):
    """The per-pixel terms of the mesh objective (rgb L1, alpha L1, TV-normal, + the patch term on rgb') as three launches of
    ``csrc/mesh_loss.cu`` -- loss AND gradient w.r.t. the antialiased (rgba, normal) in the forward call, like the NeRF stage's fused
    objective.  Opt-in (``mesh_optim(..., fused_objective=True)``): checked against the eager composition on the CPU through the host
    harness; it has not run on a GPU yet."""

    @staticmethod
    def forward(ctx, rgba, normal, gate, tgt_rgb, m_erode, m_blur, w_view, normal_bg, c_rgb, c_alpha, c_tv, patch):
        f = lambda t: t.detach().to(torch.float32).contiguous()
        rgba_c, normal_c, gate_c, tgt_c, me_c, mb_c, wv_c = f(rgba), f(normal), f(gate), f(tgt_rgb), f(m_erode), f(m_blur), f(w_view)
        bs, h, w, _ = rgba_c.shape
        dev = rgba_c.device
        out_rgb = torch.empty(bs, h, w, 3, dtype=torch.float32, device=dev)
        nfg, g_nfg = torch.empty(bs * h * w, 3, dtype=torch.float32, device=dev), torch.empty(bs * h * w, 3, dtype=torch.float32, device=dev)
        loss = torch.zeros(3, dtype=torch.float32, device=dev)
        nbg = (ctypes.c_float * 3)(*[float(v) for v in normal_bg])
        call('mve_mesh_loss_forward', ptr(rgba_c), ptr(normal_c), ptr(tgt_c), ptr(me_c), ptr(mb_c), ptr(wv_c), nbg, c_u32(bs), c_u32(h), c_u32(w),
             c_f32(c_rgb), c_f32(c_alpha), ptr(out_rgb), ptr(nfg), ptr(loss), stream())
        g_extra, lp = None, None
        if patch is not None:                                # the LPIPS patch term looks at rgb' (:784-801); its gradient is chained below
            patch_loss, pick, ps, wgt, scale = patch
            out_p = _patches(out_rgb, h, ps)[pick].permute(0, 2, 3, 1).contiguous()
            tgt_p = _patches(tgt_c, h, ps)[pick].permute(0, 2, 3, 1).contiguous()
            lp, gp, _ = patch_loss.loss_and_grad(out_p, tgt_p, wgt, scale)
            g = h // ps
            ge = torch.zeros(bs * g * g, ps, ps, 3, dtype=torch.float32, device=dev)
            ge[pick] = gp.to(torch.float32)
            g_extra = ge.reshape(bs, g, g, ps, ps, 3).permute(0, 1, 3, 2, 4, 5).reshape(bs, h, w, 3).contiguous()
        g_rgba, g_normal = torch.empty_like(rgba_c), torch.empty_like(normal_c)
        call('mve_mesh_loss_backward', ptr(rgba_c), ptr(tgt_c), ptr(me_c), ptr(mb_c), ptr(wv_c), ptr(gate_c), nbg, c_u32(bs), c_u32(h), c_u32(w),
             c_f32(c_rgb), c_f32(c_alpha), c_f32(c_tv), ptr(nfg), ptr(g_extra), ptr(g_nfg), ptr(loss), ptr(g_rgba), ptr(g_normal), stream())
        ctx.save_for_backward(g_rgba, g_normal)
        total = loss.sum()
        return total if lp is None else total + lp

    @staticmethod
    def backward(ctx, g):
        g_rgba, g_normal = ctx.saved_tensors
        return (g_rgba * g, g_normal * g) + (None,) * 10


def normalize_depth(depths, alphas, far_depth=0.25, alpha_clip=0.5, eps=1e-5):
    """Inverse depth [n,h,w] + alpha [n,h,w,1] -> the ControlNet depth image in [0,1] (``geometry_utils.normalize_depth``): per view,
    foreground inverse depth rescaled so that the farthest confident (alpha >= alpha_clip) pixel maps to ``far_depth`` and the nearest to 1."""
    a = alphas.squeeze(-1)
    d_max = depths.flatten(1).amax(dim=1)[:, None, None]
    d_fg = depths / a.clamp(min=eps)
    d_min = d_fg.masked_fill(a < alpha_clip, 1 / eps).flatten(1).amin(dim=1)[:, None, None]
    d_fg = (d_fg - d_min) / (d_max - d_min).clamp(min=eps) * (1 - far_depth) + far_depth
    return (d_fg * a).clamp(min=0, max=1)


def render_mesh_views(self, in_mesh, nerf_code, camera_poses, intrinsics, intrinsics_size, render_size, lights, ambient_light, render_bs):
    """The mesh branch of the per-step render (``mvedit_3d_pipeline.py:1341-1360,1391-1396``) -> (ctrl_images, ctrl_depths) bf16
    [n,3,rs,rs] like ``MVEdit3DStep.render_views``."""
    images, alphas, depths = [], [], []
    for pose_b, intr_b, light_b in zip(camera_poses.split(render_bs), intrinsics.split(render_bs), lights.split(render_bs)):
        out = self.mesh_renderer(
            [in_mesh], pose_b[None], intr_b[None] * (render_size / intrinsics_size), render_size, render_size,
            make_nerf_shading_fun(self.nerf.decoder, nerf_code, light_b[:, None, None, :].expand(-1, render_size, render_size, -1),
                                  ambient_light, self.tonemapping), normal_bg=self.normal_bg)
        rgba = out['rgba'].squeeze(0).detach()
        images.append(rgba[..., :3] + self.nerf.bg_color * (1 - rgba[..., 3:]))
        alphas.append(rgba[..., 3:])
        depths.append(out['depth'].squeeze(0).detach())
    images = torch.cat(images, dim=0).permute(0, 3, 1, 2).clamp(min=0, max=1).to(torch.bfloat16)
    alphas, depths = torch.cat(alphas, dim=0), torch.cat(depths, dim=0)
    depths = normalize_depth(depths, alphas).to(torch.bfloat16).unsqueeze(1).repeat(1, 3, 1, 1)
    return images.contiguous(), depths.contiguous()


def _from_rank0(t):
    """Rank 0's draw on every rank (random permutations / jitter must agree across the data-parallel replicas)."""
    import torch.distributed as dist
    t = t.contiguous()
    dist.broadcast(t, 0)
    return t


def _patches(x_nhwc, render_size, patch_size):
    """[n, rs, rs, C] -> [n * (rs/ps)^2, C, ps, ps] (``:784-792``)."""
    g = render_size // patch_size
    c = x_nhwc.shape[-1]
    return x_nhwc.reshape(-1, g, patch_size, g, patch_size, c).permute(0, 1, 3, 5, 2, 4).reshape(-1, c, patch_size, patch_size)


def mesh_optim(self, tgt_images, tgt_masks, tgt_normals,                                                   # input images
               optimizer, lr, lr_multiplier, inverse_steps, render_bs, patch_bs, mesh_simplify_texture_steps,  # optimisation settings
               patch_rgb_weight, patch_normal_weight, alpha_soften, normal_reg_weight, mesh_normal_reg_weight,  # loss weights
               nerf_code, tet_verts, deform, tet_sdf, tet_indices, dmtet, in_mesh,                          # mesh model
               render_size, intrinsics, intrinsics_size, camera_poses, cam_weights, lights, patch_size,     # cameras
               is_end, ambient_light, mesh_reduction, debug=False, perturb=True, noise=None, fused_objective=None):
    """``self``: the pipeline (``nerf``, ``mesh_renderer``, ``normal_bg``, ``tonemapping``).  ``noise`` (extension, for parity tests):
    dict with ``camera_perm`` [n], ``jitter`` [steps, render_bs, 2] in [0,1), ``patch_perm`` [steps, n_patches] replacing the draws.

    Multi-GPU (``nerf.data_parallel`` with torch.distributed up; the reference is single-GPU): data-parallel over VIEWS -- every rank
    renders its block of the iteration's ``render_bs`` views, the per-view loss terms are weighted by the block's share and the
    replicated regularisers by 1 / world, so that ONE all-reduce (sum) of the gradients reproduces the single-GPU gradient; the fused
    Adam step and the marching-tets extraction then run replicated and bit-identical.  Random draws are rank 0's (one small broadcast).

    ``fused_objective`` (default: ``self.mesh_fused_objective`` if set, else False): the per-pixel loss terms and their gradient as three
    launches of ``csrc/mesh_loss.cu`` instead of ~60 eager torch ops + autograd (same values; CPU-checked, not yet run on a GPU)."""
    use_normal = tgt_normals is not None                          # :667
    use_pn = use_normal and patch_normal_weight > 0               # :806
    nerf, dec = self.nerf, self.nerf.decoder
    device = tet_verts.device
    noise = noise or {}
    cam_weights_mean = cam_weights.mean()
    n_views = camera_poses.size(0)
    tgt_masks_blur = blur_masks(tgt_masks.square().squeeze(0).permute(0, 3, 1, 2), 9, 1.5).permute(0, 2, 3, 1).clamp(
        min=alpha_soften ** 2, max=(1 - alpha_soften) ** 2).sqrt()
    directions = pixel_directions(intrinsics * (render_size / intrinsics_size), render_size, render_size)      # [n, h, w, 3]
    normal_bg = tgt_images.new_tensor(self.normal_bg)
    decoder_training_prev = dec.training
    dec.train(True)
    fused = hasattr(optimizer, 'set_lr')
    sink_prev = dec.grad_sink
    dec.grad_sink = optimizer if fused and hasattr(optimizer, 'grad_sink') else None
    rank, world = view_shard.world() if getattr(nerf, 'data_parallel', False) else (0, 1)
    shared = (lambda t: _from_rank0(t)) if world > 1 else (lambda t: t)
    if fused_objective is None:
        fused_objective = bool(getattr(self, 'mesh_fused_objective', False))
    if fused_objective and use_normal:
        raise NotImplementedError('mesh_optim: the fused objective kernels (mesh_loss.cu) do not take target normals; use fused_objective=False')
    try:
        with torch.enable_grad():
            if fused:
                optimizer.set_lr(lr, group=0)
                optimizer.set_lr(lr * (0.04 if use_normal else 0.04 * lr_multiplier), group=1)          # :688
            else:
                optimizer.param_groups[0]['lr'] = lr
                optimizer.param_groups[1]['lr'] = lr * (0.04 if use_normal else 0.04 * lr_multiplier)
            camera_perm = noise['camera_perm'].to(device) if 'camera_perm' in noise else shared(torch.randperm(n_views, device=device))
            split = lambda x: x[camera_perm].split(render_bs, dim=0)
            pose_b, intr_b = split(camera_poses), split(intrinsics)
            img_b, mask_b, blur_b = split(tgt_images.squeeze(0)), split(tgt_masks.squeeze(0)), split(tgt_masks_blur)
            dir_b, w_b, light_b = split(directions), split(cam_weights), split(lights)
            nrm_b = split(tgt_normals.squeeze(0)) if use_normal else None
            nb = len(pose_b)
            if is_end:
                inverse_steps = max(inverse_steps, mesh_simplify_texture_steps)
            mesh_is_simplified = False                          # (:713) set by the decimation of the last call
            for step in range(inverse_steps):
                k = step % nb
                bs = img_b[k].shape[0]
                lo, hi = view_shard.local_range(bs, rank, world) if world > 1 else (0, bs)
                share = (hi - lo) / bs                       # this rank's share of the per-view means
                intrinsics_batch = intr_b[k] * (render_size / intrinsics_size)
                if perturb:                                  # +-0.5 px principal-point jitter (:733-735)
                    u = noise['jitter'][step, :bs].to(device) if 'jitter' in noise else shared(torch.rand_like(intrinsics_batch[:, 2:]))
                    intrinsics_batch = torch.cat([intrinsics_batch[:, :2], intrinsics_batch[:, 2:] + (u - 0.5) / self.mesh_renderer.ssaa], dim=1)
                geo = 0.0 if mesh_is_simplified else 1.0         # a decimated mesh is fixed: only the colour terms remain (:764-779, :806)
                loss = 0.0 if mesh_is_simplified else (laplacian_smooth_loss(in_mesh.v, in_mesh.f)
                                                        + normal_consistency(in_mesh.face_normals, in_mesh.f)) * (mesh_normal_reg_weight / world)
                if hi > lo:
                    target_rgbs, target_m, target_m_blur, target_dir = img_b[k][lo:hi], mask_b[k][lo:hi], blur_b[k][lo:hi], dir_b[k][lo:hi]
                    target_m_erode = min_pool(target_m)
                    target_w = w_b[k][lo:hi, None, None, None].expand(-1, render_size, render_size, 1)
                    target_lights = light_b[k][lo:hi, None, None, :].expand(-1, render_size, render_size, 3)
                    render_out = self.mesh_renderer(
                        [in_mesh], pose_b[k][lo:hi][None], intrinsics_batch[lo:hi][None], render_size, render_size,
                        make_nerf_shading_fun(dec, nerf_code, target_lights, ambient_light, self.tonemapping), normal_bg=self.normal_bg)
                    rgba = render_out['rgba'].squeeze(0)
                    gate = view_cosine_gate(render_out['depth'].squeeze(0).detach(), target_dir)
                    wgt = target_w / cam_weights_mean
                    grid_n = render_size // patch_size
                    def draw(key):                               # which patches an LPIPS term looks at (:793, :813)
                        if world > 1:                            # each rank draws its share of the patches among its own views
                            return torch.randperm((hi - lo) * grid_n * grid_n, device=device)[:max(patch_bs * (hi - lo) // bs, 1)]
                        perm = noise[key][step].to(device) if key in noise else torch.randperm((hi - lo) * grid_n * grid_n, device=device)
                        return perm[:patch_bs]
                    if patch_rgb_weight > 0 or use_pn:
                        pick = draw('patch_perm')
                        w_pick = _patches(target_w, render_size, patch_size)[pick, 0, 0, 0] / cam_weights_mean
                    if fused_objective:
                        n_px = (hi - lo) * render_size * render_size
                        lw = float(nerf.pixel_loss.loss_weight)
                        views = _MeshObjectiveFn.apply(
                            rgba, render_out['normal'].squeeze(0), gate.squeeze(-1), target_rgbs, target_m_erode.squeeze(-1), target_m_blur.squeeze(-1),
                            w_b[k][lo:hi] / cam_weights_mean, self.normal_bg, lw * 4.5 / (n_px * 3), geo * lw * 2.0 / n_px,
                            geo * normal_reg_weight * 2 / (n_px * 3),
                            (nerf.patch_loss, pick, patch_size, w_pick, patch_rgb_weight) if patch_rgb_weight > 0 else None)
                    else:
                        out_alphas = rgba[..., 3:]
                        out_rgbs = rgba[..., :3] / out_alphas.clamp(min=1e-3)
                        out_rgbs = out_rgbs * target_m_erode + target_rgbs * (1 - target_m_erode)
                        out_normals = render_out['normal'].squeeze(0)
                        out_normals = out_normals * gate + out_normals.detach() * (1 - gate)      # value unchanged, gradient scaled by the gate
                        out_normals_fg = (out_normals - normal_bg * (1 - out_alphas)) / out_alphas.clamp(min=1e-3)
                        views = nerf.pixel_loss(out_rgbs, target_rgbs, weight=wgt) * 4.5
                        target_n = nrm_b[k][lo:hi] if use_normal else None
                        if not mesh_is_simplified:
                            views = views + nerf.pixel_loss(out_alphas, target_m_blur, weight=wgt) * 2.0
                            views = views + tv_normal_loss(out_normals_fg.permute(0, 3, 1, 2), out_alphas.detach().permute(0, 3, 1, 2),
                                                           target=target_n.permute(0, 3, 1, 2) if use_normal else None) * (normal_reg_weight * 2)
                        if patch_rgb_weight > 0:
                            out_p, tgt_p = _patches(out_rgbs, render_size, patch_size), _patches(target_rgbs, render_size, patch_size)
                            views = views + lpips_patch_loss(nerf.patch_loss, out_p[pick], tgt_p[pick], w_pick) * patch_rgb_weight
                        if use_pn and not mesh_is_simplified:        # high-passed normal patch term (:806-821): its own patch draw, the rgb draw's weights (as the reference)
                            pick_n = draw('patch_perm_normal')
                            out_np, tgt_np = _patches(out_normals, render_size, patch_size), _patches(target_n, render_size, patch_size)
                            views = views + lpips_patch_loss(nerf.patch_loss, highpass(out_np[pick_n]), highpass(tgt_np[pick_n]), w_pick) * patch_normal_weight
                    loss = loss + views * share

                optimizer.zero_grad()
                if torch.is_tensor(loss):                    # (a rank without a view of this batch has nothing to add on a decimated mesh)
                    loss.backward()
                if world > 1:
                    if fused and hasattr(optimizer, 'flat_grad'):
                        optimizer.fold_grads()
                        view_shard.allreduce_flat(optimizer.flat_grad)
                    else:
                        view_shard.allreduce_grads([p_ for g_ in optimizer.param_groups for p_ in g_['params']])
                optimizer.step()

                if not mesh_is_simplified:
                    with torch.enable_grad():
                        mesh_verts, mesh_faces = dmtet(tet_verts + deform, tet_sdf, tet_indices)
                        if mesh_reduction < 1 and is_end and (inverse_steps - (step + 1)) <= mesh_simplify_texture_steps:
                            # (:829-844) decimate once, then fit only the texture: the field's parameters under a fresh optimiser.  The
                            # reference calls open3d on the CPU here; so does mve_mesh_simplify (host C++, identical on every rank)
                            mesh_verts, mesh_faces = simplify_mesh(mesh_verts.detach(), mesh_faces, round(mesh_faces.shape[0] * mesh_reduction))
                            mesh_verts, indices = torch.unique(mesh_verts, dim=0, return_inverse=True, sorted=False)
                            mesh_faces = indices[mesh_faces]
                            optimizer = optimizer.__class__(list(dec.parameters()), lr=lr)
                            if dec.grad_sink is not None:
                                dec.grad_sink = optimizer
                            mesh_is_simplified = True
                        in_mesh = Mesh(v=mesh_verts, f=mesh_faces.int(), device=device)
                        in_mesh.auto_normal()
                if debug:
                    print('mesh_optim step %d: loss %.5f, %d vertices, %d faces' % (step, float(loss), mesh_verts.shape[0], mesh_faces.shape[0]))
    finally:
        dec.grad_sink = sink_prev
        dec.train(decoder_training_prev)
    return in_mesh


def texture_optim(self, tgt_images,                                                   # input images
                  optimizer, lr, inverse_steps, render_bs, patch_bs,                  # optimisation settings
                  patch_rgb_weight,                                                   # loss weights
                  nerf_code, in_mesh,                                                 # mesh model
                  render_size, intrinsics, intrinsics_size, camera_poses, cam_weights_dense, patch_size,
                  debug=False, perturb=True, noise=None, patch_views=None):
    """Fit the field's albedo on a FIXED mesh to the target views (``MVEditTexturePipeline.texture_optim``,
    ``lib/pipelines/mvedit_texture_pipeline.py:93-172``; the super-resolution pipeline's copy is identical,
    ``mvedit_texture_superres_pipeline.py:89-168``): render ``render_bs`` views through ``MeshRenderer`` with the unshaded field albedo at the
    surface points, composite on ``self.bg_color``, L1 x 2 with dense per-pixel camera weights [n,h,w,1] + the LPIPS patch term.
    ``noise`` (extension, parity tests): ``camera_perm``, ``jitter`` [steps, render_bs, 2], ``patch_perm`` [steps, n_patches].
    ``patch_views``: the super-resolution variant's ``num_cameras`` -- its patch term only looks at the first ``num_cameras`` views of every
    rendered batch (``mvedit_texture_superres_pipeline.py:139-148``, "ignore regularization views")."""
    nerf, dec = self.nerf, self.nerf.decoder
    device = camera_poses.device
    noise = noise or {}
    decoder_training_prev = dec.training
    dec.train(True)
    fused = hasattr(optimizer, 'set_lr')
    sink_prev = dec.grad_sink
    dec.grad_sink = optimizer if fused and hasattr(optimizer, 'grad_sink') else None
    shading = make_nerf_albedo_shading_fun(dec, nerf_code)
    try:
        with torch.enable_grad():
            if fused:
                optimizer.set_lr(lr, group=0)
            else:
                optimizer.param_groups[0]['lr'] = lr
            camera_perm = noise['camera_perm'].to(device) if 'camera_perm' in noise else torch.randperm(camera_poses.size(0), device=device)
            split = lambda x: x[camera_perm].split(render_bs, dim=0)
            pose_b, intr_b, img_b, w_b = split(camera_poses), split(intrinsics), split(tgt_images.squeeze(0)), split(cam_weights_dense)
            nb = len(pose_b)
            for step in range(inverse_steps):
                k = step % nb
                target_rgbs, target_w = img_b[k], w_b[k]
                intrinsics_batch = intr_b[k] * (render_size / intrinsics_size)
                if perturb:
                    u = noise['jitter'][step, :target_rgbs.shape[0]].to(device) if 'jitter' in noise else torch.rand_like(intrinsics_batch[:, 2:])
                    intrinsics_batch = torch.cat([intrinsics_batch[:, :2], intrinsics_batch[:, 2:] + (u - 0.5) / self.mesh_renderer.ssaa], dim=1)
                rgba = self.mesh_renderer([in_mesh], pose_b[k][None], intrinsics_batch[None], render_size, render_size, shading)['rgba'].squeeze(0)
                out_rgbs = rgba[..., :3] + (1 - rgba[..., 3:].clamp(min=1e-3)) * self.bg_color
                loss = nerf.pixel_loss(out_rgbs, target_rgbs, weight=target_w) * 2
                if patch_rgb_weight > 0:
                    pv = slice(None, patch_views)
                    out_p, tgt_p = _patches(out_rgbs[pv], render_size, patch_size), _patches(target_rgbs[pv], render_size, patch_size)
                    w_p = _patches(target_w[pv], render_size, patch_size)
                    perm = noise['patch_perm'][step].to(device) if 'patch_perm' in noise else torch.randperm(out_p.size(0), device=device)
                    pick = perm[:patch_bs]
                    loss = loss + lpips_patch_loss(nerf.patch_loss, out_p[pick], tgt_p[pick], w_p[pick].amax(dim=(1, 2, 3))) * patch_rgb_weight
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
                if debug:
                    print('texture_optim step %d: loss %.5f' % (step, float(loss)))
    finally:
        dec.grad_sink = sink_prev
        dec.train(decoder_training_prev)

"""CPU oracle of the mesh stage -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline).

The reference's Python restated op by op, on top of oracle/raster_oracle.py for the four nvdiffrast ops:
  ``DMTetOracle``             lib/models/decoders/mesh_renderer/base_mesh_renderer.py:104-188 (per-call torch.unique, as the reference)
  ``auto_normal``             mesh_utils.py:359-382
  ``laplacian_smooth_loss``   base_mesh_renderer.py:71-101 (sparse uniform Laplacian), ``normal_consistency`` :22-68
  ``mesh_renderer_forward``   base_mesh_renderer.py:207-299, 383-395 (single-scene branch)
  ``mesh_optim``              lib/pipelines/mvedit_3d_pipeline.py:658-872, every random draw supplied
Pinned by: tests/golden/mesh_pins.npz for DMTet / auto_normal / the regularisers (the reference's own classes run in the build
container); the renderer / rasteriser part is a restatement of an absent dependency (nvdiffrast) -- **parity unpinned**.
"""
import types

import numpy as np
import torch
import torch.nn.functional as F

from . import raster_oracle as ro
from .nerf_oracle import L1LossMod, TVLoss, depth_to_normal, gaussian_blur, get_ray_directions


class DMTetOracle:
    def __init__(self, device='cpu'):
        self.device = device
        self.triangle_table = torch.tensor([
            [-1, -1, -1, -1, -1, -1], [1, 0, 2, -1, -1, -1], [4, 0, 3, -1, -1, -1], [1, 4, 2, 1, 3, 4], [3, 1, 5, -1, -1, -1],
            [2, 3, 0, 2, 5, 3], [1, 4, 0, 1, 5, 4], [4, 2, 5, -1, -1, -1], [4, 5, 2, -1, -1, -1], [4, 1, 0, 4, 5, 1],
            [3, 2, 0, 3, 5, 2], [1, 3, 5, -1, -1, -1], [4, 1, 2, 4, 3, 1], [3, 0, 4, -1, -1, -1], [2, 0, 1, -1, -1, -1],
            [-1, -1, -1, -1, -1, -1]], dtype=torch.long, device=device)
        self.num_triangles_table = torch.tensor([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0], dtype=torch.long, device=device)
        self.base_tet_edges = torch.tensor([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=torch.long, device=device)

    def __call__(self, pos_nx3, sdf_n, tet_fx4):
        with torch.no_grad():
            occ_n = sdf_n > 0
            occ_fx4 = occ_n[tet_fx4.reshape(-1)].reshape(-1, 4)
            occ_sum = torch.sum(occ_fx4, -1)
            valid_tets = (occ_sum > 0) & (occ_sum < 4)
            all_edges = tet_fx4[valid_tets][:, self.base_tet_edges].reshape(-1, 2)
            order = (all_edges[:, 0] > all_edges[:, 1]).long().unsqueeze(1)
            all_edges = torch.stack([torch.gather(all_edges, 1, order), torch.gather(all_edges, 1, 1 - order)], -1).reshape(-1, 2)
            unique_edges, idx_map = torch.unique(all_edges, dim=0, return_inverse=True)
            unique_edges = unique_edges.long()
            mask_edges = occ_n[unique_edges.reshape(-1)].reshape(-1, 2).sum(-1) == 1
            mapping = torch.ones((unique_edges.shape[0]), dtype=torch.long, device=self.device) * -1
            mapping[mask_edges] = torch.arange(mask_edges.sum(), dtype=torch.long, device=self.device)
            idx_map = mapping[idx_map]
            interp_v = unique_edges[mask_edges]
        edges_to_interp = pos_nx3[interp_v.reshape(-1)].reshape(-1, 2, 3)
        edges_to_interp_sdf = sdf_n[interp_v.reshape(-1)].reshape(-1, 2, 1)
        edges_to_interp_sdf = torch.cat([edges_to_interp_sdf[:, :1], -edges_to_interp_sdf[:, 1:]], dim=1)
        denominator = edges_to_interp_sdf.sum(1, keepdim=True)
        edges_to_interp_sdf = torch.flip(edges_to_interp_sdf, [1]) / denominator
        verts = (edges_to_interp * edges_to_interp_sdf).sum(1)
        idx_map = idx_map.reshape(-1, 6)
        v_id = torch.pow(2, torch.arange(4, dtype=torch.long, device=self.device))
        tetindex = (occ_fx4[valid_tets] * v_id.unsqueeze(0)).sum(-1)
        num_triangles = self.num_triangles_table[tetindex]
        faces = torch.cat((
            torch.gather(input=idx_map[num_triangles == 1], dim=1, index=self.triangle_table[tetindex[num_triangles == 1]][:, :3]).reshape(-1, 3),
            torch.gather(input=idx_map[num_triangles == 2], dim=1, index=self.triangle_table[tetindex[num_triangles == 2]][:, :6]).reshape(-1, 3),
        ), dim=0)
        return verts, faces


def make_mesh(v, f):
    m = types.SimpleNamespace(v=v, f=f, vn=None, fn=None, vt=None, ft=None, vc=None, albedo=None, face_normals=None)
    auto_normal(m)
    return m


def auto_normal(mesh):
    i0, i1, i2 = mesh.f[:, 0].long(), mesh.f[:, 1].long(), mesh.f[:, 2].long()
    v0, v1, v2 = mesh.v[i0, :], mesh.v[i1, :], mesh.v[i2, :]
    face_normals = F.normalize(torch.cross(v1 - v0, v2 - v0, dim=-1), dim=-1)
    vn = torch.zeros_like(mesh.v)
    vn = vn.scatter_add(0, i0[:, None].repeat(1, 3), face_normals)
    vn = vn.scatter_add(0, i1[:, None].repeat(1, 3), face_normals)
    vn = vn.scatter_add(0, i2[:, None].repeat(1, 3), face_normals)
    mesh.vn = F.normalize(vn, dim=-1)
    mesh.fn = mesh.f.to(torch.int32)
    mesh.face_normals = face_normals


def compute_edge_to_face_mapping(attr_idx):
    with torch.no_grad():
        all_edges = torch.cat((torch.stack((attr_idx[:, 0], attr_idx[:, 1]), dim=-1), torch.stack((attr_idx[:, 1], attr_idx[:, 2]), dim=-1),
                               torch.stack((attr_idx[:, 2], attr_idx[:, 0]), dim=-1)), dim=-1).view(-1, 2)
        order = (all_edges[:, 0] > all_edges[:, 1]).long().unsqueeze(dim=1)
        sorted_edges = torch.cat((torch.gather(all_edges, 1, order), torch.gather(all_edges, 1, 1 - order)), dim=-1)
        unique_edges, idx_map = torch.unique(sorted_edges, dim=0, return_inverse=True)
        tris = torch.arange(attr_idx.shape[0]).repeat_interleave(3)
        tris_per_edge = torch.zeros((unique_edges.shape[0], 2), dtype=torch.int64)
        mask0, mask1 = order[:, 0] == 0, order[:, 0] == 1
        tris_per_edge[idx_map[mask0], 0] = tris[mask0]
        tris_per_edge[idx_map[mask1], 1] = tris[mask1]
        return tris_per_edge


def normal_consistency(face_normals, t_pos_idx):
    tris_per_edge = compute_edge_to_face_mapping(t_pos_idx.long())
    n0, n1 = face_normals[tris_per_edge[:, 0], :], face_normals[tris_per_edge[:, 1], :]
    term = 1.0 - torch.clamp(torch.sum(n0 * n1, -1, keepdim=True), min=-1.0, max=1.0)
    return torch.mean(torch.abs(term))


def laplacian_smooth_loss(verts, faces):
    with torch.no_grad():
        faces = faces.long()
        V = verts.shape[0]
        ii, jj = faces[:, [1, 2, 0]].flatten(), faces[:, [2, 0, 1]].flatten()
        adj = torch.stack([torch.cat([ii, jj]), torch.cat([jj, ii])], dim=0).unique(dim=1)
        adj_values = torch.ones(adj.shape[1], dtype=verts.dtype)
        idx = torch.cat((adj, torch.stack((adj[0], adj[0]), dim=0)), dim=1)
        L = torch.sparse_coo_tensor(idx, torch.cat((-adj_values, adj_values)), (V, V)).coalesce()
    return L.mm(verts).norm(dim=1).mean()


# ---- the four rasteriser ops with autograd, built from raster_oracle ---------------------------------------------------------------

def dr_rasterize(pos, tri, resolution, grad_db=True):
    """ids / coverage from the fp32 numpy rasteriser (the discrete decision), (u, v, z/w) re-evaluated differentiably in pos's dtype."""
    rast_np, db_np = ro.rasterize(pos.detach().float().numpy(), np.asarray(tri), resolution, grad_db=grad_db)
    ids = torch.from_numpy(rast_np[..., 3]).long() - 1
    uvz = ro.barycentrics(pos, torch.as_tensor(np.asarray(tri)), ids)
    rast = torch.cat([uvz, torch.from_numpy(rast_np[..., 3:]).to(pos.dtype)], dim=-1)
    return rast, torch.from_numpy(db_np).to(pos.dtype)


def mesh_renderer_forward(mesh, poses, intrinsics, h, w, shading_fun=None, normal_bg=(0.5, 0.5, 1.0), aa=True, near=0.01, far=100.0, ssaa=1,
                          dilate_edges=0):
    """base_mesh_renderer.py:207-299,383-395 for num_scenes == 1.  poses [1,n,3|4,4], intrinsics [1,n,4]."""
    num_scenes, num_images = poses.shape[:2]
    if ssaa > 1:
        h, w, intrinsics = h * ssaa, w * ssaa, intrinsics * ssaa
    r_mat_c2w = torch.cat([poses[..., :3, :1], -poses[..., :3, 1:3]], dim=-1)
    proj = poses.new_zeros([num_scenes, num_images, 4, 4])
    proj[..., 0, 0] = 2 * intrinsics[..., 0] / w
    proj[..., 0, 2] = -2 * intrinsics[..., 2] / w + 1
    proj[..., 1, 1] = -2 * intrinsics[..., 1] / h
    proj[..., 1, 2] = -2 * intrinsics[..., 3] / h + 1
    proj[..., 2, 2] = -(far + near) / (far - near)
    proj[..., 2, 3] = -(2 * far * near) / (far - near)
    proj[..., 3, 2] = -1
    v_cam = (mesh.v - poses[0, :, :3, 3].unsqueeze(-2)) @ r_mat_c2w[0]
    v_clip = F.pad(v_cam, pad=(0, 1), mode='constant', value=1.0) @ proj[0].transpose(-1, -2)
    tri = mesh.f
    rast, rast_db = dr_rasterize(v_clip, tri, (h, w))
    fg = (rast[..., 3] > 0).unsqueeze(0)
    alpha = fg.to(v_clip.dtype).unsqueeze(-1)
    depth = 1 / ro.interpolate(-v_cam[..., 2:3], rast, tri)[0].reshape(num_scenes, num_images, h, w)
    depth = depth.masked_fill(~fg, 0)
    normal = ro.interpolate(mesh.vn.unsqueeze(0), rast, mesh.fn)[0].reshape(num_scenes, num_images, h, w, 3)
    normal = F.normalize(normal, dim=-1)
    rot_normal = (normal @ r_mat_c2w.unsqueeze(2)) / 2 + 0.5
    rot_normal = torch.where(fg.unsqueeze(-1), rot_normal, rot_normal.new_tensor(list(normal_bg)))
    if mesh.vt is not None and mesh.albedo is not None:
        texc, texc_db = ro.interpolate(mesh.vt.unsqueeze(0), rast, mesh.ft, rast_db=rast_db, diff_attrs='all')
        albedo = ro.texture(mesh.albedo.unsqueeze(0)[..., :3], texc.detach(), texc_db, filter_mode='linear-mipmap-linear').unsqueeze(0)
        albedo = torch.where(fg.unsqueeze(-1), albedo, torch.zeros_like(albedo))
    elif mesh.vc is not None:
        rgba = ro.interpolate(mesh.vc if mesh.vc.dim() == 3 else mesh.vc[None], rast, tri)[0].reshape(num_scenes, num_images, h, w, 4)
        alpha = alpha * rgba[..., 3:4]
        albedo = rgba[..., :3] * alpha
    else:
        albedo = torch.zeros_like(rot_normal)
    if shading_fun is not None:
        xyz = ro.interpolate(mesh.v.unsqueeze(0), rast, tri)[0].reshape(num_scenes, num_images, h, w, 3)
        rgb_reshade = shading_fun(world_pos=xyz[fg], albedo=albedo[fg], world_normal=normal[fg], fg_mask=fg)
        albedo = torch.zeros_like(albedo).masked_scatter(fg.unsqueeze(-1).expand_as(albedo), rgb_reshade.to(albedo.dtype))
    rgba = torch.cat([albedo, alpha], dim=-1)
    if dilate_edges > 0:
        rgba = rgba.reshape(num_scenes * num_images, h, w, 4).permute(0, 3, 1, 2)
        rgba = edge_dilation(rgba, rgba[:, 3:], dilate_edges).permute(0, 2, 3, 1).reshape(num_scenes, num_images, h, w, 4)
    if aa:
        rgba, depth, rot_normal = ro.antialias(torch.cat([rgba, depth.unsqueeze(-1), rot_normal], dim=-1).squeeze(0), rast, v_clip,
                                               np.asarray(tri)).unsqueeze(0).split([4, 1, 3], dim=-1)
        depth = depth.squeeze(-1)
    if ssaa > 1:
        def down(x):
            b = x.shape[:-3]
            y = F.interpolate(x.reshape(b.numel(), *x.shape[-3:]).permute(0, 3, 1, 2), scale_factor=1 / ssaa, mode='area').permute(0, 2, 3, 1)
            return y.reshape(*b, *y.shape[1:])
        rgba, depth, rot_normal = down(rgba), down(depth.unsqueeze(-1)).squeeze(-1), down(rot_normal)
    return dict(rgba=rgba, depth=depth, normal=rot_normal)


def mesh_optim(decoder, tgt_images, tgt_masks, optimizer, lr, lr_multiplier, inverse_steps, render_bs, patch_bs, patch_rgb_weight,
               alpha_soften, normal_reg_weight, mesh_normal_reg_weight, nerf_code, tet_verts, deform, tet_sdf, tet_indices, dmtet, in_mesh,
               render_size, intrinsics, intrinsics_size, camera_poses, cam_weights, lights, patch_size, ambient_light, noise,
               pixel_loss=None, patch_loss=None, normal_bg=(0.5, 0.5, 1.0), tonemapping=None, near=0.01, far=100.0,
               tgt_normals=None, patch_normal_weight=0.0):
    """mvedit_3d_pipeline.py:658-872 without simplification; ``noise``: camera_perm, jitter [steps, bs, 2], patch_perm (and
    patch_perm_normal for the high-passed normal patch term of ``tgt_normals``)."""
    from .nerf_oracle import highpass
    use_normal = tgt_normals is not None
    pixel_loss = pixel_loss or L1LossMod(loss_weight=1.2)
    loss_tv = TVLoss(loss_weight=1.0, power=1.5)
    cam_weights_mean = cam_weights.mean()
    tgt_masks_blur = gaussian_blur(tgt_masks.square().squeeze(0).permute(0, 3, 1, 2), 9, 1.5).permute(0, 2, 3, 1)[None].clamp(
        min=alpha_soften ** 2, max=(1 - alpha_soften) ** 2).sqrt()
    directions = get_ray_directions(render_size, render_size, intrinsics[None] * (render_size / intrinsics_size), norm=False)
    normal_bg_t = tgt_images.new_tensor(list(normal_bg))
    optimizer.param_groups[0]['lr'] = lr
    optimizer.param_groups[1]['lr'] = lr * (0.04 if use_normal else 0.04 * lr_multiplier)
    camera_perm = noise['camera_perm']
    sp = lambda x: x[camera_perm].split(render_bs, dim=0)
    pose_batches, intrinsics_batches = sp(camera_poses), sp(intrinsics)
    tgt_image_batches, tgt_mask_batches = sp(tgt_images.squeeze(0)), sp(tgt_masks.squeeze(0))
    tgt_mask_blur_batches, tgt_dir_batches = sp(tgt_masks_blur.squeeze(0)), sp(directions.squeeze(0))
    cam_weights_batches, lights_batches = sp(cam_weights), sp(lights)
    tgt_normals_batches = sp(tgt_normals.squeeze(0)) if use_normal else None
    nb = len(pose_batches)
    losses = []
    for step in range(inverse_steps):
        k = step % nb
        pose_batch, intrinsics_batch = pose_batches[k], intrinsics_batches[k]
        target_rgbs, target_m, target_m_blur = tgt_image_batches[k], tgt_mask_batches[k], tgt_mask_blur_batches[k]
        target_m_erode = -F.max_pool2d(-target_m.permute(0, 3, 1, 2), 5, stride=1, padding=2).permute(0, 2, 3, 1)
        target_dir = tgt_dir_batches[k]
        target_w = cam_weights_batches[k][:, None, None, None].expand(-1, render_size, render_size, 1)
        target_lights = lights_batches[k][:, None, None, :].expand(-1, render_size, render_size, 3)
        intrinsics_batch = intrinsics_batch * (render_size / intrinsics_size)
        intrinsics_batch = torch.cat([intrinsics_batch[:, :2], intrinsics_batch[:, 2:] + (noise['jitter'][step, :len(pose_batch)] - 0.5)], dim=1)

        def shading_fun(world_pos=None, albedo=None, world_normal=None, fg_mask=None, **kwargs):
            if len(world_pos) == 0:
                return world_pos if albedo is None else albedo
            base_albedo = decoder.point_decode([world_pos], None, nerf_code)[1]
            fg_lights = target_lights[fg_mask.squeeze(0)]
            shading = ((fg_lights[:, None, :] @ world_normal[:, :, None]).clamp(min=0) * (1 - ambient_light) + ambient_light).squeeze(-1)
            if tonemapping is None:
                return base_albedo * shading
            return tonemapping.lut(tonemapping.inverse_lut(base_albedo) + shading.clamp(min=1e-6).log2())

        render_out = mesh_renderer_forward(in_mesh, pose_batch[None], intrinsics_batch[None], render_size, render_size, shading_fun,
                                           normal_bg=normal_bg, near=near, far=far)
        out_alphas = render_out['rgba'][..., 3:].squeeze(0)
        out_rgbs = (render_out['rgba'][..., :3] / render_out['rgba'][..., 3:].clamp(min=1e-3)).squeeze(0)
        out_rgbs = out_rgbs * target_m_erode + target_rgbs * (1 - target_m_erode)
        out_normals = render_out['normal'].squeeze(0)
        out_normals_opencv = depth_to_normal(render_out['depth'].squeeze(0).detach(), target_dir, format='opencv') * 2 - 1
        out_normals_cos = (out_normals_opencv[..., None, :] @ F.normalize(target_dir[..., :, None], dim=-2)).squeeze(-1).neg().clamp(min=0)
        out_normals_cos = -F.max_pool2d(-out_normals_cos.permute(0, 3, 1, 2), 5, stride=1, padding=2).permute(0, 2, 3, 1)
        out_normals = out_normals * out_normals_cos + out_normals.detach() * (1 - out_normals_cos)
        out_normals_fg = (out_normals - normal_bg_t * (1 - out_alphas)) / out_alphas.clamp(min=1e-3)
        out_normals_fg_weight = out_alphas.detach()
        loss = pixel_loss(out_rgbs.reshape(target_rgbs.size()), target_rgbs, weight=target_w / cam_weights_mean) * 4.5
        alphas_loss = pixel_loss(out_alphas.reshape(target_m_blur.size()), target_m_blur, weight=target_w / cam_weights_mean) * 2.0
        target_n = tgt_normals_batches[k] if use_normal else None
        normal_reg_loss = loss_tv(out_normals_fg.permute(0, 3, 1, 2), target_n.permute(0, 3, 1, 2) if use_normal else None,
                                  weight=out_normals_fg_weight.permute(0, 3, 1, 2)) * (normal_reg_weight * 2)
        lapsmth_loss = laplacian_smooth_loss(in_mesh.v, in_mesh.f) * mesh_normal_reg_weight
        norm_const_loss = normal_consistency(in_mesh.face_normals, in_mesh.f) * mesh_normal_reg_weight
        loss = loss + alphas_loss + normal_reg_loss + lapsmth_loss + norm_const_loss
        g = render_size // patch_size
        pt = lambda x: x.reshape(-1, g, patch_size, g, patch_size, x.shape[-1]).permute(0, 1, 3, 5, 2, 4).reshape(-1, x.shape[-1], patch_size, patch_size)
        if patch_rgb_weight > 0 or (use_normal and patch_normal_weight > 0):
            out_rgb_patch, tgt_rgb_patch, target_w_patch = pt(out_rgbs), pt(target_rgbs), pt(target_w)
            patch_batch = noise['patch_perm'][step][:patch_bs]
            target_w_patch_ = target_w_patch[patch_batch, 0, 0, 0]
        if patch_rgb_weight > 0:
            loss = loss + patch_loss(out_rgb_patch[patch_batch], tgt_rgb_patch[patch_batch], weight=target_w_patch_ / cam_weights_mean) * patch_rgb_weight
        if use_normal and patch_normal_weight > 0:                                                # :806-821 (weights of the rgb draw)
            out_normal_patch, tgt_normal_patch = pt(out_normals), pt(target_n)
            patch_batch = noise['patch_perm_normal'][step][:patch_bs]
            loss = loss + patch_loss(highpass(out_normal_patch[patch_batch]), highpass(tgt_normal_patch[patch_batch]),
                                     weight=target_w_patch_ / cam_weights_mean) * patch_normal_weight
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        losses.append(float(loss.detach()))
        mesh_verts, mesh_faces = dmtet(tet_verts + deform, tet_sdf, tet_indices)
        in_mesh = make_mesh(mesh_verts, mesh_faces.int())
    return in_mesh, losses


# ---- texture baking (base_mesh_renderer.py:397-603), restated on the raster oracle ------------------------------------------------------

def edge_dilation(img, mask, radius=3, iters=7):
    """lib/ops/edge_dilation.py:5-49 (pinned by tests/golden/mesh_pins.npz through the product's copy of the same semantics)."""
    if radius == 0 or iters == 0:
        return img
    n, c, h, w = img.size()
    int_radius = round(radius)
    kernel_size = int(int_radius * 2 + 1)
    distance1d_sq = torch.linspace(-int_radius, int_radius, kernel_size, dtype=img.dtype).square()
    kernel_distance = (distance1d_sq.reshape(1, -1) + distance1d_sq.reshape(-1, 1)).sqrt()
    kernel_neg_distance = kernel_distance.max() - kernel_distance + 1
    for _ in range(iters):
        mask_out = F.max_pool2d(mask, kernel_size, stride=1, padding=int_radius)
        do_fill_mask = ((mask_out - mask) > 0.5).squeeze(1)
        do_fill = do_fill_mask.nonzero()
        mask_unfold = F.unfold(mask, kernel_size, padding=int_radius).reshape(n, kernel_size * kernel_size, h, w).permute(0, 2, 3, 1)
        fill_ind = (mask_unfold[do_fill_mask] * kernel_neg_distance.flatten()).argmax(dim=-1)
        do_fill_h = do_fill[:, 1] + fill_ind // kernel_size - int_radius
        do_fill_w = do_fill[:, 2] + fill_ind % kernel_size - int_radius
        img_out = img.clone()
        img_out[do_fill[:, 0], :, do_fill[:, 1], do_fill[:, 2]] = img[do_fill[:, 0], :, do_fill_h, do_fill_w]
        img, mask = img_out, mask_out
    return img


def _proj(poses_batch, intrinsics_batch, h, w, near, far):
    bs = poses_batch.size(0)
    r_mat_c2w = torch.cat([poses_batch[:, :3, :1], -poses_batch[:, :3, 1:3]], dim=-1)
    proj = poses_batch.new_zeros([bs, 4, 4])
    proj[:, 0, 0] = 2 * intrinsics_batch[:, 0] / w
    proj[:, 0, 2] = -2 * intrinsics_batch[:, 2] / w + 1
    proj[:, 1, 1] = -2 * intrinsics_batch[:, 1] / h
    proj[:, 1, 2] = -2 * intrinsics_batch[:, 3] / h + 1
    proj[:, 2, 2] = -(far + near) / (far - near)
    proj[:, 2, 3] = -(2 * far * near) / (far - near)
    proj[:, 3, 2] = -1
    return r_mat_c2w, proj


def bake_multiview(mesh, images, alphas, poses, intrinsics, map_size=64, cos_weight_pow=8.0, render_bs=8, near=0.01, far=100.0,
                   weights_only=False):
    """base_mesh_renderer.py:507-603 (``weights_only``: get_cam_weights_uv, :425-505, which runs the same loop on the weight alone).
    images [1,n,h,w,3], alphas [1,n,h,w,1], poses [1,n,3|4,4], intrinsics [1,n,4]."""
    images, alphas = images[0], alphas[0]
    n, h, w, _ = images.size()
    poses, intrinsics = poses[0].expand(n, -1, -1), intrinsics[0].expand(n, -1)
    new_albedo_map_sum = torch.zeros((map_size, map_size, 3), dtype=images.dtype)
    weights_sum = torch.zeros((map_size, map_size, 1), dtype=images.dtype)
    vt_clip = torch.cat([mesh.vt * 2 - 1, mesh.vt.new_tensor([[0., 1.]]).expand(mesh.vt.size(0), -1)], dim=-1)
    tex_rast, tex_rast_db = dr_rasterize(vt_clip[None], mesh.ft, (map_size, map_size))
    valid = (tex_rast[..., 3] > 0).reshape(map_size, map_size)
    out_weights = []
    for images_batch, alphas_batch, poses_batch, intrinsics_batch in zip(images.split(render_bs), alphas.split(render_bs),
                                                                         poses.split(render_bs), intrinsics.split(render_bs)):
        bs = images_batch.size(0)
        r_mat_c2w, proj = _proj(poses_batch, intrinsics_batch, h, w, near, far)
        v_cam = (mesh.v.detach() - poses_batch[:, :3, 3].unsqueeze(-2)) @ r_mat_c2w
        v_clip = F.pad(v_cam, pad=(0, 1), mode='constant', value=1.0) @ proj.transpose(-1, -2)
        rast, rast_db = dr_rasterize(v_clip, mesh.f, (h, w))
        texc, texc_db = ro.interpolate(mesh.vt.unsqueeze(0), rast, mesh.ft, rast_db=rast_db, diff_attrs='all')
        with torch.enable_grad():
            dummy_maps = torch.ones((bs, map_size, map_size, 1), dtype=images.dtype).requires_grad_(True)
            albedo = ro.texture(dummy_maps, texc, texc_db, filter_mode='linear-mipmap-linear')
            visibility_grad = torch.autograd.grad(albedo.sum(), dummy_maps, create_graph=False)[0]
        fg = rast[..., 3] > 0
        depth = 1 / ro.interpolate(-v_cam[..., 2:3], rast, mesh.f)[0].reshape(bs, h, w)
        depth = depth.masked_fill(~fg, 0)
        directions = get_ray_directions(h, w, intrinsics_batch, norm=True)
        normals_opencv = depth_to_normal(depth, directions, format='opencv') * 2 - 1
        normals_cos_weight = (normals_opencv[..., None, :] @ directions[..., :, None]).squeeze(-1).neg().clamp(min=0)
        img_space_weight = (normals_cos_weight ** cos_weight_pow) * alphas_batch
        img_space_weight = -F.max_pool2d(-img_space_weight.permute(0, 3, 1, 2), 5, stride=1, padding=2).permute(0, 2, 3, 1)
        v_img = v_clip[..., :2] / v_clip[..., 3:] * 0.5 + 0.5
        imgc, imgc_db = ro.interpolate(v_img, tex_rast.expand(bs, -1, -1, -1), mesh.f, rast_db=tex_rast_db.expand(bs, -1, -1, -1), diff_attrs='all')
        if weights_only:
            tex = ro.texture(img_space_weight, imgc, imgc_db, filter_mode='linear-mipmap-linear')
            out_weights.append(tex * visibility_grad)
            continue
        tex = ro.texture(torch.cat([images_batch, img_space_weight], dim=-1), imgc, imgc_db, filter_mode='linear-mipmap-linear')
        weight = tex[..., 3:4] * visibility_grad
        new_albedo_map_sum += (tex[..., :3] * weight).sum(dim=0)
        weights_sum += weight.sum(dim=0)
    if weights_only:
        return torch.cat(out_weights, dim=0)[None], valid[None]
    new_albedo_map = new_albedo_map_sum / weights_sum.clamp(min=1e-8)
    new_albedo_map = edge_dilation(new_albedo_map.permute(2, 0, 1)[None], valid[None, None].float()).squeeze(0).permute(1, 2, 0)
    return torch.cat([new_albedo_map.clamp(min=0, max=1), torch.ones_like(new_albedo_map[..., :1])], dim=-1)


def bake_xyz_shading_fun(mesh, shading_fun, map_size=64, dilation_iters=7):
    """base_mesh_renderer.py:397-423 for a mesh that already has vt / ft."""
    vt_clip = torch.cat([mesh.vt * 2 - 1, mesh.vt.new_tensor([[0., 1.]]).expand(mesh.vt.size(0), -1)], dim=-1)
    rast, _ = dr_rasterize(vt_clip[None], mesh.ft, (map_size, map_size))
    valid = (rast[..., 3] > 0).reshape(map_size, map_size)
    xyz = ro.interpolate(mesh.v[None], rast, mesh.f)[0].reshape(map_size, map_size, 3)
    new_albedo_map = xyz.new_zeros((map_size, map_size, 3))
    new_albedo_map[valid] = shading_fun(world_pos=xyz[valid])
    new_albedo_map = edge_dilation(new_albedo_map.permute(2, 0, 1)[None], valid[None, None].float(), iters=dilation_iters).squeeze(0).permute(1, 2, 0)
    return torch.cat([new_albedo_map.clamp(min=0, max=1), torch.ones_like(new_albedo_map[..., :1])], dim=-1)


def texture_optim(decoder, tgt_images, optimizer, lr, inverse_steps, render_bs, nerf_code, in_mesh, render_size, intrinsics, intrinsics_size,
                  camera_poses, cam_weights_dense, noise, bg_color=0.5, pixel_loss=None, near=0.01, far=100.0):
    """lib/pipelines/mvedit_texture_pipeline.py:93-172 without the patch term, every random draw supplied."""
    pixel_loss = pixel_loss or L1LossMod(loss_weight=1.2)
    optimizer.param_groups[0]['lr'] = lr
    camera_perm = noise['camera_perm']
    pose_batches = camera_poses[camera_perm].split(render_bs, dim=0)
    intrinsics_batches = intrinsics[camera_perm].split(render_bs, dim=0)
    tgt_image_batches = tgt_images.squeeze(0)[camera_perm].split(render_bs, dim=0)
    cam_weights_batches = cam_weights_dense[camera_perm].split(render_bs, dim=0)
    num_pose_batches = len(pose_batches)

    def shading_fun(world_pos=None, albedo=None, **kwargs):
        if len(world_pos) == 0:
            return world_pos if albedo is None else albedo
        return decoder.point_decode([world_pos], None, nerf_code)[1]

    for inverse_step_id in range(inverse_steps):
        pose_batch = pose_batches[inverse_step_id % num_pose_batches]
        intrinsics_batch = intrinsics_batches[inverse_step_id % num_pose_batches] * (render_size / intrinsics_size)
        target_rgbs = tgt_image_batches[inverse_step_id % num_pose_batches]
        target_w = cam_weights_batches[inverse_step_id % num_pose_batches]
        intrinsics_batch = torch.cat([intrinsics_batch[:, :2], intrinsics_batch[:, 2:] + (noise['jitter'][inverse_step_id, :len(pose_batch)] - 0.5)], dim=1)
        render_out = mesh_renderer_forward(in_mesh, pose_batch[None], intrinsics_batch[None], render_size, render_size, shading_fun, near=near, far=far)
        out_rgbs = (render_out['rgba'][..., :3] + (1 - render_out['rgba'][..., 3:].clamp(min=1e-3)) * bg_color).squeeze(0)
        loss = pixel_loss(out_rgbs.reshape(target_rgbs.size()), target_rgbs, weight=target_w) * 2
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()

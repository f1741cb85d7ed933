"""Op-by-op PyTorch restatement of the NeRF-adapter slice of the reference, as the reference runs it:
eager torch around the ray-marching extension.

TEST INFRASTRUCTURE ONLY (see oracle/raymarching_oracle.c header).  Imported by tests/, by bench.py's reference legs
(``cpu_baseline``, ``gpu_baseline``, ``--impl reference``) and by smoke() as the checker -- never by mvedit_b200.

Restated from (file:line under /root/reference):
  geometry helpers          lib/core/utils/geometry_utils.py:18-55,119-168
  losses                    lib/models/losses/pixelwise_loss.py:9-35, tv_loss.py:7-61 (mmgen ``weighted_loss`` semantics)
  get_noise_scales          lib/core/diffusion.py:4-21
  lib.ops.raymarching       lib/ops/raymarching/raymarching.py:67-524 (the Python wrappers: two-pass march with ``.item()``,
                            autograd Function around composite fwd/bwd, in-place inference ops)
  VolumeRenderer            lib/models/decoders/base_volume_renderer.py:105-177 (update_extra_state), :179-343 (forward)
  BaseNeRF                  lib/models/autoencoders/base_nerf.py:245-322 (ray_sample, get_raybatch_inds), :489-556 (render)
  nerf_optim                lib/pipelines/mvedit_3d_pipeline.py:452-656
  render + shading          lib/pipelines/mvedit_3d_pipeline.py:1341-1395

Two interchangeable ray-marching backends:
  * ``RefOps``  -- the reference's OWN kernels, compiled unmodified into oracle/_ref (GPU box): this is the reference path;
  * ``CpuOps``  -- the C restatement oracle/raymarching_oracle.c (pinned against the above by tests/golden), for CPU runs.
The hash grid is ``oracle/field_oracle.py`` (tiny-cuda-nn is not installable offline: plain-PyTorch gathers -- PARITY UNPINNED for
that part and slower than tcnn; stated wherever a time is reported).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import field_oracle as fo


# ------------------------------------------------------------------------------------------------ geometry (geometry_utils.py)
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


def get_noise_scales(alphas_bar, t, num_timesteps, dtype=torch.float32):
    """lib/core/diffusion.py:4-21."""
    alphas_bar = t.new_tensor(alphas_bar, dtype=torch.float32)
    if t.is_floating_point():
        int_t = t.long()
        frac_t = t - int_t
        a0 = alphas_bar[int_t]
        a1 = alphas_bar[(int_t + 1).clamp(max=num_timesteps - 1)]
        s0, s1 = torch.sqrt((1 - a0) / a0), torch.sqrt((1 - a1) / a1)
        ve = s0 * (1 - frac_t) + s1 * frac_t
        return torch.sqrt(1 / (1 + ve ** 2)).to(dtype), torch.sqrt(ve ** 2 / (1 + ve ** 2)).to(dtype)
    a = alphas_bar[t]
    return torch.sqrt(a).to(dtype), torch.sqrt(1 - a).to(dtype)


# ------------------------------------------------------------------------------------------------ losses
def _weighted(loss, weight=None, avg_factor=None):
    """mmgen ``weighted_loss`` with reduction='mean': elementwise * weight, then mean (or sum / avg_factor)."""
    if weight is not None:
        loss = loss * weight
    return loss.mean() if avg_factor is None else loss.sum() / avg_factor


class Tonemapping(nn.Module):
    """lib/models/decoders/tonemapping.py:5-52: 16 knots of x -> sigmoid(c (x + e)) g_s + c (x + e) g_l + b on a log2-exposure axis;
    ``lut`` / ``inverse_lut`` interpolate linearly between the knots (bucketize right=True, index clamped to [1, n-1])."""

    def __init__(self, exposure=0.0, contrast=0.953, bias=0.088, sigmoid_gain=0.943, log_gain=0.011, lut_logx_min=-9, lut_logx_max=3,
                 lut_steps=16):
        super().__init__()
        self.p = (exposure, contrast, bias, sigmoid_gain, log_gain)
        self.register_buffer('lut_x', torch.linspace(lut_logx_min, lut_logx_max, lut_steps))
        self.register_buffer('lut_y', self.smooth_forward(self.lut_x))

    def smooth_forward(self, x, input_mode='log'):
        e, c, b, gs, gl = self.p
        if input_mode == 'linear':
            x = x.clamp(min=1e-6).log2()
        x = (x + e) * c
        return x.sigmoid() * gs + x * gl + b

    @staticmethod
    def _pw(v, a, b):
        i = torch.bucketize(v, a, right=True).clamp(min=1, max=len(a) - 1)
        t = (v - a[i - 1]) / (a[i] - a[i - 1])
        return b[i - 1] + (b[i] - b[i - 1]) * t

    def lut(self, x, input_mode='log'):
        dt = x.dtype
        x = x.to(self.lut_x.dtype)
        if input_mode == 'linear':
            x = x.clamp(min=1e-6).log2()
        return self._pw(x, self.lut_x, self.lut_y).to(dt)

    def inverse_lut(self, y, output_mode='log'):
        dt = y.dtype
        x = self._pw(y.to(self.lut_y.dtype), self.lut_y, self.lut_x)
        return (torch.exp2(x) if output_mode == 'linear' else x).to(dt)


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


# ------------------------------------------------------------------------------------------------ ray-marching backends
class _CompositeTrain(torch.autograd.Function):
    """raymarching.py:313-368."""

    @staticmethod
    def forward(ctx, ops, sigmas, rgbs, ts, rays, T_thresh, binarize):
        sigmas, rgbs = sigmas.float().contiguous(), rgbs.float().contiguous()
        weights, weights_sum, depth, image = ops._composite_fwd(sigmas, rgbs, ts, rays, T_thresh, binarize)
        ctx.save_for_backward(sigmas, rgbs, ts, rays, weights_sum, depth, image)
        ctx.ops, ctx.dims = ops, (T_thresh, binarize)
        return weights, weights_sum, depth, image

    @staticmethod
    def backward(ctx, gw, gws, gd, gi):
        sigmas, rgbs, ts, rays, weights_sum, depth, image = ctx.saved_tensors
        gs, gc = ctx.ops._composite_bwd(gw.contiguous(), gws.contiguous(), gd.contiguous(), gi.contiguous(), sigmas, rgbs, ts, rays,
                                        weights_sum, depth, image, *ctx.dims)
        return None, gs, gc, None, None, None, None


class RefOps:
    """lib.ops.raymarching on the reference's own kernels (oracle/_ref, CUDA only)."""

    def __init__(self):
        from . import build_ref
        self.m = build_ref.load_ref()
        if self.m is None:
            raise RuntimeError('oracle/_ref is not built')
        self.device = torch.device('cuda')

    def near_far_from_aabb(self, rays_o, rays_d, aabb, min_near=0.2):
        rays_o, rays_d = rays_o.float().contiguous().view(-1, 3), rays_d.float().contiguous().view(-1, 3)
        N = rays_o.shape[0]
        nears, fars = torch.empty(N, device=rays_o.device), torch.empty(N, device=rays_o.device)
        self.m.near_far_from_aabb(rays_o, rays_d, aabb.float().contiguous(), N, min_near, nears, fars)
        return nears, fars

    def morton3D(self, coords):
        coords = coords.int().contiguous()
        out = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
        self.m.morton3D(coords, coords.shape[0], out)
        return out

    def packbits(self, grid, thresh, bitfield):
        grid = grid.contiguous()
        self.m.packbits(grid, grid.numel() // 8, float(thresh), bitfield)
        return bitfield

    def march_rays_train(self, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, perturb=False, dt_gamma=0, max_steps=1024,
                         noises=None):
        """raymarching.py:238-302: two passes with a host sync between them."""
        rays_o, rays_d = rays_o.float().contiguous().view(-1, 3), rays_d.float().contiguous().view(-1, 3)
        N, dev = rays_o.shape[0], rays_o.device
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        if noises is None:
            noises = torch.rand(N, device=dev) if perturb else torch.zeros(N, device=dev)
        rays = torch.empty(N, 2, dtype=torch.int32, device=dev)
        self.m.march_rays_train(rays_o, rays_d, density_bitfield, bound, False, dt_gamma, max_steps, N, C, H, nears, fars, None, None, None,
                                rays, counter, noises)
        M = counter.item()
        xyzs, dirs, ts = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        self.m.march_rays_train(rays_o, rays_d, density_bitfield, bound, False, dt_gamma, max_steps, N, C, H, nears, fars, xyzs, dirs, ts,
                                rays, counter, noises)
        return xyzs, dirs, ts, rays

    def _composite_fwd(self, sigmas, rgbs, ts, rays, T_thresh, binarize):
        M, N, dev = sigmas.shape[0], rays.shape[0], sigmas.device
        weights = torch.zeros(M, device=dev)
        weights_sum, depth, image = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
        self.m.composite_rays_train_forward(sigmas, rgbs, ts, rays, M, N, T_thresh, binarize, weights, weights_sum, depth, image)
        return weights, weights_sum, depth, image

    def _composite_bwd(self, gw, gws, gd, gi, sigmas, rgbs, ts, rays, weights_sum, depth, image, T_thresh, binarize):
        M, N = sigmas.shape[0], rays.shape[0]
        gs, gc = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        self.m.composite_rays_train_backward(gw, gws, gd, gi, sigmas, rgbs, ts, rays, weights_sum, depth, image, M, N, T_thresh, binarize,
                                             gs, gc)
        return gs, gc

    def composite_rays_train(self, sigmas, rgbs, ts, rays, T_thresh=1e-4, binarize=False):
        return _CompositeTrain.apply(self, sigmas, rgbs, ts, rays, T_thresh, binarize)

    def march_rays(self, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, perturb=False,
                   dt_gamma=0, max_steps=1024):
        """raymarching.py:440-483."""
        dev = rays_o.device
        M = n_alive * n_step
        xyzs, dirs, ts = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        noises = torch.rand(n_alive, device=dev) if perturb else torch.zeros(n_alive, device=dev)
        self.m.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, False, dt_gamma, max_steps, C, H, density_bitfield,
                          near, far, xyzs, dirs, ts, noises)
        return xyzs, dirs, ts

    def composite_rays(self, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh=1e-2, binarize=False):
        self.m.composite_rays(n_alive, n_step, T_thresh, binarize, rays_alive, rays_t, sigmas.float().contiguous(), rgbs.float().contiguous(),
                              ts, weights_sum, depth, image)


class CpuOps:
    """The same interface on oracle/raymarching_oracle.c (torch CPU tensors in and out)."""

    def __init__(self):
        from . import raymarching_oracle as orc
        self.o = orc
        self.device = torch.device('cpu')

    @staticmethod
    def _n(t):
        return t.detach().cpu().numpy()

    def near_far_from_aabb(self, rays_o, rays_d, aabb, min_near=0.2):
        n, f = self.o.near_far_from_aabb(self._n(rays_o), self._n(rays_d), self._n(aabb), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)

    def morton3D(self, coords):
        return torch.from_numpy(self.o.morton3D(self._n(coords)))

    def packbits(self, grid, thresh, bitfield):
        bitfield.copy_(torch.from_numpy(self.o.packbits(self._n(grid.float()), float(thresh))))
        return bitfield

    def march_rays_train(self, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, perturb=False, dt_gamma=0, max_steps=1024,
                         noises=None):
        N = rays_o.reshape(-1, 3).shape[0]
        if noises is None:
            noises = torch.rand(N) if perturb else torch.zeros(N)
        x, d, t, r = self.o.march_rays_train(self._n(rays_o), self._n(rays_d), bound, self._n(density_bitfield), C, H, self._n(nears),
                                             self._n(fars), self._n(noises), dt_gamma=float(dt_gamma), max_steps=max_steps)
        return torch.from_numpy(x), torch.from_numpy(d), torch.from_numpy(t), torch.from_numpy(r)

    def _composite_fwd(self, sigmas, rgbs, ts, rays, T_thresh, binarize):
        return tuple(torch.from_numpy(a) for a in self.o.composite_rays_train_forward(self._n(sigmas), self._n(rgbs), self._n(ts),
                                                                                      self._n(rays), T_thresh, binarize))

    def _composite_bwd(self, gw, gws, gd, gi, sigmas, rgbs, ts, rays, weights_sum, depth, image, T_thresh, binarize):
        gs, gc = self.o.composite_rays_train_backward(*[self._n(a) for a in (gw, gws, gd, gi, sigmas, rgbs, ts, rays, weights_sum, depth,
                                                                           image)], T_thresh, binarize)
        return torch.from_numpy(gs), torch.from_numpy(gc)

    def composite_rays_train(self, sigmas, rgbs, ts, rays, T_thresh=1e-4, binarize=False):
        return _CompositeTrain.apply(self, sigmas, rgbs, ts, rays, T_thresh, binarize)

    def march_rays(self, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, perturb=False,
                   dt_gamma=0, max_steps=1024):
        assert not perturb
        x, d, t = self.o.march_rays(n_alive, n_step, self._n(rays_alive), self._n(rays_t), self._n(rays_o), self._n(rays_d), bound,
                                    self._n(density_bitfield), C, H, self._n(near), self._n(far), None, dt_gamma=float(dt_gamma),
                                    max_steps=max_steps)
        return torch.from_numpy(x), torch.from_numpy(d), torch.from_numpy(t)

    def composite_rays(self, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh=1e-2, binarize=False):
        # numpy views share memory with the (contiguous, CPU) torch tensors: the C oracle updates them in place
        self.o.composite_rays(n_alive, n_step, rays_alive.numpy(), rays_t.numpy(), self._n(sigmas), self._n(rgbs), self._n(ts),
                              weights_sum.numpy(), depth.numpy(), image.numpy(), T_thresh, binarize)


# ------------------------------------------------------------------------------------------------ decoder (ingp_decoder.py + base_volume_renderer.py)
class OracleDecoder(nn.Module):
    """iNGPDecoder (ingp_decoder.py:43-125) on VolumeRenderer (base_volume_renderer.py:17-343), one scene, hash grid = field_oracle.
    State-dict keys as the reference: aabb, encoder.params, mlp.net.{0,1}.{weight,bias}."""

    def __init__(self, ops, bound=1, min_near=0.2, max_steps=256, weight_culling_th=0.0, base_resolution=16, max_resolution=320,
                 n_levels=12, sigmoid_saturation=0.001, blob_density=1.0, blob_radius=0.2):
        super().__init__()
        self.ops = ops
        self.bound, self.min_near, self.max_steps, self.weight_culling_th = bound, min_near, max_steps, weight_culling_th
        self.sigmoid_saturation, self.blob_density, self.blob_radius = sigmoid_saturation, blob_density, blob_radius
        self.levels, n_entries = fo.level_table(n_levels, base_resolution, max_resolution, bound)
        self.register_buffer('aabb', torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound]))
        self.encoder = nn.Module()
        self.encoder.params = nn.Parameter(torch.zeros(n_entries * 2))
        self.mlp = nn.Module()
        self.mlp.net = nn.ModuleList([nn.Linear(2 * n_levels, 64), nn.Linear(64, 4)])

    def point_decode(self, xyzs, dirs, code, density_only=False):
        sig, rgb = fo.point_decode(xyzs[0], self.encoder.params.view(-1, 2), self.mlp.net[0].weight, self.mlp.net[0].bias,
                                   self.mlp.net[1].weight, self.mlp.net[1].bias, self.levels, self.bound, self.sigmoid_saturation,
                                   self.blob_density, self.blob_radius)
        return sig, (None if density_only else rgb), [len(xyzs[0])]

    def point_density_decode(self, xyzs, code):
        sig, _, n = self.point_decode(xyzs, None, code, density_only=True)
        return sig, n

    def update_extra_state(self, code, density_grid, density_bitfield, iter_density, density_thresh=0.01, decay=0.9, S=128, noise=None):
        """base_volume_renderer.py:105-177, full-update branch (the only one the pipelines reach: SURVEY.md Appendix F).
        ``noise`` ([H^3,3] in [0,1), meshgrid order) replaces ``torch.rand_like`` so that two implementations see the same jitter."""
        with torch.no_grad():
            device = density_grid.device
            tmp_grid = torch.full_like(density_grid, -1)
            grid_size = int(round(density_grid.size(-1) ** (1. / 3.)))
            assert iter_density < 16
            ar = torch.arange(grid_size, dtype=torch.int32, device=device)
            xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing='ij')
            coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
            indices = self.ops.morton3D(coords).long()
            xyzs = (coords.float() - (grid_size - 1) / 2) * (2 * self.bound / grid_size)
            half_voxel_width = self.bound / grid_size
            u = torch.rand_like(xyzs) if noise is None else noise.to(xyzs)
            xyzs = xyzs + u * (2 * half_voxel_width) - half_voxel_width
            sigmas = torch.cat([self.point_density_decode([c], code)[0] for c in xyzs.split(1 << 19)]).reshape(1, -1)
            tmp_grid[:, indices] = sigmas.clamp(max=torch.finfo(tmp_grid.dtype).max).to(tmp_grid.dtype)
            valid_mask = (density_grid >= 0) & (tmp_grid >= 0)
            density_grid[:] = torch.where(valid_mask, torch.maximum(density_grid * decay, tmp_grid), density_grid)
            mean_density = torch.mean(density_grid.clamp(min=0))
            density_thresh = min(mean_density, density_thresh)
            self.ops.packbits(density_grid[0], density_thresh, density_bitfield[0])

    def forward(self, rays_o, rays_d, code, density_bitfield, grid_size, dt_gamma=0.0, perturb=False, update_extra_state=0,
                extra_args=None, extra_kwargs=None, noises=None):
        """base_volume_renderer.py:179-343 for one scene.  rays_o / rays_d (1,N,3)."""
        ops = self.ops
        for _ in range(update_extra_state):
            self.update_extra_state(code, *extra_args, **extra_kwargs)
        if isinstance(dt_gamma, torch.Tensor):
            dt_gamma = float(dt_gamma.reshape(-1)[0])              # zip over ONE scene: only element 0 is used (:212-218)
        ro, rd, bitfield = rays_o[0], rays_d[0], density_bitfield[0]
        nears, fars = ops.near_far_from_aabb(ro, rd, self.aabb.to(ro), self.min_near)
        if self.training:
            xyzs, dirs, ts, rays = ops.march_rays_train(ro, rd, self.bound, bitfield, 1, grid_size, nears, fars, perturb=perturb,
                                                        dt_gamma=float(dt_gamma), max_steps=self.max_steps, noises=noises)
            if self.weight_culling_th > 0:
                with torch.no_grad():
                    sigmas, _ = self.point_density_decode([xyzs], code)
                    weights, _, _, _ = ops.composite_rays_train(sigmas, sigmas.new_zeros(sigmas.shape[0], 3), ts, rays)
                    mask = weights > self.weight_culling_th
                filt_inds = F.pad(torch.cumsum(mask, dim=0), (1, 0), value=0)
                new_start = filt_inds[rays[:, 0].long()]
                new_end = filt_inds[(rays[:, 0] + rays[:, 1]).long()]
                xyzs, dirs, ts = xyzs[mask], dirs[mask], ts[mask]
                rays = torch.stack([new_start, new_end - new_start], dim=-1).int()
            sigmas, rgbs, _ = self.point_decode([xyzs], [dirs], code)
            weights, weights_sum, depth, image = ops.composite_rays_train(sigmas, rgbs, ts, rays)
            return dict(weights=weights, weights_sum=weights_sum[None], depth=depth[None], image=image[None], rays=[rays], ts=[ts])
        N, device = ro.shape[0], ro.device
        ro, rd = ro.float().contiguous(), rd.float().contiguous()
        weights_sum, depth, image = torch.zeros(N, device=device), torch.zeros(N, device=device), torch.zeros(N, 3, device=device)
        rays_alive = torch.arange(N, dtype=torch.int32, device=device)
        rays_t = nears.clone()
        step = 0
        with torch.no_grad():
            while step < self.max_steps:
                n_alive = rays_alive.size(0)
                if n_alive == 0:
                    break
                n_step = min(max(N // n_alive, 1), 8)
                xyzs, dirs, ts = ops.march_rays(n_alive, n_step, rays_alive, rays_t, ro, rd, self.bound, bitfield, 1, grid_size, nears, fars,
                                                perturb=perturb, dt_gamma=float(dt_gamma), max_steps=self.max_steps)
                sigmas, rgbs, _ = self.point_decode([xyzs], [dirs], code)
                ops.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image)
                rays_alive = rays_alive[rays_alive >= 0]
                step += n_step
        return dict(weights=None, weights_sum=[weights_sum], depth=[depth], image=[image], rays=None, ts=None)


# ------------------------------------------------------------------------------------------------ BaseNeRF (base_nerf.py)
class OracleNeRF:
    def __init__(self, decoder, grid_size=128, bg_color=1.0, patch_size=128, pixel_loss_weight=1.2, update_extra_interval=16,
                 update_extra_iters=1):
        self.decoder, self.grid_size, self.bg_color, self.patch_size = decoder, grid_size, bg_color, patch_size
        self.pixel_loss = L1LossMod(loss_weight=pixel_loss_weight)
        self.patch_loss = None
        self.update_extra_interval, self.update_extra_iters = update_extra_interval, update_extra_iters

    def ray_sample(self, cond_rays_o, cond_rays_d, cond_imgs, n_samples, sample_inds=None, cond_extras=None):
        """base_nerf.py:245-303, patch-wise branch (the pipelines always build BaseNeRF with a patch loss / patch_size 128)."""
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

    def render(self, density_bitfield, h, w, intrinsics, poses, cfg=dict(), normal_bg=(0.5, 0.5, 1.0)):
        """base_nerf.py:489-556 (one scene; intrinsics (1,V,4), poses (1,V,3|4,4))."""
        decoder = self.decoder
        training_prev = decoder.training
        decoder.train(False)
        dt_gamma = cfg.get('dt_gamma_scale', 0.0) * 2 / (intrinsics[..., 0] + intrinsics[..., 1]).mean(dim=-1)
        directions = get_ray_directions(h, w, intrinsics, norm=False, device=intrinsics.device)
        rays_o, rays_d = get_rays(directions, poses, norm=True)
        num_scenes, num_imgs = rays_o.shape[:2]
        outputs = decoder(rays_o.reshape(num_scenes, -1, 3), rays_d.reshape(num_scenes, -1, 3), None, density_bitfield, self.grid_size,
                          dt_gamma=dt_gamma, perturb=False)
        weights_sum = outputs['weights_sum'][0]
        out_image = torch.cat([outputs['image'][0], weights_sum.unsqueeze(-1)], dim=-1).reshape(num_scenes, num_imgs, h, w, 4)
        out_depth = outputs['depth'][0].reshape(num_scenes, num_imgs, h, w) * torch.linalg.norm(directions, dim=-1)
        decoder.train(training_prev)
        out_depth_fg = out_depth / out_image[..., 3].clamp(min=1e-6)
        out_normal_fg = depth_to_normal(out_depth_fg, directions)
        out_normal = out_normal_fg * out_image[..., 3:] + out_normal_fg.new_tensor(normal_bg) * (1 - out_image[..., 3:])
        return out_image, out_depth, out_normal, out_normal_fg


def highpass(x, std=5, offset=0.5):
    """lib/pipelines/utils.py:187-188."""
    return offset + x - gaussian_blur(x, int(round(std)) * 6 + 1, std)


# ------------------------------------------------------------------------------------------------ nerf_optim (mvedit_3d_pipeline.py:452-656)
def nerf_optim(nerf, tgt_images, tgt_masks, tgt_normals, optimizer, lr, inverse_steps, n_inverse_rays, patch_rgb_weight,
               patch_normal_weight, alpha_soften, normal_reg_weight, entropy_weight, nerf_code, density_grid, density_bitfield,
               render_size, intrinsics, intrinsics_size, camera_poses, cam_weights, cam_lights, patch_size, is_init, bg_width,
               ambient_light, dt_gamma_scale, init_shaded, alpha_blur_std=1.5, debug=False, tgt_depths=None, depth_weight=0.0,
               normal_bg=(0.5, 0.5, 1.0), raybatch_inds=None, march_noises=None, grid_noises=None, tonemapping=None):
    """The reference's reconstruction loop, term by term.  ``raybatch_inds`` / ``march_noises`` /
    ``grid_noises`` (lists, one entry per iteration / occupancy refresh) replace the internal random draws for parity runs."""
    loss_tv = TVLoss(loss_weight=1.0, power=1.5)
    use_normal = tgt_normals is not None
    use_depth = tgt_depths is not None and depth_weight > 0
    device = tgt_images.device
    num_cameras = camera_poses.shape[0]
    cam_ids_dense = torch.arange(num_cameras, device=device)[None, :, None, None, None].expand(-1, -1, render_size, render_size, -1)
    cam_weights_mean = cam_weights.mean()
    if alpha_blur_std > 0:
        kernel_size = int((alpha_blur_std * 6) // 2 * 2 + 1)
        tgt_masks_blur = gaussian_blur(tgt_masks.square().squeeze(0).permute(0, 3, 1, 2), kernel_size, alpha_blur_std
                                       ).permute(0, 2, 3, 1)[None].clamp(min=alpha_soften ** 2, max=(1 - alpha_soften) ** 2).sqrt()
    else:
        tgt_masks_blur = tgt_masks.clamp(min=alpha_soften ** 2, max=(1 - alpha_soften) ** 2).sqrt()
    directions = get_ray_directions(render_size, render_size, intrinsics[None] * (render_size / intrinsics_size), norm=False,
                                    device=intrinsics.device)
    cond_rays_o, cond_rays_d = get_rays(directions, camera_poses[None], norm=True)
    normal_bg_t = tgt_images.new_tensor(normal_bg)
    training_prev = nerf.decoder.training
    nerf.decoder.train(True)
    log = []
    with torch.enable_grad():
        optimizer.param_groups[0]['lr'] = lr
        if raybatch_inds is None:
            raybatch_inds, num_raybatch = nerf.get_raybatch_inds(tgt_images, n_inverse_rays)
        else:
            num_raybatch = len(raybatch_inds)
        iter_density = 0
        n_refresh = 0
        for inverse_step_id in range(inverse_steps):
            if inverse_step_id % nerf.update_extra_interval == 0:
                for _ in range(nerf.update_extra_iters):
                    nerf.decoder.update_extra_state(nerf_code, density_grid, density_bitfield, iter_density, density_thresh=0.1,
                                                    noise=None if grid_noises is None else grid_noises[n_refresh])
                    n_refresh += 1
            inds = raybatch_inds[inverse_step_id % num_raybatch] if raybatch_inds is not None else None
            cond_extras = [tgt_masks_blur, directions, cam_ids_dense]
            if use_normal:
                cond_extras.append(tgt_normals)
            if use_depth:
                cond_extras.append(tgt_depths)
            ray_samples = nerf.ray_sample(cond_rays_o, cond_rays_d, tgt_images, n_inverse_rays, sample_inds=inds, cond_extras=cond_extras)
            rays_o, rays_d, target_rgbs, target_m_blur, target_dir, target_cam_ids = ray_samples[:6]
            ray_samples = list(ray_samples[6:])
            if use_normal:
                target_n = ray_samples.pop(0)
            if use_depth:
                target_depth = ray_samples.pop(0)
            target_cam_ids = target_cam_ids[:, 0, 0, 0]
            target_w = cam_weights[target_cam_ids][:, None, None, None].expand(-1, patch_size, patch_size, 1)
            target_lights = cam_lights[target_cam_ids][:, None, None, :].expand(-1, patch_size, patch_size, 3)
            dt_gamma = dt_gamma_scale / (intrinsics[target_cam_ids, :2].mean(dim=-1) * render_size / intrinsics_size)
            outputs = nerf.decoder(rays_o, rays_d, nerf_code, density_bitfield, nerf.grid_size, dt_gamma=dt_gamma, perturb=True,
                                   noises=None if march_noises is None else march_noises[inverse_step_id])
            out_rgbs = outputs['image'].reshape(target_rgbs.size())
            out_alphas = outputs['weights_sum'].reshape(target_m_blur.size())
            out_depth = outputs['depth'].reshape(-1, nerf.patch_size, nerf.patch_size)
            out_depth = out_depth * torch.linalg.norm(target_dir, dim=-1).reshape(out_depth.size())  # 1/r -> 1/z
            out_depth_fg = out_depth / out_alphas.reshape(-1, nerf.patch_size, nerf.patch_size).clamp(min=1e-6)
            out_normals_fg = depth_to_normal(out_depth_fg, target_dir)
            out_normals_fg_mask = out_alphas.reshape(-1, nerf.patch_size, nerf.patch_size, 1)
            out_normals = out_normals_fg * out_normals_fg_mask + normal_bg_t * (1 - out_normals_fg_mask)
            out_normals_fg_weight = -F.max_pool2d(-out_normals_fg_mask.detach().squeeze(-1).unsqueeze(1), 3, stride=1, padding=1
                                                  ).squeeze(1).unsqueeze(-1)
            if not is_init or init_shaded:
                out_normals_fg_opencv = torch.cat([out_normals_fg[..., :1] * 2 - 1, -out_normals_fg[..., 1:3] * 2 + 1], dim=-1)
                nerf_shading = ((target_lights[..., None, :] @ out_normals_fg_opencv[..., :, None]).clamp(min=0)
                                * (1 - ambient_light) + ambient_light).squeeze(-1)
                if tonemapping is None:
                    out_rgbs = out_rgbs * nerf_shading + nerf.bg_color * (1 - out_alphas)
                else:                                                                             # :564-570
                    out_rgbs = tonemapping.lut(tonemapping.inverse_lut(out_rgbs / out_alphas.clamp(min=1e-6))
                                               + nerf_shading.clamp(min=1e-6).log2()) * out_alphas + nerf.bg_color * (1 - out_alphas)
            else:
                out_rgbs = out_rgbs + nerf.bg_color * (1 - out_alphas)
            pixel_rgb_loss = nerf.pixel_loss(out_rgbs.reshape(target_rgbs.size()), target_rgbs, weight=target_w / cam_weights_mean) * 4.5
            alphas_loss = nerf.pixel_loss(out_alphas.reshape(target_m_blur.size()), target_m_blur, weight=target_w / cam_weights_mean
                                          ) * (5.0 if is_init else 1.0)
            normal_reg_loss = loss_tv(out_normals_fg.permute(0, 3, 1, 2), target_n.permute(0, 3, 1, 2) if use_normal else None,
                                      weight=out_normals_fg_weight.permute(0, 3, 1, 2)) * (normal_reg_weight * 10)
            loss = pixel_rgb_loss + alphas_loss + normal_reg_loss
            if use_depth:
                loss = loss + nerf.pixel_loss(out_depth.reshape(target_depth.size()), target_depth, weight=target_w / cam_weights_mean
                                              ) * depth_weight
            bin_weights_sum = outputs['weights'].float()
            bin_width = outputs['ts'][0][:, 1].float()
            bg_weights_sum = 1 - outputs['weights_sum'].flatten()
            entropy_loss = -(torch.sum(bin_weights_sum * (torch.log(bin_weights_sum.clamp(min=1e-6)) - torch.log(bin_width.clamp(min=1e-6))))
                             + torch.sum(bg_weights_sum * (torch.log(bg_weights_sum.clamp(min=1e-6)) - math.log(bg_width)))
                             ) * (entropy_weight / target_rgbs.shape[:-1].numel())
            loss = loss + entropy_loss
            if patch_rgb_weight > 0 and nerf.patch_loss is not None:
                loss = loss + nerf.patch_loss(out_rgbs.reshape(target_rgbs.size()).permute(0, 3, 1, 2), target_rgbs.permute(0, 3, 1, 2),
                                              weight=target_w[:, 0, 0, 0] / cam_weights_mean) * patch_rgb_weight
            if use_normal and patch_normal_weight > 0:                                            # :619-626
                loss = loss + nerf.patch_loss(highpass(out_normals.reshape(target_n.size()).permute(0, 3, 1, 2)),
                                              highpass(target_n.permute(0, 3, 1, 2)),
                                              weight=target_w[:, 0, 0, 0] / cam_weights_mean) * patch_normal_weight
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            if debug:
                log.append(dict(loss=float(loss.detach()), pixel_rgb=float(pixel_rgb_loss.detach()), alpha=float(alphas_loss.detach()), normal_reg=float(normal_reg_loss.detach()),
                                entropy=float(entropy_loss.detach())))
    nerf.decoder.train(training_prev)
    return log if debug else None


# ------------------------------------------------------------------------------------------------ render for denoise P2 (:1341-1395)
def render_views(nerf, density_bitfield, camera_poses, intrinsics, intrinsics_size, render_size, cam_lights, ambient_light,
                 testmode_dt_gamma_scale, render_bs=6, normal_bg=(0.5, 0.5, 1.0), out_dtype=torch.bfloat16, tonemapping=None):
    images, alphas, depths = [], [], []
    for pose_batch, intr_batch, light_batch in zip(camera_poses.split(render_bs, dim=0), intrinsics.split(render_bs, dim=0),
                                                   cam_lights.split(render_bs, dim=0)):
        rgba, depth, normal, normal_fg = nerf.render(density_bitfield, render_size, render_size,
                                                     intr_batch[None] * (render_size / intrinsics_size), pose_batch[None],
                                                     cfg=dict(return_rgba=True, compute_normal=True, dt_gamma_scale=testmode_dt_gamma_scale),
                                                     normal_bg=normal_bg)
        normal_fg_opencv = torch.cat([normal_fg[..., :1] * 2 - 1, -normal_fg[..., 1:3] * 2 + 1], dim=-1)
        shading = ((light_batch[:, None, None, None, :] @ normal_fg_opencv[..., :, None]).clamp(min=0) * (1 - ambient_light)
                   + ambient_light).squeeze(-1)
        if tonemapping is None:
            image = rgba[..., :3] * shading + nerf.bg_color * (1 - rgba[..., 3:])
        else:                                                                                     # :1377-1384
            image = tonemapping.lut(tonemapping.inverse_lut(rgba[..., :3] / rgba[..., 3:].clamp(min=1e-6))
                                    + shading.clamp(min=1e-6).log2()) * rgba[..., 3:] + nerf.bg_color * (1 - rgba[..., 3:])
        images.append(image.squeeze(0)); alphas.append(rgba[..., 3:].squeeze(0)); depths.append(depth.squeeze(0))
    images = torch.cat(images, dim=0).to(out_dtype).permute(0, 3, 1, 2).clamp(min=0, max=1)
    alphas, depths = torch.cat(alphas, dim=0), torch.cat(depths, dim=0)
    depths = normalize_depth(depths, alphas).to(out_dtype).unsqueeze(1).repeat(1, 3, 1, 1)
    return images, depths

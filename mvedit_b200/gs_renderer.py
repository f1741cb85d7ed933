"""3D Gaussian-splatting rasteriser on libmvedit_b200 (BASELINE configs[2]).

The reference snapshot does not ship its 3DGS adapter (README.md:121 only names "3DGS" and ashawkey/diff-gaussian-rasterization;
``lib/models/decoders`` has no ``gs_renderer`` -- SURVEY.md §0), so there is no reference interface to mirror line by line: the call
surface follows the public ``diff_gaussian_rasterization`` package (``GaussianRasterizationSettings`` + ``GaussianRasterizer``,
ashawkey fork: colour, depth and alpha outputs), with pinhole intrinsics in place of the projection matrices.

Split of the work:
  * per-Gaussian projection (world -> camera, EWA 2-D covariance + 0.3 px low-pass, conic, 3-sigma radius, tile rect): elementwise torch
    on [P] tensors -- differentiable, so autograd carries d/d(mean2D, conic, depth) back to (means3D, scales, rotations);
  * tile binning + alpha blending forward / backward: CUDA (csrc/gs_raster.cu) behind ``_BlendFn`` -- key duplication, per-tile ranges,
    256-thread tile CTAs with shared-memory staging, warp-reduced gradient atomics.  The 64-bit key sort is ``torch.sort`` (CUB radix
    sort: library plumbing).
There is no CPU path: tensors must be CUDA, the library must be built.
"""
import ctypes
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from ._lib import call, ptr, stream, c_u32

TILE = 16


def quat_to_rotmat(q):
    q = F.normalize(q, dim=-1)
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(q.shape[:-1] + (3, 3))


def project_gaussians(means3D, scales, rotations, viewmat, K, H, W):
    """-> xy [P,2] (pixel-index coordinates), conic [P,3], depth [P], rect [P,4] int32 (tile x0,y0,x1,y1; empty = culled)."""
    fx, fy, cx, cy = [float(v) for v in K]
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    pc = means3D @ R.t() + t
    x, y, z = pc.unbind(-1)
    valid = z > 0.2                                             # near plane of the public implementation
    zs = torch.where(valid, z, torch.ones_like(z))
    limx, limy = 1.3 * W / (2 * fx), 1.3 * H / (2 * fy)
    tx = (x / zs).clamp(-limx, limx) * zs
    ty = (y / zs).clamp(-limy, limy) * zs
    zero = torch.zeros_like(zs)
    J = torch.stack([fx / zs, zero, -fx * tx / (zs * zs), zero, fy / zs, -fy * ty / (zs * zs)], dim=-1).reshape(-1, 2, 3)
    M = quat_to_rotmat(rotations) * scales[:, None, :]
    Tm = (J @ R) @ M
    cov = Tm @ Tm.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    valid = valid & (det > 0)
    dets = torch.where(valid, det, torch.ones_like(det))
    conic = torch.stack([c / dets, -b / dets, a / dets], dim=-1)
    xy = torch.stack([fx * x / zs + cx - 0.5, fy * y / zs + cy - 0.5], dim=-1)
    with torch.no_grad():
        mid = 0.5 * (a + c)
        radius = torch.ceil(3.0 * torch.sqrt(mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))))
        gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
        tr = lambda v: torch.trunc(v).to(torch.int32)
        rect = torch.stack([tr((xy[:, 0] - radius) / TILE).clamp(0, gx), tr((xy[:, 1] - radius) / TILE).clamp(0, gy),
                            tr((xy[:, 0] + radius + TILE - 1) / TILE).clamp(0, gx), tr((xy[:, 1] + radius + TILE - 1) / TILE).clamp(0, gy)], dim=-1)
        rect = torch.where(valid[:, None], rect, torch.zeros_like(rect)).contiguous()
    return xy, conic, z, rect


class _BlendFn(torch.autograd.Function):
    """Tile binning + alpha blending.  (xy, conic, opacity, colors, depth) -> (color [H,W,3], depth [H,W], alpha [H,W])."""

    @staticmethod
    def forward(ctx, xy, conic, opacity, colors, depth, rect, bg, H, W):
        dev = xy.device
        P = xy.shape[0]
        f = lambda t_: t_.detach().float().contiguous()
        xy_c = f(xy)
        co = torch.cat([f(conic), f(opacity).reshape(P, 1)], dim=1).contiguous()
        ft = torch.cat([f(colors), f(depth).reshape(P, 1)], dim=1).contiguous()
        gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
        counts = ((rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1])).to(torch.int64)
        offsets = torch.cumsum(counts, dim=0)
        L = int(offsets[-1]) if P else 0                          # one host read per render: sizes the key buffers
        ranges = torch.zeros(gx * gy, 2, dtype=torch.int32, device=dev)
        if L:
            keys = torch.empty(L, dtype=torch.int64, device=dev)
            vals = torch.empty(L, dtype=torch.int32, device=dev)
            call('mve_gs_duplicate_keys', ptr(rect), ptr(f(depth)), ptr(offsets), c_u32(P), c_u32(gx), ptr(keys), ptr(vals), stream())
            keys, perm = torch.sort(keys)
            point_list = vals[perm].contiguous()
            call('mve_gs_tile_ranges', ptr(keys), c_u32(L), ptr(ranges), stream())
        else:
            point_list = torch.zeros(1, dtype=torch.int32, device=dev)
        color = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
        odepth, alpha, final_T = (torch.empty(H, W, dtype=torch.float32, device=dev) for _ in range(3))
        n_contrib = torch.empty(H, W, dtype=torch.int32, device=dev)
        bg_h = (ctypes.c_float * 3)(*[float(v) for v in bg])
        call('mve_gs_blend_forward', ptr(ranges), ptr(point_list), ptr(xy_c), ptr(co), ptr(ft), bg_h, c_u32(W), c_u32(H), ptr(color), ptr(odepth),
             ptr(alpha), ptr(final_T), ptr(n_contrib), stream())
        ctx.save_for_backward(ranges, point_list, xy_c, co, ft, final_T, n_contrib)
        ctx.bg, ctx.hw, ctx.n_instances = bg_h, (H, W), L
        _BlendFn.last_instances = L
        return color, odepth, alpha

    @staticmethod
    def backward(ctx, g_color, g_depth, g_alpha):
        ranges, point_list, xy_c, co, ft, final_T, n_contrib = ctx.saved_tensors
        H, W = ctx.hw
        P = xy_c.shape[0]
        d_xy, d_co, d_ft = torch.zeros_like(xy_c), torch.zeros_like(co), torch.zeros_like(ft)
        z = lambda g: None if g is None else g.float().contiguous()
        gc = z(g_color) if g_color is not None else torch.zeros(H, W, 3, device=xy_c.device)
        call('mve_gs_blend_backward', ptr(ranges), ptr(point_list), ptr(xy_c), ptr(co), ptr(ft), ctx.bg, c_u32(W), c_u32(H), ptr(final_T),
             ptr(n_contrib), ptr(gc), ptr(z(g_depth)), ptr(z(g_alpha)), ptr(d_xy), ptr(d_co), ptr(d_ft), stream())
        return d_xy, d_co[:, :3], d_co[:, 3].reshape(P), d_ft[:, :3], d_ft[:, 3], None, None, None, None


class GaussianRasterizationSettings(SimpleNamespace):
    """image_height, image_width, viewmatrix [4,4] world -> camera (OpenCV), intrinsics (fx, fy, cx, cy), bg (3 floats)."""


class GaussianRasterizer(torch.nn.Module):
    """``GaussianRasterizer(raster_settings)(means3D, opacities, colors_precomp, scales, rotations)`` -> (color [3,H,W], depth [1,H,W],
    alpha [1,H,W]) like the ashawkey fork of diff_gaussian_rasterization (precomputed colours: SH degree 0)."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, opacities, colors_precomp, scales, rotations):
        s = self.raster_settings
        if not means3D.is_cuda:
            raise RuntimeError('mvedit_b200 ops need CUDA tensors (no CPU fallback)')
        H, W = int(s.image_height), int(s.image_width)
        xy, conic, depth, rect = project_gaussians(means3D, scales, rotations, s.viewmatrix.to(means3D), s.intrinsics, H, W)
        color, odepth, alpha = _BlendFn.apply(xy, conic, opacities.reshape(-1), colors_precomp, depth, rect, tuple(float(v) for v in s.bg), H, W)
        return color.permute(2, 0, 1), odepth[None], alpha[None]

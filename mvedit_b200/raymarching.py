"""Host-side mirror of the reference's ray-marching operator interface.

Same function names, argument meaning, return values and autograd behaviour as
/root/reference/lib/ops/raymarching/raymarching.py (cited per function), so this module can be
bound where the reference binds ``lib.ops.raymarching``.  The arithmetic runs in
libmvedit_b200.so on the current CUDA stream.  Extra keyword ``noises=`` (not in the reference)
lets a caller supply the perturbation draws the reference takes from ``torch.rand`` internally
(raymarching.py:279-282, 474-478) -- parity tests need identical draws.

No CPU fallback: CPU tensors are moved to the current CUDA device exactly like the reference does
(raymarching.py:46-51); without a GPU the call raises.
"""
from itertools import groupby

import torch
from torch.autograd import Function

from ._lib import call, ptr, stream, c_int, c_u32, c_f32


def _cuda_f32(t):
    if not t.is_cuda:
        t = t.cuda()
    return t.float().contiguous()


# ----------------------------------------------------------------------------------------------
# utils
# ----------------------------------------------------------------------------------------------

def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:31-67. rays_o/d [N,3] (any leading shape), aabb [6] -> nears [N], fars [N]."""
    rays_o = _cuda_f32(rays_o).view(-1, 3)
    rays_d = _cuda_f32(rays_d).view(-1, 3)
    aabb = _cuda_f32(aabb)
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    fars = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    call('mve_near_far_from_aabb', ptr(rays_o), ptr(rays_d), ptr(aabb), c_u32(N), c_f32(min_near), ptr(nears), ptr(fars), stream())
    return nears, fars


def batch_near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:70-97."""
    if isinstance(rays_o, torch.Tensor):
        assert rays_o.size() == rays_d.size()
        num_scenes, num_rays, _ = rays_o.size()
        nears, fars = near_far_from_aabb(rays_o.reshape(num_scenes * num_rays, 3), rays_d.reshape(num_scenes * num_rays, 3),
                                         aabb, min_near)
        nears = nears.reshape(num_scenes, num_rays)
        fars = fars.reshape(num_scenes, num_rays)
    else:
        if len(rays_o) == 1:
            nears, fars = near_far_from_aabb(rays_o[0], rays_d[0], aabb, min_near)
            nears, fars = [nears], [fars]
        else:
            num_rays_per_scene = [r.size(0) for r in rays_o]
            nears, fars = near_far_from_aabb(torch.cat(rays_o, dim=0), torch.cat(rays_d, dim=0), aabb, min_near)
            nears = nears.split(num_rays_per_scene)
            fars = fars.split(num_rays_per_scene)
    return nears, fars


def morton3D(coords):
    """raymarching.py:132-151. coords int [N,3] -> int32 [N]."""
    if not coords.is_cuda:
        coords = coords.cuda()
    coords = coords.int().contiguous()
    N = coords.shape[0]
    indices = torch.empty(N, dtype=torch.int32, device=coords.device)
    call('mve_morton3D', ptr(coords), c_u32(N), ptr(indices), stream())
    return indices


def morton3D_invert(indices):
    """raymarching.py:156-175. indices int [N] -> int32 [N,3]."""
    if not indices.is_cuda:
        indices = indices.cuda()
    indices = indices.int().contiguous()
    N = indices.shape[0]
    coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
    call('mve_morton3D_invert', ptr(indices), c_u32(N), ptr(coords), stream())
    return coords


def packbits(grid, thresh, bitfield=None):
    """raymarching.py:180-203. grid float [C, H^3] (fp16 or fp32; the reference casts to fp32 first, the kernel
    here reads fp16 directly) -> uint8 [C*H^3/8] (written in place if given)."""
    if not grid.is_cuda:
        grid = grid.cuda()
    if grid.dtype not in (torch.float16, torch.float32):
        grid = grid.float()
    grid = grid.contiguous()
    N = grid.numel() // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    call('mve_packbits', ptr(grid), c_int(int(grid.dtype == torch.float16)), c_u32(N), c_f32(float(thresh)), ptr(bitfield), stream())
    return bitfield


# ----------------------------------------------------------------------------------------------
# train
# ----------------------------------------------------------------------------------------------

def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars,
                     perturb=False, dt_gamma=0, max_steps=1024, contract=False, noises=None, max_points=None, zero_tail=True, want_dirs=True):
    """raymarching.py:232-311.  Returns xyzs [M,3], dirs [M,3], ts [M,2], rays int32 [N,2] = (offset, count).

    Reference protocol: count pass -> host reads M -> alloc -> write pass.  Here the count pass is the fused kernel
    without outputs; the host read of M is kept because the caller receives exactly-sized tensors.
    If ``max_points`` is given no host sync happens: capacity-sized buffers are returned together with the device
    counter as a 5th value (B200-native protocol used by VolumeRenderer)."""
    rays_o = _cuda_f32(rays_o).view(-1, 3)
    rays_d = _cuda_f32(rays_d).view(-1, 3)
    if not density_bitfield.is_cuda:
        density_bitfield = density_bitfield.cuda()
    density_bitfield = density_bitfield.contiguous()
    nears, fars = _cuda_f32(nears), _cuda_f32(fars)
    N = rays_o.shape[0]
    dev = rays_o.device
    if noises is None and perturb:
        noises = torch.rand(N, dtype=torch.float32, device=dev)
    elif noises is not None:
        noises = _cuda_f32(noises)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    rays = torch.empty(N, 2, dtype=torch.int32, device=dev)
    dtg = 0.0 if isinstance(dt_gamma, torch.Tensor) else float(dt_gamma)   # a tensor dt_gamma is read on the device (graph-safe)
    common = (ptr(rays_o), ptr(rays_d), ptr(density_bitfield), c_f32(bound), c_int(int(contract)), c_f32(dtg),
              c_u32(max_steps), c_u32(N), c_u32(C), c_u32(H), ptr(nears), ptr(fars), ptr(noises))
    if max_points is not None:
        alloc = torch.zeros if zero_tail else torch.empty
        xyzs = alloc(max_points, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(max_points, 3, dtype=torch.float32, device=dev) if want_dirs else None
        ts = alloc(max_points, 2, dtype=torch.float32, device=dev)
        # single march: the counting pass records each sample's t, the write pass is a coalesced expansion (no second grid walk)
        t_scratch = None if (contract or N * max_steps > (1 << 27)) else torch.empty(N * max_steps, dtype=torch.float32, device=dev)
        call('mve_march_rays_train', *common, ptr(xyzs), ptr(dirs), ptr(ts), c_u32(max_points), ptr(rays), ptr(counter),
             ptr(dt_gamma if isinstance(dt_gamma, torch.Tensor) else None), ptr(t_scratch), stream())
        return xyzs, dirs, ts, rays, counter
    call('mve_march_rays_train', *common, ptr(None), ptr(None), ptr(None), c_u32(0), ptr(rays), ptr(counter), ptr(None), ptr(None), stream())
    M = int(counter.item())
    xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
    ts = torch.empty(M, 2, dtype=torch.float32, device=dev)
    if M > 0:
        call('mve_march_rays_train_write', *common, ptr(xyzs), ptr(dirs), ptr(ts), c_u32(M), ptr(rays), stream())
    return xyzs, dirs, ts, rays


class _composite_rays_train(Function):
    """raymarching.py:314-368."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, ts, rays, T_thresh=1e-4, binarize=False, m_dev=None, entropy=None):
        sigmas = sigmas.float().contiguous()
        rgbs = rgbs.float().contiguous()
        ts = ts.float().contiguous()
        rays = rays.int().contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        # the kernel writes every sample covered by a ray; samples outside any ray keep the 0 the reference guarantees
        weights = torch.zeros(M, dtype=torch.float32, device=dev)
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        call('mve_composite_rays_train_forward', ptr(sigmas), ptr(rgbs), ptr(ts), ptr(rays), c_u32(M), ptr(m_dev), c_u32(N),
             c_f32(T_thresh), c_int(int(binarize)), ptr(weights), ptr(weights_sum), ptr(depth), ptr(image), stream())
        ctx.save_for_backward(sigmas, rgbs, ts, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh, binarize]
        ctx.m_dev = m_dev
        ctx.entropy = entropy        # (device scalar weight, python scale) or None: fused sample-entropy gradient
        if entropy is not None:
            ctx.mark_non_differentiable(weights)      # its gradient is produced inside the backward kernel
            ctx.set_materialize_grads(False)
        return weights, weights_sum, depth, image

    @staticmethod
    def backward(ctx, grad_weights, grad_weights_sum, grad_depth, grad_image):
        sigmas, rgbs, ts, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh, binarize = ctx.dims
        z = lambda g, like: torch.zeros_like(like) if g is None else g.float().contiguous()
        grad_weights = None if (ctx.entropy is not None or grad_weights is None) else grad_weights.float().contiguous()
        grad_weights_sum, grad_depth, grad_image = z(grad_weights_sum, weights_sum), z(grad_depth, depth), z(grad_image, image)
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        call('mve_composite_rays_train_backward', ptr(grad_weights), ptr(grad_weights_sum), ptr(grad_depth), ptr(grad_image),
             ptr(sigmas), ptr(rgbs), ptr(ts), ptr(rays), ptr(weights_sum), ptr(depth), ptr(image), c_u32(M), ptr(ctx.m_dev), c_u32(N),
             c_f32(T_thresh), c_int(int(binarize)), ptr(ctx.entropy[0] if ctx.entropy else None),
             c_f32(ctx.entropy[1] if ctx.entropy else 0.0), ptr(grad_sigmas), ptr(grad_rgbs), stream())
        return grad_sigmas, grad_rgbs, None, None, None, None, None, None


composite_rays_train = _composite_rays_train.apply


def all_equal(iterable):
    g = groupby(iterable)
    return next(g, True) and not next(g, False)


def batch_composite_rays_train(sigmas, rgbs, ts, rays, num_points, T_thresh=1e-4, binarize=False):
    """raymarching.py:376-424."""
    num_scenes = len(ts)
    if num_scenes > 1:
        ts_ = torch.cat(ts, dim=0)
        rays_, num_rays = [], []
        point_offset_total = 0
        for ray_single, num_points_per_scene in zip(rays, num_points):
            rays_.append(torch.stack([ray_single[:, 0] + point_offset_total, ray_single[:, 1]], dim=-1))
            point_offset_total += num_points_per_scene
            num_rays.append(ray_single.size(0))
        rays_ = torch.cat(rays_, dim=0)
        weights, weights_sum_, depth_, image_ = composite_rays_train(sigmas, rgbs, ts_, rays_, T_thresh, binarize)
        if all_equal(num_rays):
            weights_sum = weights_sum_.reshape(num_scenes, num_rays[0])
            depth = depth_.reshape(num_scenes, num_rays[0])
            image = image_.reshape(num_scenes, num_rays[0], 3)
        else:
            weights_sum = weights_sum_.split(num_rays, dim=0)
            depth = depth_.split(num_rays, dim=0)
            image = image_.split(num_rays, dim=0)
    else:
        weights, weights_sum, depth, image = composite_rays_train(sigmas, rgbs, ts[0], rays[0], T_thresh, binarize)
        weights_sum, depth, image = weights_sum[None], depth[None], image[None]
    return weights, weights_sum, depth, image


# ----------------------------------------------------------------------------------------------
# infer
# ----------------------------------------------------------------------------------------------

def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
               perturb=False, dt_gamma=0, max_steps=1024, contract=False, noises=None):
    """raymarching.py:431-491.  -> xyzs [n_alive*n_step,3], dirs [...,3], ts [...,2] (zeros past a ray's end)."""
    rays_o = _cuda_f32(rays_o).view(-1, 3)
    rays_d = _cuda_f32(rays_d).view(-1, 3)
    dev = rays_o.device
    M = n_alive * n_step
    xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
    ts = torch.empty(M, 2, dtype=torch.float32, device=dev)
    if noises is None and perturb:
        noises = torch.rand(n_alive, dtype=torch.float32, device=dev)
    elif noises is not None:
        noises = _cuda_f32(noises)
    call('mve_march_rays', c_u32(n_alive), c_u32(n_step), ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), c_f32(bound),
         c_int(int(contract)), c_f32(float(dt_gamma)), c_u32(max_steps), c_u32(C), c_u32(H), ptr(density_bitfield.contiguous()),
         ptr(near), ptr(far), ptr(xyzs), ptr(dirs), ptr(ts), ptr(noises), stream())
    return xyzs, dirs, ts


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh=1e-2,
                   binarize=False):
    """raymarching.py:494-524.  In place on rays_alive, rays_t, weights_sum, depth, image."""
    sigmas = sigmas.float().contiguous()
    rgbs = rgbs.float().contiguous()
    call('mve_composite_rays', c_u32(n_alive), c_u32(n_step), c_f32(T_thresh), c_int(int(binarize)), ptr(rays_alive), ptr(rays_t),
         ptr(sigmas), ptr(rgbs), ptr(ts), ptr(weights_sum), ptr(depth), ptr(image), stream())
    return tuple()

"""Seeded synthetic inputs shared by tests, smoke() and bench.py."""
import math

import numpy as np


def _expand_bits(v):
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def morton3d_np(x, y, z):
    x, y, z = (a.astype(np.uint64) for a in (x, y, z))
    return (_expand_bits(x) | (_expand_bits(y) << 1) | (_expand_bits(z) << 2)).astype(np.int64)


def sphere_density_grid(H=128, bound=1.0, radius=0.5, dtype=np.float32):
    """Morton-ordered density grid [H^3]: 1 inside the sphere |x|<radius (voxel centres), 0 outside."""
    c = np.arange(H)
    xx, yy, zz = np.meshgrid(c, c, c, indexing='ij')
    pos = (np.stack([xx, yy, zz], -1).reshape(-1, 3).astype(np.float32) - (H - 1) / 2) * (2 * bound / H)
    occ = (np.linalg.norm(pos, axis=-1) < radius).astype(dtype)
    grid = np.zeros(H ** 3, dtype)
    grid[morton3d_np(xx.reshape(-1), yy.reshape(-1), zz.reshape(-1))] = occ
    return grid


def pack_bitfield_np(grid, thresh=0.5):
    bits = (grid.reshape(-1, 8) >= thresh).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(-1).astype(np.uint8)


def surround_poses(n_views, radius=3.7, elev_lo=0.0, elev_hi=0.6, seed=0):
    """c2w [n,4,4] OpenCV convention (x right, y down, z forward), cameras on a sphere looking at the origin."""
    rng = np.random.default_rng(seed)
    az = np.linspace(0, 2 * math.pi, n_views, endpoint=False)
    el = rng.uniform(elev_lo, elev_hi, n_views)
    poses = np.zeros((n_views, 4, 4), np.float32)
    for i in range(n_views):
        pos = radius * np.array([math.cos(el[i]) * math.cos(az[i]), math.cos(el[i]) * math.sin(az[i]), math.sin(el[i])])
        fwd = -pos / np.linalg.norm(pos)
        up = np.array([0, 0, 1.0])
        right = np.cross(fwd, up); right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        poses[i, :3, 0], poses[i, :3, 1], poses[i, :3, 2], poses[i, :3, 3] = right, down, fwd, pos
        poses[i, 3, 3] = 1
    return poses


def camera_rays(poses, size, fov_deg=30.0):
    """rays_o, rays_d [n*size*size, 3] float32, unit directions; returns also focal in pixels."""
    f = 0.5 * size / math.tan(math.radians(fov_deg) / 2)
    j, i = np.meshgrid(np.arange(size, dtype=np.float32) + 0.5, np.arange(size, dtype=np.float32) + 0.5, indexing='ij')
    d_cam = np.stack([(i - size / 2) / f, (j - size / 2) / f, np.ones_like(i)], -1).reshape(-1, 3)
    d_cam /= np.linalg.norm(d_cam, axis=-1, keepdims=True)
    ro, rd = [], []
    for p in poses:
        rd.append(d_cam @ p[:3, :3].T)
        ro.append(np.broadcast_to(p[:3, 3], rd[-1].shape))
    return np.concatenate(ro).astype(np.float32), np.concatenate(rd).astype(np.float32), f

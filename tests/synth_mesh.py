"""Seeded synthetic meshes / cameras for the mesh-stage tests."""
import math

import numpy as np

from .synth import surround_poses as _surround


def icosphere(subdiv=2):
    """Unit icosphere: v [V,3] float64, f [F,3] int64, outward counter-clockwise faces, closed manifold."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.stack(v), np.array(f, np.int64)


def surround_poses(n, seed=0, radius=3.7):
    return _surround(n, radius=radius, seed=seed).astype(np.float64)


def intrinsics(size, fov_deg=30.0):
    f = 0.5 * size / math.tan(math.radians(fov_deg) / 2)
    return np.array([f, f, size / 2, size / 2], np.float64)


def project(v, poses, fov_deg=30.0, near=0.01, far=100.0, return_cam=False):
    """World vertices [V,3] -> clip space [B,V,4] the way MeshRenderer.forward does (base_mesh_renderer.py:222-237): OpenCV c2w poses,
    y / z columns flipped to OpenGL, intrinsics (f, f, s/2, s/2) of a square image."""
    B = poses.shape[0]
    s = 2.0
    fx, fy, cx, cy = intrinsics(s, fov_deg)
    r = np.concatenate([poses[:, :3, :1], -poses[:, :3, 1:3]], axis=-1)
    proj = np.zeros((B, 4, 4))
    proj[:, 0, 0] = 2 * fx / s
    proj[:, 0, 2] = -2 * cx / s + 1
    proj[:, 1, 1] = -2 * fy / s
    proj[:, 1, 2] = -2 * cy / s + 1
    proj[:, 2, 2] = -(far + near) / (far - near)
    proj[:, 2, 3] = -(2 * far * near) / (far - near)
    proj[:, 3, 2] = -1
    v_cam = (v[None] - poses[:, None, :3, 3]) @ r
    v_clip = np.concatenate([v_cam, np.ones_like(v_cam[..., :1])], axis=-1) @ proj.transpose(0, 2, 1)
    return (v_clip, v_cam) if return_cam else v_clip

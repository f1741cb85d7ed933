"""numpy/ctypes front end of oracle/raymarching_oracle.c (CPU restatement of
/root/reference/lib/ops/raymarching/src/raymarching.cu).

TEST INFRASTRUCTURE ONLY -- see the header of raymarching_oracle.c.  Function
names and argument order mirror the reference's Python wrappers
(/root/reference/lib/ops/raymarching/raymarching.py), with numpy arrays in
place of CUDA tensors and the random noise passed in explicitly
(the reference draws it inside, raymarching.py:279-282,474-478).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    so = os.path.join(_HERE, 'liboracle.so')
    src = os.path.join(_HERE, 'raymarching_oracle.c')
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, 'liboracle.so'], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_march_rays_train.restype = ctypes.c_int64
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3), _f(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), ctypes.c_uint32(N), ctypes.c_float(min_near), _p(nears), _p(fars))
    return nears, fars


def morton3D(coords):
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().orc_morton3D(_p(coords), ctypes.c_uint32(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().orc_morton3D_invert(_p(indices), ctypes.c_uint32(N), _p(out))
    return out


def packbits(grid, thresh, bitfield=None):
    grid = _f(grid)
    N = grid.size // 8
    if bitfield is None:
        bitfield = np.empty(N, np.uint8)
    lib().orc_packbits(_p(grid), ctypes.c_uint32(N), ctypes.c_float(thresh), _p(bitfield))
    return bitfield


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, noises=None, dt_gamma=0.0,
                     max_steps=1024, contract=False):
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    nears, fars = _f(nears), _f(fars)
    N = rays_o.shape[0]
    noises = np.zeros(N, np.float32) if noises is None else _f(noises)
    grid = np.ascontiguousarray(density_bitfield, dtype=np.uint8)
    rays = np.empty((N, 2), np.int32)
    args = (_p(rays_o), _p(rays_d), _p(grid), ctypes.c_float(bound), ctypes.c_int(int(contract)), ctypes.c_float(dt_gamma),
            ctypes.c_uint32(max_steps), ctypes.c_uint32(N), ctypes.c_uint32(C), ctypes.c_uint32(H), _p(nears), _p(fars), _p(noises))
    M = lib().orc_march_rays_train(*args, None, None, None, _p(rays), ctypes.c_int64(0))
    xyzs, dirs, ts = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    if M > 0:
        lib().orc_march_rays_train(*args, _p(xyzs), _p(dirs), _p(ts), _p(rays), ctypes.c_int64(M))
    return xyzs, dirs, ts, rays


def composite_rays_train_forward(sigmas, rgbs, ts, rays, T_thresh=1e-4, binarize=False):
    sigmas, rgbs, ts = _f(sigmas), _f(rgbs), _f(ts)
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    weights = np.zeros(M, np.float32)
    weights_sum, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().orc_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(ts), _p(rays), ctypes.c_uint32(M), ctypes.c_uint32(N),
                                           ctypes.c_float(T_thresh), ctypes.c_int(int(binarize)),
                                           _p(weights), _p(weights_sum), _p(depth), _p(image))
    return weights, weights_sum, depth, image


def composite_rays_train_backward(grad_weights, grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, ts, rays,
                                  weights_sum, depth, image, T_thresh=1e-4, binarize=False):
    a = [_f(x) for x in (grad_weights, grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, ts)]
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    weights_sum, depth, image = _f(weights_sum), _f(depth), _f(image)
    M, N = a[4].shape[0], rays.shape[0]
    grad_sigmas, grad_rgbs = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().orc_composite_rays_train_backward(*[_p(x) for x in a], _p(rays), _p(weights_sum), _p(depth), _p(image),
                                            ctypes.c_uint32(M), ctypes.c_uint32(N), ctypes.c_float(T_thresh),
                                            ctypes.c_int(int(binarize)), _p(grad_sigmas), _p(grad_rgbs))
    return grad_sigmas, grad_rgbs


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
               noises=None, dt_gamma=0.0, max_steps=1024, contract=False):
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    rays_alive = np.ascontiguousarray(rays_alive, dtype=np.int32)
    rays_t, near, far = _f(rays_t), _f(near), _f(far)
    noises = np.zeros(n_alive, np.float32) if noises is None else _f(noises)
    grid = np.ascontiguousarray(density_bitfield, dtype=np.uint8)
    Mp = n_alive * n_step
    xyzs, dirs, ts = np.zeros((Mp, 3), np.float32), np.zeros((Mp, 3), np.float32), np.zeros((Mp, 2), np.float32)
    lib().orc_march_rays(ctypes.c_uint32(n_alive), ctypes.c_uint32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d),
                         ctypes.c_float(bound), ctypes.c_int(int(contract)), ctypes.c_float(dt_gamma), ctypes.c_uint32(max_steps),
                         ctypes.c_uint32(C), ctypes.c_uint32(H), _p(grid), _p(near), _p(far), _p(xyzs), _p(dirs), _p(ts), _p(noises))
    return xyzs, dirs, ts


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh=1e-2,
                   binarize=False):
    """In place on rays_alive (int32), rays_t, weights_sum, depth, image (float32, C-contiguous numpy)."""
    for a in (rays_t, weights_sum, depth, image):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    assert rays_alive.dtype == np.int32 and rays_alive.flags.c_contiguous
    sigmas, rgbs, ts = _f(sigmas), _f(rgbs), _f(ts)
    lib().orc_composite_rays(ctypes.c_uint32(n_alive), ctypes.c_uint32(n_step), ctypes.c_float(T_thresh), ctypes.c_int(int(binarize)),
                             _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs), _p(ts), _p(weights_sum), _p(depth), _p(image))

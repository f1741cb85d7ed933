"""ctypes binding of libmvedit_b200.so (C ABI declared in include/mvedit_b200.h).

Fails loudly: no library -> ImportError-like RuntimeError at first use; non-zero return -> RuntimeError
carrying mve_last_error().  No fallback path exists.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmvedit_b200.so')
_lib = None

c_void_p, c_int, c_u32, c_f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_float


def get_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libmvedit_b200.so is not built (run `python -m mvedit_b200.build` or __graft_entry__.build()); '
                'mvedit_b200 has no fallback path')
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mve_last_error.restype = ctypes.c_char_p
        _lib.mve_version.restype = c_int
    return _lib


def ptr(t):
    """Raw device pointer of a contiguous CUDA tensor (or NULL for None)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError('mvedit_b200 ops need CUDA tensors (no CPU fallback); got device %s' % t.device)
    if not t.is_contiguous():
        raise RuntimeError('mvedit_b200 ops need contiguous tensors')
    return c_void_p(t.data_ptr())


def raw_ptr(t):
    """Device pointer of a (possibly strided) CUDA tensor view; the caller passes the strides explicitly."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError('mvedit_b200 ops need CUDA tensors (no CPU fallback); got device %s' % t.device)
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def check(code, name):
    if code != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, code, get_lib().mve_last_error().decode()))


# kernels launched per C-ABI call (memsets not counted); used for the bench's "gpu_launches" claim
KERNELS_PER_CALL = {'mve_groupnorm_bf16': 2, 'mve_field_backward': 2, 'mve_density_grid_update': 2, 'mve_march_rays_train': 2}
LAUNCHES = [0]
PROFILE = [None]      # set to a list to record (name, start_event, end_event, meta) per call (bench.py roofline pass)


def call(name, *args, _meta=None):
    fn = getattr(get_lib(), name)
    LAUNCHES[0] += KERNELS_PER_CALL.get(name, 1)
    prof = PROFILE[0]
    if prof is None or torch.cuda.is_current_stream_capturing():
        check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), name)
    e1.record()
    prof.append((name, e0, e1, _meta))

"""Host harness of the mesh rasteriser -- TEST INFRASTRUCTURE ONLY.

``mvedit_b200/csrc/mesh_raster.cu`` and ``mesh_loss.cu`` are written so that they also compile as plain C++ (``-DMVE_HOST_HARNESS``): the per-triangle /
per-pixel device functions are then driven by serial loops and the ``mve_*`` entry points take HOST pointers.  The CPU test-suite uses
this to check, without a GPU, (1) the kernels' arithmetic bit for bit against ``oracle/raster_oracle.py`` and (2) the Python autograd
mirror ``mvedit_b200/mesh_raster.py`` (argument order, gradient plumbing) by routing its ``call`` / ``ptr`` / ``stream`` to this
library.  The product never loads it: ``mvedit_b200._lib`` only opens ``libmvedit_b200.so`` and refuses CPU tensors.
"""
import contextlib
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'mvedit_b200', 'csrc')
SRCS = [os.path.join(CSRC, f) for f in ('mesh_raster.cu', 'mesh_loss.cu')]
DEPS = SRCS + [os.path.join(CSRC, 'host_dual.cuh')]
OUT_DIR = os.path.join(ROOT, 'tests', '_host')
LIB = os.path.join(OUT_DIR, 'libmesh_raster_host.so')
_lib = None


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in DEPS):
        cmd = ['g++', '-x', 'c++', '-std=c++17', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-fvisibility=hidden', '-DMVE_HOST_HARNESS',
               '-I', CSRC] + SRCS + ['-o', LIB]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('host harness build failed:\n' + r.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


@contextlib.contextmanager
def routed(module):
    """Route ``module``'s C-ABI calls (its ``call`` / ``ptr`` / ``stream`` names) to the host harness, for CPU tensors."""
    h = lib()

    def call(name, *args, _meta=None):
        code = getattr(h, name)(*args)
        if code != 0:
            raise RuntimeError('%s failed on the host harness (%d)' % (name, code))

    def ptr(t):
        if t is None:
            return ctypes.c_void_p(0)
        assert (not t.is_cuda) and t.is_contiguous()
        return ctypes.c_void_p(t.data_ptr())

    saved = (module.call, module.ptr, module.stream)
    module.call, module.ptr, module.stream = call, ptr, (lambda: ctypes.c_void_p(0))
    try:
        yield
    finally:
        module.call, module.ptr, module.stream = saved

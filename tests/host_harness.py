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
_shimmed = {}


def shimmed(source):
    """An element-wise kernel source of the product (e.g. ``nerf_loss.cu``) compiled UNCHANGED as C++ through tests/host_shim/cuda_host_shim.h:
    launches become serial loops, ``common.cuh`` becomes the shim.  -> ctypes library whose ``mve_*`` entry points take host pointers."""
    import re
    if source in _shimmed:
        return _shimmed[source]
    os.makedirs(OUT_DIR, exist_ok=True)
    src = os.path.join(CSRC, source)
    shim_dir = os.path.join(ROOT, 'tests', 'host_shim')
    out_cpp = os.path.join(OUT_DIR, source.replace('.cu', '_host.cpp'))
    asan = os.environ.get('MVE_HOST_ASAN') == '1'          # AddressSanitizer build (run pytest under LD_PRELOAD=$(gcc -print-file-name=libasan.so))
    out_so = os.path.join(OUT_DIR, 'lib' + source.replace('.cu', '_host_asan.so' if asan else '_host.so'))
    deps = [src, os.path.join(shim_dir, 'cuda_host_shim.h'), os.path.join(CSRC, 'tonemap.cuh'), os.path.abspath(__file__)]
    if (not os.path.exists(out_so)) or os.path.getmtime(out_so) < max(os.path.getmtime(d) for d in deps):
        text = open(src).read()
        text = text.replace('#include "common.cuh"', '#include "cuda_host_shim.h"').replace('#include "../../include/mvedit_b200.h"',
                                                                                            '#include "mvedit_b200.h"')

        def launch(m):
            cfg, depth, cur = [], 0, ''
            for ch in m.group(2):                # split the launch configuration at top-level commas only
                depth += ch in '([' and 1 or (ch in ')]' and -1 or 0)
                if ch == ',' and depth == 0:
                    cfg.append(cur.strip()); cur = ''
                else:
                    cur += ch
            cfg.append(cur.strip())
            return 'SHIM_LAUNCH(%s, %s, %s, %s);' % (m.group(1), cfg[0], cfg[1], m.group(3))
        text, n = re.subn(r'(\w+)<<<([^>]*)>>>\((.*?)\);', launch, text)
        assert n > 0 and '<<<' not in text
        open(out_cpp, 'w').write(text)
        cmd = ['g++', '-std=c++17', '-O1', '-ffp-contract=off', '-fPIC', '-shared', '-Wno-unknown-pragmas', '-I', shim_dir, '-I', CSRC,
               '-I', os.path.join(ROOT, 'include'), out_cpp, '-o', out_so] + (['-fsanitize=address', '-fno-omit-frame-pointer', '-g'] if asan else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('host shim build of %s failed:\n%s' % (source, r.stderr))
    _shimmed[source] = ctypes.CDLL(out_so)
    return _shimmed[source]


class Libraries:
    """Several host libraries behind one lookup (``routed(module, Libraries(a, b))``)."""

    def __init__(self, *libs):
        self.libs = libs

    def __getattr__(self, name):
        for l in self.libs:
            try:
                return getattr(l, name)
            except AttributeError:
                pass
        raise AttributeError(name)


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
def routed(module, library=None):
    """Route ``module``'s C-ABI calls (its ``call`` / ``ptr`` / ``stream`` names) to the host harness (or to ``library``, e.g. a
    ``shimmed`` source), for CPU tensors."""
    h = library if library is not None else lib()

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

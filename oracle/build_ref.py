"""Build recipe for ``oracle/_ref`` -- the reference's OWN ray-marching CUDA extension.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mvedit_b200/`` may import this.

The reference ships its ray-marching kernels as a JIT-built torch extension
(``/root/reference/lib/ops/raymarching/backend.py:33-41``; sources
``src/raymarching.cu`` + ``src/bindings.cpp``).  Those two files compile on
their own with nvcc 12.9 for sm_100a, so -- as the task allows for a reference
"whose path compiles from its own few source files" -- we compile them *where
they lie* under ``/root/reference`` (no copy of the sources enters this repo)
and drop only the resulting ``_raymarching_ref*.so`` into ``oracle/_ref/``.
``oracle/_ref/`` is git-ignored but NOT gpurun-ignored, so the binary travels
to the GPU box, where ``/root/reference`` does not exist.

The extension is CUDA-only (``raymarching.py:46-51`` forces ``.cuda()``): it is
used on the GPU box (a) as the bit-level pin for the CPU oracle and for our
kernels (tests/test_gpu_ref_parity.py) and (b) as the GPU baseline of BASELINE
config 5.  It cannot run in the CPU container.

Usage:  python oracle/build_ref.py            (idempotent; ~2-3 min)
"""
import os
import shutil
import sys

REF_SRC = '/root/reference/lib/ops/raymarching/src'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
NAME = '_raymarching_ref'


def built_path():
    if not os.path.isdir(OUT):
        return None
    for f in os.listdir(OUT):
        if f.startswith(NAME) and f.endswith('.so'):
            return os.path.join(OUT, f)
    return None


def build(verbose=False):
    if built_path() is not None:
        return built_path()
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: only the prebuilt file is used
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    from torch.utils.cpp_extension import load
    build_dir = os.path.join(OUT, 'build')
    os.makedirs(build_dir, exist_ok=True)
    # same flags as the reference's backend.py:6-12 (no arch there; arch from
    # TORCH_CUDA_ARCH_LIST as torch does for the reference)
    load(name=NAME,
         extra_cflags=['-O3', '-std=c++17'],
         extra_cuda_cflags=['-O3', '-std=c++17',
                            '-U__CUDA_NO_HALF_OPERATORS__', '-U__CUDA_NO_HALF_CONVERSIONS__',
                            '-U__CUDA_NO_HALF2_OPERATORS__'],
         sources=[os.path.join(REF_SRC, 'raymarching.cu'), os.path.join(REF_SRC, 'bindings.cpp')],
         build_directory=build_dir, verbose=verbose, is_python_module=False)
    so = os.path.join(build_dir, NAME + '.so')
    shutil.copy(so, os.path.join(OUT, NAME + '.so'))
    shutil.rmtree(build_dir, ignore_errors=True)
    return built_path()


def load_ref():
    """Import the prebuilt reference extension (GPU box). Returns module or None."""
    p = built_path()
    if p is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    p = build(verbose='-v' in sys.argv)
    print('oracle/_ref:', p)

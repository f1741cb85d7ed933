"""In-tree build of libmvedit_b200.so (hand-written CUDA for sm_100a, C ABI in include/mvedit_b200.h).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  `python -m mvedit_b200.build` or `__graft_entry__.build()`.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', '_obj')
LIB = os.path.join(HERE, 'libmvedit_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-ccbin', '/usr/bin/g++',
         '--expt-relaxed-constexpr', '-Xptxas', '-v']
# per-file flags: the mesh rasteriser's coverage test relies on a*b - c*d being exactly negated when the operands swap (watertight
# shared edges) and is bit-compared with the numpy oracle -> no FMA contraction there
FILE_FLAGS = {'mesh_raster.cu': ['--fmad=false']}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    m = 0
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cuh', '.h')):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    sp = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), _deps_mtime()):
        return obj, ''
    cmd = [NVCC] + ARCH + FLAGS + FILE_FLAGS.get(src, []) + ['-c', sp, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    with open(obj[:-2] + '.ptxas.log', 'w') as f:
        f.write(r.stderr)
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in res]
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart', '-ccbin', '/usr/bin/g++']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    return LIB


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))

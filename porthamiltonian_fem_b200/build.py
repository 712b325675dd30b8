"""In-tree build of libphwave.so (nvcc, sm_100a).  Used by __graft_entry__.build() and the tests."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libphwave.so')
SOURCES = ['phwave.cu', 'layout.cpp']
HEADERS = ['kernels.cuh', 'kernels_pipe.cuh', 'kernels_loop.cuh', 'kernels_ho.cuh', 'kernels_sweep.cuh', 'pipe_carve.hpp', 'layout.hpp', os.path.join('..', '..', 'include', 'phwave.h')]
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC,-O3,-fopenmp', '-shared']


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ['-o', LIB] + [os.path.join(CSRC, f) for f in SOURCES] + ['-lgomp']
    if verbose:
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libphwave.so')
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True, extra=['-Xptxas', '-v'] if '--ptxas' in sys.argv else [])

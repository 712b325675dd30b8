"""Builds tests/emu/_build/libphwemu.so: the kernels of porthamiltonian_fem_b200/csrc compiled for the host with g++ on top of the
SPMD emulator (emu_rt.hpp).  The kernel headers are copied into _build unchanged except for one textual rewrite: the classic
`extern __shared__ double smem[];` declaration of the non-pipelined kernels becomes a pointer to the emulator's buffer (the
pipelined kernels already declare their dynamic shared memory through PHW_DYN_SMEM)."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'porthamiltonian_fem_b200', 'csrc')
BUILD = os.path.join(HERE, '_build')
LIB = os.path.join(BUILD, 'libphwemu.so')
HEADERS = ['kernels.cuh', 'kernels_pipe.cuh', 'kernels_sweep.cuh', 'pipe_carve.hpp', 'layout.hpp']


def build(verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    deps = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(CSRC, 'layout.cpp')] + \
           [os.path.join(HERE, f) for f in ('emu_main.cpp', 'emu_rt.hpp', 'cuda_runtime.h', 'cooperative_groups.h', 'build_emu.py')]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    for h in HEADERS:
        src = open(os.path.join(CSRC, h)).read()
        src = re.sub(r'extern __shared__ double (\w+)\[\];', r'double* \1 = reinterpret_cast<double*>(phw_emu::dyn_smem());', src)
        open(os.path.join(BUILD, h), 'w').write(src)
    cmd = ['g++', '-std=c++17', '-O1', '-g', '-fPIC', '-shared', '-pthread', '-DPHW_HOST_EMU', '-Wno-attributes', '-Wno-unused',
           '-I', BUILD, '-I', HERE, os.path.join(HERE, 'emu_main.cpp'), os.path.join(CSRC, 'layout.cpp'), '-o', LIB]
    if verbose:
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('g++ failed building the emulator')
    return LIB


if __name__ == '__main__':
    print(build(verbose=True))

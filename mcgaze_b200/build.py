"""Build the CUDA shared library in-tree (mcgaze_b200/libmcgaze_b200.so) for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box
with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmcgaze_b200.so')
SOURCES = ['mcg_api.cu']
HEADERS = ['common.cuh', 'ptx.cuh', 'umma_gemm.cuh', 'stem_fused.cuh', 'simt_gemm.cuh', 'head_kernels.cuh',
           os.path.join('..', '..', 'include', 'mcgaze_b200.h')]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source with `-gencode arch=compute_100a,code=sm_100a -lineinfo`."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '-Xcompiler', '-fPIC', '-shared'] + (['-Xptxas', '-v'] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB + '.tmp']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed building libmcgaze_b200.so')
    if verbose:
        sys.stderr.write(r.stderr)
    os.replace(LIB + '.tmp', LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

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
OBJ = os.path.join(HERE, 'build')
API_H = os.path.join('..', '..', 'include', 'mcgaze_b200.h')
# source -> headers it depends on (each source is compiled to its own object, then linked)
SOURCES = {
    'mcg_api.cu': ['common.cuh', 'ptx.cuh', 'umma_gemm.cuh', 'bneck_fused.cuh', 'stem_fused.cuh', 'simt_gemm.cuh', 'head_kernels.cuh', API_H],
    'preprocess.cu': ['common.cuh', API_H],
    'metric.cu': ['common.cuh', API_H],
    'png_decode.cu': ['common.cuh', 'png_core.cuh', API_H],
}


def _newer(deps, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in deps)


def _run(cmd, verbose: bool) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed building libmcgaze_b200.so')
    if verbose:
        sys.stderr.write(r.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source with `-gencode arch=compute_100a,code=sm_100a -lineinfo` and link the library."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    flags = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC'] + \
            (['-Xptxas', '-v'] if verbose else [])
    os.makedirs(OBJ, exist_ok=True)
    objs, relink = [], force or not os.path.exists(LIB)
    for src, deps in SOURCES.items():
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + '.o')
        objs.append(obj)
        if force or _newer([src] + deps, obj):
            _run([nvcc] + flags + ['-c', os.path.join(CSRC, src), '-o', obj], verbose)
            relink = True
    if relink or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        _run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared'] + objs + ['-o', LIB + '.tmp'], verbose)
        os.replace(LIB + '.tmp', LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

"""CPU: the C-ABI library builds, loads and exports every symbol include/mcgaze_b200.h declares;
no compute is attempted without a GPU, and the product refuses to run without one."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from mcgaze_b200 import build, lib as L
    build.build()
    return L


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'mcgaze_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(mcg_[a-z_]+)\s*\(', hdr)))
    assert declared and set(declared) == set(lib.EXPORTS)
    so = lib.load_library()
    for name in declared:
        assert getattr(so, name) is not None
    assert b'sm_100a' in so.mcg_version()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    so = lib.load_library()
    h = ctypes.c_void_p()
    t = (lib.mcg_tensor * 1)()
    t[0].name = b'x'
    assert so.mcg_create(ctypes.byref(h), 0, t, 1, 0) == -2
    assert b'no CPU fallback' in so.mcg_last_error()
    with pytest.raises(lib.McgError):
        lib.Engine({'x': torch.zeros(1)})


def test_preprocess_has_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    so = lib.load_library()
    f = (lib.mcg_frame * 1)()
    c3 = (ctypes.c_float * 3)(1, 1, 1)
    assert so.mcg_preprocess(f, 1, c3, c3, 1, None, 32, 32, None) == -2
    assert b'no CPU fallback' in so.mcg_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'mcgaze_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f
                assert 'mcgaze_oracle' not in src, f


def test_sass_has_tcgen05_and_tma(lib):
    """The built cubin really contains tcgen05 MMA / TMEM loads / TMA (incl. im2col) instructions."""
    import shutil
    import subprocess
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run(['cuobjdump', '-sass', lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'IM2COL'):
        assert mnemonic in sass, mnemonic

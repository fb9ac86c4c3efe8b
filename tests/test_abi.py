"""The C ABI: header and library agree, the library loads without a GPU, and the product
package refuses to run without its CUDA extension."""

import ctypes
import os
import re

import numpy as np
import pytest

from sporco_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'spcsc.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(spcsc_[a-z0-9_]+)\s*\(', src)))


def test_binding_covers_header():
    assert header_functions() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    import shutil
    if shutil.which('nvcc') or os.path.exists('/usr/local/cuda/bin/nvcc'):
        from sporco_b200 import build
        build.build()                      # no-op when the sources are unchanged
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), 'libspcsc.so does not export %s' % name
    lib.spcsc_version.restype = ctypes.c_int
    assert lib.spcsc_version() >= 100


def test_no_cpu_fallback_without_device():
    """On a box without a GPU constructing a solver must raise, never compute on the host."""
    lib = _lib.load()
    if lib.spcsc_device_count() > 0:
        pytest.skip('a CUDA device is present')
    from sporco_b200.admm import cbpdn
    D = np.ones((3, 3, 2), np.float32)
    S = np.ones((8, 8), np.float32)
    with pytest.raises(_lib.SpcscError):
        cbpdn.ConvBPDN(D, S, 0.1)


def test_missing_extension_is_loud(tmp_path):
    with pytest.raises(ImportError):
        _lib.load(path=str(tmp_path / 'nope.so'))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'sporco_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run with -m gpu on a B200)')


@pytest.fixture(scope='session')
def emu_library():
    """The kernel sources compiled for the CPU emulation harness (tests/emu) -- test
    infrastructure that lets kernel logic be checked without a GPU."""
    import ctypes
    from tests.emu import build_emu
    from sporco_b200 import _lib
    return _lib._declare(ctypes.CDLL(build_emu.build()))

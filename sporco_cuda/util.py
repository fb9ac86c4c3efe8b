"""Device utility functions of the ``sporco_cuda`` interface
(docs/source/modules/sporco.cuda.rst:40-104)."""

import ctypes

from sporco_b200 import _lib

__all__ = ['device_count', 'current_device', 'memory_info', 'device_name']

_current = [0]


def device_count():
    """Number of CUDA devices (0 when the extension or a device is missing)."""
    try:
        return max(int(_lib.load().spcsc_device_count()), 0)
    except ImportError:
        return 0


def current_device(id=None):
    """Get, or set and get, the device used by functions that take no explicit device."""
    if id is not None:
        if not 0 <= int(id) < device_count():
            raise ValueError('invalid device number %r' % (id,))
        _current[0] = int(id)
    return _current[0]


def memory_info():
    """(free, total) bytes on the current device."""
    free = ctypes.c_uint64(0)
    total = ctypes.c_uint64(0)
    _lib.check(_lib.load().spcsc_memory_info(_current[0], ctypes.byref(free), ctypes.byref(total)))
    return int(free.value), int(total.value)


def device_name(dev=0):
    """Hardware model name of device `dev`."""
    buf = ctypes.create_string_buffer(256)
    _lib.check(_lib.load().spcsc_device_name(int(dev), buf, 256))
    return buf.value.decode()

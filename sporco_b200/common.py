"""Infrastructure shared by the iterative solvers.

Host-side mirror of ``sporco.common`` (sporco/common.py:84-294): a metaclass that builds
the per-class ``IterationStats`` named tuple from the ``itstat_fields_*`` attributes and
stops the ``init`` timer once construction is complete, ``set_dtype`` / ``set_attr``
helpers and the formatter of the ``Verbose`` status table.
"""

import collections
import re

import numpy as np


class _IterSolver_Meta(type):
    def __init__(cls, *args):
        type.__init__(cls, *args)
        nt = collections.namedtuple('IterationStats', cls.itstat_fields())
        nt.__module__ = cls.__module__
        nt.__qualname__ = cls.__qualname__ + '.IterationStats'
        cls.IterationStats = nt

    def __call__(cls, *args, **kwargs):
        obj = type.__call__(cls, *args, **kwargs)
        obj.timer.stop('init')
        return obj


class IterativeSolver(metaclass=_IterSolver_Meta):
    """Base of all solver classes."""

    itstat_fields_objfn = ()
    itstat_fields_alg = ()
    itstat_fields_extra = ()

    @classmethod
    def itstat_fields(cls):
        return ('Iter',) + cls.itstat_fields_objfn + cls.itstat_fields_alg + \
            cls.itstat_fields_extra + ('Time',)

    def set_dtype(self, opt, dtype):
        """``self.dtype`` from `dtype`, unless option ``DataType`` overrides it; a dtype
        that is already set wins over both (sporco/common.py:146-173)."""
        if getattr(self, 'dtype', None) is None:
            self.dtype = np.dtype(dtype if opt['DataType'] is None else opt['DataType'])

    def set_attr(self, name, val, dval=None, dtype=None, reset=False):
        """Assign attribute `name` from `val` (or the default `dval` when `val` is None),
        cast to `dtype`; existing non-None values are kept unless `reset`
        (sporco/common.py:177-226)."""
        if val is None:
            val = dval
        if dtype is not None and val is not None:
            val = dtype(val) if isinstance(dtype, type) else dtype.type(val)
        if reset or getattr(self, name, None) is None:
            setattr(self, name, val)


def solve_status_str(hdrlbl, fmtmap=None, fwdth0=4, fwdthdlt=6, fprec=2):
    """Header line, row format and separator length of the ``Verbose`` status table
    (sporco/common.py:230-294)."""
    fmtmap = fmtmap or {}
    wfloat = fprec + fwdthdlt
    fmts = []
    for i, lbl in enumerate(hdrlbl):
        if lbl in fmtmap:
            fmts.append(fmtmap[lbl])
        elif i == 0:
            fmts.append('%%%dd' % fwdth0)
        else:
            fmts.append('%%%d.%de' % (wfloat, fprec))
    widths = []
    for f in fmts:
        m = re.match(r'%-?(\d+)', f)
        if m is None:
            raise ValueError("Format string '%s' does not contain field width" % f)
        widths.append(int(m.group(1)))
    hdr = '  '.join('%-*s' % (w, t) for t, w in zip(hdrlbl, widths))
    return hdr, '  '.join(fmts), len(hdr)


def transpose_ntpl_list(lst):
    """List of named tuples -> named tuple of arrays (sporco/util.py transpose_ntpl_list)."""
    if not lst:
        return None
    cls = type(lst[0])
    return cls(*[np.array([getattr(t, f) for t in lst]) for f in cls._fields])


def real_dtype(dtype):
    """Real dtype matching the precision of `dtype` (sporco/fft.py:76-102)."""
    return np.dtype(np.float32) if np.dtype(dtype) in (np.dtype(np.float32),
                                                       np.dtype(np.complex64)) \
        else np.dtype(np.float64)


def complex_dtype(dtype):
    """Complex dtype matching the precision of `dtype` (sporco/fft.py:44-72)."""
    return np.dtype(np.complex64) if np.dtype(dtype) in (np.dtype(np.float32),
                                                         np.dtype(np.complex64)) \
        else np.dtype(np.complex128)

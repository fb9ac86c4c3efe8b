"""Array-role bookkeeping for convolutional sparse representations.

Host-side mirror of the parts of ``sporco.cnvrep`` the ConvBPDN path touches:
:class:`CSC_ConvRepIndexing` (sporco/cnvrep.py:24-199) and :func:`l1Wshape`
(sporco/cnvrep.py:492-550).  External arrays keep the reference's layout,
``S (N0, N1, C, K, 1)``, ``D (hd, wd, Cd, 1, M)``, ``X (N0, N1, Cx, K, M)``.
"""

import numpy as np


class CSC_ConvRepIndexing(object):
    """Infer which axes of `D` and `S` are spatial, channel, signal and filter axes."""

    def __init__(self, D, S, dimK=None, dimN=2):
        nd_extra = S.ndim - dimN
        self.dimCd = D.ndim - (dimN + 1)
        self.Cd = D.shape[-2] if self.dimCd else 1
        if dimK is None:
            if nd_extra == 0:
                dimC = dimK = 0
            elif nd_extra == 1:
                dimC = self.dimCd          # S is taken to have as many channel axes as D
                dimK = 1 - dimC
            else:
                dimC = dimK = 1
        else:
            dimC = nd_extra - dimK
        self.dimN, self.dimC, self.dimK = dimN, dimC, dimK
        self.C = S.shape[dimN] if dimC == 1 else 1
        if self.Cd > 1 and self.C != self.Cd:
            raise ValueError("Multi-channel dictionary with signal with mismatched number "
                             "of channels (Cd=%d, C=%d)" % (self.Cd, self.C))
        self.K = S.shape[dimN + dimC] if dimK == 1 else 1
        self.M = D.shape[-1]
        self.Nv = tuple(S.shape[:dimN])
        self.N = int(np.prod(self.Nv))
        self.axisN = tuple(range(dimN))
        self.axisC, self.axisK, self.axisM = dimN, dimN + 1, dimN + 2
        self.shpD = tuple(D.shape[:dimN]) + (self.Cd, 1, self.M)
        self.shpS = self.Nv + (self.C, self.K, 1)
        self.shpX = self.Nv + (self.C - self.Cd + 1, self.K, self.M)

    def __str__(self):
        return '\n'.join('%-6s %s' % (k, getattr(self, k)) for k in
                         ('dimN', 'dimC', 'dimK', 'C', 'Cd', 'K', 'M', 'Nv', 'shpD', 'shpS',
                          'shpX'))


def l1Wshape(W, cri):
    """Internal 5-D (broadcastable) shape of an l1 weight array given in external form."""
    sdim = cri.dimN + cri.dimC + cri.dimK
    if W.ndim < sdim:
        if W.size != 1:
            raise ValueError('weight array must be scalar or have at least the same number '
                             'of dimensions as input array')
        return (1,) * (cri.dimN + 3)
    if W.ndim == sdim:
        return W.shape + (1,) * (3 - cri.dimC - cri.dimK)
    if W.ndim == cri.dimN + 3:
        return W.shape
    return W.shape[:-1] + (1,) * (2 - cri.dimC - cri.dimK) + W.shape[-1:]


# ---- dictionary-update side (sporco/cnvrep.py:277-470, 868-1074) ----
# `dsz` is (hd, wd, M) or (hd, wd, Cd, M), or -- a multi-scale dictionary -- a tuple of such tuples, one per block of
# equally sized filters (cnvrep.py:277-360; blocks that differ in their channel count are not supported).

def _is_multiscale(dsz):
    return isinstance(dsz[0], (tuple, list))


def _blocks(dsz, dimN):
    """Normalised list of blocks [(hd, wd, [Cd,] Mb), ...] of a (single- or multi-scale) specification."""
    if not _is_multiscale(dsz):
        dsz = (dsz,)
    out = []
    for b in dsz:
        if isinstance(b[0], (tuple, list)):
            raise NotImplementedError('dictionary blocks that differ in their channel structure are not supported')
        if len(b) not in (dimN + 1, dimN + 2):
            raise ValueError('dsz must have dimN+1 or dimN+2 entries (per block)')
        out.append(tuple(int(v) for v in b))
    if len(set(len(b) for b in out)) != 1 or (len(out[0]) == dimN + 2 and len(set(b[dimN] for b in out)) != 1):
        raise NotImplementedError('all blocks of a multi-scale dictionary must have the same number of channels')
    return out


def _single_support(dsz, dimN):
    """(max hd, max wd, [Cd,] M) of the specification: the support that holds every filter."""
    bl = _blocks(dsz, dimN)
    mx = tuple(max(b[i] for b in bl) for i in range(dimN))
    return mx + tuple(bl[0][dimN:-1]) + (sum(b[-1] for b in bl),)


def filter_supports(dsz, dimN=2):
    """Per filter: its support size, as an (M, dimN) integer array."""
    rows = []
    for b in _blocks(dsz, dimN):
        rows += [list(b[:dimN])] * b[-1]
    return np.array(rows, dtype=np.int32)


class CDU_ConvRepIndexing(object):
    """Array roles for the dictionary update: `dsz` is (hd, wd, M) or (hd, wd, Cd, M) or a tuple of such blocks;
    `self.dsz` is the support that holds every filter, `self.dsz_spec` what was given."""

    def __init__(self, dsz, S, dimK=None, dimN=2):
        self.dsz_spec = dsz
        self.multiscale = _is_multiscale(dsz)
        dsz = _single_support(dsz, dimN)
        self.dsz = dsz
        self.dimCd = len(dsz) - dimN - 1
        self.Cd = dsz[dimN] if self.dimCd else 1
        self.M = dsz[-1]
        nd_extra = S.ndim - dimN
        if dimK is None:
            if nd_extra == 0:
                dimC = dimK = 0
            elif nd_extra == 1:
                dimC = self.dimCd
                dimK = 1 - dimC
            else:
                dimC = dimK = 1
        else:
            dimC = nd_extra - dimK
        self.dimN, self.dimC, self.dimK = dimN, dimC, dimK
        self.C = S.shape[dimN] if dimC == 1 else 1
        self.Cx = self.C - self.Cd + 1
        if self.Cd > 1 and self.C != self.Cd:
            raise ValueError("Multi-channel dictionary with signal with mismatched number "
                             "of channels (Cd=%d, C=%d)" % (self.Cd, self.C))
        self.K = S.shape[dimN + dimC] if dimK == 1 else 1
        self.Nv = tuple(S.shape[:dimN])
        self.N = int(np.prod(self.Nv))
        self.axisN = tuple(range(dimN))
        self.axisC, self.axisK, self.axisM = dimN, dimN + 1, dimN + 2
        self.shpD = self.Nv + (self.Cd, 1, self.M)
        self.shpS = self.Nv + (self.C, self.K, 1)
        self.shpX = self.Nv + (self.Cx, self.K, self.M)


def stdformD(D, Cd, M, dimN=2):
    """Dictionary with explicit channel and (singleton) signal axes (cnvrep.py:473-489)."""
    return D.reshape(D.shape[0:dimN] + (Cd, 1, M))


def zpad(x, Nv):
    """Zero-pad the leading (spatial) axes of `x` to `Nv` (cnvrep.py:876-891)."""
    out = np.zeros(tuple(Nv) + x.shape[len(Nv):], dtype=x.dtype)
    out[tuple(slice(0, n) for n in x.shape[:len(Nv)])] = x
    return out


def bcrop(x, dsz, dimN=2):
    """Crop the leading (spatial) axes to the filter support; with a multi-scale specification to the largest
    support, every block of filters zero outside its own (cnvrep.py:894-950)."""
    mx = _single_support(dsz, dimN)
    out = x[tuple(slice(0, n) for n in mx[:dimN])]
    if not _is_multiscale(dsz):
        return out
    out = np.array(out)
    m0 = 0
    for b in _blocks(dsz, dimN):
        keep = np.zeros(mx[:dimN], dtype=bool)
        keep[tuple(slice(0, n) for n in b[:dimN])] = True
        blk = out[..., m0:m0 + b[-1]]
        blk[~keep] = 0
        m0 += b[-1]
    return out


def zeromean(v, dsz, dimN=2):
    """Subtract, per filter and channel, the mean over the filter's own support (cnvrep.py:609-668)."""
    vz = v.copy()
    m0 = 0
    for b in _blocks(dsz, dimN):
        sl = tuple(slice(0, n) for n in b[:dimN]) + (Ellipsis, slice(m0, m0 + b[-1]))
        vz[sl] -= np.mean(v[sl], axis=tuple(range(dimN)))
        m0 += b[-1]
    return vz


def normalise(v, dimN=2):
    """Scale every filter to unit l2 norm over its first `dimN` axes; zero filters stay zero
    (cnvrep.py:823-848)."""
    ax = tuple(range(dimN))
    vn = np.sqrt(np.sum(v ** 2, ax, keepdims=True))
    vn[vn == 0] = 1.0
    return np.asarray(v / vn, dtype=v.dtype)


def Pcn(x, dsz, Nv, dimN=2, dimC=1, crp=False, zm=False):
    """Projection onto the constraint set: support `dsz`, optional zero mean, unit norm
    (cnvrep.py:953-1033)."""
    pad = (lambda a: a) if crp else (lambda a: zpad(a, Nv))
    zmf = (lambda a: zeromean(a, dsz, dimN)) if zm else (lambda a: a)
    return normalise(zmf(pad(bcrop(x, dsz, dimN))), dimN + dimC)


def getPcn(dsz, Nv, dimN=2, dimC=1, crp=False, zm=False):
    """Closure form of :func:`Pcn` (cnvrep.py:1036-1074)."""
    return lambda x: Pcn(x, dsz, Nv, dimN, dimC, crp, zm)


def mskWshape(W, cri):
    """Internal 5-D (broadcastable) shape of a data-fidelity mask given in external form
    (sporco/cnvrep.py:553-605)."""
    ck = W.ndim - cri.dimN
    if ck >= 2:
        return W.shape + (1,) if ck == 2 else W.shape
    if ck == 1:
        if cri.C == 1 and cri.K > 1:
            return W.shape[0:cri.dimN] + (1, W.shape[cri.dimN]) + (1,)
        return W.shape[0:cri.dimN] + (W.shape[cri.dimN], 1) + (1,)
    return W.shape + (1,) * (3 - ck)

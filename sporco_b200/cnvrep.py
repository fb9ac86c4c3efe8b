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

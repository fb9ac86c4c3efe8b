"""CPU oracle for the pre-processing step -- TEST INFRASTRUCTURE, not product code.
numpy restatement of ``sporco.signal.tikhonov_filter`` (sporco/signal.py:244-303), pinned to the
reference by oracle/make_golden.py (fixture tests/golden/tikhonov.npz)."""

import numpy as np

from . import cbpdn_oracle as co


def tikhonov_filter(s, lmbda, npd=16, fft=None):
    fft = fft or co.FFTBackend()
    grv = np.array([-1.0, 1.0]).reshape([2, 1])
    gcv = np.array([-1.0, 1.0]).reshape([1, 2])
    shp = (s.shape[0] + 2 * npd, s.shape[1] + 2 * npd)
    Gr = fft.rfftn(grv, shp, (0, 1))
    Gc = fft.rfftn(gcv, shp, (0, 1))
    A = 1.0 + lmbda * (np.conj(Gr) * Gr + np.conj(Gc) * Gc).real
    if s.ndim > 2:
        A = A[(slice(None),) * 2 + (np.newaxis,) * (s.ndim - 2)]
    sp = np.pad(s, ((npd, npd),) * 2 + ((0, 0),) * (s.ndim - 2), 'symmetric')
    spf = fft.rfftn(sp, None, (0, 1))
    spf /= A
    slp = fft.irfftn(spf, sp.shape[:2], (0, 1))
    sl = slp[npd:(slp.shape[0] - npd), npd:(slp.shape[1] - npd)]
    sh = s - sl
    return sl.astype(s.dtype), sh.astype(s.dtype)

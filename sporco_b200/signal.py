"""Signal pre-processing on the device (mirror of the part of ``sporco.signal`` the ConvBPDN
example scripts use either side of the solver)."""

import numpy as np

from . import _lib


def tikhonov_filter(s, lmbda, npd=16, device=0):
    r"""Lowpass filter based on Tikhonov regularisation (sporco/signal.py:244-303): returns the
    lowpass component :math:`\mathbf{x} = \mathrm{argmin} (1/2)\|\mathbf{x}-\mathbf{s}\|^2 +
    (\lambda/2) \sum_i \|G_i \mathbf{x}\|^2` of each image (computed on a symmetrically padded
    copy, then cropped) and the highpass remainder.  `s` has the two image axes first; any
    further axes index independent images, as in the reference."""
    s = np.asarray(s)
    if s.ndim < 2:
        raise ValueError('input must have at least two dimensions')
    if not np.isrealobj(s):
        raise NotImplementedError('complex input is not supported')
    dt = s.dtype if s.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    x = np.ascontiguousarray(np.moveaxis(s.reshape(s.shape[:2] + (-1,)), 2, 0), dtype=dt)
    sl, sh = _lib.tikhonov_filter(x, lmbda, npd, device)
    back = lambda a: np.moveaxis(a, 0, 2).reshape(s.shape).astype(s.dtype, copy=False)
    return back(sl), back(sh)

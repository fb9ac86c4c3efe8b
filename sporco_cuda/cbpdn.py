"""Functional ConvBPDN interface of ``sporco_cuda``
(docs/source/modules/sporco.cuda.rst:106-251).

``cbpdn(D, S, lmbda, opt, dev=0)`` is the GPU counterpart of
``sporco.admm.cbpdn.ConvBPDN(D, S, lmbda, opt).solve()`` for a single greyscale image;
float32 throughout, as the original extension.  The gradient-regularised and masked
variants of the original extension are not implemented and raise.
"""

import numpy as np

from sporco_b200.admm import cbpdn as _cbpdn

__all__ = ['cbpdn', 'cbpdngrd', 'cbpdnmsk', 'cbpdngrdmsk']


def _options(opt):
    if opt is None:
        return _cbpdn.ConvBPDN.Options()
    if isinstance(opt, _cbpdn.ConvBPDN.Options):
        return opt
    keys = _cbpdn.ConvBPDN.Options.defaults

    def prune(d, ref):
        out = {}
        for k, v in dict.items(d):
            if k not in ref:
                continue                     # options of other solver classes are ignored
            out[k] = prune(v, ref[k]) if isinstance(v, dict) and isinstance(ref[k], dict) else v
        return out
    return _cbpdn.ConvBPDN.Options(prune(opt, keys))


def cbpdn(D, S, lmbda, opt=None, dev=0):
    """Solve convolutional BPDN for one image on GPU `dev`; returns X with shape (N0, N1, M)."""
    D = np.asarray(D, dtype=np.float32)
    S = np.asarray(S, dtype=np.float32)
    if D.ndim != 3 or S.ndim != 2:
        raise ValueError('cbpdn expects a three dimensional dictionary and a two dimensional '
                         'signal (single image, single channel)')
    o = _options(opt)
    o['DataType'] = np.float32
    b = _cbpdn.ConvBPDN(D, S, np.float32(lmbda), o, dimK=0, device=dev)
    X = b.solve()
    return np.ascontiguousarray(X.reshape(S.shape + (D.shape[-1],)))


def _missing(name):
    def f(*args, **kwargs):
        raise NotImplementedError('sporco_cuda.%s is not implemented by the sporco_b200 backend'
                                  % name)
    f.__name__ = name
    return f


cbpdngrd = _missing('cbpdngrd')
cbpdnmsk = _missing('cbpdnmsk')
cbpdngrdmsk = _missing('cbpdngrdmsk')

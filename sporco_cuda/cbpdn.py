"""Functional ConvBPDN interface of ``sporco_cuda``
(docs/source/modules/sporco.cuda.rst:106-251).

``cbpdn(D, S, lmbda, opt, dev=0)`` is the GPU counterpart of
``sporco.admm.cbpdn.ConvBPDN(D, S, lmbda, opt).solve()`` for a single greyscale image;
float32 throughout, as the original extension.  ``cbpdngrd`` (ConvBPDNGradReg), ``cbpdnmsk``
(AddMaskSim about ConvBPDN) and ``cbpdngrdmsk`` (AddMaskSim about ConvBPDNGradReg) follow the
same pattern; as documented for the original extension, the masked versions extend the
``L1Weight`` / ``GradWeight`` arrays for the AMS impulse filter themselves (sporco.cuda.rst:169-251).
"""

import numpy as np

from sporco_b200.admm import cbpdn as _cbpdn

__all__ = ['cbpdn', 'cbpdngrd', 'cbpdnmsk', 'cbpdngrdmsk']


def _options(opt, cls=None):
    cls = cls or _cbpdn.ConvBPDN
    if opt is None:
        return cls.Options()
    if isinstance(opt, cls.Options):
        return cls.Options(_plain(opt))
    keys = cls.Options.defaults

    def prune(d, ref):
        out = {}
        for k, v in dict.items(d):
            if k not in ref:
                continue                     # options of other solver classes are ignored
            out[k] = prune(v, ref[k]) if isinstance(v, dict) and isinstance(ref[k], dict) else v
        return out
    return cls.Options(prune(opt, keys))


def _plain(d):
    return {k: (_plain(v) if isinstance(v, dict) else v) for k, v in dict.items(d)}


def _check(D, S):
    D = np.asarray(D, dtype=np.float32)
    S = np.asarray(S, dtype=np.float32)
    if D.ndim != 3 or S.ndim != 2:
        raise ValueError('expected a three dimensional dictionary and a two dimensional '
                         'signal (single image, single channel)')
    return D, S


def _extend(w, M, fill):
    """Per-filter weight array with an entry appended for the AMS impulse filter."""
    w = np.asarray(w, dtype=np.float32)
    if w.ndim == 0 or w.shape[-1] != M:
        return w
    return np.concatenate((w, np.full(w.shape[:-1] + (1,), fill, dtype=np.float32)), axis=-1)


def cbpdn(D, S, lmbda, opt=None, dev=0):
    """Solve convolutional BPDN for one image on GPU `dev`; returns X with shape (N0, N1, M)."""
    D, S = _check(D, S)
    o = _options(opt)
    o['DataType'] = np.float32
    b = _cbpdn.ConvBPDN(D, S, np.float32(lmbda), o, dimK=0, device=dev)
    X = b.solve()
    return np.ascontiguousarray(X.reshape(S.shape + (D.shape[-1],)))


def cbpdngrd(D, S, lmbda, mu, opt=None, dev=0):
    """ConvBPDNGradReg for one image on GPU `dev`; returns X with shape (N0, N1, M)."""
    D, S = _check(D, S)
    o = _options(opt, _cbpdn.ConvBPDNGradReg)
    o['DataType'] = np.float32
    b = _cbpdn.ConvBPDNGradReg(D, S, np.float32(lmbda), np.float32(mu), o, dimK=0, device=dev)
    X = b.solve()
    return np.ascontiguousarray(X.reshape(S.shape + (D.shape[-1],)))


def cbpdnmsk(D, s, w, lmbda, opt=None, dev=0):
    """AddMaskSim about ConvBPDN with the {0,1} mask `w`; returns the primary maps (N0, N1, M)."""
    D, s = _check(D, s)
    o = _options(opt)
    o['DataType'] = np.float32
    o['L1Weight'] = _extend(o['L1Weight'], D.shape[-1], 0.0)
    b = _cbpdn.AddMaskSim(_cbpdn.ConvBPDN, D, s, np.asarray(w, dtype=np.float32),
                          np.float32(lmbda), o, dimK=0, device=dev)
    X = b.solve()
    return np.ascontiguousarray(X.reshape(s.shape + (D.shape[-1],)))


def cbpdngrdmsk(D, s, w, lmbda, mu, opt=None, dev=0):
    """AddMaskSim about ConvBPDNGradReg; returns the primary maps (N0, N1, M)."""
    D, s = _check(D, s)
    o = _options(opt, _cbpdn.ConvBPDNGradReg)
    o['DataType'] = np.float32
    M = D.shape[-1]
    o['L1Weight'] = _extend(o['L1Weight'], M, 0.0)
    gw = np.asarray(o['GradWeight'], dtype=np.float32)
    o['GradWeight'] = np.concatenate((np.broadcast_to(gw, (M,)), np.zeros(1, np.float32)))
    b = _cbpdn.AddMaskSim(_cbpdn.ConvBPDNGradReg, D, s, np.asarray(w, dtype=np.float32),
                          np.float32(lmbda), np.float32(mu), o, dimK=0, device=dev)
    X = b.solve()
    return np.ascontiguousarray(X.reshape(s.shape + (D.shape[-1],)))

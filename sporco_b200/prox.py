"""Proximal operators of the path on the GPU, under the reference's names (``sporco.prox.prox_l1``
prox/_lp.py:144-183, ``prox_sl1l2`` prox/_l21.py:51-88): stand-alone launches of the arithmetic that
the fused row kernels carry out in registers."""

from . import _lib


def prox_l1(v, alpha, device=0):
    return _lib.prox_l1(v, alpha, device)


def prox_sl1l2(v, alpha, beta, axis=None, device=0):
    return _lib.prox_sl1l2(v, alpha, beta, axis, device)

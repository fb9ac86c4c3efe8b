"""Level-1 linear algebra of the path on the GPU, under the reference's names
(``sporco.linalg.solvedbi_sm`` linalg.py:232-297, ``solvemdbi_ism`` linalg.py:370-444): stand-alone
launches of the arithmetic that the fused column kernels carry out in registers."""

from . import _lib


def solvedbi_sm(ah, rho, b, c=None, axis=4, device=0):
    """Solve ``(rho I + a a^H) x = b`` along `axis` (the filter axis M) for arrays in the
    (N0, N1f, 1, K, M) layout; `c` (the reference's cached component) is accepted and ignored."""
    if ah.shape[2] != 1:
        raise ValueError('solvedbi_sm takes a single-channel ah; see solvemdbi_ism')
    return _lib.solvedbi_sm(ah, rho, b, c, axis, device)


def solvemdbi_ism(ah, rho, b, axisM=4, axisK=2, device=0):
    """Solve ``(rho I + sum_c a_c a_c^H) x = b`` with the channels of `ah` on axis `axisK`
    (the reference iterates Sherman-Morrison over them; the device solves the equivalent C x C
    Hermitian system per frequency, equal to the reference to rounding)."""
    if axisM != 4 or axisK != 2:
        raise ValueError('expected axisM=4, axisK=2 (the (N0, N1f, C, K, M) layout)')
    return _lib.solvedbi_sm(ah, rho, b, None, 4, device)

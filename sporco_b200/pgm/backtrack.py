"""Step-size search by backtracking (mirror of sporco/pgm/backtrack.py)."""

import math


class BacktrackBase(object):
    def update(self, solverobj):
        raise NotImplementedError()


class BacktrackStandard(BacktrackBase):
    """Standard FISTA backtracking (sporco/pgm/backtrack.py:45-107): repeat the proximal step
    with L *= gamma_u until F(x) <= Q_L(x, y) or `maxiter` trials.  The trials themselves are
    device work (``spcsc_pgm_trial``); only the scalar test runs here, in the solver's
    working precision like the reference."""

    def __init__(self, gamma_u=1.2, maxiter=50):
        self.gamma_u = gamma_u
        self.maxiter = maxiter

    def update(self, solverobj):
        it = 0
        search = True
        while search and it < self.maxiter:
            f, q = solverobj._trial()
            if f <= q:
                search = False
            else:
                solverobj.L *= self.gamma_u
            it += 1
        solverobj.F = f
        solverobj.Q = q
        solverobj.iterBTrack = it
        solverobj.ystep()


class BacktrackRobust(BacktrackBase):
    """Robust backtracking of Florea and Vorobyov (sporco/pgm/backtrack.py:110-210): L is first decreased by
    gamma_d, the auxiliary point is y = (T_k x_prev + t z) / (T_k + t) with t = (1 + sqrt(1 + 4 L T_k)) / (2 L),
    and after the search z += t L (x - y).  The sequences x, y, z stay on the device (``spcsc_pgm_combine_y`` /
    ``spcsc_pgm_trial`` / ``spcsc_pgm_finish``); T_k, t and the F <= Q test are host scalars."""

    def __init__(self, gamma_d=0.9, gamma_u=2.0, maxiter=50):
        self.gamma_d, self.gamma_u, self.maxiter = gamma_d, gamma_u, maxiter
        self.Tk = 0.
        self.Zrb = None            # lives on the device

    def update(self, solverobj):
        solverobj.L *= self.gamma_d
        it = 0
        search = True
        while search and it < self.maxiter:
            L = float(solverobj.L)
            t = float(1. + math.sqrt(1. + 4. * L * self.Tk)) / (2. * L)
            T = self.Tk + t
            solverobj._combine_y(self.Tk / T, t / T, first=(it == 0))
            f, q = solverobj._trial()
            if f <= q:
                search = False
            else:
                solverobj.L *= self.gamma_u
            it += 1
        self.Tk = T
        solverobj._finish_robust(t * float(solverobj.L))
        solverobj.F = f
        solverobj.Q = q
        solverobj.iterBTrack = it

"""Step-size search by backtracking (mirror of sporco/pgm/backtrack.py)."""


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
    """Robust backtracking of sporco/pgm/backtrack.py:110-192 -- not implemented on the device
    yet; constructing a solver with it raises."""

    def __init__(self, gamma_d=0.9, gamma_u=2.0, maxiter=50):
        self.gamma_d, self.gamma_u, self.maxiter = gamma_d, gamma_u, maxiter

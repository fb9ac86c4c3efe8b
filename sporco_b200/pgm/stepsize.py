"""Step-size policies for the PGM solvers (mirror of sporco/pgm/stepsize.py:18-145).

The reference evaluates the policies on full gradient arrays; here the device reduces everything a
policy needs to a handful of scalars (``spcsc_pgm_policy_stats``: with R the residual spectrum at the
auxiliary point and G the Gram matrix of the dictionary spectra, ||grad||^2 = sum R^H G R,
<grad, Hess grad> = sum |G R|^2, and the two-point quantities of Barzilai-Borwein likewise from
per-frequency sums), so ``update`` receives the solver and those scalars.
"""


class StepSizePolicyBase(object):
    """Interface: ``update(solverobj, stats) -> L`` (the inverse step size)."""

    def update(self, solverobj, stats=None):
        raise NotImplementedError()


class StepSizePolicyCauchy(StepSizePolicyBase):
    r"""Cauchy step: alpha = ||grad||^2 / (grad^T Hess grad), L = 1 / alpha
    (sporco/pgm/stepsize.py:50-87)."""

    def update(self, solverobj, stats=None):
        if stats is None:
            stats = solverobj.policy_stats()
        ty = solverobj.dtype.type
        den = ty(stats[0])
        num = ty(stats[1])
        return num / den


class StepSizePolicyBB(StepSizePolicyBase):
    r"""Barzilai-Borwein step: alpha = dx^T dg / ||dg||^2 with dx, dg the changes of iterate and gradient
    since the previous proximal step, L = 1 / alpha; a negative L keeps the current one
    (sporco/pgm/stepsize.py:90-145).  The previous state lives on the device
    (``spcsc_pgm_policy_stats(store=1)``)."""

    def __init__(self):
        super(StepSizePolicyBB, self).__init__()
        self.xprv = 0.0
        self.gradprv = 0.0

    def store_prev_state(self, xprv=None, gradprv=None):
        """Kept for interface parity: the device remembers the previous iterate and gradient itself."""
        self.xprv, self.gradprv = xprv, gradprv

    def update(self, solverobj, stats=None):
        if stats is None:
            stats = solverobj.policy_stats()
        ty = solverobj.dtype.type
        den = ty(stats[3])
        num = ty(stats[2])
        L = num / den
        if L < 0.:
            L = solverobj.L
        return L

"""Momentum coefficient rules for the accelerated proximal gradient solvers.

Each rule maps the solver's momentum variable (the Nesterov sequence value t, or the iteration
count k for the linear rules, see ``PGM.var_momentum``) to the next value t+; the extrapolation
weight of the accelerated step is then ``(t - 1) / t+``.  Same class names and constructor
arguments as ``sporco.pgm.momentum`` (sporco/pgm/momentum.py:19-132).
"""

import math


class MomentumBase(object):
    """Interface: ``update(var) -> t+``."""

    def update(self, var):
        raise NotImplementedError()


class MomentumNesterov(MomentumBase):
    """t+ = (1 + sqrt(1 + 4 t^2)) / 2."""

    def update(self, t):
        return (1.0 + math.sqrt(1.0 + 4.0 * float(t) ** 2)) / 2.0


class MomentumGenLinear(MomentumBase):
    """t+ = (k + a) / b with the iteration count k."""

    def __init__(self, a=50., b=2.):
        self.a, self.b = a, b

    def update(self, k):
        return (k + self.a) / self.b


class MomentumLinear(MomentumGenLinear):
    """t+ = (k + b) / b: the generalised rule with a = b."""

    def __init__(self, b=2.):
        MomentumGenLinear.__init__(self, a=b, b=b)

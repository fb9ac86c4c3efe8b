"""Momentum coefficient rules (mirror of sporco/pgm/momentum.py)."""

import numpy as np


class MomentumBase(object):
    def update(self, var):
        raise NotImplementedError()


class MomentumNesterov(MomentumBase):
    """t+ = (1 + sqrt(1 + 4 t^2)) / 2   (sporco/pgm/momentum.py:45-48)."""

    def update(self, t):
        return 0.5 * float(1. + np.sqrt(1. + 4. * t ** 2))


class MomentumLinear(MomentumBase):
    """t+ = (k + b) / b   (sporco/pgm/momentum.py:78-101)."""

    def __init__(self, b=2.):
        self.b = b

    def update(self, k):
        return (k + self.b) / self.b


class MomentumGenLinear(MomentumBase):
    """t+ = (k + a) / b   (sporco/pgm/momentum.py:104-132)."""

    def __init__(self, a=50., b=2.):
        self.a = a
        self.b = b

    def update(self, k):
        return (k + self.a) / self.b

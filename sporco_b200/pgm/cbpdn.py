"""Convolutional BPDN by PGM / FISTA on the B200 engine.

Drop-in counterpart of ``sporco.pgm.cbpdn.ConvBPDN`` (sporco/pgm/cbpdn.py:29-384): same
constructor, ``Options`` tree, ``IterationStats`` fields (ObjFun, DFid, RegL1, Rsdl, F_Btrack,
Q_Btrack, IterBTrack, L) and ``solve / getcoef / setdict / reconstruct`` surface.
Supported on the device: fixed step 1/L, ``BacktrackStandard`` / ``BacktrackRobust``,
``StepSizePolicyCauchy`` / ``StepSizePolicyBB``, ``Monotone``; Nesterov / linear momentum.
"""

import copy

import numpy as np

from .. import _lib, cnvrep as cr
from . import pgm
from .backtrack import BacktrackStandard, BacktrackRobust     # noqa: F401  (re-exported for user code)


class ConvBPDN(pgm.PGMDFT):
    class Options(pgm.PGMDFT.Options):
        defaults = copy.deepcopy(pgm.PGMDFT.Options.defaults)
        defaults.update({'NonNegCoef': False, 'NoBndryCross': False})
        defaults.update({'L1Weight': 1.0})
        defaults.update({'L': 500.0})

        def __init__(self, opt=None):
            pgm.PGMDFT.Options.__init__(self, {} if opt is None else opt)

    itstat_fields_objfn = ('ObjFun', 'DFid', 'RegL1')
    hdrtxt_objfn = ('Fnc', 'DFid', u'Regℓ1')
    hdrval_objfun = {'Fnc': 'ObjFun', 'DFid': 'DFid', u'Regℓ1': 'RegL1'}

    def __init__(self, D, S, lmbda=None, opt=None, dimK=None, dimN=2, device=0):
        if dimN != 2:
            raise NotImplementedError('sporco_b200 implements the dimN=2 (image) case only')
        if not (np.isrealobj(D) and np.isrealobj(S)):
            raise NotImplementedError('complex-valued dictionaries / signals are not supported')
        opt = self._coerce_options(opt)
        if not hasattr(self, 'cri'):
            self.cri = cr.CSC_ConvRepIndexing(D, S, dimK=dimK, dimN=dimN)
        cri = self.cri
        self.set_dtype(opt, S.dtype)
        self.D = np.asarray(D.reshape(cri.shpD), dtype=self.dtype)
        self.S = np.asarray(S.reshape(cri.shpS), dtype=self.dtype)
        self._device = device
        self._cache = {}
        _lib.require_device()
        self._h = _lib.Handle(cri.Nv[0], cri.Nv[1], cri.C, cri.Cd, cri.K, cri.M,
                              self.D.shape[0], self.D.shape[1], self.dtype, device)
        self._h.set_signal(self.S[..., 0])
        self._h.set_dict(self.D[:, :, :, 0, :])

        if lmbda is None:                                # sporco/pgm/cbpdn.py:194-199
            b = np.conj(self.Df) * self.Sf
            lmbda = 0.1 * abs(b).max()
        self.lmbda = self.dtype.type(lmbda)
        self.wl1 = np.asarray(opt['L1Weight'], dtype=self.dtype)
        w = self.wl1.reshape(cr.l1Wshape(self.wl1, cri))
        self._h.set_l1_weight(np.ascontiguousarray(w))

        super(ConvBPDN, self).__init__(cri.shpX, cri.Nv, cri.axisN, S.dtype, opt)
        if opt['Monotone'] and opt['Backtrack'] is not None:
            raise NotImplementedError('Monotone together with backtracking is not supported on the device')
        x0 = opt['X0']
        self._h.pgm_reset(None if x0 is None else np.asarray(x0, dtype=self.dtype).reshape(cri.shpX))
        self._stats = None
        self._rsdl = None
        self._rejected = False

    # ---- state on the device
    def _fetch(self, which):
        if which not in self._cache:
            self._cache[which] = self._h.get_array(which)
        return self._cache[which]

    @property
    def X(self):
        return self._fetch(_lib.ARR_PGM_X)

    @property
    def Xf(self):
        return self._fetch(_lib.ARR_PGM_XF)

    @property
    def Yf(self):
        return self._fetch(_lib.ARR_PGM_YF)

    @property
    def Df(self):
        return self._fetch(_lib.ARR_DF)

    @property
    def Sf(self):
        return self._fetch(_lib.ARR_SF)

    def setdict(self, D=None):
        if D is not None:
            self.D = np.asarray(D, dtype=self.dtype).reshape(self.cri.shpD)
        self._cache.pop(_lib.ARR_DF, None)
        self._h.set_dict(self.D[:, :, :, 0, :])

    def attach_process_group(self, dist, group=None):
        """Shard the images over the ranks of a ``torch.distributed`` group: every rank keeps the images it
        was built with; the sums behind F, Q, the residual and the objective of every proximal step are
        reduced over the ranks on the device, so all ranks take the same backtracking / stopping decisions."""
        from .. import _dist
        self._world, self._p2p = _dist.attach(self._h, self._device, float(self.cri.K), dist, group)

    def getcoef(self):
        return self.X

    def reconstruct(self, X=None):
        if X is None:
            X = self.X
        return self._h.reconstruct(np.asarray(X, dtype=self.dtype).reshape(self.cri.shpX))

    # ---- one proximal step on the device; returns (F, Q) in the working precision
    def _trial(self):
        o = self.opt
        self._h.pgm_configure(self.lmbda, o['NonNegCoef'], o['NoBndryCross'])
        s = self._h.pgm_trial(float(self.L))
        for k in (_lib.ARR_PGM_X, _lib.ARR_PGM_XF, _lib.ARR_PGM_YF):
            self._cache.pop(k, None)
        self._stats = s
        self._rsdl = s[4]
        ty = self.dtype.type
        f = ty(s[0])
        # Q = f(y) + <x - y, grad f(y)> + (L/2) ||x - y||^2      (sporco/pgm/backtrack.py:93-95)
        q = ty(s[1]) + ty(s[2]) + (self.L / 2.) * ty(s[3])
        return f, q

    def policy_stats(self, store=False):
        """Scalars of the step-size policies / objective at the current state (spcsc_pgm_policy_stats)."""
        return self._h.pgm_policy_stats(store)

    def on_iteration_start(self):
        """sporco/pgm/pgm.py:835-846: with ``Monotone`` the objective of the previous iterate is what
        the next proximal step is compared with (the reference takes it from the initial state
        for the first TWO iterations: its ``objfn`` is only refreshed by steps with k > 0)."""
        self._rejected = False
        if self.opt['Monotone']:
            if self.k == 0:
                st = self.policy_stats()
                self.objfn = (st[4] + self.lmbda * st[5], st[4], st[5])
            self.objfn_prev = self.objfn

    def xstep(self):
        """Proximal step of sporco/pgm/pgm.py:779-811 with its optional step-size policy and the
        monotone safeguard: the policy sets L from the gradient at the auxiliary point before the step;
        a step that raises the objective is taken back (Xf only, as in the reference)."""
        if self.stepsizepolicy is not None:
            from .stepsize import StepSizePolicyBB
            bb = isinstance(self.stepsizepolicy, StepSizePolicyBB) or any(
                c.__name__ == 'StepSizePolicyBB' for c in type(self.stepsizepolicy).__mro__)
            st = self.policy_stats(store=bb)
            if self.k > 1:
                self.L = self.stepsizepolicy.update(self, st)
        self._trial()
        if self.opt['Monotone'] and self.k > 0:
            s = self._stats
            objfn = (s[5] + self.lmbda * s[6], s[5], s[6])
            if self.objfn_prev[0] < objfn[0]:
                self._rejected = True
                self.objfn = self.objfn_prev
            else:
                self.objfn = objfn

    def _combine_y(self, a, b, first):
        self._h.pgm_combine_y(a, b, first)
        self._cache.pop(_lib.ARR_PGM_YF, None)

    def _finish_robust(self, c0):
        self._rsdl = self._h.pgm_finish(_lib.PGM_FINISH_ROBUST, c0)
        for k in (_lib.ARR_PGM_XF,):
            self._cache.pop(k, None)

    def ystep(self):
        """Momentum step (sporco/pgm/pgm.py:815-831) on the device."""
        tprv = self.t
        self.t = self.momentum.update(self.var_momentum())
        if self._rejected:
            # monotone: Xf = Xfprv, so Yf = Xf + (tprv / t)(ZZf - Xf)
            self._rsdl = self._h.pgm_finish(_lib.PGM_FINISH_REJECT, tprv / self.t)
            self._cache.pop(_lib.ARR_PGM_XF, None)
        else:
            self._h.pgm_accept((tprv - 1.) / self.t)
        self._cache.pop(_lib.ARR_PGM_YF, None)

    def rsdl(self):
        return self.dtype.type(self._rsdl)

    def eval_objfn(self):
        if self.opt['Monotone']:         # the record carries the tracked objective (sporco/pgm/pgm.py:546-549)
            return self.objfn
        dfd = self._stats[5]
        rl1 = self._stats[6]
        return (dfd + self.lmbda * rl1, dfd, rl1)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None:
            h.close()


class ConvBPDNMask(ConvBPDN):
    """FISTA for convolutional BPDN with a spatial mask in the data fidelity term (mirror of
    sporco/pgm/cbpdn.py:387-508)::

        argmin_x (1/2) || W (sum_m d_m * x_m - s) ||_2^2 + lambda sum_m || x_m ||_1

    The gradient, the backtracking functional and DFid take the residual through the signal
    domain (``rfft(W^2 irfft(.))``); on the device these are transforms of the K*C residual
    planes only, next to the M-times larger coefficient arrays.
    """

    def __init__(self, D, S, lmbda, W=None, opt=None, dimK=None, dimN=2, device=0):
        super(ConvBPDNMask, self).__init__(D, S, lmbda, opt, dimK=dimK, dimN=dimN, device=device)
        if W is None:
            W = np.array([1.0], dtype=self.dtype)
        W = np.asarray(W)
        self.W = np.asarray(W.reshape(cr.mskWshape(W, self.cri)), dtype=self.dtype)
        self._h.pgm_set_mask(np.ascontiguousarray(self.W[..., 0]))

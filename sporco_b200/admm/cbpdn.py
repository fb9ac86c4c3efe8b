"""Convolutional BPDN by ADMM on the B200 engine.

Drop-in counterparts of ``sporco.admm.cbpdn.GenericConvBPDN``, ``ConvBPDN`` and
``ConvBPDNJoint`` (sporco/admm/cbpdn.py:30-808): same constructors, ``Options`` trees,
``IterationStats`` fields and ``solve / getcoef / setdict / reconstruct / getitstat``
surface.  All array arithmetic of the iteration -- the batched 2-D real FFTs, the
frequency-domain Sherman-Morrison solve, relaxation, the l1 / l2,1 proximal step, the dual
update and the residual norms that drive the automatic penalty update -- runs as CUDA
kernels inside libspcsc (see sporco_b200/csrc); this module only marshals options and
arrays across the C ABI.  There is no CPU fallback: unsupported configurations raise.
"""

import copy
import os

import numpy as np

from .. import _lib, cnvrep as cr, common
from . import admm




class GenericConvBPDN(admm.ADMMEqual):
    """Base of the convolutional BPDN solvers (mirror of sporco/admm/cbpdn.py:30-380)."""

    class Options(admm.ADMMEqual.Options):
        """Options of ``sporco.admm.cbpdn.GenericConvBPDN.Options`` (cbpdn.py:93-164).
        ``HighMemSolve`` is accepted and ignored: the device solve never caches the
        Sherman-Morrison vector and is always exact for the current rho."""

        defaults = copy.deepcopy(admm.ADMMEqual.Options.defaults)
        defaults.update({'AuxVarObj': False, 'fEvalX': True, 'gEvalY': False,
                         'ReturnX': False, 'HighMemSolve': False, 'LinSolveCheck': False,
                         'RelaxParam': 1.8, 'NonNegCoef': False, 'NoBndryCross': False})
        defaults['AutoRho'].update({'Enabled': True, 'Period': 1, 'AutoScaling': True,
                                    'Scaling': 1000.0, 'RsdlRatio': 1.2})

        def __init__(self, opt=None):
            admm.ADMMEqual.Options.__init__(self, {} if opt is None else opt)

        def __setitem__(self, key, value):
            admm.ADMMEqual.Options.__setitem__(self, key, value)
            if key == 'AuxVarObj':
                self['fEvalX'] = value is not True
                self['gEvalY'] = value is True

    itstat_fields_objfn = ('ObjFun', 'DFid', 'Reg')
    itstat_fields_extra = ('XSlvRelRes',)
    hdrtxt_objfn = ('Fnc', 'DFid', 'Reg')
    hdrval_objfun = {'Fnc': 'ObjFun', 'DFid': 'DFid', 'Reg': 'Reg'}

    _joint = False
    _two_reg = False          # a second regularisation column in the rows (RegL2 of ConvElasticNet)

    def __init__(self, D, S, opt=None, dimK=None, dimN=2, device=0):
        if dimN != 2:
            raise NotImplementedError('sporco_b200 implements the dimN=2 (image) case only')
        if not (np.isrealobj(D) and np.isrealobj(S)):
            raise NotImplementedError('complex-valued dictionaries / signals are not supported')
        opt = self._coerce_options(opt)
        self.set_dtype(opt, S.dtype)
        if not hasattr(self, 'cri'):
            self.cri = cr.CSC_ConvRepIndexing(D, S, dimK=dimK, dimN=dimN)
        cri = self.cri
        self._cache = {}
        self._h = None
        self._rho = None
        super(GenericConvBPDN, self).__init__(cri.shpX, S.dtype, opt)

        self.D = np.asarray(D.reshape(cri.shpD), dtype=self.dtype)
        self.S = np.asarray(S.reshape(cri.shpS), dtype=self.dtype)
        self._device = device
        self._open_handle()

    # ---- device handle -----------------------------------------------------------------
    def _open_handle(self):
        cri = self.cri
        _lib.require_device()
        self._h = _lib.Handle(cri.Nv[0], cri.Nv[1], cri.C, cri.Cd, cri.K, cri.M,
                              self.D.shape[0], self.D.shape[1], self.dtype, self._device)
        self._h.set_signal(self.S[..., 0])
        self._h.set_dict(self.D[:, :, :, 0, :])
        self._h.admm_reset(float(self._rho))
        y0, u0 = self.opt['Y0'], self.opt['U0']
        if y0 is not None:
            self._h.set_array(_lib.ARR_Y, np.asarray(y0, dtype=self.dtype).reshape(cri.shpX))
        if u0 is not None:
            self._h.set_array(_lib.ARR_U, np.asarray(u0, dtype=self.dtype).reshape(cri.shpX))

    @property
    def rho(self):
        return self._rho

    @rho.setter
    def rho(self, value):
        self._rho = value
        if getattr(self, '_h', None) is not None and value is not None:
            self._h.admm_set_rho(float(value))

    def _fetch(self, which):
        if which not in self._cache:
            self._cache[which] = self._h.get_array(which)
        return self._cache[which]

    @property
    def Y(self):
        return self._fetch(_lib.ARR_Y)

    @Y.setter
    def Y(self, value):
        self._cache.pop(_lib.ARR_Y, None)
        self._h.set_array(_lib.ARR_Y, np.asarray(value, dtype=self.dtype).reshape(self.cri.shpX))

    @property
    def U(self):
        return self._fetch(_lib.ARR_U)

    @U.setter
    def U(self, value):
        self._cache.pop(_lib.ARR_U, None)
        self._h.set_array(_lib.ARR_U, np.asarray(value, dtype=self.dtype).reshape(self.cri.shpX))

    @property
    def X(self):
        if self.k == 0 and not self._h_has_x():
            return None
        return self._fetch(_lib.ARR_X)

    @property
    def Xf(self):
        if self.k == 0 and not self._h_has_x():
            return None
        return self._fetch(_lib.ARR_XF)

    def _h_has_x(self):
        return self._h.admm_scalars()[1] > 0

    @property
    def Df(self):
        return self._fetch(_lib.ARR_DF)

    @property
    def Sf(self):
        return self._fetch(_lib.ARR_SF)

    # ---- reference surface -------------------------------------------------------------
    def setdict(self, D=None):
        """Set the dictionary (sporco/admm/cbpdn.py:242-256): its spectrum and the
        per-frequency Gram terms of the x-step are recomputed on the device."""
        if D is not None:
            self.D = np.asarray(D, dtype=self.dtype).reshape(self.cri.shpD)
        self._cache.pop(_lib.ARR_DF, None)
        self._h.set_dict(self.D[:, :, :, 0, :])

    def getcoef(self):
        return self.getmin()

    def reconstruct(self, X=None):
        """irfftn(sum_m Df * rfftn(X)) on the device (sporco/admm/cbpdn.py:373-380)."""
        if X is not None:
            X = np.asarray(X, dtype=self.dtype).reshape(self.cri.shpX)
        return self._h.reconstruct(X)

    def itstat_extra(self):
        return (self.xrrs,)

    def rhochange(self):
        """Nothing to refresh: the device solve forms 1/(g + rho) on the fly."""

    def _admm_config(self):
        o = self.opt
        ar = o['AutoRho']
        return dict(
            lmbda=float(getattr(self, 'lmbda', 0.0)), mu=float(getattr(self, 'mu', 0.0)),
            rlx=float(self.rlx), abs_tol=float(o['AbsStopTol']), rel_tol=float(o['RelStopTol']),
            ar_scaling=float(self.rho_tau), ar_rsdl_ratio=float(self.rho_mu),
            ar_rsdl_target=float(self.rho_xi), ar_enabled=int(bool(ar['Enabled'])),
            ar_period=int(ar['Period']), ar_autoscaling=int(bool(ar['AutoScaling'])),
            ar_std_residuals=int(bool(ar['StdResiduals'])), joint=int(self._joint),
            nonneg=int(bool(o['NonNegCoef'])), no_bndry_cross=int(bool(o['NoBndryCross'])),
            fast_solve=int(bool(o['FastSolve'])), aux_var_obj=int(bool(o['gEvalY'])),
            linsolve_check=int(bool(o['LinSolveCheck'])),
            l2_weight=float(getattr(self, '_l2_weight', 0.0)),
            ams_maps=int(getattr(self, '_ams_maps', 0)))

    def _device_iterate(self, n, want_rows):
        if bool(self.opt['gEvalY']) != (not bool(self.opt['fEvalX'])):
            raise NotImplementedError('fEvalX / gEvalY must be set together through AuxVarObj')
        self._h.admm_configure(**self._admm_config())
        rows, done, stopped = self._h.admm_iterate(n, want_rows)
        self._cache.pop(_lib.ARR_Y, None)
        self._cache.pop(_lib.ARR_U, None)
        self._cache.pop(_lib.ARR_X, None)
        self._cache.pop(_lib.ARR_XF, None)
        rho, _ = self._h.admm_scalars()
        self._rho = common.real_dtype(self.dtype).type(rho)
        if want_rows and done > 0:
            x = rows[done - 1].xslv_relres
            self.xrrs = None if x < 0 else x
        return rows, done, stopped

    def _make_itstat(self, row, t):
        rdt = common.real_dtype(self.dtype).type
        xr = None if row.xslv_relres < 0 else row.xslv_relres
        reg = (row.regl1, row.regl21) if (self._joint or self._two_reg) else (row.regl1,)
        tpl = (int(row.iter), row.objfun, row.dfid) + reg + \
            (rdt(row.primal_rsdl), rdt(row.dual_rsdl), row.eps_primal, row.eps_dual,
             rdt(row.rho), xr, t)
        return type(self).IterationStats(*tpl)

    # ---- multi-GPU ---------------------------------------------------------------------
    def attach_process_group(self, dist, group=None):
        """Join the solvers of all ranks of a ``torch.distributed`` process group into one
        batch-sharded problem: this rank keeps its own images (the `S` it was built with), and
        the squared norms behind r, s, rho and the stopping test -- global over all images in
        the reference (sporco/admm/admm.py:462-486) -- are summed over the ranks once per
        iteration on the device.  Coefficient maps never leave their GPU.  ``torch.distributed``
        only carries the 128-byte NCCL id, the peer-memory handles and the element count."""
        from .. import _dist
        self._world, self._p2p = _dist.attach(self._h, self._device, float(self.Nx), dist, group)

    # ---- pickling: device state travels as host arrays
    def __getstate__(self):
        st = {k: v for k, v in self.__dict__.items() if k not in ('_h', '_cache')}
        st['_saved'] = {'Y': self.Y.copy(), 'U': self.U.copy(), 'k_dev': self._h.admm_scalars()[1]}
        return st

    def __setstate__(self, st):
        saved = st.pop('_saved')
        self.__dict__.update(st)
        self._cache = {}
        self._h = None
        self._open_handle()
        self._after_open()
        self._h.set_array(_lib.ARR_Y, saved['Y'])
        self._h.set_array(_lib.ARR_U, saved['U'])
        self._h.lib.spcsc_admm_set_iter(self._h.h, int(saved['k_dev']))

    def _after_open(self):
        pass

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None:
            h.close()


class ConvBPDN(GenericConvBPDN):
    """ADMM solver for convolutional BPDN (mirror of sporco/admm/cbpdn.py:386-630)::

        argmin_x (1/2) || sum_m d_m * x_m - s ||_2^2 + lambda sum_m || x_m ||_1
    """

    class Options(GenericConvBPDN.Options):
        defaults = copy.deepcopy(GenericConvBPDN.Options.defaults)
        defaults.update({'L1Weight': 1.0})

        def __init__(self, opt=None):
            GenericConvBPDN.Options.__init__(self, {} if opt is None else opt)

    itstat_fields_objfn = ('ObjFun', 'DFid', 'RegL1')
    hdrtxt_objfn = ('Fnc', 'DFid', u'Regℓ1')
    hdrval_objfun = {'Fnc': 'ObjFun', 'DFid': 'DFid', u'Regℓ1': 'RegL1'}

    def __init__(self, D, S, lmbda=None, opt=None, dimK=None, dimN=2, device=0):
        opt = self._coerce_options(opt)
        self.set_dtype(opt, S.dtype)
        rdt = common.real_dtype(self.dtype)
        self.xrrs = None
        super(ConvBPDN, self).__init__(D, S, opt, dimK, dimN, device=device)

        if lmbda is None:                      # sporco/admm/cbpdn.py:573-578
            b = np.conj(self.Df) * self.Sf
            lmbda = 0.1 * abs(b).max()
        self.lmbda = rdt.type(lmbda)
        self.set_attr('rho', opt['rho'], dval=(50.0 * self.lmbda + 1.0), dtype=rdt, reset=True)
        if self.lmbda != 0.0:                  # sporco/admm/cbpdn.py:588-593
            rho_xi = float((1.0 + (18.3) ** (np.log10(self.lmbda) + 1.0)))
        else:
            rho_xi = 1.0
        self.set_attr('rho_xi', opt['AutoRho', 'RsdlTarget'], dval=rho_xi, dtype=rdt,
                      reset=True)
        self.wl1 = np.asarray(opt['L1Weight'], dtype=rdt)
        self.wl1 = self.wl1.reshape(cr.l1Wshape(self.wl1, self.cri))
        self._after_open()
        if opt['Y0'] is not None and opt['U0'] is None:
            # intent of sporco/admm/cbpdn.py:601-610 (the reference itself raises
            # AttributeError on this path because lmbda is not yet set when uinit runs)
            self.U = (self.lmbda / self.rho) * np.sign(self.Y)

    def _after_open(self):
        w = np.ascontiguousarray(self.wl1, dtype=self.dtype)
        if w.ndim != 5:
            raise ValueError('L1Weight does not reduce to a 5-D internal shape')
        self._h.set_l1_weight(w)

    def uinit(self, ushape):
        return np.zeros(ushape, dtype=self.dtype)


class ConvBPDNJoint(ConvBPDN):
    """ADMM solver for convolutional BPDN with an l2,1 joint sparsity term over the channel
    axis (mirror of sporco/admm/cbpdn.py:636-808)::

        argmin_x (1/2) sum_c || sum_m d_m * x_cm - s_c ||_2^2
                 + lambda sum_c sum_m || x_cm ||_1 + mu || {x_cm} ||_2,1
    """

    class Options(ConvBPDN.Options):
        defaults = copy.deepcopy(ConvBPDN.Options.defaults)
        defaults.update({'L21Weight': 1.0})

        def __init__(self, opt=None):
            ConvBPDN.Options.__init__(self, {} if opt is None else opt)

    itstat_fields_objfn = ('ObjFun', 'DFid', 'RegL1', 'RegL21')
    hdrtxt_objfn = ('Fnc', 'DFid', u'Regℓ1', u'Regℓ2,1')
    hdrval_objfun = {'Fnc': 'ObjFun', 'DFid': 'DFid', u'Regℓ1': 'RegL1',
                     u'Regℓ2,1': 'RegL21'}
    _joint = True

    def __init__(self, D, S, lmbda=None, mu=0.0, opt=None, dimK=None, dimN=2, device=0):
        opt = self._coerce_options(opt)
        self.set_dtype(opt, S.dtype)
        self.mu = self.dtype.type(mu)
        self.wl21 = np.asarray(opt['L21Weight'], dtype=self.dtype)
        super(ConvBPDNJoint, self).__init__(D, S, lmbda, opt, dimK=dimK, dimN=dimN,
                                            device=device)

    def _after_open(self):
        super(ConvBPDNJoint, self)._after_open()
        w = np.asarray(self.wl21, dtype=self.dtype)
        K, M = self.cri.K, self.cri.M
        while w.ndim > 2 and w.shape[0] == 1:
            w = w.reshape(w.shape[1:])
        if w.ndim == 0:
            w = w.reshape(1, 1)
        elif w.ndim == 1:
            w = w.reshape(1, -1)
        if w.ndim != 2 or w.shape[0] not in (1, K) or w.shape[1] not in (1, M):
            raise NotImplementedError('L21Weight must broadcast over (K, M) only')
        self._h.set_l21_weight(np.ascontiguousarray(w))


class ConvElasticNet(ConvBPDN):
    """ADMM solver for the convolutional elastic net (mirror of sporco/admm/cbpdn.py:810-990)::

        argmin_x (1/2) || sum_m d_m * x_m - s ||_2^2 + lambda sum_m || x_m ||_1
                 + (mu/2) sum_m || x_m ||_2^2

    Same iteration as :class:`ConvBPDN` with ``mu + rho`` on the diagonal of the x-step system
    (admm/cbpdn.py:948-955) and the extra term in the objective; ``IterationStats`` gains ``RegL2``.
    """

    itstat_fields_objfn = ('ObjFun', 'DFid', 'RegL1', 'RegL2')
    hdrtxt_objfn = ('Fnc', 'DFid', u'Regℓ1', u'Regℓ2')
    hdrval_objfun = {'Fnc': 'ObjFun', 'DFid': 'DFid', u'Regℓ1': 'RegL1', u'Regℓ2': 'RegL2'}
    _two_reg = True

    def __init__(self, D, S, lmbda=None, mu=0.0, opt=None, dimK=None, dimN=2, device=0):
        opt = self._coerce_options(opt)
        self.set_dtype(opt, S.dtype)
        self.mu = self.dtype.type(mu)
        self._l2_weight = float(self.mu)
        super(ConvElasticNet, self).__init__(D, S, lmbda, opt, dimK=dimK, dimN=dimN, device=device)


class ConvBPDNGradReg(ConvBPDN):
    """ADMM solver for convolutional BPDN with an l2 penalty on the gradient of the coefficient
    maps (mirror of sporco/admm/cbpdn.py:993-1216)::

        argmin_x (1/2) || sum_m d_m * x_m - s ||_2^2 + lambda sum_m || x_m ||_1
                 + (mu/2) sum_i sum_m w_m || G_i x_m ||_2^2

    The x-step system has the diagonal ``mu w_m GHG + rho`` (linalg.solvedbd_sm) instead of
    ``rho``; ``IterationStats`` gains ``RegGrad``.  With a multi-channel dictionary the system is the
    rank-C update of that diagonal (``linalg.solvemdbi_ism`` in the reference, one C x C solve per frequency
    here); ``LinSolveCheck`` / ``AuxVarObj`` are then not available.
    """

    class Options(ConvBPDN.Options):
        defaults = copy.deepcopy(ConvBPDN.Options.defaults)
        defaults.update({'GradWeight': 1.0})

        def __init__(self, opt=None):
            ConvBPDN.Options.__init__(self, {} if opt is None else opt)

    itstat_fields_objfn = ('ObjFun', 'DFid', 'RegL1', 'RegGrad')
    hdrtxt_objfn = ('Fnc', 'DFid', u'Regℓ1', u'Regℓ2∇')
    hdrval_objfun = {'Fnc': 'ObjFun', 'DFid': 'DFid', u'Regℓ1': 'RegL1', u'Regℓ2∇': 'RegGrad'}
    _two_reg = True

    def __init__(self, D, S, lmbda=None, mu=0.0, opt=None, dimK=None, dimN=2, device=0):
        opt = self._coerce_options(opt)
        self.cri = cr.CSC_ConvRepIndexing(D, S, dimK=dimK, dimN=dimN)
        if self.cri.Cd != 1 and (opt['LinSolveCheck'] or opt['AuxVarObj']):
            raise NotImplementedError('ConvBPDNGradReg with a multi-channel dictionary: LinSolveCheck and '
                                      'AuxVarObj are not supported')
        self.set_dtype(opt, S.dtype)
        self.mu = self.dtype.type(mu)
        gw = opt['GradWeight']
        wm = np.asarray(gw, dtype=self.dtype)
        if wm.ndim > 1 or (wm.ndim == 1 and wm.size != self.cri.M):
            raise ValueError('GradWeight must be a scalar or an M-vector')
        self._wgrd = np.ascontiguousarray(np.broadcast_to(wm, (self.cri.M,)), dtype=self.dtype)
        self.Wgrd = wm.reshape((1,) * (dimN + 2) + wm.shape) if wm.ndim else wm
        super(ConvBPDNGradReg, self).__init__(D, S, lmbda, opt, dimK=dimK, dimN=dimN, device=device)

    def _after_open(self):
        # sum_i |G_i|^2 of the forward-difference filters, as signal.gradient_filters forms it
        # (signal.py:196-240): transforms of the two 2-tap filters through the device FFT
        N0, N1 = self.cri.Nv
        g0 = np.zeros((N0, N1), dtype=self.dtype)
        g1 = np.zeros((N0, N1), dtype=self.dtype)
        g0[0, 0], g0[1 % N0, 0] = 1, -1
        g1[0, 0], g1[0, 1 % N1] = 1, -1
        gf = _lib.rfft2(np.stack((g0, g1)), device=self._device)
        self._ghg = np.ascontiguousarray(np.sum((np.conj(gf) * gf).real, axis=0), dtype=self.dtype)
        self.GHGf = self.Wgrd * self._ghg.reshape(self._ghg.shape + (1, 1, 1))
        self._h.set_gradreg(self._ghg, self._wgrd)
        super(ConvBPDNGradReg, self)._after_open()


class AddMaskSim(object):
    """Boundary / missing-data masking by additive mask simulation (mirror of
    sporco/admm/cbpdn.py:2287-2485): a wrapper about a :class:`ConvBPDN` (or :class:`ConvElasticNet`)
    object whose dictionary gets an impulse filter appended; the impulse's coefficient map absorbs
    the signal where the mask `W` is zero (one impulse filter per channel for a multi-channel dictionary).

    On the device the hook the reference installs on ``ystep`` / ``obfn_gvar`` is a property of the
    prox kernel: the mask enters as the l1 weight of the impulse map (0 where ``W == 0``: the map
    is ``AX + U``; a huge weight elsewhere: the map is 0), and maps flagged as additive-mask maps are
    neither clipped by ``NonNegCoef`` / ``NoBndryCross`` nor counted in ``RegL1``.
    """

    def __init__(self, cbpdnclass, D, S, W, *args, **kwargs):
        dimK = kwargs.get('dimK', None)
        dimN = kwargs.get('dimN', 2)
        if not (isinstance(cbpdnclass, type) and issubclass(cbpdnclass, ConvBPDN)) or \
                issubclass(cbpdnclass, ConvBPDNJoint):
            raise NotImplementedError('AddMaskSim wraps sporco_b200 ConvBPDN / ConvElasticNet objects')
        self.cri = cr.CSC_ConvRepIndexing(D, S, dimK=dimK, dimN=dimN)
        Cd = self.cri.Cd
        if Cd > 1 and issubclass(cbpdnclass, ConvBPDNGradReg):
            raise NotImplementedError('AddMaskSim about ConvBPDNGradReg with a multi-channel dictionary '
                                      'is not supported')
        # one impulse filter, or one per channel (each non-zero in its own channel) for a multi-channel
        # dictionary (sporco/admm/cbpdn.py:2337-2345)
        if Cd == 1:
            self.imp = np.zeros(D.shape[0:dimN] + (1,))
            self.imp[(0,) * dimN] = 1.0
        else:
            self.imp = np.zeros(D.shape[0:dimN] + (Cd,) * 2)
            for c in range(Cd):
                self.imp[(0,) * dimN + (c, c)] = 1.0
        Di = np.concatenate((D, self.imp.astype(D.dtype)), axis=D.ndim - 1)
        self.cbpdn = cbpdnclass(Di, S, *args, **kwargs)
        self.IterationStats = self.cbpdn.IterationStats
        inner = self.cbpdn
        self.W = np.asarray(np.asarray(W).reshape(cr.mskWshape(np.asarray(W), self.cri)),
                            dtype=inner.dtype)
        # a mask with a channel axis applies channel c to the map of impulse c (:2361-2362)
        if Cd > 1 and self.W.shape[self.cri.dimN] > 1:
            self.W = np.swapaxes(self.W, self.cri.axisC, self.cri.axisM)
        # positions of the impulse maps that are forced to zero: exactly the reference's
        # ``Yi[np.where(self.W.astype(bool))] = 0.0`` on Yi of shape (N0, N1, Cx, K, Cd)
        icri = inner.cri
        kdim = icri.K if self.W.shape[icri.axisK] > 1 else 1
        cdim = icri.shpX[icri.axisC] if self.W.shape[icri.axisC] > 1 else 1
        on = np.zeros(icri.Nv + (cdim, kdim, Cd), dtype=bool)
        # (index arrays, not broadcasting: a mask without a channel axis reaches the first impulse map only --
        # the reference's behaviour, kept)
        on[np.where(self.W.astype(bool))] = True
        # combined l1 weight: the inner object's own weight on the primary maps ...
        w1 = inner.wl1
        shp = tuple(max(a, b) for a, b in zip(w1.shape[:4], on.shape[:4])) + (icri.M,)
        wfull = np.empty(shp, dtype=inner.dtype)
        wfull[...] = np.broadcast_to(w1, shp)
        # ... and the mask on the impulse maps
        huge = inner.dtype.type(1e30 if inner.dtype == np.float32 else 1e300)
        for c in range(Cd):
            wfull[..., icri.M - Cd + c] = np.where(np.broadcast_to(on[..., c], shp[:4]), huge,
                                                  inner.dtype.type(0))
        inner._ams_maps = Cd
        inner.wl1 = wfull
        inner._after_open()
        self.timer = inner.timer
        self.itstat = inner.itstat

    def solve(self):
        """Solve with the inner object and strip the additive-mask map from the result."""
        Xi = self.cbpdn.solve()
        self.timer = self.cbpdn.timer
        self.itstat = self.cbpdn.itstat
        return Xi[self.index_primary()]

    def setdict(self, D=None):
        Di = np.concatenate((D, self.imp.astype(D.dtype)), axis=D.ndim - 1)
        self.cbpdn.setdict(Di)

    def getcoef(self):
        return self.cbpdn.getcoef()[self.index_primary()]

    def index_primary(self):
        return np.s_[..., 0:-self.cri.Cd]

    def index_addmsk(self):
        return np.s_[..., -self.cri.Cd:]

    def reconstruct(self, X=None):
        """Reconstruction from the primary maps only (admm/cbpdn.py:2461-2475)."""
        inner = self.cbpdn
        if X is None:
            X = inner.Y[self.index_primary()]
        X = np.asarray(X, dtype=inner.dtype)
        Xi = np.concatenate((X, np.zeros(X.shape[:-1] + (self.cri.Cd,), dtype=inner.dtype)), axis=X.ndim - 1)
        return inner.reconstruct(Xi)

    def getitstat(self):
        return self.cbpdn.getitstat()

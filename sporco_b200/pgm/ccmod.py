"""Constrained convolutional MOD (dictionary update) by PGM on the B200 engine.

Counterpart of ``sporco.pgm.ccmod.ConvCnstrMOD`` (sporco/pgm/ccmod.py:28-404): same constructor
and ``Options`` tree, ``IterationStats`` fields (DFid, Cnstr, Rsdl, F_Btrack, Q_Btrack,
IterBTrack, L) and the ``setcoef / getdict / solve / reconstruct`` surface.  The gradient, the
inverse transform, the constraint projection, the forward transform, the momentum step and all
sums run on the device (``spcsc_ccmod_*``); the host keeps the step size and the momentum
sequence, as the reference does.  When the object is built by
:class:`sporco_b200.dictlrn.cbpdndl.ConvBPDNDictLearn` it shares the X step's handle, so
coefficient maps and dictionary pass between the two steps without leaving the GPU.

Supported: greyscale and multi-channel signals with a single-channel or a multi-channel
dictionary, single- or multi-scale filter supports, step 1/L fixed or found by ``BacktrackStandard``, Nesterov (or
linear) momentum.  ``BacktrackRobust``, ``Monotone`` and ``StepSizePolicy`` raise ``NotImplementedError``.
"""

import copy

import numpy as np

from .. import _lib, cnvrep as cr
from . import pgm


class ConvCnstrMOD(pgm.PGMDFT):
    class Options(pgm.PGMDFT.Options):
        defaults = copy.deepcopy(pgm.PGMDFT.Options.defaults)
        defaults.update({'ZeroMean': False})

        def __init__(self, opt=None):
            pgm.PGMDFT.Options.__init__(self, {} if opt is None else opt)

    itstat_fields_objfn = ('DFid', 'Cnstr')
    hdrtxt_objfn = ('DFid', 'Cnstr')
    hdrval_objfun = {'DFid': 'DFid', 'Cnstr': 'Cnstr'}

    def __init__(self, Z, S, dsz, opt=None, dimK=1, dimN=2, device=0, handle=None):
        if dimN != 2:
            raise NotImplementedError('sporco_b200 implements the dimN=2 (image) case only')
        opt = self._coerce_options(opt)
        self.cri = cr.CDU_ConvRepIndexing(dsz, S, dimK=dimK, dimN=dimN)
        cri = self.cri
        from .backtrack import BacktrackStandard
        if opt['Backtrack'] is not None and not (isinstance(opt['Backtrack'], BacktrackStandard) or
                                                  type(opt['Backtrack']).__name__ == 'BacktrackStandard'):
            raise NotImplementedError('only BacktrackStandard is implemented for the device dictionary update')
        if opt['Monotone'] or opt['StepSizePolicy'] is not None:
            raise NotImplementedError('Monotone / StepSizePolicy are not implemented for the device '
                                      'dictionary update')
        super(ConvCnstrMOD, self).__init__(cri.shpD, cri.Nv, cri.axisN, S.dtype, opt)
        if self.backtrack is not None and not isinstance(self.backtrack, BacktrackStandard):
            # a parameter holder of the same name built by the reference package
            self.backtrack = BacktrackStandard(gamma_u=self.backtrack.gamma_u, maxiter=self.backtrack.maxiter)
        # NB the reference passes dval = 14 K here (pgm/ccmod.py:218) but PGM.__init__ has already
        # set L = 1 (pgm/pgm.py:242) and set_attr keeps a value that is set: the effective
        # default is 1, which is what this class reproduces.
        self.set_attr('L', opt['L'], dval=cri.K * 14.0, dtype=self.dtype)
        self.S = np.asarray(S.reshape(cri.shpS), dtype=self.dtype)
        self.dsz = cri.dsz
        self._cache = {}
        self._stats = None
        self._owns_handle = handle is None
        if handle is None:
            _lib.require_device()
            handle = _lib.Handle(cri.Nv[0], cri.Nv[1], cri.C, cri.Cd, cri.K, cri.M,
                                 cri.dsz[0], cri.dsz[1], self.dtype, device)
            handle.set_signal(self.S[..., 0])
        self._h = handle
        # multi-scale dictionary: every filter is projected over its own support (cnvrep.py:277-360)
        self._h.ccmod_set_supports(cr.filter_supports(dsz) if cri.multiscale else None)
        x0 = opt['X0']
        if x0 is None:
            d0 = np.zeros((cri.dsz[0], cri.dsz[1], cri.Cd, cri.M), dtype=self.dtype)
        else:
            x0 = np.asarray(x0, dtype=self.dtype).reshape(cri.shpD)
            d0 = np.ascontiguousarray(x0[0:cri.dsz[0], 0:cri.dsz[1], :, 0, :])
        self._h.ccmod_reset(d0, opt['ZeroMean'])
        if Z is not None:
            self.setcoef(Z)

    # ---- reference surface
    def setcoef(self, Z):
        """Set the coefficient maps (pgm/ccmod.py:264-281); a host array is sent to the GPU and
        transformed there."""
        Z = np.asarray(Z, dtype=self.dtype).reshape(self.cri.shpX)
        self._h.ccmod_setcoef(Z)

    def setcoef_from_xstep(self, source=_lib.COEF_ADMM_Y):
        """The X step on the same handle supplies its current iterate (device to device)."""
        self._h.ccmod_setcoef_device(source)

    def getdict(self, crop=True):
        """Current dictionary in the internal layout, cropped to the filter support
        (hd, wd, Cd, 1, M) or zero-padded (N0, N1, Cd, 1, M), as in pgm/ccmod.py:283-291."""
        if 'D' not in self._cache:
            self._cache['D'] = self._h.ccmod_get_dict()
        d = self._cache['D']                                  # (hd, wd, Cd, M)
        cri = self.cri
        if crop:
            return d.reshape(cri.dsz[0], cri.dsz[1], cri.Cd, 1, cri.M)
        return cr.zpad(d.reshape(cri.dsz[0], cri.dsz[1], cri.Cd, 1, cri.M), cri.Nv)

    @property
    def X(self):
        return self.getdict(crop=False)

    @property
    def Xf(self):
        return self._h.ccmod_get_spectrum(0)

    @property
    def Yf(self):
        return self._h.ccmod_get_spectrum(1)

    def getmin(self):
        return self.getdict(crop=False)

    def reconstruct(self, D=None):
        raise NotImplementedError('use ConvBPDNDictLearn.reconstruct or ConvBPDN.reconstruct')

    # ---- one PGM iteration: PGMDFT.xstep + ystep of pgm/pgm.py:779-831 in one device call
    def _trial(self):
        """One proximal trial at the current L (backtracking only): F = f(X) and the quadratic model
        Q = f(Y) + <grad f(Y), X - Y> + (L / 2) ||X - Y||^2 in the DFT scaling (pgm/backtrack.py:88-97)."""
        if self.backtrack is None:
            return None
        f, fy, lin, dxy2 = self._h.ccmod_trial(float(self.L))
        self._tried = True
        rdt = self.dtype.type
        return rdt(f), rdt(fy) + rdt(lin) + (self.L / 2.) * rdt(dxy2)

    def ystep(self):
        tprv = self.t
        self.t = self.momentum.update(self.var_momentum())
        # FastSolve: no iteration record is built (pgm/pgm.py:347), so the two statistics that
        # need a pass of their own are skipped
        flags = 0 if self.opt['FastSolve'] else 3
        if self.backtrack is not None and getattr(self, '_tried', False):
            self._stats = self._h.ccmod_accept((tprv - 1.) / self.t, flags)
            self._tried = False
        else:
            self._stats = self._h.ccmod_step(float(self.L), (tprv - 1.) / self.t, flags)
        self._cache.clear()

    def rsdl(self):
        return self._stats[2]

    def eval_objfn(self):
        return (self._stats[0], self._stats[1])

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and getattr(self, '_owns_handle', False):
            h.close()

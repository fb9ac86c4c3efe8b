"""Constrained convolutional MOD (dictionary update) by consensus ADMM on the B200 engine.

Counterpart of ``sporco.admm.ccmod.ConvCnstrMOD_Consensus`` (sporco/admm/ccmod.py:613-911 over
``ADMMConsensus``, sporco/admm/admm.py:1419-1707): same constructor and ``Options`` tree,
``IterationStats`` fields (DFid, Cnstr, PrimalRsdl, DualRsdl, EpsPrimal, EpsDual, Rho, XSlvRelRes)
and the ``setcoef / getdict / solve`` surface.  One block per (image, coefficient channel): the
x step of a block is the Sherman-Morrison solve of the ConvBPDN x step with the block's coefficient
spectra in the role of the dictionary, so it runs on the same column kernel
(``spcsc_ccmod_cns_step``); the y step needs only the filter supports of the block mean (cropping is
linear), which is also all that is exchanged when the images are sharded over GPUs.  The host keeps
the scalar arithmetic of ``compute_residuals`` / ``update_rho`` (admm.py:462-486, 549-575), as the
reference does.

Supported: greyscale and multi-channel signals with a single-channel dictionary (channels become
further blocks, ccmod.py:697-705) or a dictionary with the signal's channels; the objective on the
consensus variable (``AuxVarObj`` True, the class default) or on the block variables (``AuxVarObj`` False; one GPU),
``LinSolveCheck``.  The ``ism`` and ``cg`` solvers
are not provided (the factory functions ``ConvCnstrMOD`` / ``ConvCnstrMODOptions`` accept ``method='cns'``,
their default in the reference).
"""

import copy

import numpy as np

from .. import _lib, common, cnvrep as cr
from . import admm


class ConvCnstrMOD_Consensus(admm.ADMM):
    class Options(admm.ADMM.Options):
        """Keys and defaults of ``ConvCnstrMOD_Consensus.Options`` (ccmod.py:632-655: the
        ``ConvCnstrMODBase`` defaults overridden by the ``ADMMConsensus`` ones, RelaxParam 1.8)."""

        defaults = copy.deepcopy(admm.ADMM.Options.defaults)
        defaults.update({'fEvalX': False, 'gEvalY': True, 'AuxVarObj': True, 'ReturnX': False,
                         'RelaxParam': 1.8, 'ZeroMean': False, 'LinSolveCheck': False})

        def __init__(self, opt=None):
            admm.ADMM.Options.__init__(self, {} if opt is None else opt)
            if self['AutoRho', 'RsdlTarget'] is None:
                self['AutoRho', 'RsdlTarget'] = 1.0

        def __setitem__(self, key, value):
            admm.ADMM.Options.__setitem__(self, key, value)
            if key == 'AuxVarObj':
                if value is True:
                    self['fEvalX'] = False
                    self['gEvalY'] = True
                else:
                    self['fEvalX'] = True
                    self['gEvalY'] = False

    itstat_fields_objfn = ('DFid', 'Cnstr')
    itstat_fields_extra = ('XSlvRelRes',)
    hdrtxt_objfn = ('DFid', 'Cnstr')
    hdrval_objfun = {'DFid': 'DFid', 'Cnstr': 'Cnstr'}

    def __init__(self, Z, S, dsz, opt=None, dimK=1, dimN=2, device=0, handle=None):
        if dimN != 2:
            raise NotImplementedError('sporco_b200 implements the dimN=2 (image) case only')
        if not (np.isrealobj(S) and (Z is None or np.isrealobj(Z))):
            raise NotImplementedError('complex-valued data is not supported')
        opt = self._coerce_options(opt)
        if bool(opt['fEvalX']) == bool(opt['gEvalY']):
            raise NotImplementedError('fEvalX / gEvalY must be set together through AuxVarObj')
        self.cri = cr.CDU_ConvRepIndexing(dsz, S, dimK=dimK, dimN=dimN)
        cri = self.cri
        # a single-channel dictionary with a multi-channel signal: the channels are further blocks
        self.Nb = cri.K if cri.C == cri.Cd else cri.C * cri.K
        yshape = tuple(cri.shpD)
        self.yshape = yshape
        self.xshape = yshape + (self.Nb,)
        Nx = self.Nb * int(np.prod(yshape))
        super(ConvCnstrMOD_Consensus, self).__init__(Nx, yshape, self.xshape, S.dtype, opt)
        # NB the reference means the number of images as default penalty parameter (ccmod.py:691-692) but
        # ADMM.__init__ has already set rho = 1 (admm.py:247) and set_attr keeps a value that is set: the
        # effective default is 1, which the base class here has reproduced
        self.S = np.asarray(S.reshape(cri.shpS), dtype=self.dtype)
        self.dsz = cri.dsz
        self._cache = {}
        self._stats = None
        self._device = device
        self._udiv = 1.0
        self._owns_handle = handle is None
        if handle is None:
            _lib.require_device()
            handle = _lib.Handle(cri.Nv[0], cri.Nv[1], cri.C, cri.Cd, cri.K, cri.M,
                                 cri.dsz[0], cri.dsz[1], self.dtype, device)
            handle.set_signal(self.S[..., 0])
        self._h = handle
        # multi-scale dictionary: every filter is projected over its own support (cnvrep.py:277-360)
        self._h.ccmod_set_supports(cr.filter_supports(dsz) if cri.multiscale else None)
        y0 = opt['Y0']
        if y0 is None:
            d0 = np.zeros((cri.dsz[0], cri.dsz[1], cri.Cd, cri.M), dtype=self.dtype)
        else:
            y0 = np.asarray(y0, dtype=self.dtype).reshape(cri.shpD)
            d0 = np.ascontiguousarray(y0[0:cri.dsz[0], 0:cri.dsz[1], :, 0, :])
            if np.any(y0[cri.dsz[0]:]) or np.any(y0[:, cri.dsz[1]:]):
                raise NotImplementedError('Y0 must vanish outside the filter support')
        if opt['U0'] is not None:
            raise NotImplementedError('U0 is not supported; the duals start from Y0 / rho (or 0)')
        self._d0, self._y0_given = d0, y0 is not None
        self._h.ccmod_reset(d0, opt['ZeroMean'])
        self._h.ccmod_cns_init(float(self.rho), self._y0_given)
        if Z is not None:
            self.setcoef(Z)

    # ---- reference surface
    def setcoef(self, Z):
        """Set the coefficient maps (ccmod.py:753-769); a host array goes to the GPU and is
        transformed there."""
        cri = self.cri
        Z = np.asarray(Z, dtype=self.dtype).reshape(cri.shpX)
        self._h.ccmod_setcoef(Z)

    def setcoef_from_xstep(self, source=_lib.COEF_ADMM_Y):
        """The X step on the same handle supplies its current iterate (device to device)."""
        self._h.ccmod_setcoef_device(source)

    def getdict(self, crop=True):
        """The consensus variable Y, cropped to the filter support or zero-padded (ccmod.py:849-857)."""
        if 'D' not in self._cache:
            self._cache['D'] = self._h.ccmod_get_dict()
        d = self._cache['D']                                  # (hd, wd, Cd, M)
        cri = self.cri
        d = d.reshape(cri.dsz[0], cri.dsz[1], cri.Cd, 1, cri.M)
        return d if crop else cr.zpad(d, cri.Nv)

    @property
    def Y(self):
        return self.getdict(crop=False)

    def var_y(self):
        return self.Y

    def _blocks(self, which):
        """Block variables in the reference's layout (N0, N1, Cd, 1, M, Nb) (ccmod.py:728-735)."""
        cri = self.cri
        a = self._h.ccmod_cns_get(which, cri.K * cri.C, cri.M)            # (K*C, M, N0, N1), batch = (k, c)
        nbl = (cri.K * cri.C) // cri.Cd
        a = a.reshape(nbl, cri.Cd, cri.M, cri.Nv[0], cri.Nv[1])
        return np.ascontiguousarray(a.transpose(3, 4, 1, 2, 0))[:, :, :, np.newaxis, :, :]

    @property
    def X(self):
        return self._blocks(0)

    @property
    def U(self):
        u = self._blocks(1)
        return u if self._udiv == 1.0 else (u / self.dtype.type(self._udiv)).astype(self.dtype)

    def getmin(self):
        return self.Y

    def reconstruct(self, D=None):
        raise NotImplementedError('use ConvBPDNDictLearn.reconstruct or ConvBPDN.reconstruct')

    # ---- iterations: array work on the device, scalar control flow here (admm.py:331-377)
    def _device_iterate(self, n, want_rows):
        opt = self.opt
        rdt = common.real_dtype(self.dtype).type
        ar = opt['AutoRho']
        need_rsdl = ar['Enabled'] or not opt['FastSolve']
        flags = (0 if opt['FastSolve'] else 3) | (4 if opt['LinSolveCheck'] else 0) | (8 if opt['fEvalX'] else 0)
        rows, done, stopped = [], 0, False
        for _ in range(n):
            k = self.k + done
            st = self._h.ccmod_cns_step(float(self.rho), float(self._udiv), float(self.rlx), flags)
            self._stats = st
            self._udiv = 1.0
            self._cache.clear()
            done += 1
            if not need_rsdl:
                continue
            nb = float(self.Nb)
            nX, nR, nU = (rdt(np.sqrt(st[i])) for i in (2, 3, 4))
            nY, nS = rdt(np.sqrt(st[5])), rdt(np.sqrt(st[6]))
            rho = self.rho
            rn = max(nX, rdt(np.sqrt(nb)) * nY)             # rsdl_rn (admm.py:1693-1700)
            sn = rho * nU                                    # rsdl_sn (:1704-1707)
            r = nR                                           # ||X - Y||           (:1673-1676)
            s = rdt(np.sqrt(nb)) * rho * nS                  # sqrt(Nb) rho ||Yprev - Y||  (:1680-1689)
            if ar['StdResiduals']:
                epri = np.sqrt(self.Nc) * opt['AbsStopTol'] + rn * opt['RelStopTol']
                edua = np.sqrt(self.Nx) * opt['AbsStopTol'] + sn * opt['RelStopTol']
            else:
                if rn == 0.0:
                    rn = 1.0
                if sn == 0.0:
                    sn = 1.0
                r = r / rn
                s = s / sn
                epri = np.sqrt(self.Nc) * opt['AbsStopTol'] / rn + opt['RelStopTol']
                edua = np.sqrt(self.Nx) * opt['AbsStopTol'] / sn + opt['RelStopTol']
            if want_rows:
                rows.append((k, st[0], st[1], r, s, epri, edua, rho, st[7] if st[7] >= 0.0 else None))
            # update_rho (admm.py:549-575); U /= rsf is applied by the next reads of U on the device
            if ar['Enabled'] and k != 0 and np.mod(k + 1, ar['Period']) == 0:
                tau, mu, xi = self.rho_tau, self.rho_mu, self.rho_xi
                if ar['AutoScaling']:
                    if s == 0.0 or r == 0.0:
                        rhomlt = tau
                    else:
                        rhomlt = np.sqrt(r / (s * xi) if r > s * xi else (s * xi) / r)
                        if rhomlt > tau:
                            rhomlt = tau
                else:
                    rhomlt = tau
                rsf = 1.0
                if r > xi * mu * s:
                    rsf = rhomlt
                elif s > (mu / xi) * r:
                    rsf = 1.0 / rhomlt
                self.rho = self.rho * rdt(rsf)
                self._udiv = float(rsf)
            if r < epri and s < edua:
                stopped = True
                break
        return rows, done, stopped

    def _make_itstat(self, row, t):
        k, dfd, cns, r, s, epri, edua, rho, xrrs = row
        return type(self).IterationStats(int(k), dfd, cns, r, s, epri, edua, rho, xrrs, t)

    def attach_process_group(self, dist, group=None):
        """Shard the blocks (images) over the ranks of a ``torch.distributed`` group: every rank owns
        the block variables of its images; the filter supports of the block mean, the residual norms
        and the data-fidelity value are summed over the ranks on the device."""
        from .. import _dist
        import torch
        if self.k != 0:
            raise RuntimeError('attach_process_group must precede the first iteration')
        if self._owns_handle:        # a shared handle is attached by the X step that owns it
            _dist.attach(self._h, self._device, self.Nb * int(np.prod(self.yshape)), dist, group)
        t = torch.tensor([float(self.Nb)], dtype=torch.float64, device=torch.device('cuda', self._device))
        dist.all_reduce(t, group=group)
        self.Nb = int(round(t.item()))
        self.Nx = self.Nb * int(np.prod(self.yshape))
        self.Nc = self.Nx
        self._h.ccmod_cns_init(float(self.rho), self._y0_given, self.Nb)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and getattr(self, '_owns_handle', False):
            h.close()


def ConvCnstrMODOptions(opt=None, method='cns'):
    """Options object of the selected dictionary update (sporco/admm/ccmod.py:970-1001)."""
    if method != 'cns':
        raise NotImplementedError("dictionary update method '%s' is not provided; use 'cns' (or pgm.ccmod)" % method)
    return ConvCnstrMOD_Consensus.Options(opt)


def ConvCnstrMOD(*args, **kwargs):
    """Dictionary update object of the selected class (sporco/admm/ccmod.py:914-966; default 'cns')."""
    method = kwargs.pop('method', 'cns')
    if method != 'cns':
        raise NotImplementedError("dictionary update method '%s' is not provided; use 'cns' (or pgm.ccmod)" % method)
    return ConvCnstrMOD_Consensus(*args, **kwargs)

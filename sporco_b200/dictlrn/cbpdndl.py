"""Convolutional dictionary learning on the B200 engine.

Counterpart of ``sporco.dictlrn.cbpdndl.ConvBPDNDictLearn`` (sporco/dictlrn/cbpdndl.py:231-524):
alternation of a convolutional sparse coding step (``xmethod`` 'admm' or 'pgm': the ConvBPDN
solvers of this package) and a dictionary update (``dmethod`` 'pgm':
:class:`sporco_b200.pgm.ccmod.ConvCnstrMOD`, or 'cns': the consensus ADMM update
:class:`sporco_b200.admm.ccmod.ConvCnstrMOD_Consensus`), same constructor, ``Options`` tree
(``CBPDN`` / ``CCMOD`` sub-trees, ``DictSize``, ``AccurateDFid``) and ``IterationStats`` fields.
Both steps share one engine handle: the coefficient maps go from the X step to the D step, and
the dictionary spectrum back, as device arrays; per outer iteration the host sees the two
records of scalars only.  The ADMM dictionary updates ``dmethod`` 'ism' and 'cg' are not provided.
"""

import copy

import numpy as np

from .. import _lib, cdict, cnvrep as cr
from ..admm import cbpdn as admm_cbpdn
from ..admm import ccmod as admm_ccmod
from ..pgm import cbpdn as pgm_cbpdn
from ..pgm import ccmod as pgm_ccmod
from . import common as dc
from . import dictlrn


def cbpdn_class_label_lookup(label):
    clsmod = {'admm': admm_cbpdn.ConvBPDN, 'pgm': pgm_cbpdn.ConvBPDN}
    if label in clsmod:
        return clsmod[label]
    raise ValueError('Unknown ConvBPDN solver method %s' % label)


def ccmod_class_label_lookup(label):
    if label == 'pgm':
        return pgm_ccmod.ConvCnstrMOD
    if label == 'cns':
        return admm_ccmod.ConvCnstrMOD_Consensus
    if label in ('ism', 'cg'):
        raise NotImplementedError("dictionary update method '%s' is not provided; use 'pgm'"
                                  % label)
    raise ValueError('Unknown ConvCnstrMOD solver method %s' % label)


_X_OVERRIDES = {
    'admm': {'MaxMainIter': 1, 'AutoRho': {'Period': 10, 'AutoScaling': False,
                                           'RsdlRatio': 10.0, 'Scaling': 2.0,
                                           'RsdlTarget': 1.0}},
    'pgm': {'MaxMainIter': 1},
}


def ConvBPDNOptionsDefaults(method='admm'):
    """Defaults of the X step inside dictionary learning (dictlrn/cbpdndl.py:43-56)."""
    dflt = copy.deepcopy(cbpdn_class_label_lookup(method).Options.defaults)
    for k, v in _X_OVERRIDES[method].items():
        if isinstance(v, dict):
            dflt[k].update(v)
        else:
            dflt[k] = v
    return dflt


def ConvBPDNOptions(opt=None, method='admm'):
    """An ``Options`` object of the selected X-step class with the dictionary-learning
    defaults applied (dictlrn/cbpdndl.py:60-84)."""
    o = cbpdn_class_label_lookup(method).Options(copy.deepcopy(_X_OVERRIDES[method]))
    if opt is not None:
        o.update(cdict._plain(opt))
    return o


def ConvBPDN(*args, **kwargs):
    """X-step object of the selected class (dictlrn/cbpdndl.py:88-121)."""
    method = kwargs.pop('method', 'admm')
    return cbpdn_class_label_lookup(method)(*args, **kwargs)


_D_OVERRIDES = {
    'pgm': {'MaxMainIter': 1},
    'cns': {'MaxMainIter': 1, 'AutoRho': {'Period': 10, 'AutoScaling': False, 'RsdlRatio': 10.0,
                                          'Scaling': 2.0, 'RsdlTarget': 1.0}},
}


def ConvCnstrMODOptionsDefaults(method='pgm'):
    """Defaults of the D step inside dictionary learning (dictlrn/cbpdndl.py:139-152)."""
    dflt = copy.deepcopy(ccmod_class_label_lookup(method).Options.defaults)
    for k, v in _D_OVERRIDES[method].items():
        if isinstance(v, dict):
            dflt[k].update(v)
        else:
            dflt[k] = v
    return dflt


def ConvCnstrMODOptions(opt=None, method='pgm'):
    o = ccmod_class_label_lookup(method).Options(copy.deepcopy(_D_OVERRIDES[method]))
    if opt is not None:
        o.update(cdict._plain(opt))
    return o


def ConvCnstrMOD(*args, **kwargs):
    method = kwargs.pop('method', 'pgm')
    return ccmod_class_label_lookup(method)(*args, **kwargs)


class ConvBPDNDictLearn(dictlrn.DictLearn):
    class Options(dictlrn.DictLearn.Options):
        defaults = copy.deepcopy(dictlrn.DictLearn.Options.defaults)
        defaults.update({'DictSize': None, 'AccurateDFid': False})

        def __init__(self, opt=None, xmethod=None, dmethod=None):
            self.xmethod = 'admm' if xmethod is None else xmethod
            self.dmethod = 'pgm' if dmethod is None else dmethod
            tree = copy.deepcopy(type(self).defaults)
            tree.update({'CBPDN': ConvBPDNOptionsDefaults(self.xmethod),
                         'CCMOD': ConvCnstrMODOptionsDefaults(self.dmethod)})
            top = {k: v for k, v in tree.items() if k not in ('CBPDN', 'CCMOD')}
            top['CBPDN'] = ConvBPDNOptions(None, method=self.xmethod)
            top['CCMOD'] = ConvCnstrMODOptions(None, method=self.dmethod)
            dict.__init__(self)
            self.pth = ()
            self.dflt = tree
            for k, v in top.items():
                dict.__setitem__(self, k, v)
            self.update({} if opt is None else opt)

        def __reduce__(self):
            return (_rebuild_options, (cdict._plain(self), self.xmethod, self.dmethod))

    def __init__(self, D0, S, lmbda=None, opt=None, xmethod=None, dmethod=None, dimK=1, dimN=2,
                 device=0):
        if opt is None:
            opt = ConvBPDNDictLearn.Options(xmethod=xmethod, dmethod=dmethod)
        elif not isinstance(opt, ConvBPDNDictLearn.Options):
            opt = ConvBPDNDictLearn.Options(
                cdict._plain(opt), xmethod=getattr(opt, 'xmethod', xmethod),
                dmethod=getattr(opt, 'dmethod', dmethod))
        if xmethod is None:
            xmethod = opt.xmethod
        if dmethod is None:
            dmethod = opt.dmethod
        if opt.xmethod != xmethod or opt.dmethod != dmethod:
            raise ValueError('Parameters xmethod and dmethod must have the same values used to '
                             'initialise the Options object')
        self.opt = opt
        self._dist = None
        self.xmethod, self.dmethod = xmethod, dmethod
        dsz = D0.shape if opt['DictSize'] is None else opt['DictSize']
        cri = cr.CDU_ConvRepIndexing(dsz, S, dimK, dimN)
        # normalised initial dictionary, also the first iterate of the D step  (cbpdndl.py:448-454)
        D0 = cr.Pcn(np.asarray(D0), dsz, cri.Nv, dimN, cri.dimCd, crp=True,
                    zm=opt['CCMOD', 'ZeroMean'])
        optname = 'X0' if dmethod == 'pgm' else 'Y0'
        opt['CCMOD'].update({optname: cr.zpad(cr.stdformD(D0, cri.Cd, cri.M, dimN), cri.Nv)})
        xstep = ConvBPDN(D0, S, lmbda, opt['CBPDN'], method=xmethod, dimK=dimK, dimN=dimN,
                         device=device)
        dstep = ConvCnstrMOD(None, S, dsz, opt['CCMOD'], method=dmethod, dimK=dimK, dimN=dimN,
                             device=device, handle=xstep._h)
        if dstep.dtype != xstep.dtype:
            raise ValueError('X step and D step must use the same data type')
        self._coef_source = _lib.COEF_ADMM_Y if xmethod == 'admm' else _lib.COEF_PGM_X
        isc = dictlrn.IterStatsConfig(
            isfld=dc.isfld(xmethod, dmethod, opt), isxmap=dc.isxmap(xmethod, opt),
            isdmap=dc.isdmap(dmethod), evlmap=dc.evlmap(opt['AccurateDFid']),
            hdrtxt=dc.hdrtxt(xmethod, dmethod, opt), hdrmap=dc.hdrmap(xmethod, dmethod, opt),
            fmtmap={'It_X': '%4d', 'It_D': '%4d'})
        super(ConvBPDNDictLearn, self).__init__(xstep, dstep, opt, isc)

    # ---- both steps run without copying their minimiser to the host ...
    def run_xstep(self):
        self.xstep.run()

    def run_dstep(self):
        self.dstep.run()

    # ---- ... and the two hand-overs stay on the device (dictlrn/dictlrn.py:379-389)
    def post_xstep(self):
        self.dstep.setcoef_from_xstep(self._coef_source)

    def post_dstep(self):
        xs = self.xstep
        xs._h.ccmod_push_dict()
        xs._cache.pop(_lib.ARR_DF, None)
        self._xdict_stale = True

    def solve(self):
        d = super(ConvBPDNDictLearn, self).solve()
        if getattr(self, '_xdict_stale', False):        # host copy of the X step's dictionary
            self.xstep.D = np.asarray(self.getdict(crop=True), dtype=self.xstep.dtype).reshape(
                self.xstep.cri.shpD)
            self._xdict_stale = False
        return d

    def getdict(self, crop=True):
        return self.dstep.getdict(crop=crop)

    def reconstruct(self, D=None, X=None):
        """sum_m d_m * x_m for the current (or the given) dictionary and coefficient maps."""
        xs = self.xstep
        if D is not None:
            D = np.asarray(D, dtype=xs.dtype)
            D = cr.bcrop(D, self.dstep.dsz).reshape(xs.cri.shpD)
            xs._h.set_dict(np.ascontiguousarray(D[:, :, :, 0, :]))
        if X is None:
            X = self.getcoef()
        try:
            return xs.reconstruct(X)
        finally:
            if D is not None:
                xs._h.ccmod_push_dict()

    def evaluate(self):
        """Functional value for the pair (new dictionary, current coefficient maps) when option
        ``AccurateDFid`` is set (dictlrn/cbpdndl.py:502-524); the data fidelity is the one the
        D step has just computed on the device."""
        if not self.opt['AccurateDFid']:
            return None
        X = self.xstep.getcoef() if self.xmethod == 'pgm' else self.xstep.var_y()
        dfd = self.dstep._stats[0]
        rl1 = float(np.sum(np.abs(X), dtype=np.float64))
        if self._dist is not None:       # sharded images: DFid is already global, RegL1 is not
            import torch
            t = torch.tensor([rl1], dtype=torch.float64, device=torch.device('cuda', self.xstep._device))
            self._dist[0].all_reduce(t, group=self._dist[1])
            rl1 = float(t.item())
        return dict(DFid=dfd, RegL1=rl1, ObjFun=dfd + self.xstep.lmbda * rl1)

    def attach_process_group(self, dist, group=None):
        """Shard the training images over the ranks of a ``torch.distributed`` group: every rank
        codes its own images (ADMM or PGM X step); of the dictionary gradient only the filter supports of
        the gradient step are summed over ranks (crop before reduce: h w Cd M values), together with the
        data-fidelity value, on the device over peer memory (NCCL where peers cannot be mapped), so all
        ranks hold the same dictionary."""
        self.xstep.attach_process_group(dist, group)
        if self.dmethod == 'cns':       # the block mean of the consensus update runs over all ranks' images
            self.dstep.attach_process_group(dist, group)
        self._dist = (dist, group)


def _rebuild_options(content, xmethod, dmethod):
    return ConvBPDNDictLearn.Options(content, xmethod=xmethod, dmethod=dmethod)

"""Host-side frame of the PGM / FISTA solvers (mirror of sporco/pgm/pgm.py:37-560, 712-894).

The scalar control flow of the reference -- step size L, momentum sequence t, the F <= Q
backtracking test, the stopping test -- stays on the host, as in the reference; every array
operation (gradient in the DFT domain, inverse transform, proximal step, forward transform,
the sums entering F, Q and the residual) is a batch of CUDA kernels behind
``spcsc_pgm_trial`` / ``spcsc_pgm_accept``."""

import copy

from .. import cdict, common, util
from .momentum import MomentumNesterov


class PGM(common.IterativeSolver):
    class Options(cdict.ConstrainedDict):
        """Keys and defaults of ``sporco.pgm.pgm.PGM.Options`` (pgm.py:157-166)."""

        defaults = {'FastSolve': False, 'Verbose': False, 'StatusHeader': True,
                    'DataType': None, 'X0': None, 'Callback': None, 'MaxMainIter': 1000,
                    'IterTimer': 'solve', 'RelStopTol': 1e-3, 'L': None,
                    'AutoStop': {'Enabled': False, 'Tau0': 1e-2}, 'Monotone': False,
                    'Momentum': MomentumNesterov(), 'StepSizePolicy': None, 'Backtrack': None}

        def __init__(self, opt=None):
            cdict.ConstrainedDict.__init__(self, {} if opt is None else opt)

    fwiter = 4
    fpothr = 2
    itstat_fields_objfn = ('ObjFun', 'FVal', 'GVal')
    itstat_fields_alg = ('Rsdl', 'F_Btrack', 'Q_Btrack', 'IterBTrack', 'L')
    itstat_fields_extra = ()
    hdrtxt_objfn = ('Fnc', 'f', 'g')
    hdrval_objfun = {'Fnc': 'ObjFun', 'f': 'FVal', 'g': 'GVal'}

    def __new__(cls, *args, **kwargs):
        obj = super(PGM, cls).__new__(cls)
        obj.timer = util.Timer(['init', 'solve', 'solve_wo_func', 'solve_wo_rsdl',
                                'solve_wo_btrack'])
        obj.timer.start('init')
        return obj

    @classmethod
    def _coerce_options(cls, opt):
        if opt is None:
            return cls.Options()
        if isinstance(opt, PGM.Options):
            return opt
        if isinstance(opt, dict):
            return cls.Options({k: (dict(v) if isinstance(v, dict) else v)
                                for k, v in dict.items(opt)})
        raise TypeError('Parameter opt must be an instance of PGM.Options')

    def __init__(self, xshape, dtype, opt=None):
        self.opt = self._coerce_options(opt)
        self.set_dtype(self.opt, dtype)
        self.set_attr('L', self.opt['L'], dval=1.0, dtype=self.dtype)
        o = self.opt
        # step-size policy is switched off when backtracking is enabled (sporco/pgm/pgm.py:236-239)
        self.stepsizepolicy = o['StepSizePolicy']
        if o['Backtrack'] is not None:
            self.stepsizepolicy = None
        self.momentum = o['Momentum']
        if o['AutoStop', 'Enabled']:
            self.tau0 = o['AutoStop', 'Tau0']
        self.F = None
        self.Q = None
        self.iterBTrack = None
        self.backtrack = o['Backtrack']
        self.itstat = []
        self.k = 0
        self.t = 1

    # device-backed subclass provides _trial(), ystep(), _stats()
    def solve(self):
        """Outer loop of sporco/pgm/pgm.py:284-383."""
        self.run()
        return self.getmin()

    def run(self):
        """:meth:`solve` without the final host copy of the minimiser."""
        fmtstr, nsep = self.display_start()
        self.timer.start(['solve', 'solve_wo_func', 'solve_wo_rsdl', 'solve_wo_btrack'])
        for self.k in range(self.k, self.k + self.opt['MaxMainIter']):
            self.on_iteration_start()
            if self.backtrack is not None:
                self.timer.stop('solve_wo_btrack')
                self.backtrack.update(self)
                self.timer.start('solve_wo_btrack')
            else:
                self.xstep()
                self.ystep()
            if not self.opt['FastSolve']:
                frcxd = self.rsdl()
                tol = self.opt['RelStopTol']
                if self.opt['AutoStop', 'Enabled']:
                    tol = self.tau0 / (1. + self.k)
                itst = self.iteration_stats(self.k, frcxd)
                self.itstat.append(itst)
                self.display_status(fmtstr, itst)
            if self.opt['Callback'] is not None:
                if self.opt['Callback'](self):
                    break
            if not self.opt['FastSolve']:
                if frcxd < tol:
                    break
        self.k += 1
        self.timer.stop(['solve', 'solve_wo_func', 'solve_wo_rsdl', 'solve_wo_btrack'])
        self.display_end(nsep)

    def getmin(self):
        return self.X

    def on_iteration_start(self):
        """Hook at the top of an iteration (sporco/pgm/pgm.py:835-846); the copies of the previous
        iterates the reference makes here are implicit on the device."""

    def xstep(self):
        """One proximal step at the current L."""
        self._trial()

    def var_momentum(self):
        # sporco/pgm/pgm.py:662-668: the Nesterov rule takes t, the linear rules the iteration count.
        # Decided by the rule's class name so that a rule object built by the reference package
        # (sporco.pgm.momentum.MomentumNesterov) is recognised as well.
        nesterov = isinstance(self.momentum, MomentumNesterov) or any(
            c.__name__ == 'MomentumNesterov' for c in type(self.momentum).__mro__)
        return self.t if nesterov else self.k

    def iteration_stats(self, k, frcxd):
        tk = self.timer.elapsed(self.opt['IterTimer'])
        tpl = (k,) + self.eval_objfn() + (frcxd, self.F, self.Q, self.iterBTrack, self.L) + \
            self.itstat_extra() + (tk,)
        return type(self).IterationStats(*tpl)

    def itstat_extra(self):
        return ()

    def getitstat(self):
        return common.transpose_ntpl_list(self.itstat)

    @classmethod
    def hdrtxt(cls):
        return ('Itn',) + cls.hdrtxt_objfn + ('Rsdl', 'F', 'Q', 'It_Bt', 'L')

    @classmethod
    def hdrval(cls):
        hdr = {'Itn': 'Iter'}
        hdr.update(cls.hdrval_objfun)
        hdr.update({'Rsdl': 'Rsdl', 'F': 'F_Btrack', 'Q': 'Q_Btrack', 'It_Bt': 'IterBTrack',
                    'L': 'L'})
        return hdr

    def display_start(self):
        if not self.opt['Verbose']:
            return '', 0
        hdr = type(self).hdrtxt()
        if self.opt['Backtrack'] is None:
            hdr = hdr[:-4]
        hdrstr, fmtstr, nsep = common.solve_status_str(hdr, fmtmap={'It_Bt': '%5d'},
                                                       fwdth0=type(self).fwiter,
                                                       fprec=type(self).fpothr)
        if self.opt['StatusHeader']:
            print(hdrstr)
            print('-' * nsep)
        return fmtstr, nsep

    def display_status(self, fmtstr, itst):
        if self.opt['Verbose']:
            hv = type(self).hdrval()
            vals = tuple(getattr(itst, hv[c]) for c in type(self).hdrtxt())
            if self.opt['Backtrack'] is None:
                vals = vals[:-4]
            print(fmtstr % vals)

    def display_end(self, nsep):
        if self.opt['Verbose'] and self.opt['StatusHeader']:
            print('-' * nsep)


class PGMDFT(PGM):
    """PGM with updates in the DFT domain (sporco/pgm/pgm.py:712-894)."""

    class Options(PGM.Options):
        defaults = copy.deepcopy(PGM.Options.defaults)

        def __init__(self, opt=None):
            PGM.Options.__init__(self, {} if opt is None else opt)

    def __init__(self, xshape, Nv, axisN, dtype, opt=None):
        super(PGMDFT, self).__init__(xshape, dtype, opt)
        self.Nv = Nv
        self.axisN = axisN

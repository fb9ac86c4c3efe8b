"""Host-side frame of the ADMM solvers.

Mirror of the parts of ``sporco.admm.admm`` that surround the ConvBPDN iteration: the
``Options`` trees (sporco/admm/admm.py:72-173, 814-846), timers, ``IterationStats``
bookkeeping, the ``Verbose`` table (sporco/admm/admm.py:579-625) and the outer ``solve``
loop (sporco/admm/admm.py:293-389).  The iteration body itself -- xstep, relaxation, ystep,
ustep, residuals, rho update -- does not run here: a derived class forwards batches of
iterations to libspcsc, which executes them as CUDA kernels and hands back one statistics
row per iteration.
"""

import copy

import numpy as np

from .. import cdict, common, util


class ADMM(common.IterativeSolver):
    """Outer loop and bookkeeping of an ADMM solver whose iterations run on the device."""

    class Options(cdict.ConstrainedDict):
        """ADMM options; keys and defaults as ``sporco.admm.admm.ADMM.Options``."""

        defaults = {'FastSolve': False, 'Verbose': False, 'StatusHeader': True,
                    'DataType': None, 'MaxMainIter': 1000, 'IterTimer': 'solve',
                    'AbsStopTol': 0.0, 'RelStopTol': 1e-3, 'RelaxParam': 1.0, 'rho': None,
                    'AutoRho': {'Enabled': False, 'Period': 10, 'Scaling': 2.0,
                                'RsdlRatio': 10.0, 'RsdlTarget': None, 'AutoScaling': False,
                                'StdResiduals': False},
                    'Y0': None, 'U0': None, 'Callback': None}

        def __init__(self, opt=None):
            cdict.ConstrainedDict.__init__(self, {} if opt is None else opt)

    fwiter = 4
    fpothr = 2
    itstat_fields_objfn = ('ObjFun', 'FVal', 'GVal')
    itstat_fields_alg = ('PrimalRsdl', 'DualRsdl', 'EpsPrimal', 'EpsDual', 'Rho')
    itstat_fields_extra = ()
    hdrtxt_objfn = ('Fnc', 'f', 'g')
    hdrval_objfun = {'Fnc': 'ObjFun', 'f': 'FVal', 'g': 'GVal'}

    #: iterations submitted to the device per host round trip when nothing (Verbose,
    #: Callback) needs the host after every iteration
    batch_iters = 50

    def __new__(cls, *args, **kwargs):
        obj = super(ADMM, cls).__new__(cls)
        obj.timer = util.Timer(['init', 'solve', 'solve_wo_func', 'solve_wo_rsdl'])
        obj.timer.start('init')
        return obj

    @classmethod
    def _coerce_options(cls, opt):
        """Accept this class's Options, or any dict carrying the same key tree (e.g. an
        Options object built by the reference package)."""
        if opt is None:
            return cls.Options()
        if isinstance(opt, ADMM.Options):
            return opt
        if isinstance(opt, dict):
            return cls.Options(cdict._plain(opt))
        raise TypeError('Parameter opt must be an instance of ADMM.Options')

    def __init__(self, Nx, yshape, ushape, dtype, opt=None):
        self.opt = self._coerce_options(opt)
        self.Nx = Nx
        self.Nc = int(np.prod(np.array(ushape)))
        self.set_dtype(self.opt, dtype)
        rdt = common.real_dtype(self.dtype)
        self.set_attr('rho', self.opt['rho'], dval=1.0, dtype=rdt)
        self.set_attr('rho_tau', self.opt['AutoRho', 'Scaling'], dval=2.0, dtype=rdt)
        self.set_attr('rho_mu', self.opt['AutoRho', 'RsdlRatio'], dval=10.0, dtype=rdt)
        self.set_attr('rho_xi', self.opt['AutoRho', 'RsdlTarget'], dval=1.0, dtype=rdt)
        self.set_attr('rlx', self.opt['RelaxParam'], dval=1.0, dtype=rdt)
        self.itstat = []
        self.k = 0

    # ---- to be provided by the device-backed subclass
    def _device_iterate(self, n, want_rows):
        """Run up to `n` iterations; return (rows, n_done, stopped)."""
        raise NotImplementedError()

    def _make_itstat(self, row, t):
        raise NotImplementedError()

    # ---- outer loop (sporco/admm/admm.py:293-389)
    def solve(self):
        self.run()
        return self.getmin()

    def run(self):
        """The body of :meth:`solve` without the final host copy of the minimiser: callers
        that keep the result on the device (dictionary learning) use this."""
        fmtstr, nsep = self.display_start()
        self.timer.start(['solve', 'solve_wo_func', 'solve_wo_rsdl'])
        per_iter = self.opt['Verbose'] or self.opt['Callback'] is not None
        want_rows = not self.opt['FastSolve']
        remaining = int(self.opt['MaxMainIter'])
        stop = False
        while remaining > 0 and not stop:
            n = 1 if per_iter else min(self.batch_iters, remaining)
            t0 = self.timer.elapsed(self.opt['IterTimer'])
            rows, done, stopped = self._device_iterate(n, want_rows)
            t1 = self.timer.elapsed(self.opt['IterTimer'])
            if want_rows:
                for i in range(done):
                    itst = self._make_itstat(rows[i], t0 + (t1 - t0) * (i + 1) / max(done, 1))
                    self.itstat.append(itst)
                    self.display_status(fmtstr, itst)
            self.k += done
            remaining -= done
            if self.opt['Callback'] is not None and done > 0:
                self.k -= 1                       # the reference's loop variable during the call
                try:
                    if self.opt['Callback'](self):
                        stop = True
                finally:
                    self.k += 1
            if stopped or done == 0:
                stop = True
        self.timer.stop(['solve', 'solve_wo_func', 'solve_wo_rsdl'])
        self.display_end(nsep)

    def getmin(self):
        return self.X

    def getitstat(self):
        return common.transpose_ntpl_list(self.itstat)

    # ---- status table
    @classmethod
    def hdrtxt(cls):
        return ('Itn',) + cls.hdrtxt_objfn + ('r', 's', u'ρ')

    @classmethod
    def hdrval(cls):
        m = {'Itn': 'Iter'}
        m.update(cls.hdrval_objfun)
        m.update({'r': 'PrimalRsdl', 's': 'DualRsdl', u'ρ': 'Rho'})
        return m

    def display_start(self):
        if not self.opt['Verbose']:
            return '', 0
        hdr = type(self).hdrtxt()
        if not self.opt['AutoRho', 'Enabled']:
            hdr = hdr[:-1]
        hdrstr, fmtstr, nsep = common.solve_status_str(hdr, fwdth0=type(self).fwiter,
                                                       fprec=type(self).fpothr)
        if self.opt['StatusHeader']:
            print(hdrstr)
            print('-' * nsep)
        return fmtstr, nsep

    def display_status(self, fmtstr, itst):
        if self.opt['Verbose']:
            hv = type(self).hdrval()
            vals = tuple(getattr(itst, hv[c]) for c in type(self).hdrtxt())
            if not self.opt['AutoRho', 'Enabled']:
                vals = vals[:-1]
            print(fmtstr % vals)

    def display_end(self, nsep):
        if self.opt['Verbose'] and self.opt['StatusHeader']:
            print('-' * nsep)

    def var_x(self):
        return self.X

    def var_y(self):
        return self.Y


class ADMMEqual(ADMM):
    """ADMM with the constraint x = y (sporco/admm/admm.py:791-983)."""

    class Options(ADMM.Options):
        defaults = copy.deepcopy(ADMM.Options.defaults)
        defaults.update({'fEvalX': True, 'gEvalY': True, 'ReturnX': True})

        def __init__(self, opt=None):
            ADMM.Options.__init__(self, {} if opt is None else opt)

    def __init__(self, xshape, dtype, opt=None):
        Nx = int(np.prod(np.array(xshape)))
        super(ADMMEqual, self).__init__(Nx, xshape, xshape, dtype, opt)

    def getmin(self):
        return self.X if self.opt['ReturnX'] else self.Y

    def obfn_fvar(self):
        return self.X if self.opt['fEvalX'] else self.Y

    def obfn_gvar(self):
        return self.Y if self.opt['gEvalY'] else self.X

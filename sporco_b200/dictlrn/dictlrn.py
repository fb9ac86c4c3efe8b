"""Alternation frame of dictionary learning (mirror of sporco/dictlrn/dictlrn.py:29-418)."""

import collections

from .. import cdict, common, util


class IterStatsConfig(object):
    """How the per-iteration record of the learner is assembled from the records of the X step
    and the D step, and how it is displayed (dictlrn/dictlrn.py:29-159)."""

    fwiter = 4
    fpothr = 2

    def __init__(self, isfld, isxmap, isdmap, evlmap, hdrtxt, hdrmap, fmtmap=None):
        self.IterationStats = collections.namedtuple('IterationStats', isfld)
        self.isxmap, self.isdmap, self.evlmap = isxmap, isdmap, evlmap
        self.hdrtxt, self.hdrmap = hdrtxt, hdrmap
        self.hdrstr, self.fmtstr, self.nsep = common.solve_status_str(
            hdrtxt, fmtmap=fmtmap, fwdth0=type(self).fwiter, fprec=type(self).fpothr)

    def iterstats(self, j, t, isx, isd, evl):
        vals = []
        for f in self.IterationStats._fields:
            if f in self.isxmap:
                vals.append(getattr(isx, self.isxmap[f]))
            elif f in self.isdmap:
                vals.append(getattr(isd, self.isdmap[f]))
            elif f in self.evlmap:
                vals.append(evl[f])
            elif f == 'Iter':
                vals.append(j)
            elif f == 'Time':
                vals.append(t)
            else:
                vals.append(None)
        return self.IterationStats._make(vals)

    def printheader(self):
        print(self.hdrstr)
        self.printseparator()

    def printseparator(self):
        print('-' * self.nsep)

    def printiterstats(self, itst):
        print(self.fmtstr % tuple(getattr(itst, self.hdrmap[c]) for c in self.hdrtxt))


class DictLearn(object):
    """Alternate ``xstep.solve()`` and ``dstep.solve()`` (dictlrn/dictlrn.py:187-418)."""

    class Options(cdict.ConstrainedDict):
        defaults = {'Verbose': False, 'StatusHeader': True, 'IterTimer': 'solve',
                    'MaxMainIter': 1000, 'Callback': None}

        def __init__(self, opt=None):
            cdict.ConstrainedDict.__init__(self, {} if opt is None else opt)

    def __new__(cls, *args, **kwargs):
        obj = super(DictLearn, cls).__new__(cls)
        obj.timer = util.Timer(['init', 'solve', 'solve_wo_eval'])
        obj.timer.start('init')
        return obj

    def __init__(self, xstep, dstep, opt=None, isc=None):
        self.opt = DictLearn.Options() if opt is None else opt
        if isc is None:
            isc = IterStatsConfig(
                isfld=['Iter', 'ObjFunX', 'XPrRsdl', 'XDlRsdl', 'XRho', 'ObjFunD', 'DPrRsdl',
                       'DDlRsdl', 'DRho', 'Time'],
                isxmap={'ObjFunX': 'ObjFun', 'XPrRsdl': 'PrimalRsdl', 'XDlRsdl': 'DualRsdl',
                        'XRho': 'Rho'},
                isdmap={'ObjFunD': 'DFid', 'DPrRsdl': 'PrimalRsdl', 'DDlRsdl': 'DualRsdl',
                        'DRho': 'Rho'},
                evlmap={},
                hdrtxt=['Itn', 'FncX', 'r_X', 's_X', u'ρ_X', 'FncD', 'r_D', 's_D', u'ρ_D'],
                hdrmap={'Itn': 'Iter', 'FncX': 'ObjFunX', 'r_X': 'XPrRsdl', 's_X': 'XDlRsdl',
                        u'ρ_X': 'XRho', 'FncD': 'ObjFunD', 'r_D': 'DPrRsdl', 's_D': 'DDlRsdl',
                        u'ρ_D': 'DRho'})
        self.isc = isc
        self.xstep, self.dstep = xstep, dstep
        self.itstat = []
        self.j = 0
        self.timer.stop('init')

    def solve(self):
        if self.opt['Verbose'] and self.opt['StatusHeader']:
            self.isc.printheader()
        self.timer.start(['solve', 'solve_wo_eval'])
        for self.j in range(self.j, self.j + self.opt['MaxMainIter']):
            self.run_xstep()
            self.post_xstep()
            self.run_dstep()
            self.post_dstep()
            self.timer.stop('solve_wo_eval')
            evl = self.evaluate()
            self.timer.start('solve_wo_eval')
            t = self.timer.elapsed(self.opt['IterTimer'])
            xs, ds = self.xstep, self.dstep
            xit = xs.itstat[-1] if xs.itstat else \
                xs.IterationStats(*([0.0] * len(xs.IterationStats._fields)))
            dit = ds.itstat[-1] if ds.itstat else \
                ds.IterationStats(*([0.0] * len(ds.IterationStats._fields)))
            itst = self.isc.iterstats(self.j, t, xit, dit, evl)
            self.itstat.append(itst)
            if self.opt['Verbose']:
                self.isc.printiterstats(itst)
            if self.opt['Callback'] is not None:
                if self.opt['Callback'](self):
                    break
        self.j += 1
        self.timer.stop(['solve', 'solve_wo_eval'])
        if self.opt['Verbose'] and self.opt['StatusHeader']:
            self.isc.printseparator()
        return self.getdict()

    def run_xstep(self):
        self.xstep.solve()

    def run_dstep(self):
        self.dstep.solve()

    def post_xstep(self):
        self.dstep.setcoef(self.xstep.getcoef())

    def post_dstep(self):
        self.xstep.setdict(self.dstep.getdict())

    def evaluate(self):
        return None

    def getdict(self):
        return self.dstep.getdict()

    def getcoef(self):
        return self.xstep.getcoef()

    def getitstat(self):
        return common.transpose_ntpl_list(self.itstat)

"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python oracle/make_golden.py

Imports bwohlberg/sporco read-only from /root/reference (with the stub modules in
oracle/shims for its missing optional imports), runs its ConvBPDN / ConvBPDNJoint / PGM
solvers on small seeded problems, asserts that the numpy restatement in
oracle/cbpdn_oracle.py reproduces every output bit for bit, and stores inputs + outputs as
fixtures.  /root/reference does not exist on the GPU box; the fixtures travel instead.
"""

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
sys.path[:0] = [os.path.join(HERE, 'shims'), REF, ROOT]
warnings.filterwarnings('ignore')

from sporco.admm import cbpdn as rcbpdn          # noqa: E402
from sporco.pgm import cbpdn as rpgm             # noqa: E402
from sporco.pgm.backtrack import BacktrackStandard, BacktrackRobust  # noqa: E402
from sporco.pgm.stepsize import StepSizePolicyBB, StepSizePolicyCauchy  # noqa: E402
from sporco import linalg as rlinalg, prox as rprox, fft as rfft   # noqa: E402
from oracle import cbpdn_oracle as orc           # noqa: E402
from oracle import cbpdndl_oracle as orcdl       # noqa: E402
from sporco.dictlrn import cbpdndl as rcbpdndl   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def same(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        raise AssertionError('oracle differs from the reference: %s' % what)


def stat(its, name):
    return np.asarray(getattr(its, name), dtype=np.float64)


def admm_case(tag, dt, D, S, lmbda, opt, dimK=None, mu=None, enet_mu=None, grad_mu=None):
    if grad_mu is not None:
        b = rcbpdn.ConvBPDNGradReg(D, S, lmbda, grad_mu, rcbpdn.ConvBPDNGradReg.Options(opt), dimK=dimK)
    elif enet_mu is not None:
        b = rcbpdn.ConvElasticNet(D, S, lmbda, enet_mu, rcbpdn.ConvBPDN.Options(opt), dimK=dimK)
    elif mu is None:
        b = rcbpdn.ConvBPDN(D, S, lmbda, rcbpdn.ConvBPDN.Options(opt), dimK=dimK)
    else:
        b = rcbpdn.ConvBPDNJoint(D, S, lmbda, mu, rcbpdn.ConvBPDNJoint.Options(opt), dimK=dimK)
    b.solve()
    r = orc.admm_convbpdn(D, S, lmbda, mu=mu, opt=opt, dimK=dimK, enet_mu=enet_mu, grad_mu=grad_mu)
    same(b.Y, r.Y, tag + ' Y')
    same(b.U, r.U, tag + ' U')
    same(b.X, r.X, tag + ' X')
    its = b.getitstat()
    col = {'ObjFun': 1, 'DFid': 2, 'RegL1': 3}
    off = 1 if (mu is not None or enet_mu is not None or grad_mu is not None) else 0
    col.update({'PrimalRsdl': 4 + off, 'DualRsdl': 5 + off, 'Rho': 8 + off})
    for name, c in col.items():
        same(stat(its, name), np.array([row[c] for row in r.itstat], dtype=np.float64),
             tag + ' ' + name)
    rec = b.reconstruct()
    out = dict(D=D, S=S, lmbda=np.float64(lmbda), Y=b.Y, U=b.U, X=b.X, recon=rec,
               Rho=stat(its, 'Rho'), ObjFun=stat(its, 'ObjFun'), DFid=stat(its, 'DFid'),
               RegL1=stat(its, 'RegL1'), PrimalRsdl=stat(its, 'PrimalRsdl'),
               DualRsdl=stat(its, 'DualRsdl'))
    if mu is not None:
        out['mu'] = np.float64(mu)
        out['RegL21'] = stat(its, 'RegL21')
    if grad_mu is not None:
        out['mu'] = np.float64(grad_mu)
        out['RegGrad'] = stat(its, 'RegGrad')
        same(out['RegGrad'], np.array([row[4] for row in r.itstat], dtype=np.float64), tag + ' RegGrad')
    if enet_mu is not None:
        out['mu'] = np.float64(enet_mu)
        out['RegL2'] = stat(its, 'RegL2')
        same(out['RegL2'], np.array([row[4] for row in r.itstat], dtype=np.float64), tag + ' RegL2')
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), **out)
    print('wrote', tag, 'Y nnz', int(np.count_nonzero(b.Y)), 'final rho', float(b.rho))


AMS_OPT = {'ams_gry': {'MaxMainIter': 30, 'RelStopTol': 0.0},
           'ams_k3': {'MaxMainIter': 20, 'RelStopTol': 0.0, 'NonNegCoef': True, 'NoBndryCross': True,
                      'AuxVarObj': True}}


GRD_OPT = {'grd_k3': lambda dt: {'MaxMainIter': 30, 'RelStopTol': 0.0},
           'grd_aux': lambda dt: {'MaxMainIter': 20, 'RelStopTol': 0.0, 'AuxVarObj': True,
                                  'LinSolveCheck': True,
                                  'GradWeight': np.linspace(0.2, 2.0, 6).astype(dt)}}


def ams_case(tag, dt, D, S, W, lmbda, opt, dimK=None, grad_mu=None):
    """AddMaskSim about ConvBPDN or ConvBPDNGradReg (admm/cbpdn.py:2287-2485)."""
    if grad_mu is None:
        b = rcbpdn.AddMaskSim(rcbpdn.ConvBPDN, D, S, W, lmbda, rcbpdn.ConvBPDN.Options(opt), dimK=dimK)
    else:
        b = rcbpdn.AddMaskSim(rcbpdn.ConvBPDNGradReg, D, S, W, lmbda, grad_mu,
                              rcbpdn.ConvBPDNGradReg.Options(opt), dimK=dimK)
    X = b.solve()
    r = orc.admm_addmasksim(D, S, W, lmbda, opt=opt, dimK=dimK, grad_mu=grad_mu)
    same(b.cbpdn.Y, r.Y, tag + ' Y')
    same(b.cbpdn.X, r.X, tag + ' X')
    its = b.getitstat()
    o_ = 0 if grad_mu is None else 1
    for name, c in (('ObjFun', 1), ('DFid', 2), ('RegL1', 3), ('PrimalRsdl', 4 + o_), ('DualRsdl', 5 + o_),
                    ('Rho', 8 + o_)):
        same(stat(its, name), np.array([row[c] for row in r.itstat], dtype=np.float64), tag + ' ' + name)
    out = dict(D=D, S=S, W=W, lmbda=np.float64(lmbda), Y=b.cbpdn.Y, Xprimary=X, recon=b.reconstruct(),
               Rho=stat(its, 'Rho'), ObjFun=stat(its, 'ObjFun'), DFid=stat(its, 'DFid'),
               RegL1=stat(its, 'RegL1'), PrimalRsdl=stat(its, 'PrimalRsdl'), DualRsdl=stat(its, 'DualRsdl'))
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), **out)
    print('wrote', tag)


def pgm_case(tag, dt, D, S, lmbda, opt_ref, opt_orc, dimK=None):
    b = rpgm.ConvBPDN(D, S, lmbda, rpgm.ConvBPDN.Options(opt_ref), dimK=dimK)
    b.solve()
    r = orc.pgm_convbpdn(D, S, lmbda, opt=opt_orc, dimK=dimK)
    same(b.X, r.X, tag + ' X')
    its = b.getitstat()
    same(stat(its, 'L'), np.array([row[8] for row in r.itstat], dtype=np.float64), tag + ' L')
    same(stat(its, 'Rsdl'), np.array([row[4] for row in r.itstat], dtype=np.float64), tag + ' Rsdl')
    out = dict(D=D, S=S, lmbda=np.float64(lmbda), X=b.X, L=stat(its, 'L'),
               Rsdl=stat(its, 'Rsdl'), ObjFun=stat(its, 'ObjFun'), DFid=stat(its, 'DFid'),
               RegL1=stat(its, 'RegL1'))
    if opt_ref.get('Backtrack') is not None:
        out['IterBTrack'] = stat(its, 'IterBTrack')
        out['F_Btrack'] = stat(its, 'F_Btrack')
        out['Q_Btrack'] = stat(its, 'Q_Btrack')
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), **out)
    print('wrote', tag)


CDL_FIELDS = ('ObjFun', 'DFid', 'RegL1', 'Cnstr', 'XPrRsdl', 'XDlRsdl', 'XRho', 'D_L', 'D_Rsdl')


CDL_OPT = {'MaxMainIter': 25, 'CBPDN': {'rho': 5.0, 'AutoRho': {'Period': 4}}, 'CCMOD': {'L': 40.0}}
CDL_OPT_ZM = {'MaxMainIter': 20, 'CBPDN': {'NonNegCoef': True},
              'CCMOD': {'L': 60.0, 'ZeroMean': True}}


CDL_OPT_CLR = {'MaxMainIter': 15, 'CBPDN': {'rho': 5.0, 'AutoRho': {'Period': 4}},
               'CCMOD': {'L': 60.0, 'ZeroMean': True}}
MS_DSZ = ((4, 4, 3), (7, 6, 2))
CDL_OPT_MS = {'MaxMainIter': 15, 'DictSize': MS_DSZ, 'CBPDN': {'rho': 5.0}, 'CCMOD': {'L': 50.0, 'ZeroMean': True}}
CDL_OPT_MS_CNS = {'MaxMainIter': 15, 'DictSize': MS_DSZ, 'CBPDN': {'rho': 5.0}, 'CCMOD': {'rho': 2.0, 'ZeroMean': True}}
CDL_OPT_CNS = {'MaxMainIter': 15, 'CBPDN': {'rho': 5.0}, 'CCMOD': {'rho': 2.0, 'ZeroMean': True}}
CDL_OPT_PGMX = {'MaxMainIter': 20, 'CBPDN': {'L': 80.0}, 'CCMOD': {'L': 40.0}}


def cdl_case(tag, dt, D0, S, lmbda, opt):
    """ConvBPDNDictLearn with its default solvers (ADMM X step, PGM D step)."""
    b = rcbpdndl.ConvBPDNDictLearn(D0, S, lmbda, rcbpdndl.ConvBPDNDictLearn.Options(
        opt, xmethod='admm', dmethod='pgm'), xmethod='admm', dmethod='pgm', dimK=1)
    D1 = b.solve()
    r = orcdl.cbpdndl(D0, S, lmbda, opt)
    same(D1.squeeze(), r['D'], tag + ' D')            # the reference returns (hd, wd, 1, 1, M)
    same(b.getcoef(), r['X'], tag + ' X')
    its = b.getitstat()
    for name in CDL_FIELDS:
        same(stat(its, name), r[name], tag + ' ' + name)
    out = dict(D0=D0, S=S, lmbda=np.float64(lmbda), D=D1, X=b.getcoef())
    out.update({name: stat(its, name) for name in CDL_FIELDS})
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), **out)
    print('wrote', tag)


def cdl_ref_case(tag, dt, D0, S, lmbda, opt, xmethod, dmethod='pgm'):
    """Variants the numpy oracle does not restate as a whole (PGM X step, AccurateDFid, consensus D step -- whose
    D step alone is pinned by cns_case): the fixture holds the reference's own outputs."""
    b = rcbpdndl.ConvBPDNDictLearn(D0, S, lmbda, rcbpdndl.ConvBPDNDictLearn.Options(
        opt, xmethod=xmethod, dmethod=dmethod), xmethod=xmethod, dmethod=dmethod, dimK=1)
    D1 = b.solve()
    its = b.getitstat()
    out = dict(D0=D0, S=S, lmbda=np.float64(lmbda), D=D1, X=b.getcoef())
    for name in its._fields:
        if name not in ('Iter', 'Time') and getattr(its, name)[0] is not None:
            out[name] = stat(its, name)
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), **out)
    print('wrote', tag, sorted(out))


def cns_case(tag, dt, Z, S, dsz, opt):
    """admm.ccmod.ConvCnstrMOD_Consensus with given coefficient maps (ccmod.py:613-911)."""
    from sporco.admm import ccmod as rccmod
    c = rccmod.ConvCnstrMOD_Consensus(Z, S, dsz, rccmod.ConvCnstrMOD_Consensus.Options(dict(opt, Verbose=False)))
    c.solve()
    r = orcdl.ConsensusCCMOD(S, dsz, opt)
    r.setcoef(Z)
    r.solve()
    same(c.Y, r.Y, tag + ' Y')
    its = c.getitstat()
    ref = np.array(r.itstat, dtype=np.float64)
    cols = ('DFid', 'Cnstr', 'PrimalRsdl', 'DualRsdl', 'EpsPrimal', 'EpsDual', 'Rho')
    for i, name in enumerate(cols):
        same(stat(its, name), ref[:, i + 1], tag + ' ' + name)
    out = dict(Z=Z, S=S, dsz=np.array(dsz), Y=c.Y)
    out.update({name: stat(its, name) for name in cols})
    for k, v in opt.items():
        if k == 'Y0':
            out['Y0'] = v
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), **out)
    print('wrote', tag)


def ccmod_bt_case(tag, dt, Z, S, dsz, opt_ref):
    """pgm.ccmod.ConvCnstrMOD with BacktrackStandard (pgm/backtrack.py:74-107): the fixture holds the reference's outputs."""
    from sporco.pgm import ccmod as rpccmod
    c = rpccmod.ConvCnstrMOD(Z, S, dsz, rpccmod.ConvCnstrMOD.Options(dict(opt_ref, Verbose=False)))
    c.solve()
    its = c.getitstat()
    out = dict(Z=Z, S=S, dsz=np.array(dsz), D=c.getdict())
    for name in ('DFid', 'Cnstr', 'Rsdl', 'F_Btrack', 'Q_Btrack', 'IterBTrack', 'L'):
        out[name] = stat(its, name)
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), **out)
    print('wrote', tag)


def tikhonov():
    """sporco.signal.tikhonov_filter on a few shapes (padded sizes: power of two, composite, odd)."""
    from sporco import signal as rsignal
    from oracle import signal_oracle as sorc
    rng = np.random.default_rng(11)
    out = {}
    for dt, sfx in ((np.float64, 'f64'), (np.float32, 'f32')):
        for i, (shape, lm, npd) in enumerate((((32, 32), 5.0, 16), ((40, 36, 3), 2.0, 16),
                                              ((31, 33, 2, 2), 10.0, 8))):
            s = rng.standard_normal(shape).astype(dt)
            sl, sh = rsignal.tikhonov_filter(s, lm, npd)
            ol, oh = sorc.tikhonov_filter(s, lm, npd)
            same(sl, ol, 'tikhonov sl')
            same(sh, oh, 'tikhonov sh')
            out['s%d_%s' % (i, sfx)] = s
            out['sl%d_%s' % (i, sfx)] = sl
            out['sh%d_%s' % (i, sfx)] = sh
    np.savez_compressed(os.path.join(OUT, 'tikhonov.npz'), **out)
    print('wrote tikhonov')


def pgm_mask_case(tag, dt, D, S, W, lmbda, opt_ref, opt_orc, dimK=None):
    """pgm.cbpdn.ConvBPDNMask (pgm/cbpdn.py:387-508)."""
    b = rpgm.ConvBPDNMask(D, S, lmbda, W, rpgm.ConvBPDNMask.Options(opt_ref), dimK=dimK)
    b.solve()
    r = orc.pgm_convbpdn(D, S, lmbda, opt=opt_orc, dimK=dimK, W=W)
    same(b.X, r.X, tag + ' X')
    its = b.getitstat()
    same(stat(its, 'L'), np.array([row[8] for row in r.itstat], dtype=np.float64), tag + ' L')
    same(stat(its, 'ObjFun'), np.array([row[1] for row in r.itstat], dtype=np.float64), tag + ' ObjFun')
    out = dict(D=D, S=S, W=W, lmbda=np.float64(lmbda), X=b.X, L=stat(its, 'L'), Rsdl=stat(its, 'Rsdl'),
               ObjFun=stat(its, 'ObjFun'), DFid=stat(its, 'DFid'), RegL1=stat(its, 'RegL1'),
               IterBTrack=stat(its, 'IterBTrack'), F_Btrack=stat(its, 'F_Btrack'), Q_Btrack=stat(its, 'Q_Btrack'))
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), **out)
    print('wrote', tag)


def level1():
    """Known-answer vectors for the level-1 functions from the reference itself."""
    rng = np.random.default_rng(7)
    ah = (rng.standard_normal((8, 5, 1, 1, 6)) + 1j * rng.standard_normal((8, 5, 1, 1, 6)))
    b = (rng.standard_normal((8, 5, 1, 3, 6)) + 1j * rng.standard_normal((8, 5, 1, 3, 6)))
    x = rlinalg.solvedbi_sm(ah, 0.7, b, None, 4)
    same(x, orc.solvedbi_sm(ah, 0.7, b, 4), 'solvedbi_sm')
    ah3 = (rng.standard_normal((8, 5, 3, 1, 6)) + 1j * rng.standard_normal((8, 5, 3, 1, 6)))
    b3 = (rng.standard_normal((8, 5, 1, 2, 6)) + 1j * rng.standard_normal((8, 5, 1, 2, 6)))
    x3 = rlinalg.solvemdbi_ism(ah3, 0.7, b3, 4, 2)
    same(x3, orc.solvemdbi_ism(ah3, 0.7, b3, 4, 2), 'solvemdbi_ism')
    v = rng.standard_normal((9, 7, 3, 2, 4))
    p1 = rprox.prox_l1(v, 0.4)
    same(p1, orc.prox_l1(v, 0.4), 'prox_l1')
    p21 = rprox.prox_sl1l2(v, 0.3, 0.25, axis=2)
    same(p21, orc.prox_sl1l2(v, 0.3, 0.25, axis=2), 'prox_sl1l2')
    xr = rng.standard_normal((16, 12, 2))
    xf = rfft.rfftn(xr, None, (0, 1))
    n2 = rfft.rfl2norm2(xf, xr.shape, axis=(0, 1))
    same(n2, orc.rfl2norm2(xf, xr.shape, axis=(0, 1)), 'rfl2norm2')
    np.savez_compressed(os.path.join(OUT, 'level1.npz'), ah=ah, b=b, x=x, ah3=ah3, b3=b3, x3=x3,
                        v=v, prox_l1=p1, prox_sl1l2=p21, xr=xr, xf=xf, rfl2norm2=np.float64(n2))
    print('wrote level1')


def main():
    os.makedirs(OUT, exist_ok=True)
    only = [a[7:] for a in sys.argv[1:] if a.startswith('--only=')]
    if only:        # every case still runs (the random stream stays the same); only matching fixtures are written
        prefixes = tuple(only[0].split(','))
        real_save = np.savez_compressed

        def filtered(path, **kw):
            if os.path.basename(path).startswith(prefixes):
                real_save(path, **kw)
        np.savez_compressed = filtered
    level1()
    tikhonov()
    for dt, sfx in ((np.float64, 'f64'), (np.float32, 'f32')):
        rng = np.random.default_rng(12345)
        D = rng.standard_normal((5, 5, 6)).astype(dt)
        S = rng.standard_normal((32, 32, 3)).astype(dt)
        S3 = rng.standard_normal((32, 32, 3, 2)).astype(dt)
        D3 = rng.standard_normal((5, 5, 3, 6)).astype(dt)
        admm_case('admm_k3_' + sfx, dt, D, S, 0.1, {'MaxMainIter': 30, 'RelStopTol': 0.0}, dimK=1)
        admm_case('admm_fixedrho_' + sfx, dt, D, S[..., 0], 0.05,
                  {'MaxMainIter': 40, 'RelStopTol': 0.0, 'rho': 2.0, 'AutoRho': {'Enabled': False},
                   'RelaxParam': 1.0})
        admm_case('admm_nonneg_nobc_' + sfx, dt, D3, S3, 0.1,
                  {'MaxMainIter': 20, 'RelStopTol': 0.0, 'NonNegCoef': True, 'NoBndryCross': True})
        admm_case('joint_c3_' + sfx, dt, D, S3, 0.1, {'MaxMainIter': 20, 'RelStopTol': 0.0}, mu=0.05)
        admm_case('admm_stop_' + sfx, dt, D, S, 0.2, {'MaxMainIter': 200, 'RelStopTol': 5e-3}, dimK=1)
        Wm = (rng.random((32, 32)) > 0.3).astype(dt)
        Wk = (rng.random((32, 32, 3)) > 0.3).astype(dt)
        ams_case('ams_gry_' + sfx, dt, D, S[..., 0], Wm, 0.1, AMS_OPT['ams_gry'])
        ams_case('ams_k3_' + sfx, dt, D, S, Wk, 0.1, AMS_OPT['ams_k3'], dimK=1)
        pgm_mask_case('pgm_mask_' + sfx, dt, D, S, Wk, 0.1,
                      {'MaxMainIter': 20, 'RelStopTol': 0.0, 'L': 5.0,
                       'Backtrack': BacktrackStandard(gamma_u=1.3, maxiter=8)},
                      {'MaxMainIter': 20, 'RelStopTol': 0.0, 'L': 5.0,
                       'Backtrack': {'gamma_u': 1.3, 'maxiter': 8}}, dimK=1)
        gw7 = np.concatenate((np.linspace(0.2, 2.0, 6), [0.0])).astype(dt)
        ams_case('ams_grd_' + sfx, dt, D, S[..., 0], Wm, 0.1,
                 dict(AMS_OPT['ams_gry'], GradWeight=gw7, rho=3.0, AutoRho={'Enabled': False}), grad_mu=0.4)
        admm_case('grd_k3_' + sfx, dt, D, S, 0.1, GRD_OPT['grd_k3'](dt), dimK=1, grad_mu=0.4)
        admm_case('grd_aux_' + sfx, dt, D, S[..., 0], 0.1, GRD_OPT['grd_aux'](dt), grad_mu=0.2)
        admm_case('enet_k3_' + sfx, dt, D, S, 0.1, {'MaxMainIter': 30, 'RelStopTol': 0.0}, dimK=1,
                  enet_mu=0.3)
        admm_case('enet_c3_' + sfx, dt, D3, S3, 0.1, {'MaxMainIter': 20, 'RelStopTol': 0.0,
                                                      'AuxVarObj': True}, enet_mu=0.5)
        pgm_case('pgm_bt_' + sfx, dt, D, S, 0.1,
                 {'MaxMainIter': 25, 'RelStopTol': 0.0, 'L': 10.0,
                  'Backtrack': BacktrackStandard(gamma_u=1.3, maxiter=8)},
                 {'MaxMainIter': 25, 'RelStopTol': 0.0, 'L': 10.0,
                  'Backtrack': {'gamma_u': 1.3, 'maxiter': 8}}, dimK=1)
        pgm_case('pgm_fixed_' + sfx, dt, D, S, 0.1,
                 {'MaxMainIter': 30, 'RelStopTol': 0.0, 'L': 400.0},
                 {'MaxMainIter': 30, 'RelStopTol': 0.0, 'L': 400.0}, dimK=1)
        # multi-channel dictionaries in AddMaskSim (one impulse per channel, channel mask on the impulse maps) and
        # in the masked PGM solver; own generator so that the cases above keep their inputs
        rng3 = np.random.default_rng(4321)
        W3 = (rng3.random((32, 32, 3, 2)) > 0.3).astype(dt)
        ams_case('ams_c3_' + sfx, dt, D3, S3, W3, 0.1, {'MaxMainIter': 20, 'RelStopTol': 0.0})
        pgm_mask_case('pgm_mask_c3_' + sfx, dt, D3, S3, W3, 0.1,
                      {'MaxMainIter': 15, 'RelStopTol': 0.0, 'L': 5.0,
                       'Backtrack': BacktrackStandard(gamma_u=1.3, maxiter=8)},
                      {'MaxMainIter': 15, 'RelStopTol': 0.0, 'L': 5.0,
                       'Backtrack': {'gamma_u': 1.3, 'maxiter': 8}})
        # row a17: step-size policies, monotone FISTA, robust backtracking
        pb = {'MaxMainIter': 25, 'RelStopTol': 0.0}
        pgm_case('pgm_cauchy_' + sfx, dt, D, S, 0.1, dict(pb, L=50.0, StepSizePolicy=StepSizePolicyCauchy()),
                 dict(pb, L=50.0, StepSizePolicy='cauchy'), dimK=1)
        pgm_case('pgm_bb_' + sfx, dt, D, S, 0.1, dict(pb, L=50.0, StepSizePolicy=StepSizePolicyBB()),
                 dict(pb, L=50.0, StepSizePolicy='bb'), dimK=1)
        pgm_case('pgm_mono_' + sfx, dt, D, S, 0.1, dict(pb, L=150.0, Monotone=True),
                 dict(pb, L=150.0, Monotone=True), dimK=1)
        pgm_case('pgm_robust_' + sfx, dt, D, S, 0.1,
                 dict(pb, L=5.0, Backtrack=BacktrackRobust(gamma_d=0.95, gamma_u=1.8, maxiter=10)),
                 dict(pb, L=5.0, Backtrack={'kind': 'robust', 'gamma_d': 0.95, 'gamma_u': 1.8, 'maxiter': 10}),
                 dimK=1)
        D0 = rng.standard_normal((6, 6, 5)).astype(dt)
        S4 = rng.standard_normal((32, 32, 4)).astype(dt)
        cdl_case('cdl_' + sfx, dt, D0, S4, 0.1, CDL_OPT)
        cdl_case('cdl_zm_' + sfx, dt, D0, S4, 0.2, CDL_OPT_ZM)
        Sc = rng.standard_normal((32, 32, 3, 2)).astype(dt)
        D0c = rng.standard_normal((6, 6, 3, 5)).astype(dt)
        cdl_case('cdl_clr1_' + sfx, dt, D0, Sc, 0.1, CDL_OPT_CLR)      # greyscale dictionary, colour signals
        cdl_case('cdl_clr3_' + sfx, dt, D0c, Sc, 0.1, CDL_OPT_CLR)     # colour dictionary
        cdl_ref_case('cdl_accdfid_' + sfx, dt, D0, S4, 0.1, dict(CDL_OPT, AccurateDFid=True), 'admm')
        cdl_ref_case('cdl_pgmx_' + sfx, dt, D0, S4, 0.1, CDL_OPT_PGMX, 'pgm')
        # consensus dictionary update: alone (oracle pinned) and as the D step of dictionary learning
        Zc = rng.standard_normal((32, 32, 1, 3, 6)).astype(dt)
        Zc[np.abs(Zc) < 1.0] = 0
        Sc3 = rng.standard_normal((32, 32, 3)).astype(dt)
        cns_case('cns_zm_' + sfx, dt, Zc, Sc3, (5, 5, 6), {'MaxMainIter': 15, 'ZeroMean': True})
        cns_case('cns_arho_' + sfx, dt, Zc, Sc3, (5, 5, 6),
                 {'MaxMainIter': 15, 'rho': 2.0, 'AutoRho': {'Enabled': True, 'Period': 3, 'AutoScaling': True,
                                                              'Scaling': 10.0}})
        cdl_ref_case('cdl_cns_' + sfx, dt, D0, S4, 0.1, CDL_OPT_CNS, 'admm', 'cns')
        cdl_ref_case('cdl_cns_clr1_' + sfx, dt, D0, Sc, 0.1, CDL_OPT_CNS, 'admm', 'cns')
        # backtracking in the D step, alone and inside dictionary learning (PGM X step with backtracking, colour,
        # multi-scale: the configuration of examples/scripts/cdl/cbpdndl_pgm_clr.py in small)
        Zb = (rng.standard_normal((16, 32, 1, 3, 5)) * (rng.random((16, 32, 1, 3, 5)) < 0.2)).astype(dt)
        Sb = rng.standard_normal((16, 32, 3)).astype(dt)
        ccmod_bt_case('ccmod_bt_' + sfx, dt, Zb, Sb, (4, 6, 5),
                      {'MaxMainIter': 12, 'L': 2.0, 'Backtrack': BacktrackStandard(gamma_u=1.3, maxiter=10),
                       'RelStopTol': 0.0, 'ZeroMean': True})
        D0e = rng.standard_normal((7, 6, 3, 5)).astype(dt)
        cdl_ref_case('cdl_bt_clr_ms_' + sfx, dt, D0e, Sc, 0.1,
                     {'MaxMainIter': 12, 'DictSize': ((4, 4, 3, 3), (7, 6, 3, 2)),
                      'CBPDN': {'Backtrack': BacktrackStandard(gamma_u=1.1), 'L': 10.0},
                      'CCMOD': {'Backtrack': BacktrackStandard(), 'L': 5.0}}, 'pgm')
        # multi-scale dictionaries (DictSize a tuple of blocks): PGM and consensus D steps
        D0m = rng.standard_normal((7, 6, 5)).astype(dt)
        cdl_case('cdl_ms_' + sfx, dt, D0m, S4, 0.1, CDL_OPT_MS)
        cdl_ref_case('cdl_ms_cns_' + sfx, dt, D0m, S4, 0.1, CDL_OPT_MS_CNS, 'admm', 'cns')


if __name__ == '__main__':
    main()

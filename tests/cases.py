"""Parity cases shared by the GPU tests (tests/test_parity_gpu.py, real libspcsc.so) and the
CPU emulation tests (tests/test_kernels_emu.py, same kernel source under tests/emu)."""

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# Tolerances (relative l2 error of the coefficient maps against the reference's output).
# float64: rounding only.  float32: the reference's own float32-vs-float64 drift is ~1e-5
# after 10 AutoRho iterations and ~1e-4 after 50-200, because the rho
# schedule amplifies rounding; fixed-rho runs stay at ~1e-6.  north_star asks for rtol 1e-4.
TOL = {
    ('f64', 'auto'): 1e-9, ('f64', 'fixed'): 1e-10,
    ('f32', 'auto'): 3e-4, ('f32', 'fixed'): 2e-5,
}


def rel(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    wide = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = a.astype(wide)
    b = b.astype(wide)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def load(tag):
    return np.load(os.path.join(GOLDEN, tag + '.npz'))


ADMM_CASES = {
    # tag: (options, dimK, joint, kind)
    'admm_k3': ({'MaxMainIter': 30, 'RelStopTol': 0.0}, 1, False, 'auto'),
    'admm_fixedrho': ({'MaxMainIter': 40, 'RelStopTol': 0.0, 'rho': 2.0,
                       'AutoRho': {'Enabled': False}, 'RelaxParam': 1.0}, None, False, 'fixed'),
    'admm_nonneg_nobc': ({'MaxMainIter': 20, 'RelStopTol': 0.0, 'NonNegCoef': True,
                          'NoBndryCross': True}, None, False, 'auto'),
    'joint_c3': ({'MaxMainIter': 20, 'RelStopTol': 0.0}, None, True, 'auto'),
    'admm_stop': ({'MaxMainIter': 200, 'RelStopTol': 5e-3}, 1, False, 'auto'),
    # ConvElasticNet (third field 'enet'): single- and multi-channel dictionary
    'enet_k3': ({'MaxMainIter': 30, 'RelStopTol': 0.0}, 1, 'enet', 'auto'),
    'enet_c3': ({'MaxMainIter': 20, 'RelStopTol': 0.0, 'AuxVarObj': True}, None, 'enet', 'auto'),
    # ConvBPDNGradReg (third field 'grd'; GradWeight filled in by _grd_opt)
    'grd_k3': ({'MaxMainIter': 30, 'RelStopTol': 0.0}, 1, 'grd', 'auto'),
    'grd_aux': ({'MaxMainIter': 20, 'RelStopTol': 0.0, 'AuxVarObj': True, 'LinSolveCheck': True,
                 'GradWeight': 'linspace'}, None, 'grd', 'auto'),
}


def grd_opt(opt, dtype):
    o = dict(opt)
    if isinstance(o.get('GradWeight'), str):
        o['GradWeight'] = np.linspace(0.2, 2.0, 6).astype(dtype)
    return o


def run_admm_case(tag, sfx):
    """Solve the golden problem `tag` with sporco_b200 and compare with the reference's
    stored outputs.  Returns the solver for extra checks."""
    from sporco_b200.admm import cbpdn
    g = load('%s_%s' % (tag, sfx))
    opt, dimK, joint, kind = ADMM_CASES[tag]
    tol = TOL[(sfx, kind)]
    D, S = g['D'], g['S']
    enet, grd = joint == 'enet', joint == 'grd'
    joint = joint is True
    if grd:
        b = cbpdn.ConvBPDNGradReg(D, S, float(g['lmbda']), float(g['mu']),
                                  cbpdn.ConvBPDNGradReg.Options(grd_opt(opt, D.dtype)), dimK=dimK)
    elif enet:
        b = cbpdn.ConvElasticNet(D, S, float(g['lmbda']), float(g['mu']),
                                 cbpdn.ConvBPDN.Options(opt), dimK=dimK)
    elif joint:
        b = cbpdn.ConvBPDNJoint(D, S, float(g['lmbda']), float(g['mu']),
                                cbpdn.ConvBPDNJoint.Options(opt), dimK=dimK)
    else:
        b = cbpdn.ConvBPDN(D, S, float(g['lmbda']), cbpdn.ConvBPDN.Options(opt), dimK=dimK)
    Y = b.solve()
    its = b.getitstat()
    n = len(g['Rho'])
    assert len(its.Rho) == n, 'iteration count %d != reference %d' % (len(its.Rho), n)
    assert Y.dtype == D.dtype and Y.shape == g['Y'].shape
    assert rel(Y, g['Y']) <= tol, 'Y: %.3e' % rel(Y, g['Y'])
    assert rel(b.U, g['U']) <= 4 * tol, 'U: %.3e' % rel(b.U, g['U'])
    assert rel(b.X, g['X']) <= 4 * tol, 'X: %.3e' % rel(b.X, g['X'])
    stol = max(tol, 1e-6 if sfx == 'f32' else 1e-11)
    assert rel(its.Rho, g['Rho']) <= 10 * stol
    assert rel(its.ObjFun, g['ObjFun']) <= 10 * stol
    assert rel(its.DFid, g['DFid']) <= 40 * stol
    assert rel(its.RegL1, g['RegL1']) <= 10 * stol
    assert rel(its.PrimalRsdl, g['PrimalRsdl']) <= 40 * stol
    assert rel(its.DualRsdl, g['DualRsdl']) <= 40 * stol
    if joint:
        assert rel(its.RegL21, g['RegL21']) <= 10 * stol
    if enet:
        assert rel(its.RegL2, g['RegL2']) <= 10 * stol
    if grd:
        assert rel(its.RegGrad, g['RegGrad']) <= 10 * stol
        if opt.get('LinSolveCheck'):
            assert max(its.XSlvRelRes) < (1e-10 if sfx == 'f64' else 1e-4)
    assert rel(b.reconstruct().reshape(g['recon'].shape), g['recon']) <= 4 * tol
    return b


def run_fresh_case(N0, N1, M, K, C=None, mu=None, iters=8, extra=None, dt=np.float32, seed=1,
                   tol=3e-4):
    """Seeded random problem solved by sporco_b200 and by the oracle; sizes chosen so that the
    register-plan (v2) kernels, including the cluster-split column kernel, are exercised."""
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(seed)
    D = rng.standard_normal((5, 5, M)).astype(dt)
    shp = (N0, N1) + ((C,) if C else ()) + ((K,) if K else ())
    S = rng.standard_normal(shp).astype(dt)
    o = {'MaxMainIter': iters, 'RelStopTol': 0.0}
    o.update(extra or {})
    dimK = None if (C and K) else (1 if K else 0)
    if mu is None:
        b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o), dimK=dimK)
    else:
        b = cbpdn.ConvBPDNJoint(D, S, 0.1, mu, cbpdn.ConvBPDNJoint.Options(o), dimK=dimK)
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, mu=mu, opt=o, dimK=dimK)
    its = b.getitstat()
    rc = 9 if mu is not None else 8
    assert rel(Y, r.Y) < tol, 'Y %.3e' % rel(Y, r.Y)
    assert rel(b.U, r.U) < 2 * tol
    assert rel(its.Rho, [x[rc] for x in r.itstat]) < tol
    assert rel(its.ObjFun, [x[1] for x in r.itstat]) < tol
    return b, r


FRESH_CASES = [
    # N0, N1, M, K, C, mu, extra        (what it exercises)
    (256, 64, 40, 2, None, None, None),                    # column cluster of 2 (ragged), rows E=8
    (64, 256, 12, 2, None, None, None),                    # rows E=16
    (128, 128, 70, 1, None, None, {'NonNegCoef': True}),   # cluster of 2, M not a multiple
    (64, 64, 8, 2, 2, 0.05, None),                         # joint prox, CX=2
    (32, 512, 5, 1, None, None, None),                     # long rows (H=256)
    (512, 64, 9, 1, None, None, None),                     # long columns (full-warp plan)
]


FRESH_CASES_F64 = [
    # float64 on the register plans: 8 elements per lane, one column per lane group
    (64, 128, 12, 2, None, None, None),                    # rows H=64, 8-column CTAs: cluster of 2
    (256, 256, 20, 1, None, None, {'NonNegCoef': True}),   # the benchmark's transform sizes, cluster of 3
    (128, 64, 6, 2, 2, 0.05, None),                        # joint prox, CX=2
]


def run_pgm_cases(sfx):
    """PGM / FISTA golden problems (fixed step and standard backtracking)."""
    from sporco_b200.pgm import cbpdn as pcbpdn
    from sporco_b200.pgm.backtrack import BacktrackStandard
    tol = 1e-11 if sfx == 'f64' else 5e-5
    g = load('pgm_bt_' + sfx)
    opt = pcbpdn.ConvBPDN.Options({'MaxMainIter': 25, 'RelStopTol': 0.0, 'L': 10.0,
                                   'Backtrack': BacktrackStandard(gamma_u=1.3, maxiter=8)})
    b = pcbpdn.ConvBPDN(g['D'], g['S'], float(g['lmbda']), opt, dimK=1)
    X = b.solve()
    its = b.getitstat()
    assert X.dtype == g['X'].dtype and rel(X, g['X']) < tol
    assert np.array_equal(np.asarray(its.IterBTrack, dtype=np.float64), g['IterBTrack'])
    assert rel(its.L, g['L']) < 1e-6
    assert rel(its.F_Btrack, g['F_Btrack']) < 10 * tol and rel(its.Q_Btrack, g['Q_Btrack']) < 10 * tol
    assert rel(its.Rsdl, g['Rsdl']) < 10 * tol and rel(its.ObjFun, g['ObjFun']) < 10 * tol
    assert rel(b.reconstruct().squeeze(), np.zeros(1)) >= 0      # runs
    g = load('pgm_fixed_' + sfx)
    opt = pcbpdn.ConvBPDN.Options({'MaxMainIter': 30, 'RelStopTol': 0.0, 'L': 400.0})
    b = pcbpdn.ConvBPDN(g['D'], g['S'], float(g['lmbda']), opt, dimK=1)
    X = b.solve()
    its = b.getitstat()
    assert rel(X, g['X']) < tol
    assert rel(its.ObjFun, g['ObjFun']) < 10 * tol and rel(its.DFid, g['DFid']) < 10 * tol
    assert rel(its.RegL1, g['RegL1']) < 10 * tol and rel(its.Rsdl, g['Rsdl']) < 10 * tol
    return b


def run_level1_cases():
    """The level-1 entry points (spcsc_solvedbi_sm, spcsc_prox_l1, spcsc_prox_sl1l2) against the
    reference's own outputs in tests/golden/level1.npz, both precisions, and the algebraic pins of
    the reference's unit tests (tests/test_linalg.py:147-207)."""
    from sporco_b200 import linalg, prox
    g = load('level1')
    for cdt, tol in ((np.complex128, 1e-13), (np.complex64, 2e-6)):
        rdt = np.float64 if cdt == np.complex128 else np.float32
        x = linalg.solvedbi_sm(g['ah'].astype(cdt), 0.7, g['b'].astype(cdt))
        assert x.dtype == cdt and rel(x, g['x']) < tol
        a = np.conj(g['ah'])
        lhs = a * np.sum(g['ah'] * x, axis=4, keepdims=True) + 0.7 * x
        assert rel(lhs, g['b']) < 10 * tol                       # (rho I + a a^H) x = b
        x3 = linalg.solvemdbi_ism(g['ah3'].astype(cdt), 0.7, g['b3'].astype(cdt), 4, 2)
        assert rel(x3, g['x3']) < tol
        p1 = prox.prox_l1(g['v'].astype(rdt), 0.4)
        assert p1.dtype == rdt and rel(p1, g['prox_l1']) < tol
        w = np.linspace(0.2, 1.0, 4).astype(rdt)
        pw = prox.prox_l1(g['v'].astype(rdt), 0.4 * w)
        vv = g['v']
        assert rel(pw, np.sign(vv) * np.maximum(np.abs(vv) - 0.4 * w, 0)) < tol
        p21 = prox.prox_sl1l2(g['v'].astype(rdt), 0.3, 0.25, axis=2)
        assert rel(p21, g['prox_sl1l2']) < tol


PGM_VARIANTS = ('cauchy', 'bb', 'mono', 'robust')


def run_pgm_variant_case(name, sfx):
    """Row a17: StepSizePolicyCauchy / StepSizePolicyBB, Monotone, BacktrackRobust against the reference's
    stored outputs (fixtures pgm_<name>_<sfx>, generated from the imported reference)."""
    from sporco_b200.pgm import cbpdn as pcbpdn
    from sporco_b200.pgm.backtrack import BacktrackRobust
    from sporco_b200.pgm.stepsize import StepSizePolicyBB, StepSizePolicyCauchy
    g = load('pgm_%s_%s' % (name, sfx))
    extra = {'cauchy': {'L': 50.0, 'StepSizePolicy': StepSizePolicyCauchy()},
             'bb': {'L': 50.0, 'StepSizePolicy': StepSizePolicyBB()},
             'mono': {'L': 150.0, 'Monotone': True},
             'robust': {'L': 5.0, 'Backtrack': BacktrackRobust(gamma_d=0.95, gamma_u=1.8, maxiter=10)}}[name]
    opt = pcbpdn.ConvBPDN.Options(dict({'MaxMainIter': 25, 'RelStopTol': 0.0}, **extra))
    b = pcbpdn.ConvBPDN(g['D'], g['S'], float(g['lmbda']), opt, dimK=1)
    X = b.solve()
    its = b.getitstat()
    # the Barzilai-Borwein quotient amplifies rounding (its denominator is a difference of nearly equal
    # sums), so float32 trajectories of that policy are compared more loosely
    tol = 1e-10 if sfx == 'f64' else (2e-3 if name == 'bb' else 1e-4)
    assert X.dtype == g['X'].dtype
    assert rel(its.L, g['L']) < tol, 'L %.3e' % rel(its.L, g['L'])
    assert rel(X, g['X']) < tol, 'X %.3e' % rel(X, g['X'])
    assert rel(its.ObjFun, g['ObjFun']) < tol and rel(its.DFid, g['DFid']) < 4 * tol
    assert rel(its.RegL1, g['RegL1']) < tol and rel(its.Rsdl, g['Rsdl']) < 10 * tol
    if name == 'robust':
        assert np.array_equal(np.asarray(its.IterBTrack, dtype=np.float64), g['IterBTrack'])
        assert rel(its.F_Btrack, g['F_Btrack']) < 10 * tol and rel(its.Q_Btrack, g['Q_Btrack']) < 10 * tol
    return b


def run_fusion_cases():
    """The optimistic cross-iteration fusion (the prox kernel also emits the next iteration's
    row spectra): consumed when rho is constant, redone when it changed; X must stay
    retrievable; batches of odd/even length and a device-side stop inside a batch."""
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(21)
    D = rng.standard_normal((5, 5, 12)).astype(np.float32)
    S = rng.standard_normal((64, 256, 2)).astype(np.float32)
    fixed = {'RelStopTol': 0.0, 'rho': 5.0, 'AutoRho': {'Enabled': False}}
    # (a) constant rho: every fused spectrum is used
    o = dict(fixed, MaxMainIter=13)
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o), dimK=1)
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, opt=o, dimK=1)
    assert rel(Y, r.Y) < 5e-5 and rel(b.X, r.X) < 5e-5 and rel(b.U, r.U) < 5e-5
    # (b) the same in two calls of 7 + 6 iterations (odd, then even number of buffer swaps)
    o1 = dict(fixed, MaxMainIter=7)
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o1), dimK=1)
    b.solve()
    x7 = b.X.copy()
    r7 = orc.admm_convbpdn(D, S, 0.1, opt=o1, dimK=1)
    assert rel(x7, r7.X) < 5e-5
    b.opt['MaxMainIter'] = 6
    Y = b.solve()
    assert b.k == 13 and rel(Y, r.Y) < 5e-5 and rel(b.X, r.X) < 5e-5
    # (c) AutoRho with a stop inside a batch
    o = {'MaxMainIter': 60, 'RelStopTol': 3e-2}
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o), dimK=1)
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, opt=o, dimK=1)
    assert 2 < r.k < 60, r.k
    assert b.k == r.k and len(b.itstat) == r.k
    assert rel(Y, r.Y) < 3e-4 and rel(b.X, r.X) < 3e-4
    Y2 = b.solve()            # re-entering runs one more iteration, as the reference does
    assert b.k == r.k + 1


def run_auxvarobj_case(dt=np.float64):
    """AuxVarObj=True: objective evaluated on the auxiliary variable Y (admm/cbpdn.py:151-164,
    315-344): DFid from rfftn(Y), RegL1 on Y."""
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(8)
    D = rng.standard_normal((4, 4, 5)).astype(dt)
    S = rng.standard_normal((32, 64, 2)).astype(dt)
    o = {'MaxMainIter': 12, 'RelStopTol': 0.0, 'AuxVarObj': True}
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o), dimK=1)
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, opt={'MaxMainIter': 12, 'RelStopTol': 0.0, 'AuxVarObj': True}, dimK=1)
    its = b.getitstat()
    tol = 1e-9 if dt == np.float64 else 3e-4
    assert rel(Y, r.Y) < tol
    assert rel(its.DFid, [x[2] for x in r.itstat]) < 10 * tol
    assert rel(its.RegL1, [x[3] for x in r.itstat]) < 10 * tol
    assert rel(its.ObjFun, [x[1] for x in r.itstat]) < 10 * tol


def run_reproducibility_case():
    """Two solves of the same problem give bit-identical results: the sums that steer the
    algorithm are accumulated order-independently (integer bins), whatever the order in which
    thread blocks finish (cf. the reference's pickle test, tests/admm/test_cbpdn.py:631-645)."""
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(5)
    D = rng.standard_normal((5, 5, 8))
    S = rng.standard_normal((64, 64, 3))
    outs = []
    for _ in range(2):
        b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options({'MaxMainIter': 25, 'RelStopTol': 0.0}),
                           dimK=1)
        Y = b.solve().copy()
        outs.append((Y, np.array(b.getitstat().Rho), np.array(b.getitstat().PrimalRsdl)))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])

FRESH_CASES += [
    (256, 64, 40, 2, 3, None, None),                       # 3-channel signal, Cd=1: CX=3 fused async prox
    (64, 128, 10, 2, 3, 0.04, None),                       # joint l2,1 over 3 channels, fused
]


def run_multichannel_dict_cases():
    """Multi-channel dictionary (Cd=3, Woodbury solve in the cluster column kernel), plain and
    with the joint penalty, and the joint penalty with a single coefficient channel."""
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(2)
    dt = np.float32
    D3 = rng.standard_normal((5, 5, 3, 40)).astype(dt)
    S3 = rng.standard_normal((256, 64, 3, 2)).astype(dt)
    o = {'MaxMainIter': 8, 'RelStopTol': 0.0}
    for mu in (None, 0.03):
        if mu is None:
            b = cbpdn.ConvBPDN(D3, S3, 0.1, cbpdn.ConvBPDN.Options(o))
        else:
            b = cbpdn.ConvBPDNJoint(D3, S3, 0.1, mu, cbpdn.ConvBPDNJoint.Options(o))
        Y = b.solve()
        r = orc.admm_convbpdn(D3, S3, 0.1, mu=mu, opt=o)
        assert b._h.admm_schedule_info()['col_v2']
        assert rel(Y, r.Y) < 3e-4 and rel(b.getitstat().ObjFun, [x[1] for x in r.itstat]) < 1e-4
    D = rng.standard_normal((5, 5, 12)).astype(dt)
    S = rng.standard_normal((64, 256, 2)).astype(dt)
    o = {'MaxMainIter': 8, 'RelStopTol': 0.0, 'L21Weight': np.linspace(0.5, 1.5, 12).astype(dt)}
    b = cbpdn.ConvBPDNJoint(D, S, 0.1, 0.05, cbpdn.ConvBPDNJoint.Options(o), dimK=1)
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, mu=0.05, opt=o, dimK=1)
    assert rel(Y, r.Y) < 3e-4 and rel(b.getitstat().RegL21, [x[4] for x in r.itstat]) < 1e-4


def run_option_cases(capsys=None):
    """Smaller option paths of the reference frame: StdResiduals, default lambda, Verbose status
    table, Callback stop, AbsStopTol, AutoRho period / fixed scaling."""
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(31)
    D = rng.standard_normal((4, 4, 6))
    S = rng.standard_normal((32, 32, 2))
    # StdResiduals + AbsStopTol (admm/admm.py:465-471)
    o = {'MaxMainIter': 15, 'RelStopTol': 1e-4, 'AbsStopTol': 1e-6,
         'AutoRho': {'StdResiduals': True, 'Period': 3, 'AutoScaling': False, 'Scaling': 2.0,
                     'RsdlRatio': 5.0}}
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o), dimK=1)
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, opt=o, dimK=1)
    its = b.getitstat()
    assert len(its.Rho) == r.k
    assert rel(Y, r.Y) < 1e-9 and rel(its.Rho, [x[8] for x in r.itstat]) < 1e-12
    assert rel(its.PrimalRsdl, [x[4] for x in r.itstat]) < 1e-9
    assert rel(its.EpsPrimal, [x[6] for x in r.itstat]) < 1e-9
    assert rel(its.EpsDual, [x[7] for x in r.itstat]) < 1e-9
    # default lambda = 0.1 max |D^H s| (admm/cbpdn.py:573-578) and default rho / rho_xi
    o = {'MaxMainIter': 5, 'RelStopTol': 0.0}
    b = cbpdn.ConvBPDN(D, S, None, cbpdn.ConvBPDN.Options(o), dimK=1)
    b.solve()
    r = orc.admm_convbpdn(D, S, None, opt=o, dimK=1)
    assert abs(float(b.lmbda) - float(r.lmbda)) < 1e-12 * float(r.lmbda)
    assert rel(b.Y, r.Y) < 1e-9
    # Callback: stop after 4 iterations (admm/admm.py:370-372); k counts like the reference
    seen = []

    def cb(obj):
        seen.append(obj.k)
        return len(seen) >= 4
    o = {'MaxMainIter': 20, 'RelStopTol': 0.0, 'Callback': cb}
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o), dimK=1)
    b.solve()
    assert seen == [0, 1, 2, 3] and b.k == 4 and len(b.itstat) == 4
    # Verbose table: header, separator, one line per iteration (admm/admm.py:579-625)
    if capsys is not None:
        capsys.readouterr()
        o = {'MaxMainIter': 3, 'RelStopTol': 0.0, 'Verbose': True}
        b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o), dimK=1)
        b.solve()
        out = capsys.readouterr().out.strip().splitlines()
        assert out[0].split() == ['Itn', 'Fnc', 'DFid', u'Regℓ1', 'r', 's', u'ρ']
        assert set(out[1]) == {'-'} and set(out[-1]) == {'-'} and len(out) == 6
        assert out[2].split()[0] == '0' and len(out[2].split()) == 7
    # timers exist and advanced
    assert b.timer.elapsed('solve') > 0 and b.timer.elapsed('init') > 0


# ---- convolutional dictionary learning (tests/golden/cdl_*.npz, generated from the reference's
# ConvBPDNDictLearn by oracle/make_golden.py)
CDL_OPT = {'MaxMainIter': 25, 'CBPDN': {'rho': 5.0, 'AutoRho': {'Period': 4}}, 'CCMOD': {'L': 40.0}}
CDL_CASES = {
    # tag: (options, xmethod, lambda)
    'cdl': (CDL_OPT, 'admm', 0.1),
    'cdl_zm': ({'MaxMainIter': 20, 'CBPDN': {'NonNegCoef': True},
                'CCMOD': {'L': 60.0, 'ZeroMean': True}}, 'admm', 0.2),
    'cdl_accdfid': (dict(CDL_OPT, AccurateDFid=True), 'admm', 0.1),
    'cdl_clr1': ({'MaxMainIter': 15, 'CBPDN': {'rho': 5.0, 'AutoRho': {'Period': 4}},
                  'CCMOD': {'L': 60.0, 'ZeroMean': True}}, 'admm', 0.1),       # colour signals, greyscale dictionary
    'cdl_clr3': ({'MaxMainIter': 15, 'CBPDN': {'rho': 5.0, 'AutoRho': {'Period': 4}},
                  'CCMOD': {'L': 60.0, 'ZeroMean': True}}, 'admm', 0.1),       # colour dictionary
    'cdl_pgmx': ({'MaxMainIter': 20, 'CBPDN': {'L': 80.0}, 'CCMOD': {'L': 40.0}}, 'pgm', 0.1),
    # consensus ADMM dictionary update (dmethod 'cns'): greyscale; colour signals with a greyscale dictionary
    'cdl_cns': ({'MaxMainIter': 15, 'CBPDN': {'rho': 5.0}, 'CCMOD': {'rho': 2.0, 'ZeroMean': True}}, 'admm', 0.1, 'cns'),
    'cdl_cns_clr1': ({'MaxMainIter': 15, 'CBPDN': {'rho': 5.0}, 'CCMOD': {'rho': 2.0, 'ZeroMean': True}}, 'admm', 0.1,
                     'cns'),
    # multi-scale dictionary (DictSize a tuple of blocks: three 4x4 and two 7x6 filters), PGM and consensus D steps
    'cdl_ms': ({'MaxMainIter': 15, 'DictSize': ((4, 4, 3), (7, 6, 2)), 'CBPDN': {'rho': 5.0},
                'CCMOD': {'L': 50.0, 'ZeroMean': True}}, 'admm', 0.1),
    'cdl_ms_cns': ({'MaxMainIter': 15, 'DictSize': ((4, 4, 3), (7, 6, 2)), 'CBPDN': {'rho': 5.0},
                    'CCMOD': {'rho': 2.0, 'ZeroMean': True}}, 'admm', 0.1, 'cns'),
    # backtracking in both steps, colour signals, multi-scale colour dictionary (examples/scripts/cdl/cbpdndl_pgm_clr.py
    # in small); the Backtrack objects are filled in by run_cdl_case
    'cdl_bt_clr_ms': ({'MaxMainIter': 12, 'DictSize': ((4, 4, 3, 3), (7, 6, 3, 2)),
                       'CBPDN': {'Backtrack': ('std', 1.1), 'L': 10.0}, 'CCMOD': {'Backtrack': ('std', 1.2), 'L': 5.0}},
                      'pgm', 0.1),
}


def run_cdl_case(tag, sfx):
    """Learn the golden dictionary with sporco_b200 and compare with the reference's outputs."""
    from sporco_b200.dictlrn import cbpdndl
    g = load('%s_%s' % (tag, sfx))
    o, xmethod, lmbda = CDL_CASES[tag][:3]
    dmethod = CDL_CASES[tag][3] if len(CDL_CASES[tag]) > 3 else 'pgm'
    if any(isinstance(o.get(k, {}).get('Backtrack'), tuple) for k in ('CBPDN', 'CCMOD')):
        from sporco_b200.pgm.backtrack import BacktrackStandard
        o = {k: (dict(v) if isinstance(v, dict) else v) for k, v in o.items()}
        for k in ('CBPDN', 'CCMOD'):
            bt = o[k].get('Backtrack')
            if isinstance(bt, tuple):
                o[k]['Backtrack'] = BacktrackStandard(gamma_u=bt[1])
    assert float(g['lmbda']) == lmbda
    # float32: north_star's rtol 1e-4, except the PGM X step and the consensus D step cases, where the
    # reference's own float32 and float64 runs drift apart by 2e-4 ... 4e-4 (D) over these alternations
    tol = 1e-10 if sfx == 'f64' else (1e-3 if (xmethod == 'pgm' or dmethod == 'cns') else 1e-4)
    opt = cbpdndl.ConvBPDNDictLearn.Options(o, xmethod=xmethod, dmethod=dmethod)
    b = cbpdndl.ConvBPDNDictLearn(g['D0'], g['S'], lmbda, opt, xmethod=xmethod, dmethod=dmethod)
    D = b.solve()
    its = b.getitstat()
    assert D.dtype == g['D'].dtype and D.shape == g['D'].shape
    assert rel(D, g['D']) <= tol, 'D: %.3e' % rel(D, g['D'])
    X = b.getcoef()
    assert X.shape == g['X'].shape and rel(X, g['X']) <= 3 * tol, 'X: %.3e' % rel(X, g['X'])
    names = [f for f in its._fields if f in g.files and f != 'Cnstr']
    assert len(names) >= 6
    if dmethod == 'cns':
        assert 'DPrRsdl' in names and 'DDlRsdl' in names and 'DRho' in names
    for f in names:
        assert len(getattr(its, f)) == len(g[f])
        assert rel(getattr(its, f), g[f]) <= 10 * tol, '%s: %.3e' % (f, rel(getattr(its, f), g[f]))
    # constraint violation of a projected iterate: rounding level in both implementations
    lim = 1e-12 if sfx == 'f64' else 1e-5
    assert np.max(np.abs(its.Cnstr)) < lim and np.max(np.abs(g['Cnstr'])) < lim
    # learned filters: unit norm, support respected, dictionary handed to the X step
    assert np.allclose(np.sqrt(np.sum(D.astype(np.float64) ** 2, (0, 1, 2))), 1.0,      # over support and channels
                       atol=1e-12 if sfx == 'f64' else 1e-5)
    full = b.getdict(crop=False)
    assert full.shape[:2] == g['S'].shape[:2] and not np.any(full[D.shape[0]:]) \
        and not np.any(full[:, D.shape[1]:])
    if o.get('DictSize') is not None:                   # every block zero outside its own support
        m0 = 0
        for blk in o['DictSize']:
            assert not np.any(D[blk[0]:, ..., m0:m0 + blk[-1]]) and not np.any(D[:, blk[1]:, ..., m0:m0 + blk[-1]])
            m0 += blk[-1]
    rec = b.reconstruct()
    assert rec.shape[:2] == g['S'].shape[:2]
    assert rel(b.xstep.D.squeeze(), D.squeeze()) == 0.0
    return b


def run_ccmod_standalone(dt):
    """sporco_b200.pgm.ccmod.ConvCnstrMOD on its own (coefficients from the host) against the
    oracle's D-step algebra."""
    from oracle import cbpdn_oracle as co
    from oracle import cbpdndl_oracle as oc
    from sporco_b200.pgm import ccmod
    rng = np.random.default_rng(5)
    N0, N1, K, M, hd, wd = 16, 32, 3, 5, 4, 6
    S = rng.standard_normal((N0, N1, K)).astype(dt)
    Z = (rng.standard_normal((N0, N1, 1, K, M)) * (rng.random((N0, N1, 1, K, M)) < 0.2)).astype(dt)
    X0 = np.zeros((N0, N1, 1, 1, M), dt)
    X0[:hd, :wd] = oc.pcn(rng.standard_normal((N0, N1, 1, 1, M)).astype(dt), (hd, wd, M),
                          (N0, N1))[:hd, :wd]
    L = 25.0
    opt = ccmod.ConvCnstrMOD.Options({'MaxMainIter': 7, 'L': L, 'X0': X0, 'RelStopTol': 0.0})
    c = ccmod.ConvCnstrMOD(Z, S, (hd, wd, M), opt)
    c.solve()
    its = c.getitstat()
    # numpy restatement of pgm/pgm.py:328-370 + pgm/ccmod.py:295-376
    fft = co.FFTBackend()
    ax = (0, 1)
    Sf = fft.rfftn(S.reshape(N0, N1, 1, K, 1), None, ax)
    Zf = fft.rfftn(Z, None, ax)
    Xd = X0.copy()
    Xf = fft.rfftn(Xd, None, ax)
    Yf = Xf
    t = 1.0
    dfd, rs = [], []
    for _ in range(7):
        Xfp, Yfp = Xf, Yf
        g = co.inner(np.conj(Zf), co.inner(Zf, Yf, 4) - Sf, 3)
        V = fft.irfftn(Yf - g / dt(L), (N0, N1), ax)
        Xd = oc.pcn(V, (hd, wd, M), (N0, N1))
        Xf = fft.rfftn(Xd, None, ax)
        tp = t
        t = 0.5 * (1. + np.sqrt(1. + 4. * t ** 2))
        Yf = Xf + ((tp - 1.) / t) * (Xf - Xfp)
        rs.append(co.rfl2norm2(Xf - Yfp, Xd.shape, axis=ax))
        dfd.append(co.rfl2norm2(co.inner(Zf, Xf, 4) - Sf, (N0, N1, 1, K, 1), axis=ax) / 2.0)
    tol = 1e-11 if dt == np.float64 else 3e-5
    assert rel(c.getdict(crop=False), Xd) < tol
    assert c.getdict().shape == (hd, wd, 1, 1, M)
    assert rel(its.DFid, dfd) < 10 * tol and rel(its.Rsdl, rs) < 10 * tol
    assert all(v == dt(L) for v in its.L)
    return c


def run_long_determinism_case(iters=400, K=8):
    """At the benchmark's transform sizes (256x256, 64 filters) a long run is bit-reproducible and,
    once AutoRho has settled (iteration ~40 on this input), rho stays put: the residual ratio sits
    well inside the dead band, so any change later on means corrupted data (this is how a race
    between the bulk-copy engine and in-flight shared-memory loads showed up)."""
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(12345)
    D = rng.standard_normal((8, 8, 64)).astype(np.float32)
    D /= np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))
    S = rng.standard_normal((256, 256, K)).astype(np.float32)
    outs = []
    for _ in range(2):
        b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options({'MaxMainIter': iters, 'RelStopTol': 0.0}),
                           dimK=1)
        b.solve()
        its = b.getitstat()
        outs.append((np.array(its.Rho), np.array(its.PrimalRsdl), np.array(its.DualRsdl)))
        del b
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    rho, r, s = outs[0]
    assert np.all(rho[60:] == rho[60]), 'rho changed after settling: %s' % np.nonzero(np.diff(rho[60:]))[0][:5]
    assert np.all(np.diff(r[60:]) < 0) and np.all(np.diff(s[60:]) < 0)      # monotone decrease


AMS_CASES = {'ams_gry': ({'MaxMainIter': 30, 'RelStopTol': 0.0}, None),
             'ams_grd': ({'MaxMainIter': 30, 'RelStopTol': 0.0, 'rho': 3.0, 'AutoRho': {'Enabled': False},
                          'GradWeight': 'ams7'}, None),
             'ams_k3': ({'MaxMainIter': 20, 'RelStopTol': 0.0, 'NonNegCoef': True, 'NoBndryCross': True,
                         'AuxVarObj': True}, 1),
             # multi-channel dictionary: one impulse filter per channel, the mask's channels on the impulse maps
             'ams_c3': ({'MaxMainIter': 20, 'RelStopTol': 0.0}, None)}


def run_ams_case(tag, sfx):
    """AddMaskSim about ConvBPDN against the reference's outputs (tests/golden/ams_*.npz)."""
    from sporco_b200.admm import cbpdn
    g = load('%s_%s' % (tag, sfx))
    opt, dimK = AMS_CASES[tag]
    tol = TOL[(sfx, 'auto')]
    if tag == 'ams_grd':
        o = dict(opt, GradWeight=np.concatenate((np.linspace(0.2, 2.0, 6), [0.0])).astype(g['D'].dtype))
        b = cbpdn.AddMaskSim(cbpdn.ConvBPDNGradReg, g['D'], g['S'], g['W'], float(g['lmbda']), 0.4,
                             cbpdn.ConvBPDNGradReg.Options(o), dimK=dimK)
    else:
        b = cbpdn.AddMaskSim(cbpdn.ConvBPDN, g['D'], g['S'], g['W'], float(g['lmbda']),
                             cbpdn.ConvBPDN.Options(opt), dimK=dimK)
    X = b.solve()
    its = b.getitstat()
    assert X.shape == g['Xprimary'].shape and rel(X, g['Xprimary']) <= 4 * tol
    assert rel(b.cbpdn.Y, g['Y']) <= tol, 'Y: %.3e' % rel(b.cbpdn.Y, g['Y'])
    stol = max(tol, 1e-6 if sfx == 'f32' else 1e-11)
    assert len(its.Rho) == len(g['Rho']) and rel(its.Rho, g['Rho']) <= 10 * stol
    assert rel(its.ObjFun, g['ObjFun']) <= 10 * stol and rel(its.RegL1, g['RegL1']) <= 10 * stol
    assert rel(its.DFid, g['DFid']) <= 40 * stol
    assert rel(its.PrimalRsdl, g['PrimalRsdl']) <= 40 * stol and rel(its.DualRsdl, g['DualRsdl']) <= 40 * stol
    assert rel(b.reconstruct().reshape(g['recon'].shape), g['recon']) <= 4 * tol
    assert b.getcoef().shape == X.shape
    return b


TIKHONOV_CASES = (((32, 32), 5.0, 16), ((40, 36, 3), 2.0, 16), ((31, 33, 2, 2), 10.0, 8))


def run_tikhonov_cases():
    """sporco_b200.signal.tikhonov_filter against the reference's outputs."""
    from sporco_b200 import signal
    g = load('tikhonov')
    for sfx, tol in (('f64', 1e-12), ('f32', 2e-6)):
        for i, (shape, lm, npd) in enumerate(TIKHONOV_CASES):
            s = g['s%d_%s' % (i, sfx)]
            sl, sh = signal.tikhonov_filter(s, lm, npd)
            assert sl.shape == s.shape and sl.dtype == s.dtype and sh.dtype == s.dtype
            assert rel(sl, g['sl%d_%s' % (i, sfx)]) < tol, (sfx, i, rel(sl, g['sl%d_%s' % (i, sfx)]))
            assert rel(sh, g['sh%d_%s' % (i, sfx)]) < 10 * tol
            assert np.allclose(sl + sh, s, atol=1e-6 if sfx == 'f32' else 1e-14)


def run_pgm_mask_case(sfx, tag='pgm_mask'):
    """pgm.cbpdn.ConvBPDNMask with backtracking against the reference's outputs (`pgm_mask`: single-channel
    dictionary, three images; `pgm_mask_c3`: multi-channel dictionary)."""
    from sporco_b200.pgm import cbpdn as pcbpdn
    from sporco_b200.pgm.backtrack import BacktrackStandard
    tol = 1e-11 if sfx == 'f64' else 5e-5
    g = load('%s_%s' % (tag, sfx))
    c3 = tag.endswith('c3')
    opt = pcbpdn.ConvBPDN.Options({'MaxMainIter': 15 if c3 else 20, 'RelStopTol': 0.0, 'L': 5.0,
                                   'Backtrack': BacktrackStandard(gamma_u=1.3, maxiter=8)})
    b = pcbpdn.ConvBPDNMask(g['D'], g['S'], float(g['lmbda']), g['W'], opt, dimK=None if c3 else 1)
    X = b.solve()
    its = b.getitstat()
    assert X.dtype == g['X'].dtype and rel(X, g['X']) < tol
    assert np.array_equal(np.asarray(its.IterBTrack, dtype=np.float64), g['IterBTrack'])
    assert rel(its.L, g['L']) < 1e-6
    assert rel(its.F_Btrack, g['F_Btrack']) < 10 * tol and rel(its.Q_Btrack, g['Q_Btrack']) < 10 * tol
    assert rel(its.ObjFun, g['ObjFun']) < 10 * tol and rel(its.DFid, g['DFid']) < 10 * tol
    assert rel(its.Rsdl, g['Rsdl']) < 10 * tol
    return b


# ---- consensus dictionary update (admm.ccmod.ConvCnstrMOD_Consensus) against the pinned oracle
CNS_CASES = [
    # N0, N1, C, Cd, K, M, h, options
    (32, 32, 1, 1, 3, 6, 5, {'MaxMainIter': 12, 'ZeroMean': True}),
    (32, 64, 3, 1, 2, 5, 4, {'MaxMainIter': 10, 'rho': 2.0,
                             'AutoRho': {'Enabled': True, 'Period': 3, 'AutoScaling': True, 'Scaling': 10.0}}),
    (64, 32, 3, 3, 2, 6, 5, {'MaxMainIter': 10, 'Y0': 'pcn', 'AutoRho': {'StdResiduals': True}}),
    (16, 17, 1, 1, 4, 4, 3, {'MaxMainIter': 8, 'rho': 0.5, 'RelaxParam': 1.0}),      # any-size transforms
    (32, 32, 3, 1, 2, 5, 4, {'MaxMainIter': 8, 'rho': 2.0, 'AuxVarObj': False, 'ZeroMean': True}),   # objective on the blocks
]


def run_cns_case(case, dt=np.float32):
    """ConvCnstrMOD_Consensus on the device against oracle/cbpdndl_oracle.ConsensusCCMOD (bit-identical to
    the reference): dictionary, residual / penalty / objective trajectories."""
    from oracle import cbpdndl_oracle as dlo
    from sporco_b200 import cnvrep as cr
    from sporco_b200.admm import ccmod
    N0, N1, C, Cd, K, M, h, o = case
    o = dict(o)
    rng = np.random.default_rng(11)
    Cx = C - Cd + 1
    Z = rng.standard_normal((N0, N1, Cx, K, M)).astype(dt)
    Z[np.abs(Z) < 1.0] = 0
    S = rng.standard_normal((N0, N1, C, K) if C > 1 else (N0, N1, K)).astype(dt)
    dsz = (h, h, M) if Cd == 1 else (h, h, Cd, M)
    if o.get('Y0') == 'pcn':
        D0 = rng.standard_normal(dsz).astype(dt)
        cri = cr.CDU_ConvRepIndexing(dsz, S, 1, 2)
        o['Y0'] = cr.zpad(cr.stdformD(cr.Pcn(D0, dsz, cri.Nv, 2, cri.dimCd, crp=True), cri.Cd, cri.M, 2), cri.Nv)
    c = ccmod.ConvCnstrMOD_Consensus(Z, S, dsz, ccmod.ConvCnstrMOD_Consensus.Options(o))
    Y = c.solve()
    r = dlo.ConsensusCCMOD(S, dsz, o)
    r.setcoef(Z)
    r.solve()
    its = c.getitstat()
    tol = 1e-9
    if dt == np.float32:
        # The reference's own float32 run drifts from its float64 run by 1e-4 ... 6e-4 on these problems (the
        # block solves amplify rounding); the device must agree with the float64 oracle on the same inputs much
        # better than that, and with the float32 oracle within that drift.
        o64 = dict(o)
        if o64.get('Y0') is not None:
            o64['Y0'] = np.asarray(o64['Y0'], dtype=np.float64)
        r64 = dlo.ConsensusCCMOD(S.astype(np.float64), dsz, o64)
        r64.setcoef(Z.astype(np.float64))
        r64.solve()
        drift = rel(r.Y, r64.Y)
        assert rel(Y, r64.Y) < 5e-5, rel(Y, r64.Y)
        tol = max(2e-4, 1.5 * drift)
    ref = np.array(r.itstat, dtype=np.float64)
    assert len(its.Iter) == len(ref)
    assert rel(Y, r.Y) < tol, rel(Y, r.Y)
    for name, col in (('DFid', 1), ('PrimalRsdl', 3), ('DualRsdl', 4), ('EpsPrimal', 5), ('EpsDual', 6), ('Rho', 7)):
        e = rel(getattr(its, name), ref[:, col])
        assert e < 5 * tol, (name, e)
    if o.get('AuxVarObj', True):
        assert np.all(np.asarray(its.Cnstr) < 1e-5)
    else:                                   # constraint violation of the block mean of X: not small
        assert rel(its.Cnstr, ref[:, 2]) < 5 * tol, rel(its.Cnstr, ref[:, 2])
    assert rel(c.getdict(), r.getdict()) < tol
    return c


def run_cns_golden(tag, sfx):
    """ConvCnstrMOD_Consensus on the fixture's inputs against the reference's own outputs."""
    from sporco_b200.admm import ccmod
    g = load('%s_%s' % (tag, sfx))
    o = dict(CNS_GOLDEN[tag])
    dsz = tuple(int(x) for x in g['dsz'])
    c = ccmod.ConvCnstrMOD_Consensus(g['Z'], g['S'], dsz, ccmod.ConvCnstrMOD_Consensus.Options(o))
    Y = c.solve()
    its = c.getitstat()
    # float32: the reference's float32 run is itself 3e-4 ... 6e-4 away from its float64 run here
    tol = 1e-9 if sfx == 'f64' else 1e-3
    assert Y.shape == g['Y'].shape and rel(Y, g['Y']) < tol, rel(Y, g['Y'])
    for name in ('DFid', 'PrimalRsdl', 'DualRsdl', 'EpsPrimal', 'EpsDual', 'Rho'):
        assert len(getattr(its, name)) == len(g[name])
        assert rel(getattr(its, name), g[name]) < 10 * tol, (name, rel(getattr(its, name), g[name]))
    return c


CNS_GOLDEN = {
    'cns_zm': {'MaxMainIter': 15, 'ZeroMean': True},
    'cns_arho': {'MaxMainIter': 15, 'rho': 2.0,
                 'AutoRho': {'Enabled': True, 'Period': 3, 'AutoScaling': True, 'Scaling': 10.0}},
}


def run_ccmod_bt(sfx):
    """pgm.ccmod.ConvCnstrMOD with BacktrackStandard against the reference's outputs: dictionary, L trajectory,
    backtracking counts, F, Q, residual, data fidelity."""
    from sporco_b200.pgm import ccmod
    from sporco_b200.pgm.backtrack import BacktrackStandard
    g = load('ccmod_bt_' + sfx)
    opt = ccmod.ConvCnstrMOD.Options({'MaxMainIter': 12, 'L': 2.0, 'Backtrack': BacktrackStandard(gamma_u=1.3, maxiter=10),
                                      'RelStopTol': 0.0, 'ZeroMean': True})
    c = ccmod.ConvCnstrMOD(g['Z'], g['S'], tuple(int(x) for x in g['dsz']), opt)
    c.solve()
    its = c.getitstat()
    tol = 1e-10 if sfx == 'f64' else 2e-5
    assert rel(c.getdict(), g['D']) < tol, rel(c.getdict(), g['D'])
    assert np.array_equal(np.asarray(its.IterBTrack, dtype=np.float64), g['IterBTrack'])
    for name in ('L', 'F_Btrack', 'Q_Btrack', 'Rsdl', 'DFid'):
        assert rel(getattr(its, name), g[name]) < 10 * tol, (name, rel(getattr(its, name), g[name]))
    return c


def run_gradreg_multichannel_dict(dt=np.float32):
    """ConvBPDNGradReg with a 3-channel dictionary (C x C solve with the diagonal mu w_m GHG + rho per frequency)
    against the oracle (pinned to the reference: solvemdbi_ism with the diagonal as rho)."""
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(4)
    D = rng.standard_normal((5, 5, 3, 6)).astype(dt)
    S = rng.standard_normal((32, 32, 3, 2)).astype(dt)
    gw = np.linspace(0.2, 2.0, 6).astype(dt)
    for extra in ({}, {'rho': 3.0, 'AutoRho': {'Enabled': False}, 'NonNegCoef': True}):
        opt = dict({'MaxMainIter': 15, 'RelStopTol': 0.0, 'GradWeight': gw}, **extra)
        b = cbpdn.ConvBPDNGradReg(D, S, 0.1, 0.4, cbpdn.ConvBPDNGradReg.Options(opt))
        Y = b.solve()
        r = orc.admm_convbpdn(D, S, 0.1, opt=opt, grad_mu=0.4)
        tol = 1e-9 if dt == np.float64 else 3e-4
        assert rel(Y, r.Y) < tol, rel(Y, r.Y)
        its = b.getitstat()
        assert rel(its.ObjFun, [x[1] for x in r.itstat]) < tol
        assert rel(its.RegGrad, [x[4] for x in r.itstat]) < tol
        assert rel(its.Rho, [x[9] for x in r.itstat]) < tol

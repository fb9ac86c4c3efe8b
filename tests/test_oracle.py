"""The oracle against its pins: the committed golden vectors (any box) and the live
reference (only where /root/reference exists, i.e. the build container)."""

import os
import sys
import warnings

import numpy as np
import pytest

from oracle import cbpdn_oracle as orc
from tests import cases

REF = '/root/reference'


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
@pytest.mark.parametrize('tag', sorted(cases.ADMM_CASES))
def test_oracle_reproduces_golden_admm(tag, sfx):
    g = cases.load('%s_%s' % (tag, sfx))
    opt, dimK, joint, _ = cases.ADMM_CASES[tag]
    enet, grd = joint == 'enet', joint == 'grd'
    joint = joint is True
    if grd:
        opt = cases.grd_opt(opt, g['D'].dtype)
    r = orc.admm_convbpdn(g['D'], g['S'], float(g['lmbda']),
                          mu=float(g['mu']) if joint else None, opt=opt, dimK=dimK,
                          enet_mu=float(g['mu']) if enet else None,
                          grad_mu=float(g['mu']) if grd else None)
    assert np.array_equal(r.Y, g['Y'])
    assert np.array_equal(r.U, g['U'])
    assert np.array_equal(r.X, g['X'])
    rho_col = 9 if (joint or enet or grd) else 8
    assert np.array_equal(np.array([row[rho_col] for row in r.itstat], dtype=np.float64), g['Rho'])
    assert np.array_equal(np.array([row[1] for row in r.itstat], dtype=np.float64), g['ObjFun'])


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
def test_oracle_reproduces_golden_pgm(sfx):
    g = cases.load('pgm_bt_' + sfx)
    r = orc.pgm_convbpdn(g['D'], g['S'], float(g['lmbda']), dimK=1,
                         opt={'MaxMainIter': 25, 'RelStopTol': 0.0, 'L': 10.0,
                              'Backtrack': {'gamma_u': 1.3, 'maxiter': 8}})
    assert np.array_equal(r.X, g['X'])
    assert np.array_equal(np.array([row[8] for row in r.itstat], dtype=np.float64), g['L'])
    assert np.array_equal(np.array([row[7] for row in r.itstat], dtype=np.float64), g['IterBTrack'])
    g = cases.load('pgm_fixed_' + sfx)
    r = orc.pgm_convbpdn(g['D'], g['S'], float(g['lmbda']), dimK=1,
                         opt={'MaxMainIter': 30, 'RelStopTol': 0.0, 'L': 400.0})
    assert np.array_equal(r.X, g['X'])
    assert np.array_equal(np.array([row[1] for row in r.itstat], dtype=np.float64), g['ObjFun'])


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
@pytest.mark.parametrize('tag', ['cdl', 'cdl_zm', 'cdl_clr1', 'cdl_clr3', 'cdl_ms'])
def test_oracle_reproduces_golden_dictionary_learning(tag, sfx):
    from oracle import cbpdndl_oracle as ocdl
    g = cases.load('%s_%s' % (tag, sfx))
    o, _, lmbda = cases.CDL_CASES[tag][:3]
    r = ocdl.cbpdndl(g['D0'], g['S'], lmbda, o)
    assert np.array_equal(r['D'], g['D'].squeeze())
    assert np.array_equal(r['X'], g['X'])
    for f in ('ObjFun', 'DFid', 'RegL1', 'Cnstr', 'XPrRsdl', 'XDlRsdl', 'XRho', 'D_L', 'D_Rsdl'):
        assert np.array_equal(r[f], g[f]), f


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
@pytest.mark.parametrize('tag', sorted(cases.AMS_CASES))
def test_oracle_reproduces_golden_addmasksim(tag, sfx):
    g = cases.load('%s_%s' % (tag, sfx))
    opt, dimK = cases.AMS_CASES[tag]
    gm = None
    if tag == 'ams_grd':
        gm = 0.4
        opt = dict(opt, GradWeight=np.concatenate((np.linspace(0.2, 2.0, 6), [0.0])).astype(g['D'].dtype))
    r = orc.admm_addmasksim(g['D'], g['S'], g['W'], float(g['lmbda']), opt=opt, dimK=dimK, grad_mu=gm)
    assert np.array_equal(r.Y, g['Y'])
    assert np.array_equal(np.array([row[1] for row in r.itstat], dtype=np.float64), g['ObjFun'])
    assert np.array_equal(np.array([row[9 if gm else 8] for row in r.itstat], dtype=np.float64), g['Rho'])


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
@pytest.mark.parametrize('tag', sorted(cases.CNS_GOLDEN))
def test_oracle_reproduces_golden_consensus_ccmod(tag, sfx):
    """oracle/cbpdndl_oracle.ConsensusCCMOD against the reference's outputs (bit for bit)."""
    from oracle import cbpdndl_oracle as ocdl
    g = cases.load('%s_%s' % (tag, sfx))
    r = ocdl.ConsensusCCMOD(g['S'], tuple(int(x) for x in g['dsz']), cases.CNS_GOLDEN[tag])
    r.setcoef(g['Z'])
    r.solve()
    assert np.array_equal(r.Y, g['Y'])
    ref = np.array(r.itstat, dtype=np.float64)
    for i, name in enumerate(('DFid', 'Cnstr', 'PrimalRsdl', 'DualRsdl', 'EpsPrimal', 'EpsDual', 'Rho')):
        assert np.array_equal(ref[:, i + 1], g[name]), name


def test_oracle_reproduces_golden_tikhonov():
    from oracle import signal_oracle as sorc
    g = cases.load('tikhonov')
    for sfx in ('f64', 'f32'):
        for i, (shape, lm, npd) in enumerate(cases.TIKHONOV_CASES):
            sl, sh = sorc.tikhonov_filter(g['s%d_%s' % (i, sfx)], lm, npd)
            assert np.array_equal(sl, g['sl%d_%s' % (i, sfx)]) and np.array_equal(sh, g['sh%d_%s' % (i, sfx)])


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
def test_oracle_reproduces_golden_pgm_mask(sfx):
    g = cases.load('pgm_mask_' + sfx)
    r = orc.pgm_convbpdn(g['D'], g['S'], float(g['lmbda']), dimK=1, W=g['W'],
                         opt={'MaxMainIter': 20, 'RelStopTol': 0.0, 'L': 5.0,
                              'Backtrack': {'gamma_u': 1.3, 'maxiter': 8}})
    assert np.array_equal(r.X, g['X'])
    assert np.array_equal(np.array([row[8] for row in r.itstat], dtype=np.float64), g['L'])


def test_oracle_level1_known_answers():
    g = cases.load('level1')
    assert np.array_equal(orc.solvedbi_sm(g['ah'], 0.7, g['b'], 4), g['x'])
    assert np.array_equal(orc.solvemdbi_ism(g['ah3'], 0.7, g['b3'], 4, 2), g['x3'])
    assert np.array_equal(orc.prox_l1(g['v'], 0.4), g['prox_l1'])
    assert np.array_equal(orc.prox_sl1l2(g['v'], 0.3, 0.25, axis=2), g['prox_sl1l2'])
    assert orc.rfl2norm2(g['xf'], g['xr'].shape, axis=(0, 1)) == float(g['rfl2norm2'])
    # algebraic pins used by the reference's own tests (tests/test_linalg.py:147-207)
    a = np.conj(g['ah'])
    lhs = a * orc.inner(g['ah'], g['x'], 4) + 0.7 * g['x']
    assert orc.rrs(lhs, g['b']) < 1e-11


def test_oracle_sharded_norms_match_unsharded():
    """K-sharding with summed squared norms follows the single-object rho schedule."""
    g = cases.load('admm_k3_f64')
    opt = {'MaxMainIter': 12, 'RelStopTol': 0.0}
    full = orc.admm_convbpdn(g['D'], g['S'], 0.1, opt=opt, dimK=1, norm_reduce=lambda v: v)
    ref = orc.admm_convbpdn(g['D'], g['S'], 0.1, opt=opt, dimK=1)
    assert cases.rel(full.Y, ref.Y) < 1e-10
    assert cases.rel([r[8] for r in full.itstat], [r[8] for r in ref.itstat]) < 1e-12


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present on this box')
def test_oracle_matches_live_reference():
    sys.path[:0] = [os.path.join(os.path.dirname(orc.__file__), 'shims'), REF]
    warnings.filterwarnings('ignore')
    from sporco.admm import cbpdn as rcbpdn
    rng = np.random.default_rng(99)
    D = rng.standard_normal((4, 4, 5)).astype(np.float32)
    S = rng.standard_normal((16, 16, 2)).astype(np.float32)
    opt = {'MaxMainIter': 15, 'RelStopTol': 0.0, 'L1Weight': np.linspace(0.5, 1.5, 5).astype(np.float32).reshape(1, 1, 1, 5)}
    b = rcbpdn.ConvBPDN(D, S, 0.05, rcbpdn.ConvBPDN.Options(opt), dimK=1)
    b.solve()
    r = orc.admm_convbpdn(D, S, 0.05, opt=opt, dimK=1)
    assert np.array_equal(b.Y, r.Y) and np.array_equal(b.U, r.U)

"""Host-side mirror of the reference interface: option trees, dimension inference, weights."""

import pickle

import numpy as np
import pytest

from sporco_b200 import cdict, cnvrep, util
from sporco_b200.admm import cbpdn, admm


def test_options_defaults_and_tuple_keys():
    o = cbpdn.ConvBPDN.Options({'MaxMainIter': 7, 'AutoRho': {'Period': 3}})
    assert o['MaxMainIter'] == 7 and o['AutoRho', 'Period'] == 3
    assert o['AutoRho', 'Enabled'] is True and o['RelaxParam'] == 1.8 and o['ReturnX'] is False
    assert o['L1Weight'] == 1.0
    o['AutoRho', 'Scaling'] = 5.0
    assert o['AutoRho']['Scaling'] == 5.0
    assert admm.ADMM.Options()['AutoRho', 'Enabled'] is False
    assert cbpdn.ConvBPDNJoint.Options()['L21Weight'] == 1.0


def test_options_reject_unknown_keys():
    with pytest.raises(cdict.UnknownKeyError):
        cbpdn.ConvBPDN.Options({'NoSuchKey': 1})
    with pytest.raises(cdict.UnknownKeyError):
        cbpdn.ConvBPDN.Options({'AutoRho': {'Bogus': 1}})
    with pytest.raises(cdict.InvalidValueError):
        cbpdn.ConvBPDN.Options({'AutoRho': 3})
    o = cbpdn.ConvBPDN.Options()
    with pytest.raises(cdict.UnknownKeyError):
        o['AutoRho', 'Nope']


def test_auxvarobj_couples_eval_flags():
    o = cbpdn.ConvBPDN.Options()
    o['AuxVarObj'] = True
    assert o['fEvalX'] is False and o['gEvalY'] is True
    o['AuxVarObj'] = False
    assert o['fEvalX'] is True and o['gEvalY'] is False


def test_options_pickle_roundtrip():
    o = cbpdn.ConvBPDN.Options({'MaxMainIter': 5, 'AutoRho': {'Period': 2}})
    p = pickle.loads(pickle.dumps(o))
    assert isinstance(p, cbpdn.ConvBPDN.Options) and p['AutoRho', 'Period'] == 2


@pytest.mark.parametrize('dshape,sshape,dimK,expect', [
    ((8, 8, 32), (64, 64), None, dict(C=1, Cd=1, K=1, M=32, shpX=(64, 64, 1, 1, 32))),
    ((8, 8, 32), (64, 64, 5), None, dict(C=1, Cd=1, K=5, M=32, shpX=(64, 64, 1, 5, 32))),
    ((8, 8, 32), (64, 64, 3), 0, dict(C=3, Cd=1, K=1, M=32, shpX=(64, 64, 3, 1, 32))),
    ((8, 8, 3, 32), (64, 64, 3), None, dict(C=3, Cd=3, K=1, M=32, shpX=(64, 64, 1, 1, 32))),
    ((8, 8, 32), (64, 64, 3, 4), None, dict(C=3, Cd=1, K=4, M=32, shpX=(64, 64, 3, 4, 32))),
    ((8, 8, 3, 32), (64, 64, 3, 4), None, dict(C=3, Cd=3, K=4, M=32, shpX=(64, 64, 1, 4, 32))),
])
def test_convrep_indexing(dshape, sshape, dimK, expect):
    cri = cnvrep.CSC_ConvRepIndexing(np.zeros(dshape), np.zeros(sshape), dimK=dimK)
    for k, v in expect.items():
        assert getattr(cri, k) == v
    assert cri.axisN == (0, 1) and (cri.axisC, cri.axisK, cri.axisM) == (2, 3, 4)


def test_convrep_channel_mismatch():
    with pytest.raises(ValueError):
        cnvrep.CSC_ConvRepIndexing(np.zeros((4, 4, 3, 5)), np.zeros((16, 16, 2, 1)))


def test_l1wshape():
    cri = cnvrep.CSC_ConvRepIndexing(np.zeros((4, 4, 5)), np.zeros((16, 16, 2)), dimK=1)
    assert cnvrep.l1Wshape(np.array(1.0), cri) == (1, 1, 1, 1, 1)
    assert cnvrep.l1Wshape(np.zeros((1, 1, 1, 5)), cri) == (1, 1, 1, 1, 5)
    assert cnvrep.l1Wshape(np.zeros((16, 16, 2)), cri) == (16, 16, 2, 1, 1)   # as the reference
    assert cnvrep.l1Wshape(np.zeros((16, 16, 1, 2, 5)), cri) == (16, 16, 1, 2, 5)


def test_timer():
    t = util.Timer(['a', 'b'])
    t.start('a')
    t.stop('a')
    assert t.elapsed('a') >= 0.0 and t.elapsed('b') == 0.0
    with pytest.raises(KeyError):
        t.stop('zzz')


def test_iterationstats_fields():
    assert cbpdn.ConvBPDN.IterationStats._fields == (
        'Iter', 'ObjFun', 'DFid', 'RegL1', 'PrimalRsdl', 'DualRsdl', 'EpsPrimal', 'EpsDual',
        'Rho', 'XSlvRelRes', 'Time')
    assert cbpdn.ConvBPDNJoint.IterationStats._fields[3:5] == ('RegL1', 'RegL21')


# ---- host helpers added with the dictionary-learning / sibling classes -------------------------
def test_cnvrep_dictionary_helpers_match_the_oracle():
    from oracle import cbpdndl_oracle as ocdl, cbpdn_oracle as orc
    from sporco_b200 import cnvrep as cr
    rng = np.random.default_rng(0)
    x = rng.standard_normal((16, 12, 3, 1, 4))
    for zm in (False, True):
        a = cr.Pcn(x, (5, 4, 3, 4), (16, 12), dimN=2, dimC=1, crp=False, zm=zm)
        assert np.allclose(a, ocdl.pcn(x, (5, 4, 3, 4), (16, 12), zm=zm), atol=1e-15)
        c = cr.Pcn(x, (5, 4, 3, 4), (16, 12), dimN=2, dimC=1, crp=True, zm=zm)
        assert c.shape == (5, 4, 3, 1, 4) and np.allclose(cr.zpad(c, (16, 12)), a)
    assert np.allclose(np.sum(a ** 2, (0, 1, 2)), 1.0)
    cri = cr.CDU_ConvRepIndexing((5, 4, 3, 4), np.zeros((16, 12, 3, 7)), dimK=1)
    assert (cri.Cd, cri.C, cri.Cx, cri.K, cri.M) == (3, 3, 1, 7, 4) and cri.shpD == (16, 12, 3, 1, 4)
    cri = cr.CDU_ConvRepIndexing((5, 4, 4), np.zeros((16, 12, 3, 7)), dimK=1)
    assert (cri.Cd, cri.C, cri.Cx) == (1, 3, 3) and cri.shpX == (16, 12, 3, 7, 4)
    # multi-scale specification: the largest support, the filters of all blocks, per-filter supports; projection
    # of every block over its own support, equal to the oracle's restatement
    ms = ((5, 5, 2), (3, 4, 3))
    cri = cr.CDU_ConvRepIndexing(ms, np.zeros((16, 12, 7)))
    assert cri.multiscale and cri.dsz == (5, 5, 5) and cri.M == 5 and cri.shpD == (16, 12, 1, 1, 5)
    assert cr.filter_supports(ms).tolist() == [[5, 5], [5, 5], [3, 4], [3, 4], [3, 4]]
    xm = rng.standard_normal((16, 12, 1, 1, 5))
    for zm in (False, True):
        pm = cr.Pcn(xm, ms, (16, 12), dimN=2, dimC=1, crp=False, zm=zm)
        assert np.allclose(pm, ocdl.pcn(xm, ms, (16, 12), zm=zm), atol=1e-15)
        assert not np.any(pm[3:, :, :, :, 2:]) and not np.any(pm[:, 4:, :, :, 2:]) and np.any(pm[3:5, :5, :, :, :2])
    with pytest.raises(NotImplementedError):
        cr.CDU_ConvRepIndexing((((5, 5, 1, 2), (5, 5, 2, 2)), (3, 3, 3, 2)), np.zeros((16, 12, 3, 7)))
    # mask shapes: same decisions as the oracle's restatement of cnvrep.mskWshape
    for S, W in ((np.zeros((8, 8)), np.zeros((8, 8))), (np.zeros((8, 8, 3)), np.zeros((8, 8, 3))),
                 (np.zeros((8, 8, 3, 2)), np.zeros((8, 8, 1, 2))), (np.zeros((8, 8, 3, 2)), np.zeros((8, 8, 3)))):
        for dimK in ((None,) if S.ndim != 3 else (0, 1)):
            c = cr.CSC_ConvRepIndexing(np.zeros((3, 3, 2)), S, dimK=dimK)
            d = orc.Dims(np.zeros((3, 3, 2)), S, dimK=dimK)
            assert cr.mskWshape(W, c) == orc.msk_shape(W, d)


def test_dictionary_learning_options_tree():
    from sporco_b200 import cdict
    from sporco_b200.dictlrn import cbpdndl
    from sporco_b200.admm import cbpdn
    from sporco_b200.pgm import ccmod
    o = cbpdndl.ConvBPDNDictLearn.Options({'CBPDN': {'AuxVarObj': True}, 'CCMOD': {'ZeroMean': True}})
    assert isinstance(o['CBPDN'], cbpdn.ConvBPDN.Options) and isinstance(o['CCMOD'], ccmod.ConvCnstrMOD.Options)
    assert o['CBPDN', 'MaxMainIter'] == 1 and o['CCMOD', 'MaxMainIter'] == 1
    assert o['CBPDN', 'AutoRho', 'Period'] == 10 and o['CBPDN', 'AutoRho', 'Enabled'] is True
    assert o['CBPDN', 'gEvalY'] is True and o['CBPDN', 'fEvalX'] is False       # AuxVarObj took effect
    with pytest.raises(cdict.UnknownKeyError):
        o['CCMOD', 'NoSuchKey'] = 1
    p = cbpdndl.ConvBPDNDictLearn.Options(xmethod='pgm')
    assert p.xmethod == 'pgm' and 'Backtrack' in p['CBPDN'] and 'AutoRho' not in p['CBPDN']
    with pytest.raises(NotImplementedError):
        cbpdndl.ccmod_class_label_lookup('ism')
    from sporco_b200.admm import ccmod as accmod
    assert cbpdndl.ccmod_class_label_lookup('cns') is accmod.ConvCnstrMOD_Consensus
    q = cbpdndl.ConvBPDNDictLearn.Options({'CCMOD': {'rho': 3.0}}, dmethod='cns')
    assert isinstance(q['CCMOD'], accmod.ConvCnstrMOD_Consensus.Options) and q['CCMOD', 'MaxMainIter'] == 1
    assert q['CCMOD', 'AutoRho', 'Period'] == 10 and q['CCMOD', 'AutoRho', 'Enabled'] is False
    assert q['CCMOD', 'RelaxParam'] == 1.8 and q['CCMOD', 'AuxVarObj'] is True and q['CCMOD', 'rho'] == 3.0
    from sporco_b200.dictlrn import common as dc
    assert dc.isfld('admm', 'pgm', o) == ['Iter', 'ObjFun', 'DFid', 'RegL1', 'Cnstr', 'XPrRsdl', 'XDlRsdl',
                                          'XRho', 'D_L', 'D_Rsdl', 'Time']
    assert dc.isfld('admm', 'cns', q) == ['Iter', 'ObjFun', 'DFid', 'RegL1', 'Cnstr', 'XPrRsdl', 'XDlRsdl',
                                          'XRho', 'DPrRsdl', 'DDlRsdl', 'DRho', 'Time']
    assert dc.isfld('pgm', 'pgm', p) == ['Iter', 'ObjFun', 'DFid', 'RegL1', 'Cnstr', 'X_L', 'X_Rsdl', 'D_L',
                                         'D_Rsdl', 'Time']


def test_momentum_rules_and_weight_extension():
    from sporco_b200.pgm.momentum import MomentumNesterov, MomentumLinear, MomentumGenLinear
    t = 1
    for _ in range(6):
        nxt = MomentumNesterov().update(t)
        assert nxt == 0.5 * float(1. + np.sqrt(1. + 4. * t ** 2))             # the reference's expression
        t = nxt
    assert MomentumLinear(2.).update(3) == 2.5 and MomentumGenLinear(50., 2.).update(4) == 27.0
    from sporco_cuda.cbpdn import _extend
    assert _extend(1.0, 4, 0.0).shape == ()
    w = _extend(np.arange(4, dtype=np.float32).reshape(1, 1, 4), 4, 0.0)
    assert w.shape == (1, 1, 5) and w[0, 0, -1] == 0.0 and w[0, 0, 2] == 2.0

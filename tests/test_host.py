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

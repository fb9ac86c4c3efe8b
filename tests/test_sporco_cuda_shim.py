"""The `sporco_cuda`-compatible package: import surface on any box; on the GPU box the
functional cbpdn() against the oracle, and -- where the reference tree exists -- the reference's
own `sporco.cuda` picking the backend up through its import seam."""

import os
import sys

import numpy as np
import pytest

from tests import cases


def test_import_surface():
    import sporco_cuda
    from sporco_cuda import util, cbpdn
    assert set(util.__all__) == {'device_count', 'current_device', 'memory_info', 'device_name'}
    for name in ('cbpdn', 'cbpdngrd', 'cbpdnmsk', 'cbpdngrdmsk'):
        assert callable(getattr(cbpdn, name))
    assert isinstance(util.device_count(), int)
    with pytest.raises(ValueError):
        cbpdn.cbpdngrd(np.zeros((4, 4)), np.zeros((8, 8)), 0.1, 0.1, {})


def _functional_variants():
    """cbpdngrd / cbpdnmsk / cbpdngrdmsk against the oracle (same weights convention as the
    original extension: per-filter arrays are given for the M filters of D only)."""
    from oracle import cbpdn_oracle as orc
    from sporco_cuda import cbpdn as cu
    rng = np.random.default_rng(4)
    M = 8
    D = rng.standard_normal((6, 6, M)).astype(np.float32)
    S = rng.standard_normal((64, 64)).astype(np.float32)
    W = (rng.random((64, 64)) > 0.25).astype(np.float32)
    gw = np.linspace(0.2, 1.5, M).astype(np.float32)
    opt = {'MaxMainIter': 15, 'RelStopTol': 0.0, 'rho': 4.0, 'AutoRho': {'Enabled': False}}
    X = cu.cbpdngrd(D, S, 0.1, 0.3, dict(opt, GradWeight=gw))
    r = orc.admm_convbpdn(D, S, 0.1, opt=dict(opt, GradWeight=gw), grad_mu=0.3)
    assert X.shape == (64, 64, M) and cases.rel(X, r.Y[:, :, 0, 0, :]) < 1e-4
    X = cu.cbpdnmsk(D, S, W, 0.1, opt)
    r = orc.admm_addmasksim(D, S, W, 0.1, opt=opt)
    assert X.shape == (64, 64, M) and cases.rel(X, r.Y[:, :, 0, 0, :M]) < 1e-4
    X = cu.cbpdngrdmsk(D, S, W, 0.1, 0.3, dict(opt, GradWeight=gw))
    r = orc.admm_addmasksim(D, S, W, 0.1, opt=dict(opt, GradWeight=np.concatenate((gw, [0.0])).astype(np.float32)),
                            grad_mu=0.3)
    assert X.shape == (64, 64, M) and cases.rel(X, r.Y[:, :, 0, 0, :M]) < 1e-4


def test_functional_variants_emulated(emu_library):
    from sporco_b200 import _lib
    _lib.use_library(emu_library)
    try:
        _functional_variants()
    finally:
        _lib.use_library(None)


@pytest.mark.gpu
def test_functional_variants_match_oracle():
    _functional_variants()


@pytest.mark.gpu
def test_functional_cbpdn_matches_oracle():
    from oracle import cbpdn_oracle as orc
    from sporco_cuda import cbpdn as cu, util
    assert util.device_count() >= 1 and len(util.device_name(0)) > 0
    free, total = util.memory_info()
    assert 0 < free <= total
    rng = np.random.default_rng(4)
    D = rng.standard_normal((8, 8, 16)).astype(np.float32)
    S = rng.standard_normal((128, 128)).astype(np.float32)
    opt = {'MaxMainIter': 20, 'RelStopTol': 0.0, 'AutoRho': {'Enabled': True}}
    X = cu.cbpdn(D, S, 0.1, opt, dev=0)
    r = orc.admm_convbpdn(D, S, 0.1, opt=opt)
    assert X.shape == (128, 128, 16) and X.dtype == np.float32
    assert cases.rel(X, r.Y[:, :, 0, 0, :]) < 3e-4


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference tree not present')
def test_reference_import_seam():
    sys.path[:0] = [os.path.join(os.path.dirname(os.path.dirname(__file__)), 'oracle', 'shims'),
                    '/root/reference']
    from sporco import cuda
    assert cuda.have_cuda and cuda.device_count() >= 1

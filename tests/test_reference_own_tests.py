"""Drop-in check in the build container: the REFERENCE'S OWN test files
(/root/reference/tests/admm/test_cbpdn.py, /root/reference/tests/pgm/test_cbpdn.py) executed
unmodified, with `sporco.admm.cbpdn.{GenericConvBPDN,ConvBPDN,ConvBPDNJoint,ConvElasticNet,ConvBPDNGradReg,AddMaskSim}` and
`sporco.pgm.cbpdn.ConvBPDN` replaced by the sporco_b200 classes (kernels run through the CPU
emulation harness here; the same replacement works on a GPU box where the reference is
installed).  Tests of other reference classes in those files are left alone; tests that need
features sporco_b200 does not implement are listed explicitly below -- nothing is skipped
silently.  Skipped entirely where /root/reference does not exist (the GPU box)."""

import inspect
import os
import sys
import types
import warnings

import pytest

from sporco_b200 import _lib

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')

OTHER_CLASSES = ('ConvBPDNProjL1', 'ConvMinL1InL2Ball',
                 'ConvBPDNMaskDcpl', 'ConvL1L1Grd', 'MultiDictConvBPDN',
                 'ConvTwoBlockCnstrnt')
# reference tests that exercise the replaced classes but need something not implemented
NOT_IMPLEMENTED = {
    'admm': {'test_10cplx': 'complex-valued data'},
    'pgm': {'test_10cplx': 'complex-valued data'},
    # tests/admm/test_ccmod.py with ConvCnstrMOD_Consensus and the factory functions replaced (default method 'cns')
    'ccmod': {'test_03cplx': 'complex-valued data',
              'test_13': 'multi-channel coefficient maps together with a multi-channel dictionary (runs in the '
                         'reference through numpy broadcasting only; not a documented configuration)'},
    # tests/pgm/test_ccmod.py with pgm.ccmod.ConvCnstrMOD replaced
    'pgmccmod': {'test_10': 'multi-channel coefficient maps together with a multi-channel dictionary (numpy broadcasting '
                            'only; not a documented configuration)',
                 'test_13': 'BacktrackRobust in the dictionary update',
                 'test_16': 'StepSizePolicyBB in the dictionary update',
                 'test_17': 'StepSizePolicyCauchy in the dictionary update',
                 'test_18': 'Monotone in the dictionary update'},
}
# tests of the reference's other dictionary-update classes / its own option classes in that file
CCMOD_OTHER = ('ConvCnstrMOD_IterSM', 'ConvCnstrMOD_CG', 'ConvCnstrMODBase')


def _load(kind):
    sys.path[:0] = [os.path.join(os.path.dirname(os.path.dirname(__file__)), 'oracle', 'shims'), REF]
    warnings.filterwarnings('ignore')
    import sporco.admm.cbpdn as ref_admm
    import sporco.pgm.cbpdn as ref_pgm
    from sporco_b200.admm import cbpdn as my_admm
    from sporco_b200.pgm import cbpdn as my_pgm
    proxy = types.ModuleType('cbpdn_proxy')
    if kind == 'ccmod':
        import sporco.admm.ccmod as ref_ccmod
        from sporco_b200.admm import ccmod as my_ccmod
        proxy.__dict__.update(ref_ccmod.__dict__)
        for name in ('ConvCnstrMOD_Consensus', 'ConvCnstrMOD', 'ConvCnstrMODOptions'):
            setattr(proxy, name, getattr(my_ccmod, name))
        path = os.path.join(REF, 'tests', 'admm', 'test_ccmod.py')
        src = open(path).read().replace('from sporco.admm import ccmod', 'ccmod = __proxy__')
        ns = {'__proxy__': proxy, '__name__': 'ref_tests_ccmod'}
        exec(compile(src, path, 'exec'), ns)
        return ns['TestSet01']
    if kind == 'pgmccmod':
        import sporco.pgm.ccmod as ref_pccmod
        from sporco_b200.pgm import ccmod as my_pccmod
        proxy.__dict__.update(ref_pccmod.__dict__)
        proxy.ConvCnstrMOD = my_pccmod.ConvCnstrMOD
        path = os.path.join(REF, 'tests', 'pgm', 'test_ccmod.py')
        src = open(path).read().replace('from sporco.pgm import ccmod', 'ccmod = __proxy__')
        src = src.replace('from sporco.pgm.momentum import', 'from sporco_b200.pgm.momentum import')
        src = src.replace('from sporco.pgm.backtrack import', 'from sporco_b200.pgm.backtrack import')
        src = src.replace('from sporco.pgm.stepsize import', 'from sporco_b200.pgm.stepsize import')
        ns = {'__proxy__': proxy, '__name__': 'ref_tests_pgmccmod'}
        exec(compile(src, path, 'exec'), ns)
        return ns['TestSet01']
    if kind == 'admm':
        proxy.__dict__.update(ref_admm.__dict__)
        for name in ('GenericConvBPDN', 'ConvBPDN', 'ConvBPDNJoint', 'ConvElasticNet', 'ConvBPDNGradReg',
                     'AddMaskSim'):
            setattr(proxy, name, getattr(my_admm, name))
        path = os.path.join(REF, 'tests', 'admm', 'test_cbpdn.py')
    else:
        proxy.__dict__.update(ref_pgm.__dict__)
        proxy.ConvBPDN = my_pgm.ConvBPDN
        proxy.ConvBPDNMask = my_pgm.ConvBPDNMask
        path = os.path.join(REF, 'tests', 'pgm', 'test_cbpdn.py')
    src = open(path).read()
    src = src.replace('from sporco.admm import cbpdn', 'cbpdn = __proxy__')
    src = src.replace('from sporco.pgm import cbpdn', 'cbpdn = __proxy__')
    if kind == 'pgm':
        # the reference's momentum / backtracking objects are plain parameter holders; the
        # sporco_b200 solver expects its own classes of the same names
        src = src.replace('from sporco.pgm.momentum import', 'from sporco_b200.pgm.momentum import')
        src = src.replace('from sporco.pgm.backtrack import', 'from sporco_b200.pgm.backtrack import')
        src = src.replace('from sporco.pgm.stepsize import', 'from sporco_b200.pgm.stepsize import')
    ns = {'__proxy__': proxy, '__name__': 'ref_tests_' + kind}
    exec(compile(src, path, 'exec'), ns)
    return ns['TestSet01']


def _cases(kind):
    if not os.path.isdir(REF):
        return []
    cls = _load(kind)
    out = []
    for name, fn in sorted(inspect.getmembers(cls, inspect.isfunction)):
        if not name.startswith('test_'):
            continue
        body = inspect.getsource(fn)
        if any(c in body for c in (CCMOD_OTHER if kind == 'ccmod' else (('ConvCnstrMODMask',) if kind == 'pgmccmod'
                                                                         else OTHER_CLASSES))):
            continue                                      # a test of another reference class
        out.append(name)
    return out


@pytest.fixture(autouse=True, scope='module')
def _use_emulated_kernels(emu_library):
    _lib.use_library(emu_library)
    yield
    _lib.use_library(None)


@pytest.mark.parametrize('name', _cases('admm'))
def test_reference_admm_suite(name):
    if name in NOT_IMPLEMENTED['admm']:
        pytest.xfail('not implemented: ' + NOT_IMPLEMENTED['admm'][name])
    cls = _load('admm')
    obj = cls()
    obj.setup_method(None)
    getattr(obj, name)()


# 2000-iteration recovery tests: ~1 min each under emulation; they pass (run them with
# SPCSC_LONG_TESTS=1) but are kept out of the default CPU suite to keep it short
LONG = {'pgm': ('test_10', 'test_11'), 'ccmod': ('test_03', 'test_04', 'test_05'),
        # default MaxMainIter (1000) or 3000-iteration recovery runs
        'pgmccmod': ('test_01', 'test_02', 'test_11', 'test_12', 'test_14', 'test_15')}       # up to 500 / 1000 / 1000 iterations, 4-8 min emulated each


@pytest.mark.parametrize('name', _cases('pgm'))
def test_reference_pgm_suite(name):
    if name in LONG['pgm'] and not os.environ.get('SPCSC_LONG_TESTS'):
        pytest.skip('long emulated run; set SPCSC_LONG_TESTS=1')
    if name in NOT_IMPLEMENTED['pgm']:
        pytest.xfail('not implemented: ' + NOT_IMPLEMENTED['pgm'][name])
    cls = _load('pgm')
    obj = cls()
    obj.setup_method(None)
    getattr(obj, name)()


@pytest.mark.parametrize('name', _cases('ccmod'))
def test_reference_ccmod_suite(name):
    """/root/reference/tests/admm/test_ccmod.py: the consensus dictionary update and the factory functions."""
    if name in LONG['ccmod'] and not os.environ.get('SPCSC_LONG_TESTS'):
        pytest.skip('long emulated run; set SPCSC_LONG_TESTS=1')
    if name in NOT_IMPLEMENTED['ccmod']:
        pytest.xfail('not implemented: ' + NOT_IMPLEMENTED['ccmod'][name])
    cls = _load('ccmod')
    obj = cls()
    obj.setup_method(None)
    getattr(obj, name)()


@pytest.mark.parametrize('name', _cases('pgmccmod'))
def test_reference_pgm_ccmod_suite(name):
    """/root/reference/tests/pgm/test_ccmod.py: the PGM dictionary update (fixed step and BacktrackStandard, linear /
    generalised-linear momentum, multi-scale and multi-channel dictionaries, DataType)."""
    if name in LONG['pgmccmod'] and not os.environ.get('SPCSC_LONG_TESTS'):
        pytest.skip('long emulated run; set SPCSC_LONG_TESTS=1')
    if name in NOT_IMPLEMENTED['pgmccmod']:
        pytest.xfail('not implemented: ' + NOT_IMPLEMENTED['pgmccmod'][name])
    cls = _load('pgmccmod')
    obj = cls()
    obj.setup_method(None)
    getattr(obj, name)()

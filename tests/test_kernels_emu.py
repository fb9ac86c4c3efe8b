"""Kernel logic on the CPU: the CUDA sources compiled for the emulation harness in
tests/emu, driven through the same C ABI and Python classes as on the GPU, and compared
with the reference's outputs.  This is how `-m "not gpu"` covers the kernels' index
arithmetic and control flow; performance and the real hardware are the GPU tests' job."""

import os

import numpy as np
import pytest

from sporco_b200 import _lib
from tests import cases


@pytest.fixture(autouse=True, scope='module')
def _use_emulated_kernels(emu_library):
    _lib.use_library(emu_library)
    yield
    _lib.use_library(None)


# Both precisions of every case run on the GPU (tests/test_parity_gpu.py).  Under emulation the float32 instantiations of
# the heavier families are opt-in (SPCSC_LONG_TESTS=1) so that the CPU suite stays short; the kernel source is the same template.
_BOTH = ['f64', 'f32'] if os.environ.get('SPCSC_LONG_TESTS') else ['f64']
_BOTH_DT = [np.float64, np.float32] if os.environ.get('SPCSC_LONG_TESTS') else [np.float64]


@pytest.mark.parametrize('dt', [np.float32, np.float64])
@pytest.mark.parametrize('shape', [(2, 4), (8, 8), (16, 64), (64, 32), (128, 256),
                                   # not powers of two: mixed-radix passes (radices 2..31), direct DFT for 37
                                   (36, 40), (45, 63), (34, 62), (37, 74), (17, 9)])
def test_rfft2_irfft2(shape, dt):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2,) + shape).astype(dt)
    xf = _lib.rfft2(x)
    ref = np.fft.rfftn(x.astype(np.float64), axes=(1, 2))
    eps = 2e-6 if dt == np.float32 else 1e-14
    assert cases.rel(xf, ref) < eps
    assert cases.rel(_lib.irfft2(xf, shape[1]), x) < eps


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
@pytest.mark.parametrize('tag', sorted(cases.ADMM_CASES))
def test_admm_golden(tag, sfx):
    cases.run_admm_case(tag, sfx)


def test_linsolve_check_and_weights():
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(5)
    D = rng.standard_normal((4, 4, 5))
    S = rng.standard_normal((16, 16, 2))
    w = np.linspace(0.5, 1.5, 5).reshape(1, 1, 1, 5)
    opt = cbpdn.ConvBPDN.Options({'MaxMainIter': 10, 'LinSolveCheck': True, 'L1Weight': w,
                                  'RelStopTol': 0.0})
    b = cbpdn.ConvBPDN(D, S, 0.05, opt, dimK=1)
    b.solve()
    assert b.getitstat().XSlvRelRes.max() < 1e-10
    from oracle import cbpdn_oracle as orc
    r = orc.admm_convbpdn(D, S, 0.05, dimK=1, opt={'MaxMainIter': 10, 'L1Weight': w,
                                                  'RelStopTol': 0.0})
    assert cases.rel(b.Y, r.Y) < 1e-9


def test_resume_and_pickle():
    import pickle
    from sporco_b200.admm import cbpdn
    g = cases.load('admm_k3_f64')
    opt = cbpdn.ConvBPDN.Options({'MaxMainIter': 15, 'RelStopTol': 0.0})
    b = cbpdn.ConvBPDN(g['D'], g['S'], 0.1, opt, dimK=1)
    b.solve()
    b2 = pickle.loads(pickle.dumps(b))
    b.solve()
    b2.solve()
    assert b.k == 30 and b2.k == 30
    assert cases.rel(b.Y, g['Y']) < 1e-9
    assert cases.rel(b2.Y, g['Y']) < 1e-9


@pytest.mark.parametrize('case', cases.FRESH_CASES)
def test_register_plan_kernels_vs_oracle(case):
    N0, N1, M, K, C, mu, extra = case
    cases.run_fresh_case(N0, N1, M, K, C=C, mu=mu, extra=extra)


@pytest.mark.parametrize('case', cases.FRESH_CASES_F64)
def test_register_plan_kernels_float64_vs_oracle(case):
    N0, N1, M, K, C, mu, extra = case
    b, _ = cases.run_fresh_case(N0, N1, M, K, C=C, mu=mu, extra=extra, dt=np.float64, tol=1e-9)
    info = b._h.admm_schedule_info()
    assert info['col_v2'] and info['prox_v2']


def test_general_kernels_on_the_same_problem(monkeypatch):
    monkeypatch.setenv('SPCSC_KERNELS', 'v1')
    cases.run_fresh_case(256, 64, 40, 2)


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
def test_pgm_golden(sfx):
    cases.run_pgm_cases(sfx)


def test_cross_iteration_fusion():
    cases.run_fusion_cases()


@pytest.mark.parametrize('dt', [np.float64, np.float32])
def test_aux_var_obj(dt):
    cases.run_auxvarobj_case(dt)


def test_bit_reproducible_runs():
    cases.run_reproducibility_case()


def test_multichannel_dictionary_fast_path():
    cases.run_multichannel_dict_cases()


def test_option_paths(capsys):
    cases.run_option_cases(capsys)


@pytest.mark.parametrize('tag,sfx', [(t, 'f64') for t in sorted(cases.CDL_CASES)] +
                         [(t, 'f32') for t in sorted(cases.CDL_CASES) if t in ('cdl', 'cdl_cns') or os.environ.get('SPCSC_LONG_TESTS')])
def test_dictionary_learning_golden(tag, sfx):
    cases.run_cdl_case(tag, sfx)


@pytest.mark.parametrize('dt', [np.float64, np.float32])
def test_ccmod_standalone(dt):
    cases.run_ccmod_standalone(dt)


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
@pytest.mark.parametrize('tag', sorted(cases.AMS_CASES))
def test_additive_mask_simulation_golden(tag, sfx):
    cases.run_ams_case(tag, sfx)


def test_tikhonov_filter_golden():
    cases.run_tikhonov_cases()


@pytest.mark.parametrize('sfx', _BOTH)
def test_pgm_mask_golden(sfx):
    cases.run_pgm_mask_case(sfx)
    cases.run_pgm_mask_case(sfx, 'pgm_mask_c3')          # multi-channel dictionary


@pytest.mark.parametrize('wave', ['1,1', '2,2', '2,1'] if os.environ.get('SPCSC_LONG_TESTS') else ['2,2'])
@pytest.mark.parametrize('keep', ['0', '1', 'fused'])
def test_wavefront_schedule_vs_oracle(wave, keep, monkeypatch):
    """The wavefront schedule (groups of images through an L2-sized scratch, SPCSC_WAVE=g,s): ragged
    last group, several streams, X retrievable only when the X spectra are kept."""
    monkeypatch.setenv('SPCSC_WAVE', wave)
    if keep == 'fused':       # groups through column + prox kernels only, cross-iteration fusion kept
        monkeypatch.setenv('SPCSC_WAVE_FUSED', '1')
        keep = '1'
    monkeypatch.setenv('SPCSC_WAVE_KEEP', keep)
    from sporco_b200 import _lib as L
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(7)
    D = rng.standard_normal((5, 5, 12)).astype(np.float32)
    S = rng.standard_normal((64, 64, 3)).astype(np.float32)
    o = {'MaxMainIter': 9, 'RelStopTol': 0.0}
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o), dimK=1)
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, opt=o, dimK=1)
    assert cases.rel(Y, r.Y) < 3e-4 and cases.rel(b.U, r.U) < 6e-4
    its = b.getitstat()
    assert cases.rel(its.Rho, [x[8] for x in r.itstat]) < 3e-4
    assert cases.rel(its.ObjFun, [x[1] for x in r.itstat]) < 3e-4
    if keep == '1':
        assert cases.rel(b.X, r.X) < 6e-4
    else:
        with pytest.raises(L.SpcscError):
            b.X


def test_setting_y_between_solves_is_seen_by_the_fused_schedule():
    """Warm start: `b.Y = ...` between two solve() calls.  With the cross-iteration fusion the next x-step
    would otherwise reuse the row spectra of the OLD Y - U that the last prox kernel wrote."""
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(11)
    D = rng.standard_normal((4, 4, 4)).astype(np.float32)
    S = rng.standard_normal((64, 64, 2)).astype(np.float32)
    opt = {'MaxMainIter': 4, 'RelStopTol': 0.0, 'rho': 3.0, 'AutoRho': {'Enabled': False}}
    sol = []
    Ynew = None
    for touch_u in (False, True):
        b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(opt), dimK=1)
        assert b._h.admm_schedule_info()['fused']
        b.solve()
        if Ynew is None:
            Ynew = (b.Y + rng.standard_normal(b.Y.shape)).astype(np.float32)
        b.Y = Ynew
        if touch_u:
            b.U = b.U.copy()          # setting U has always invalidated the spectra
        sol.append(b.solve().copy())
    assert np.abs(sol[0]).max() > 0
    assert np.array_equal(sol[0], sol[1])


def _column_variant_cases():
    """The default column kernel (col6 = k_col5 with one thread group) on every fresh case; the opt-in variants
    on the cases that exercise their cluster / ragged / multi-channel-signal paths (keeps the CPU suite short)."""
    every = cases.FRESH_CASES + [(64, 64, 8, 5, None, None, None)]
    few = [cases.FRESH_CASES[0], cases.FRESH_CASES[6], (64, 64, 8, 5, None, None, None)]
    out = [pytest.param(c, 'col6', id='col6-case%d' % i) for i, c in enumerate(every)]
    for v in (False, True, 'cpg1', 'col4', 'col5', 'col7'):
        sel = few if (v in (False, 'col4', 'col5') or os.environ.get('SPCSC_LONG_TESTS')) else few[:1]
        out += [pytest.param(c, v, id='%s-case%d' % (v, every.index(c))) for c in sel]
    return out


@pytest.mark.parametrize('case,pair', _column_variant_cases())
def test_push_exchange_column_kernel_vs_oracle(case, pair, monkeypatch):
    """k_col3 (SPCSC_COL3=1): persistent clusters over (frequency column, run of images) items, the
    per-frequency sums pushed into the peers' shared memory and awaited on an mbarrier."""
    monkeypatch.setenv('SPCSC_COL3', {False: '1', True: '2', 'cpg1': '3', 'col4': '4', 'col5': '5', 'col6': '6', 'col7': '7'}[pair])
    N0, N1, M, K, C, mu, extra = case
    b, _ = cases.run_fresh_case(N0, N1, M, K, C=C, mu=mu, extra=extra)
    want = {False: 3, True: 4, 'cpg1': 5, 'col4': 6, 'col5': 7, 'col6': 8, 'col7': 9}[pair]
    if pair == 'col5' and N0 > 256:
        want = 2        # 512-point columns do not fit two groups' stages: k_col2 takes over
    assert b._h.admm_schedule_info()['col_kernel'] == want


@pytest.mark.parametrize('case', cases.FRESH_CASES_F64[:2])
def test_staged_column_kernel_float64(case, monkeypatch):
    """k_col4 (SPCSC_COL3=4: slab, dictionary columns and signal row staged by bulk copies) in float64:
    8 elements per lane, 16 columns per CTA, clusters of up to 8."""
    monkeypatch.setenv('SPCSC_COL3', '4')
    N0, N1, M, K, C, mu, extra = case
    b, _ = cases.run_fresh_case(N0, N1, M, K, C=C, mu=mu, extra=extra, dt=np.float64, tol=1e-9)
    # 256-point float64 columns do not fit the staged layout (239 KB): k_col2 takes over
    assert b._h.admm_schedule_info()['col_kernel'] == (6 if N0 <= 128 else 2)


def test_push_exchange_column_kernel_float64_and_colour_dictionary(monkeypatch):
    monkeypatch.setenv('SPCSC_COL3', '1')
    N0, N1, M, K, C, mu, extra = cases.FRESH_CASES_F64[0]
    cases.run_fresh_case(N0, N1, M, K, C=C, mu=mu, extra=extra, dt=np.float64, tol=1e-9)
    # multi-channel dictionary (Woodbury solve), cluster of 2
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(9)
    D = rng.standard_normal((5, 5, 3, 40)).astype(np.float32)
    S = rng.standard_normal((64, 64, 3, 2)).astype(np.float32)
    o = {'MaxMainIter': 6, 'RelStopTol': 0.0}
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(o))
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, opt=o)
    assert cases.rel(Y, r.Y) < 3e-4
    assert cases.rel(b.getitstat().ObjFun, [x[1] for x in r.itstat]) < 3e-4


@pytest.mark.parametrize('sfx', _BOTH)
@pytest.mark.parametrize('name', cases.PGM_VARIANTS)
def test_pgm_step_size_policies_monotone_and_robust_backtracking(name, sfx):
    cases.run_pgm_variant_case(name, sfx)


def test_level1_entry_points():
    cases.run_level1_cases()


@pytest.mark.parametrize('dt', [np.float32, np.float64])
@pytest.mark.parametrize('case', cases.CNS_CASES)
def test_consensus_dictionary_update_vs_oracle(case, dt):
    cases.run_cns_case(case, dt)


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
@pytest.mark.parametrize('tag', sorted(cases.CNS_GOLDEN))
def test_consensus_dictionary_update_golden(tag, sfx):
    cases.run_cns_golden(tag, sfx)


@pytest.mark.parametrize('sfx', ['f64', 'f32'])
def test_dictionary_update_backtracking_golden(sfx):
    """pgm.ccmod.ConvCnstrMOD with BacktrackStandard (trial / accept on the device, F <= Q on the host)."""
    cases.run_ccmod_bt(sfx)


@pytest.mark.parametrize('dt', [np.float64, np.float32])
def test_gradient_regularisation_with_a_multichannel_dictionary(dt):
    cases.run_gradreg_multichannel_dict(dt)


def test_bench_mode_helper_runs_on_a_small_problem(monkeypatch):
    """bench.measure_modes (timing modes A and C of SURVEY.md section 8d) with the problem shrunk for the emulation."""
    import bench
    monkeypatch.setattr(bench, 'N0', 32)
    monkeypatch.setattr(bench, 'N1', 32)
    monkeypatch.setattr(bench, 'M', 4)
    monkeypatch.setattr(bench, 'HD', 5)
    monkeypatch.setattr(bench, 'K_PER_GPU', 2)
    r = bench.measure_modes(0, steps=4, warm=3)
    assert set(r) == {'A_default', 'C_fastsolve_fixed_rho'}
    for v in r.values():
        assert v['steps'] == 4 and v['it_per_s'] > 0

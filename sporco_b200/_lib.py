"""ctypes binding of libspcsc.so (the C ABI declared in include/spcsc.h).

The library is the product: if it is missing, or if it reports no CUDA device, everything
that needs it raises -- there is no CPU fallback in this package.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libspcsc.so')

F32, F64 = 0, 1
ARR_Y, ARR_U, ARR_X, ARR_XF, ARR_DF, ARR_SF, ARR_PGM_X, ARR_PGM_XF, ARR_PGM_YF = range(9)
PGM_FINISH_REJECT, PGM_FINISH_ROBUST = 1, 2
COEF_ADMM_Y, COEF_PGM_X = 0, 1
ERR_UNSUPPORTED = -4

# every symbol include/spcsc.h declares (tests check the list against the header)
SYMBOLS = (
    'spcsc_version', 'spcsc_device_count', 'spcsc_device_name', 'spcsc_memory_info',
    'spcsc_last_error', 'spcsc_create', 'spcsc_destroy', 'spcsc_synchronize',
    'spcsc_set_dict', 'spcsc_set_signal', 'spcsc_set_l1_weight', 'spcsc_set_l21_weight',
    'spcsc_admm_configure', 'spcsc_admm_reset', 'spcsc_admm_set_rho', 'spcsc_admm_set_iter',
    'spcsc_admm_iterate',
    'spcsc_admm_get_scalars', 'spcsc_admm_last_timing', 'spcsc_admm_profile', 'spcsc_admm_schedule_info', 'spcsc_get_array', 'spcsc_set_array', 'spcsc_reconstruct',
    'spcsc_rfft2', 'spcsc_irfft2', 'spcsc_solvedbi_sm', 'spcsc_prox_l1', 'spcsc_prox_sl1l2', 'spcsc_comm_unique_id', 'spcsc_comm_create',
    'spcsc_comm_destroy', 'spcsc_attach_comm', 'spcsc_host_alloc', 'spcsc_host_free',
    'spcsc_trim_pools', 'spcsc_pgm_configure', 'spcsc_pgm_reset', 'spcsc_pgm_trial',
    'spcsc_pgm_accept', 'spcsc_pgm_policy_stats', 'spcsc_pgm_combine_y', 'spcsc_pgm_finish', 'spcsc_set_gradreg', 'spcsc_tikhonov_filter', 'spcsc_pgm_set_mask', 'spcsc_p2p_export',
    'spcsc_p2p_attach', 'spcsc_ccmod_reset', 'spcsc_ccmod_setcoef_device', 'spcsc_ccmod_setcoef',
    'spcsc_ccmod_step', 'spcsc_ccmod_trial', 'spcsc_ccmod_accept', 'spcsc_ccmod_get_dict', 'spcsc_ccmod_push_dict',
    'spcsc_ccmod_cns_init', 'spcsc_ccmod_cns_step', 'spcsc_ccmod_cns_get', 'spcsc_ccmod_set_supports', 'spcsc_ccmod_get_spectrum',
)


class SpcscError(RuntimeError):
    """Error reported by libspcsc (status code in .status)."""

    def __init__(self, status, msg):
        RuntimeError.__init__(self, 'libspcsc error %d: %s' % (status, msg))
        self.status = status


class Problem(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ('N0', 'N1', 'C', 'Cd', 'K', 'M', 'hd', 'wd', 'dtype', 'device')]


class AdmmOpts(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ('lmbda', 'mu', 'rlx', 'abs_tol', 'rel_tol', 'ar_scaling', 'ar_rsdl_ratio',
                 'ar_rsdl_target')] + \
               [(n, ctypes.c_int32) for n in
                ('ar_enabled', 'ar_period', 'ar_autoscaling', 'ar_std_residuals', 'joint',
                 'nonneg', 'no_bndry_cross', 'fast_solve', 'aux_var_obj', 'linsolve_check')] + \
               [('l2_weight', ctypes.c_double), ('ams_maps', ctypes.c_int32),
                ('reserved_', ctypes.c_int32)]


class PgmOpts(ctypes.Structure):
    _fields_ = [('lmbda', ctypes.c_double), ('nonneg', ctypes.c_int32),
                ('no_bndry_cross', ctypes.c_int32)]


class ItStat(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ('iter', 'objfun', 'dfid', 'regl1', 'regl21', 'primal_rsdl', 'dual_rsdl',
                 'eps_primal', 'eps_dual', 'rho', 'xslv_relres', 'reserved')]


_lib = None


def _declare(lib):
    vp, i32, i64p = ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)
    lib.spcsc_version.restype = ctypes.c_int
    lib.spcsc_device_count.restype = ctypes.c_int
    lib.spcsc_device_name.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    lib.spcsc_memory_info.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_uint64),
                                      ctypes.POINTER(ctypes.c_uint64)]
    lib.spcsc_last_error.argtypes = [vp]
    lib.spcsc_last_error.restype = ctypes.c_char_p
    lib.spcsc_create.argtypes = [ctypes.POINTER(Problem), ctypes.POINTER(vp)]
    lib.spcsc_destroy.argtypes = [vp]
    lib.spcsc_synchronize.argtypes = [vp]
    lib.spcsc_set_dict.argtypes = [vp, vp]
    lib.spcsc_set_signal.argtypes = [vp, vp]
    lib.spcsc_set_l1_weight.argtypes = [vp, vp, i64p]
    lib.spcsc_set_l21_weight.argtypes = [vp, vp, i64p]
    lib.spcsc_admm_configure.argtypes = [vp, ctypes.POINTER(AdmmOpts)]
    lib.spcsc_admm_reset.argtypes = [vp, ctypes.c_double]
    lib.spcsc_admm_set_rho.argtypes = [vp, ctypes.c_double]
    lib.spcsc_admm_set_iter.argtypes = [vp, i32]
    lib.spcsc_admm_iterate.argtypes = [vp, i32, ctypes.POINTER(ItStat), ctypes.POINTER(i32),
                                       ctypes.POINTER(i32)]
    lib.spcsc_admm_get_scalars.argtypes = [vp, ctypes.POINTER(ctypes.c_double),
                                           ctypes.POINTER(i32)]
    lib.spcsc_admm_last_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_float),
                                           ctypes.POINTER(ctypes.c_int64)]
    lib.spcsc_admm_profile.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_float)]
    lib.spcsc_admm_schedule_info.argtypes = [vp, ctypes.POINTER(i32)]
    i64 = ctypes.c_int64
    lib.spcsc_solvedbi_sm.argtypes = [i32, i32, i64, i32, i32, i32, ctypes.c_double, vp, vp, vp]
    lib.spcsc_prox_l1.argtypes = [i32, i32, i64, ctypes.c_double, vp, vp, vp]
    lib.spcsc_prox_sl1l2.argtypes = [i32, i32, i64, i32, i64, ctypes.c_double, ctypes.c_double, vp, vp]
    lib.spcsc_get_array.argtypes = [vp, i32, vp]
    lib.spcsc_set_array.argtypes = [vp, i32, vp]
    lib.spcsc_reconstruct.argtypes = [vp, vp, vp]
    lib.spcsc_pgm_configure.argtypes = [vp, ctypes.POINTER(PgmOpts)]
    lib.spcsc_pgm_reset.argtypes = [vp, vp]
    lib.spcsc_pgm_trial.argtypes = [vp, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    lib.spcsc_pgm_accept.argtypes = [vp, ctypes.c_double]
    lib.spcsc_pgm_policy_stats.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double)]
    lib.spcsc_pgm_combine_y.argtypes = [vp, ctypes.c_double, ctypes.c_double, i32]
    lib.spcsc_pgm_finish.argtypes = [vp, i32, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    lib.spcsc_set_gradreg.argtypes = [vp, vp, vp]
    lib.spcsc_p2p_export.argtypes = [vp, vp]
    lib.spcsc_p2p_attach.argtypes = [vp, i32, i32, vp]
    lib.spcsc_pgm_set_mask.argtypes = [vp, vp, i64p]
    lib.spcsc_tikhonov_filter.argtypes = [i32, i32, i32, i32, i32, ctypes.c_double, i32, vp, vp, vp]
    lib.spcsc_ccmod_reset.argtypes = [vp, vp, i32]
    lib.spcsc_ccmod_setcoef_device.argtypes = [vp, i32]
    lib.spcsc_ccmod_setcoef.argtypes = [vp, vp]
    lib.spcsc_ccmod_step.argtypes = [vp, ctypes.c_double, ctypes.c_double, i32, ctypes.POINTER(ctypes.c_double)]
    lib.spcsc_ccmod_trial.argtypes = [vp, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    lib.spcsc_ccmod_accept.argtypes = [vp, ctypes.c_double, i32, ctypes.POINTER(ctypes.c_double)]
    lib.spcsc_ccmod_get_dict.argtypes = [vp, vp]
    lib.spcsc_ccmod_push_dict.argtypes = [vp]
    lib.spcsc_ccmod_cns_init.argtypes = [vp, ctypes.c_double, i32, ctypes.c_int64]
    lib.spcsc_ccmod_cns_get.argtypes = [vp, i32, vp]
    lib.spcsc_ccmod_set_supports.argtypes = [vp, vp]
    lib.spcsc_ccmod_get_spectrum.argtypes = [vp, i32, vp]
    lib.spcsc_ccmod_cns_step.argtypes = [vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, i32,
                                         ctypes.POINTER(ctypes.c_double)]
    lib.spcsc_comm_unique_id.argtypes = [ctypes.c_char_p, vp]
    lib.spcsc_comm_create.argtypes = [ctypes.c_char_p, vp, i32, i32, i32, ctypes.POINTER(vp)]
    lib.spcsc_comm_destroy.argtypes = [vp]
    lib.spcsc_attach_comm.argtypes = [vp, vp, ctypes.c_double]
    lib.spcsc_host_alloc.argtypes = [ctypes.c_uint64, ctypes.POINTER(vp)]
    lib.spcsc_host_free.argtypes = [vp]
    lib.spcsc_rfft2.argtypes = [i32, i32, i32, i32, i32, vp, vp]
    lib.spcsc_irfft2.argtypes = [i32, i32, i32, i32, i32, vp, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int and name not in ('spcsc_version', 'spcsc_device_count'):
            fn.restype = ctypes.c_int
    return lib


def load(path=None):
    """Load (once) and return the ctypes library.  Raises if the extension is not built."""
    global _lib
    if path is None and _lib is not None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            'libspcsc.so not found at %s -- build the CUDA extension first with '
            '`python -m sporco_b200.build` (nvcc, sm_100a).  sporco_b200 has no CPU path.' % p)
    lib = _declare(ctypes.CDLL(p))
    if path is None:
        _lib = lib
    return lib


def use_library(lib):
    """Install an already loaded library object as the process-wide binding.  Exists for
    the test-suite's kernel emulation harness; product code never calls it."""
    global _lib
    _lib = lib


def check(status, handle=None):
    if status != 0:
        msg = load().spcsc_last_error(handle)
        raise SpcscError(status, msg.decode() if msg else 'unknown error')


def dtype_code(dtype):
    dt = np.dtype(dtype)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise SpcscError(-4, 'unsupported data type %s (float32 and float64 only)' % dt)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def require_device():
    lib = load()
    n = lib.spcsc_device_count()
    if n <= 0:
        raise SpcscError(-2, 'no CUDA device visible: sporco_b200 runs on the GPU only')
    return n


class Handle(object):
    """Thin object wrapper over a spcsc_handle."""

    def __init__(self, N0, N1, C, Cd, K, M, hd, wd, dtype, device=0):
        self.lib = load()
        self.dtype = np.dtype(dtype)
        self.cdtype = np.dtype(np.complex64 if self.dtype == np.float32 else np.complex128)
        self.dims = dict(N0=N0, N1=N1, C=C, Cd=Cd, K=K, M=M, hd=hd, wd=wd)
        self.Cx = C - Cd + 1
        pb = Problem(N0, N1, C, Cd, K, M, hd, wd, dtype_code(dtype), device)
        h = ctypes.c_void_p()
        check(self.lib.spcsc_create(ctypes.byref(pb), ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, 'h', None):
            self.lib.spcsc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _c(self, status):
        check(status, self.h)

    def _host(self, a, shape=None, complex_=False):
        dt = self.cdtype if complex_ else self.dtype
        a = np.ascontiguousarray(a, dtype=dt)
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError('array has shape %s, expected %s' % (a.shape, tuple(shape)))
        return a

    # shapes in the reference's layout
    def xshape(self):
        d = self.dims
        return (d['N0'], d['N1'], self.Cx, d['K'], d['M'])

    def set_dict(self, D):
        d = self.dims
        D = self._host(D, (d['hd'], d['wd'], d['Cd'], d['M']))
        self._c(self.lib.spcsc_set_dict(self.h, _ptr(D)))

    def set_signal(self, S):
        d = self.dims
        S = self._host(S, (d['N0'], d['N1'], d['C'], d['K']))
        self._c(self.lib.spcsc_set_signal(self.h, _ptr(S)))

    def set_l1_weight(self, W):
        W = self._host(W)
        if W.ndim != 5:
            raise ValueError('l1 weight must be given in its 5-D internal shape')
        shp = (ctypes.c_int64 * 5)(*W.shape)
        self._c(self.lib.spcsc_set_l1_weight(self.h, _ptr(W), shp))

    def set_l21_weight(self, W):
        W = self._host(W)
        if W.ndim != 2:
            raise ValueError('l2,1 weight must be given with shape (K|1, M|1)')
        shp = (ctypes.c_int64 * 2)(*W.shape)
        self._c(self.lib.spcsc_set_l21_weight(self.h, _ptr(W), shp))

    def admm_configure(self, **kw):
        o = AdmmOpts()
        for k, v in kw.items():
            setattr(o, k, v)
        self._c(self.lib.spcsc_admm_configure(self.h, ctypes.byref(o)))

    def admm_reset(self, rho):
        self._c(self.lib.spcsc_admm_reset(self.h, float(rho)))

    def admm_set_rho(self, rho):
        self._c(self.lib.spcsc_admm_set_rho(self.h, float(rho)))

    def admm_iterate(self, n, want_rows=True):
        rows = (ItStat * n)() if want_rows else None
        nd = ctypes.c_int32(0)
        st = ctypes.c_int32(0)
        self._c(self.lib.spcsc_admm_iterate(self.h, n, rows, ctypes.byref(nd), ctypes.byref(st)))
        return (rows, nd.value, bool(st.value))

    def admm_scalars(self):
        rho = ctypes.c_double(0)
        k = ctypes.c_int32(0)
        self._c(self.lib.spcsc_admm_get_scalars(self.h, ctypes.byref(rho), ctypes.byref(k)))
        return rho.value, k.value

    def admm_last_timing(self):
        ms = ctypes.c_float(0)
        n = ctypes.c_int64(0)
        self._c(self.lib.spcsc_admm_last_timing(self.h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def admm_schedule_info(self):
        info = (ctypes.c_int32 * 8)()
        self._c(self.lib.spcsc_admm_schedule_info(self.h, info))
        return {'row_fwd_v2': bool(info[0]), 'col_v2': bool(info[1]), 'prox_v2': bool(info[2]),
                'fused': bool(info[3]), 'col_kernel': int(info[4]), 'wave_group': int(info[5]),
                'wave_streams': int(info[6])}

    def admm_profile(self, n):
        ms = (ctypes.c_float * 4)()
        self._c(self.lib.spcsc_admm_profile(self.h, n, ms))
        return [ms[i] for i in range(4)]

    def get_array(self, which):
        d = self.dims
        N1f = d['N1'] // 2 + 1
        if which in (ARR_Y, ARR_U, ARR_X, ARR_PGM_X):
            out = pinned_empty(self.xshape(), self.dtype)
        elif which in (ARR_XF, ARR_PGM_XF, ARR_PGM_YF):
            out = np.empty((d['N0'], N1f, self.Cx, d['K'], d['M']), dtype=self.cdtype)
        elif which == ARR_DF:
            out = np.empty((d['N0'], N1f, d['Cd'], 1, d['M']), dtype=self.cdtype)
        elif which == ARR_SF:
            out = np.empty((d['N0'], N1f, d['C'], d['K'], 1), dtype=self.cdtype)
        else:
            raise ValueError('unknown array id')
        self._c(self.lib.spcsc_get_array(self.h, which, _ptr(out)))
        return out

    def set_array(self, which, a):
        a = self._host(a, self.xshape())
        self._c(self.lib.spcsc_set_array(self.h, which, _ptr(a)))

    def reconstruct(self, X=None):
        d = self.dims
        out = np.empty((d['N0'], d['N1'], d['C'], d['K']), dtype=self.dtype)
        if X is None:
            self._c(self.lib.spcsc_reconstruct(self.h, None, _ptr(out)))
        else:
            X = self._host(X, self.xshape())
            self._c(self.lib.spcsc_reconstruct(self.h, _ptr(X), _ptr(out)))
        return out

    def synchronize(self):
        self._c(self.lib.spcsc_synchronize(self.h))

    def pgm_configure(self, lmbda, nonneg, no_bndry_cross):
        o = PgmOpts(float(lmbda), int(bool(nonneg)), int(bool(no_bndry_cross)))
        self._c(self.lib.spcsc_pgm_configure(self.h, ctypes.byref(o)))

    def pgm_reset(self, X0=None):
        if X0 is None:
            self._c(self.lib.spcsc_pgm_reset(self.h, None))
        else:
            X0 = self._host(X0, self.xshape())
            self._c(self.lib.spcsc_pgm_reset(self.h, _ptr(X0)))

    def pgm_trial(self, L):
        out = (ctypes.c_double * 8)()
        self._c(self.lib.spcsc_pgm_trial(self.h, float(L), out))
        return [out[i] for i in range(8)]

    def pgm_accept(self, coef):
        self._c(self.lib.spcsc_pgm_accept(self.h, float(coef)))

    def pgm_policy_stats(self, store=False):
        out = (ctypes.c_double * 8)()
        self._c(self.lib.spcsc_pgm_policy_stats(self.h, 1 if store else 0, out))
        return [out[i] for i in range(8)]

    def pgm_combine_y(self, a, b, save_prev):
        self._c(self.lib.spcsc_pgm_combine_y(self.h, float(a), float(b), 1 if save_prev else 0))

    def pgm_finish(self, mode, c0):
        out = (ctypes.c_double * 2)()
        self._c(self.lib.spcsc_pgm_finish(self.h, int(mode), float(c0), out))
        return out[0]

    def pgm_set_mask(self, W):
        if W is None:
            self._c(self.lib.spcsc_pgm_set_mask(self.h, None, None))
            return
        W = np.ascontiguousarray(W, dtype=self.dtype)
        if W.ndim != 4:
            raise ValueError('mask must be given as a 4-D (N0, N1, C, K) broadcastable array')
        shp = (ctypes.c_int64 * 4)(*W.shape)
        self._c(self.lib.spcsc_pgm_set_mask(self.h, _ptr(W), shp))

    def set_gradreg(self, ghg, wgrd):
        d = self.dims
        if ghg is None:
            self._c(self.lib.spcsc_set_gradreg(self.h, None, None))
            return
        ghg = self._host(ghg, (d['N0'], d['N1'] // 2 + 1))
        wgrd = self._host(wgrd, (d['M'],))
        self._c(self.lib.spcsc_set_gradreg(self.h, _ptr(ghg), _ptr(wgrd)))

    # ---- dictionary update
    def ccmod_reset(self, D0, zero_mean):
        d = self.dims
        D0 = self._host(D0, (d['hd'], d['wd'], d['Cd'], d['M']))
        self._c(self.lib.spcsc_ccmod_reset(self.h, _ptr(D0), int(bool(zero_mean))))

    def ccmod_setcoef_device(self, source):
        self._c(self.lib.spcsc_ccmod_setcoef_device(self.h, int(source)))

    def ccmod_setcoef(self, Z):
        Z = self._host(Z, self.xshape())
        self._c(self.lib.spcsc_ccmod_setcoef(self.h, _ptr(Z)))

    def ccmod_step(self, L, coef, flags=3):
        out = (ctypes.c_double * 4)()
        self._c(self.lib.spcsc_ccmod_step(self.h, float(L), float(coef), int(flags), out))
        return [out[i] for i in range(4)]

    def ccmod_get_dict(self):
        d = self.dims
        out = np.empty((d['hd'], d['wd'], d['Cd'], d['M']), dtype=self.dtype)
        self._c(self.lib.spcsc_ccmod_get_dict(self.h, _ptr(out)))
        return out

    def ccmod_push_dict(self):
        self._c(self.lib.spcsc_ccmod_push_dict(self.h))

    def ccmod_trial(self, L):
        out = (ctypes.c_double * 4)()
        self._c(self.lib.spcsc_ccmod_trial(self.h, float(L), out))
        return [float(x) for x in out]

    def ccmod_accept(self, coef, flags=3):
        out = (ctypes.c_double * 4)()
        self._c(self.lib.spcsc_ccmod_accept(self.h, float(coef), int(flags), out))
        return [float(x) for x in out]

    def ccmod_cns_init(self, rho, y0_given, nb_global=0):
        self._c(self.lib.spcsc_ccmod_cns_init(self.h, float(rho), 1 if y0_given else 0, int(nb_global)))

    def ccmod_get_spectrum(self, which):
        """Xf (which 0) or Yf (1) of the PGM dictionary update as (N0, N1f, Cd, 1, M)."""
        d = self.dims
        n1f = d['N1'] // 2 + 1
        out = np.empty((d['Cd'], n1f, d['M'], d['N0']), dtype=self.cdtype)
        self._c(self.lib.spcsc_ccmod_get_spectrum(self.h, int(which), out.ctypes.data_as(ctypes.c_void_p)))
        return np.ascontiguousarray(out.transpose(3, 1, 0, 2))[:, :, :, np.newaxis, :]

    def ccmod_set_supports(self, hw):
        """Per-filter supports (M, 2) of a multi-scale dictionary, or None for one support."""
        if hw is None:
            self._c(self.lib.spcsc_ccmod_set_supports(self.h, None))
            return
        a = np.ascontiguousarray(hw, dtype=np.int32)
        self._c(self.lib.spcsc_ccmod_set_supports(self.h, a.ctypes.data_as(ctypes.c_void_p)))

    def ccmod_cns_get(self, which, nb, m):
        """Block variables X (which 0) or U (1) of the consensus update in device order (nb, m, N0, N1)."""
        out = np.empty((nb, m, self.dims['N0'], self.dims['N1']), dtype=self.dtype)
        self._c(self.lib.spcsc_ccmod_cns_get(self.h, int(which), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def ccmod_cns_step(self, rho, udiv, rlx, flags=3):
        out = (ctypes.c_double * 8)()
        self._c(self.lib.spcsc_ccmod_cns_step(self.h, float(rho), float(udiv), float(rlx), int(flags), out))
        return [float(x) for x in out]

    def p2p_export(self):
        buf = ctypes.create_string_buffer(64)
        self._c(self.lib.spcsc_p2p_export(self.h, buf))
        return buf.raw

    def p2p_attach(self, rank, nranks, handles):
        """handles: nranks * 64 bytes.  Returns False when peer mapping is not possible here (the
        NCCL all-reduce then stays in use)."""
        rc = self.lib.spcsc_p2p_attach(self.h, int(rank), int(nranks), handles)
        if rc == ERR_UNSUPPORTED:
            return False
        self._c(rc)
        return True

    def p2p_detach(self):
        self._c(self.lib.spcsc_p2p_attach(self.h, 0, 0, b''))

    def attach_comm(self, comm, global_nx):
        self._comm = comm                      # keep the communicator alive
        self._c(self.lib.spcsc_attach_comm(self.h, comm.c if comm is not None else None,
                                           float(global_nx)))


class Comm(object):
    """An NCCL communicator owned by libspcsc (one per process group and device)."""

    def __init__(self, nccl_lib, uid, rank, nranks, device):
        self.lib = load()
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(uid))
        c = ctypes.c_void_p()
        check(self.lib.spcsc_comm_create(nccl_lib.encode(), buf, rank, nranks, device,
                                         ctypes.byref(c)))
        self.c = c
        self.rank, self.nranks = rank, nranks
        self.p2p = None        # None: peer blocks not exchanged yet; True / False: the group's decision

    def close(self):
        if getattr(self, 'c', None):
            self.lib.spcsc_comm_destroy(self.c)
            self.c = None


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked memory from the library's pool (returned to the pool
    when the array is garbage collected)."""
    import weakref
    lib = load()
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize
    p = ctypes.c_void_p()
    check(lib.spcsc_host_alloc(max(nbytes, 1), ctypes.byref(p)))
    raw = (ctypes.c_byte * max(nbytes, 1)).from_address(p.value)
    weakref.finalize(raw, lib.spcsc_host_free, ctypes.c_void_p(p.value))
    return np.frombuffer(raw, dtype=dt, count=int(np.prod(shape))).reshape(shape)


def rfft2(x, device=0):
    """rfftn over the last two axes of a (batch, N0, N1) real array, on the GPU."""
    lib = load()
    x = np.ascontiguousarray(x)
    b, n0, n1 = x.shape
    cdt = np.complex64 if x.dtype == np.float32 else np.complex128
    out = np.empty((b, n0, n1 // 2 + 1), dtype=cdt)
    check(lib.spcsc_rfft2(dtype_code(x.dtype), device, b, n0, n1, _ptr(x), _ptr(out)))
    return out


def solvedbi_sm(ah, rho, b, c=None, axis=4, device=0):
    """``sporco.linalg.solvedbi_sm`` (single-channel ``ah``) / ``solvemdbi_ism`` (``ah`` with several
    channels on axis 2) on the GPU: arrays in the reference's (N0, N1f, C, K, M) layout, the system
    solved along the last axis.  `c` (the reference's cached component) is accepted and ignored."""
    lib = load()
    if axis not in (4, -1) or ah.ndim != 5 or b.ndim != 5:
        raise ValueError('expected 5-d arrays (N0, N1f, C, K, M) and axis=4')
    cdt = np.complex64 if np.dtype(b.dtype) == np.complex64 else np.complex128
    n0, n1f, cd, ka, m = ah.shape
    nk = b.shape[3]
    if ka != 1 or b.shape[2] != 1 or b.shape[:2] != (n0, n1f) or b.shape[4] != m:
        raise ValueError('ah must be (N0, N1f, Cd, 1, M) and b (N0, N1f, 1, K, M)')
    ahc = np.ascontiguousarray(ah[:, :, :, 0, :], dtype=cdt)
    bc = np.ascontiguousarray(b[:, :, 0, :, :], dtype=cdt)
    out = np.empty_like(bc)
    code = dtype_code(np.float32 if cdt == np.complex64 else np.float64)
    check(lib.spcsc_solvedbi_sm(code, device, n0 * n1f, nk, cd, m, float(rho), _ptr(ahc), _ptr(bc), _ptr(out)))
    return out.reshape(b.shape)


def prox_l1(v, alpha, device=0):
    """``sporco.prox.prox_l1`` on the GPU; `alpha` a scalar or an array broadcastable to `v`."""
    lib = load()
    v = np.ascontiguousarray(v)
    out = np.empty_like(v)
    w = None
    a = alpha
    if np.ndim(alpha) > 0:
        w = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, dtype=v.dtype), v.shape))
        a = 1.0
    check(lib.spcsc_prox_l1(dtype_code(v.dtype), device, v.size, float(a), None if w is None else _ptr(w),
                            _ptr(v), _ptr(out)))
    return out


def prox_sl1l2(v, alpha, beta, axis=None, device=0):
    """``sporco.prox.prox_sl1l2`` on the GPU, the l2 norm taken along one `axis` (all axes if None)."""
    lib = load()
    v = np.ascontiguousarray(v)
    if axis is None:
        outer, c, inner = 1, v.size, 1
    else:
        axis = axis % v.ndim
        outer = int(np.prod(v.shape[:axis], dtype=np.int64))
        c = v.shape[axis]
        inner = int(np.prod(v.shape[axis + 1:], dtype=np.int64))
    out = np.empty_like(v)
    check(lib.spcsc_prox_sl1l2(dtype_code(v.dtype), device, outer, c, inner, float(alpha), float(beta),
                               _ptr(v), _ptr(out)))
    return out


def tikhonov_filter(x, lmbda, npd, device=0):
    """Lowpass / highpass split of a (batch, N0, N1) real array on the GPU."""
    lib = load()
    x = np.ascontiguousarray(x)
    b, n0, n1 = x.shape
    sl = np.empty_like(x)
    sh = np.empty_like(x)
    check(lib.spcsc_tikhonov_filter(dtype_code(x.dtype), device, b, n0, n1, float(lmbda), int(npd),
                                    _ptr(x), _ptr(sl), _ptr(sh)))
    return sl, sh


def irfft2(xf, n1, device=0):
    """irfftn over the last two axes of a (batch, N0, N1/2+1) complex array, on the GPU."""
    lib = load()
    xf = np.ascontiguousarray(xf)
    b, n0, n1f = xf.shape
    if n1f != n1 // 2 + 1:
        raise ValueError('inconsistent last-axis length')
    rdt = np.float32 if xf.dtype == np.complex64 else np.float64
    out = np.empty((b, n0, n1), dtype=rdt)
    check(lib.spcsc_irfft2(dtype_code(rdt), device, b, n0, n1, _ptr(xf), _ptr(out)))
    return out


def nccl_library_path():
    """Path of the libnccl this process has loaded (torch's bundled copy when torch is
    imported), else the bare soname for the dynamic loader to resolve."""
    try:
        with open('/proc/self/maps') as f:
            for line in f:
                if 'libnccl' in line and '.so' in line:
                    return line.split()[-1]
    except OSError:
        pass
    return 'libnccl.so.2'


def comm_unique_id(nccl_lib):
    lib = load()
    buf = (ctypes.c_char * 128)()
    check(lib.spcsc_comm_unique_id(nccl_lib.encode(), buf))
    return bytes(buf.raw)

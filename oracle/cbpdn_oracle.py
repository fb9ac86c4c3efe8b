"""CPU oracle for the ConvBPDN hot path -- TEST INFRASTRUCTURE, not product code.

This file restates, in plain numpy and in the reference's own array layout
``(N0, N1, C, K, M)``, the algorithms of the ConvBPDN ADMM loop, its joint-sparsity
variant and the PGM/FISTA sibling of bwohlberg/sporco.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``; nothing under ``sporco_b200/`` may import it.

Parity pin: ``oracle/make_golden.py`` runs the real reference (imported read-only from
/root/reference with the stubs in ``oracle/shims``) next to this restatement and checks
bit-level agreement before writing ``tests/golden/*.npz``; ``tests/test_oracle.py``
re-checks the restatement against those committed fixtures on any box, and against the
live reference whenever /root/reference is present.

Reference anchors (file:line under /root/reference/sporco):
  dimension inference      cnvrep.py:33-198, l1Wshape cnvrep.py:492-550
  Sherman-Morrison solves  linalg.py:232-297 (rank one), linalg.py:370-444 (iterated)
  proximal operators       prox/_lp.py:144-183, prox/_lp.py:252-290, prox/_l21.py:51-88
  half-spectrum norm       fft.py:449-484
  ADMM loop                admm/admm.py:293-389, 434-486, 549-575, 877-885, 959-983
  ConvBPDN pieces          admm/cbpdn.py:242-344, 563-630, 785-807
  PGM loop                 pgm/pgm.py:284-370, 779-894; pgm/cbpdn.py:147-370
  backtracking / momentum  pgm/backtrack.py:74-107, pgm/momentum.py:45-48
"""

import time

import numpy as np
import numpy.fft as npfft

try:                                    # optional multi-threaded FFT (FFTW stand-in)
    import scipy.fft as spfft
except Exception:                       # pragma: no cover
    spfft = None


# ----------------------------------------------------------------------------------
# FFT backend.  'numpy' reproduces what the reference does when pyfftw is absent
# (fft.py:631-639: numpy.fft + astype); 'scipy' uses scipy.fft with worker threads as a
# stand-in for the reference's multi-threaded pyfftw path (fft.py:257-314).
# ----------------------------------------------------------------------------------
class FFTBackend(object):
    def __init__(self, kind='numpy', workers=1):
        if kind == 'scipy' and spfft is None:
            raise RuntimeError('scipy.fft not available')
        self.kind = kind
        self.workers = workers

    def rfftn(self, a, s, axes):
        cdt = np.complex64 if a.dtype == np.float32 else np.complex128
        if self.kind == 'numpy':
            return npfft.rfftn(a, s, axes).astype(cdt)
        return spfft.rfftn(a, s, axes, workers=self.workers).astype(cdt, copy=False)

    def irfftn(self, a, s, axes):
        rdt = np.float32 if a.dtype == np.complex64 else np.float64
        if self.kind == 'numpy':
            return npfft.irfftn(a, s, axes).astype(rdt)
        return spfft.irfftn(a, s, axes, workers=self.workers).astype(rdt, copy=False)


# ----------------------------------------------------------------------------------
# Dimension inference (cnvrep.py:33-198)
# ----------------------------------------------------------------------------------
class Dims(object):
    """Problem dimensions for a 2-D (dimN=2) convolutional representation."""

    def __init__(self, D, S, dimK=None, dimN=2):
        self.dimN = dimN
        self.dimCd = D.ndim - (dimN + 1)
        self.Cd = 1 if self.dimCd == 0 else D.shape[-2]
        if dimK is None:
            extra = S.ndim - dimN
            if extra == 0:
                dimC, dimK = 0, 0
            elif extra == 1:
                dimC = self.dimCd
                dimK = S.ndim - dimN - dimC
            else:
                dimC, dimK = 1, 1
        else:
            dimC = S.ndim - dimN - dimK
        self.dimC, self.dimK = dimC, dimK
        self.C = S.shape[dimN] if dimC == 1 else 1
        if self.Cd > 1 and self.C != self.Cd:
            raise ValueError('Multi-channel dictionary with signal with mismatched '
                             'number of channels (Cd=%d, C=%d)' % (self.Cd, self.C))
        self.Cx = self.C - self.Cd + 1
        self.K = S.shape[dimN + dimC] if dimK == 1 else 1
        self.M = D.shape[-1]
        self.Nv = tuple(S.shape[0:dimN])
        self.N = int(np.prod(self.Nv))
        self.axisN = tuple(range(dimN))
        self.axisC, self.axisK, self.axisM = dimN, dimN + 1, dimN + 2
        self.shpD = tuple(D.shape[0:dimN]) + (self.Cd, 1, self.M)
        self.shpS = self.Nv + (self.C, self.K, 1)
        self.shpX = self.Nv + (self.Cx, self.K, self.M)


def l1_weight_shape(W, dims):
    """Internal broadcast shape for an l1 weight array (cnvrep.py:492-550)."""
    sdim = dims.dimN + dims.dimC + dims.dimK
    if W.ndim < sdim:
        if W.size != 1:
            raise ValueError('weight array must be scalar or have at least the same '
                             'number of dimensions as input array')
        return (1,) * (dims.dimN + 3)
    if W.ndim == sdim:
        return W.shape + (1,) * (3 - dims.dimC - dims.dimK)
    if W.ndim == dims.dimN + 3:
        return W.shape
    return W.shape[0:-1] + (1,) * (2 - dims.dimC - dims.dimK) + W.shape[-1:]


# ----------------------------------------------------------------------------------
# Level-1 pieces
# ----------------------------------------------------------------------------------
def inner(x, y, axis):
    """sum(x*y) along `axis`, keepdims (linalg.py:41-88; einsum over the moved axis)."""
    xr = np.moveaxis(x, axis, 0)
    yr = np.moveaxis(y, axis, 0)
    ip = np.einsum(xr, [0, Ellipsis], yr, [0, Ellipsis])[np.newaxis, ...]
    return np.moveaxis(ip, 0, axis)


def solvedbi_sm(ah, rho, b, axis=4):
    """(rho I + a a^H) x = b per frequency (linalg.py:232-297)."""
    a = np.conj(ah)
    c = ah / (inner(ah, a, axis) + rho)
    return (b - (a * inner(c, b, axis))) / rho


def solvedbd_sm(ah, d, b, axis=4):
    """(diag(d) + a a^H) x = b per frequency (linalg.py:301-366)."""
    a = np.conj(ah)
    c = (ah / d) / (inner(ah, (a / d), axis) + 1.0)
    return (b - (a * inner(c, b, axis))) / d


def gradient_filters(ndim, axes, axshp, dtype, fft):
    """DFTs of the forward-difference filters and the sum of their squared magnitudes
    (signal.py:196-240)."""
    g = np.zeros([2 if k in axes else 1 for k in range(ndim)] + [len(axes), ], dtype)
    for k in axes:
        g[(0,) * k + (slice(None),) + (0,) * (g.ndim - 2 - k) + (k,)] = np.array([1, -1])
    Gf = fft.rfftn(g, axshp, axes)
    GHGf = np.sum(np.conj(Gf) * Gf, axis=-1).real
    return Gf, GHGf


def solvemdbi_ism(ah, rho, b, axisM, axisK):
    """(rho I + sum_k a_k a_k^H) x = b by iterated rank-one updates (linalg.py:370-444)."""
    nk = ah.shape[axisK]
    a = np.conj(ah)
    gamma = np.zeros(a.shape, a.dtype)
    dshape = list(a.shape)
    dshape[axisM] = 1
    delta = np.zeros(dshape, a.dtype)
    pre = (slice(None),) * axisK
    alpha = np.take(a, [0], axisK) / rho
    beta = b / rho
    for k in range(nk):
        sk = pre + (slice(k, k + 1),)
        gamma[sk] = alpha
        delta[sk] = 1.0 + inner(ah[sk], gamma[sk], axisM)
        beta = beta - (gamma[sk] * inner(ah[sk], beta, axisM)) / delta[sk]
        if k < nk - 1:
            alpha = np.take(a, [k + 1], axisK) / rho
            for l in range(k + 1):
                sl_ = pre + (slice(l, l + 1),)
                alpha = alpha - (gamma[sl_] * inner(ah[sl_], alpha, axisM)) / delta[sl_]
    return beta


def prox_l1(v, alpha):
    """Soft threshold (prox/_lp.py:144-183, real branch without numexpr)."""
    return np.sign(v) * (np.clip(np.abs(v) - alpha, 0, float('Inf')))


def prox_l2(v, alpha, axis=None):
    """Vector shrinkage with 0/0 := 0 (prox/_lp.py:252-290, array.py:119-136)."""
    a = np.sqrt(np.sum(v ** 2, axis=axis, keepdims=True))
    b = np.maximum(0, a - alpha)
    b = np.divide(b, a, out=np.zeros_like(b), where=(a != 0))
    return np.asarray(b * v, dtype=v.dtype)


def prox_sl1l2(v, alpha, beta, axis=None):
    """prox of alpha*l1 + beta*l2 (prox/_l21.py:51-88)."""
    return prox_l2(prox_l1(v, alpha), beta, axis)


def rfl2norm2(xf, xs, axis=(0, 1)):
    """Squared l2 norm of the spatial array from its rfftn (fft.py:449-484)."""
    scl = 1.0 / np.prod(np.array([xs[k] for k in axis]))
    pre = (slice(None),) * axis[-1]
    n0 = np.linalg.norm(xf[pre + (0,)])
    i1 = (xs[axis[-1]] + 1) // 2
    n1 = np.linalg.norm(xf[pre + (slice(1, i1),)])
    n2 = np.linalg.norm(xf[pre + (slice(-1, None),)]) if xs[axis[-1]] % 2 == 0 else 0.0
    return scl * (n0 ** 2 + 2.0 * n1 ** 2 + n2 ** 2)


def rrs(ax, b):
    """Relative residual (linalg.py:883-910)."""
    nrm = max(np.linalg.norm(ax.ravel()), np.linalg.norm(b.ravel()))
    return 0.0 if nrm == 0.0 else np.linalg.norm((ax - b).ravel()) / nrm


def _rdt(dtype):
    return np.dtype(np.float32) if np.dtype(dtype) in (np.dtype(np.float32),
                                                       np.dtype(np.complex64)) \
        else np.dtype(np.float64)


# ----------------------------------------------------------------------------------
# ADMM ConvBPDN / ConvBPDNJoint
# ----------------------------------------------------------------------------------
ADMM_DEFAULTS = {
    # admm/admm.py:148-161 merged with admm/cbpdn.py:127-134
    'MaxMainIter': 1000, 'AbsStopTol': 0.0, 'RelStopTol': 1e-3, 'RelaxParam': 1.8,
    'rho': None, 'FastSolve': False, 'DataType': None, 'AuxVarObj': False,
    'LinSolveCheck': False, 'NonNegCoef': False, 'NoBndryCross': False,
    'L1Weight': 1.0, 'L21Weight': 1.0, 'GradWeight': 1.0, 'Y0': None, 'U0': None,
    'AutoRho': {'Enabled': True, 'Period': 1, 'Scaling': 1000.0, 'RsdlRatio': 1.2,
                'RsdlTarget': None, 'AutoScaling': True, 'StdResiduals': False},
}


def _merge(defaults, opt):
    out = {}
    for k, v in defaults.items():
        out[k] = dict(v) if isinstance(v, dict) else v
    for k, v in (opt or {}).items():
        if k not in out:
            raise KeyError('unknown option %r' % (k,))
        if isinstance(out[k], dict):
            for kk, vv in v.items():
                if kk not in out[k]:
                    raise KeyError('unknown option %r' % ((k, kk),))
                out[k][kk] = vv
        else:
            out[k] = v
    return out


class ADMMResult(object):
    """State after `solve`; attribute names follow the reference object's."""
    pass


def msk_shape(W, dims):
    """Internal 5-D shape of a data-fidelity mask given in external form (cnvrep.py:553-605)."""
    ck = W.ndim - 2
    if ck >= 2:
        return W.shape + (1,) if ck == 2 else W.shape
    if ck == 1:
        if dims.C == 1 and dims.K > 1:
            return W.shape[0:2] + (1, W.shape[2]) + (1,)
        return W.shape[0:2] + (W.shape[2], 1) + (1,)
    return W.shape + (1,) * 3


def admm_addmasksim(D, S, W, lmbda=None, opt=None, dimK=None, fft=None, grad_mu=None):
    """AddMaskSim(ConvBPDN, D, S, W, lmbda, opt) (admm/cbpdn.py:2287-2485): impulse filters are appended
    (one for a single-channel dictionary, one per channel -- each non-zero in its own channel -- for a
    multi-channel one, :2337-2345), their coefficient maps are set to AX + U off the mask and to zero on it,
    and are left out of the regulariser.  A mask with a channel axis moves that axis onto the filter axis of
    the impulse maps (:2361-2362).  Returns the ADMMResult of the inner solver (all M + Cd maps)."""
    dims = Dims(D, S, dimK=dimK)
    if dims.Cd == 1:
        imp = np.zeros(D.shape[0:2] + (1,), dtype=D.dtype)
        imp[0, 0] = 1.0
    else:
        imp = np.zeros(D.shape[0:2] + (dims.Cd,) * 2, dtype=D.dtype)
        for c in range(dims.Cd):
            imp[0, 0, c, c] = 1.0
    Di = np.concatenate((D, imp), axis=D.ndim - 1)
    dtype = np.dtype(S.dtype) if (opt or {}).get('DataType') is None else np.dtype(opt['DataType'])
    W5 = np.asarray(W.reshape(msk_shape(W, dims)), dtype=dtype)
    if dims.Cd > 1 and W5.shape[2] > 1:
        W5 = np.swapaxes(W5, dims.axisC, dims.axisM)
    return admm_convbpdn(Di, S, lmbda, opt=opt, dimK=dimK, fft=fft, ams=(W5, dims.Cd), grad_mu=grad_mu)


def admm_convbpdn(D, S, lmbda=None, mu=None, opt=None, dimK=None, fft=None,
                  norm_reduce=None, record=False, timing=None, enet_mu=None, ams=None,
                  grad_mu=None):
    """Run the ConvBPDN (mu is None) or ConvBPDNJoint (mu given) ADMM loop; with `enet_mu` the
    ConvElasticNet variant (admm/cbpdn.py:810-990: x-step with mu + rho on the diagonal, extra
    (mu/2)||x||^2 term; rows then carry RegL2 where the joint solver has RegL21).
    `grad_mu`: ConvBPDNGradReg (admm/cbpdn.py:993-1206; option GradWeight; rows carry RegGrad).
    `ams` = (W5, Cd): the additive-mask hook of AddMaskSim (admm/cbpdn.py:2377-2409) for a
    dictionary whose last Cd filters are the appended impulses; see :func:`admm_addmasksim`.

    `norm_reduce`, if given, maps a float64 vector of local sums to global sums; it is
    how the K-sharded multi-rank form of the algorithm is expressed (every rank then
    takes the same rho decision).  With `norm_reduce=None` all norms are taken exactly
    as the reference takes them (np.linalg.norm on the full arrays).
    """
    fft = fft or FFTBackend()
    o = _merge(ADMM_DEFAULTS, opt)
    ar = o['AutoRho']
    dims = Dims(D, S, dimK=dimK)
    dtype = np.dtype(o['DataType']) if o['DataType'] is not None else np.dtype(S.dtype)
    rdt = _rdt(dtype)
    axN, axC, axM = dims.axisN, dims.axisC, dims.axisM

    # scalars carried in the working precision (admm/admm.py:243-254)
    rho = rdt.type(o['rho']) if o['rho'] is not None else rdt.type(1.0)
    tau = rdt.type(ar['Scaling'])
    mur = rdt.type(ar['RsdlRatio'])
    rlx = rdt.type(o['RelaxParam'])

    Y = np.zeros(dims.shpX, dtype) if o['Y0'] is None else o['Y0'].astype(dtype, copy=True)

    Dm = np.asarray(D.reshape(dims.shpD), dtype=dtype)
    Sm = np.asarray(S.reshape(dims.shpS), dtype=dtype)
    Sf = fft.rfftn(Sm, None, axN)
    Df = fft.rfftn(Dm, dims.Nv, axN)
    DSf = np.conj(Df) * Sf
    if dims.Cd > 1:
        DSf = np.sum(DSf, axis=axC, keepdims=True)

    if lmbda is None:                                   # admm/cbpdn.py:573-578
        lmbda = 0.1 * abs(np.conj(Df) * Sf).max()
    lmbda = rdt.type(lmbda)
    if o['rho'] is None:                                # admm/cbpdn.py:584-585
        rho = rdt.type(50.0 * lmbda + 1.0)
    if ar['RsdlTarget'] is not None:
        xi = rdt.type(ar['RsdlTarget'])
    elif lmbda != 0.0:                                  # admm/cbpdn.py:588-593
        xi = rdt.type(float(1.0 + (18.3) ** (np.log10(lmbda) + 1.0)))
    else:
        xi = rdt.type(1.0)
    wl1 = np.asarray(o['L1Weight'], dtype=rdt)
    wl1 = wl1.reshape(l1_weight_shape(wl1, dims))
    joint = mu is not None
    if joint:
        mu_ = dtype.type(mu)
        wl21 = np.asarray(o['L21Weight'], dtype=dtype)
    enet = enet_mu is not None
    if enet:
        assert not joint
        emu = dtype.type(enet_mu)
    grd = grad_mu is not None
    if grd:
        assert not joint and not enet
        gmu = dtype.type(grad_mu)
        gw = o.get('GradWeight', 1.0)
        if hasattr(gw, 'ndim'):
            Wgrd = np.asarray(gw.reshape((1,) * 4 + gw.shape), dtype=dtype)
        else:
            Wgrd = np.asarray(gw, dtype=dtype)
        _, GHGf0 = gradient_filters(5, axN, dims.Nv, dtype, fft)
        GHGf = Wgrd * GHGf0

    if o['U0'] is not None:
        U = o['U0'].astype(dtype, copy=True)
    elif o['Y0'] is None:
        U = np.zeros(dims.shpX, dtype)
    else:                                               # admm/cbpdn.py:601-610
        # the reference evaluates this inside ADMM.__init__, i.e. with rho still at its
        # base-class value (opt['rho'] or 1.0) and lmbda not yet set -> it actually
        # raises AttributeError there; callers that pass Y0 also pass U0.  We restate
        # the documented intent.
        U = (lmbda / rho) * np.sign(Y)

    Nx = np.prod(np.array(dims.shpX))
    Nc = Nx
    hD = Dm.shape[0:2]

    def norm(a):
        return np.linalg.norm(a)

    res = ADMMResult()
    itstat = []
    trace = []
    X = None
    Xf = None
    xrrs = None
    k = 0
    t_start = time.perf_counter()
    for k in range(0, o['MaxMainIter']):
        Yprev = Y.copy()
        # ---- xstep (admm/cbpdn.py:267-293)
        YU = Y - U
        b = DSf + rho * fft.rfftn(YU, None, axN)
        rho_x = (emu + rho) if enet else rho           # admm/cbpdn.py:948-955
        if grd:                                         # admm/cbpdn.py:1173-1201
            rho_x = gmu * GHGf + rho
            if dims.Cd == 1:
                Xf = solvedbd_sm(Df, rho_x, b, axM)
            else:                                       # admm/cbpdn.py:1181-1184: the diagonal goes in as "rho"
                Xf = solvemdbi_ism(Df, rho_x, b, axM, axC)
        elif dims.Cd == 1:
            Xf = solvedbi_sm(Df, rho_x, b, axM)
        else:
            Xf = solvemdbi_ism(Df, rho_x, b, axM, axC)
        X = fft.irfftn(Xf, dims.Nv, axN)
        if o['LinSolveCheck']:
            dx = inner(Df, Xf, axM)
            if dims.Cd == 1:
                ax = np.conj(Df) * dx + rho_x * Xf          # rho_x: scalar, or the GradReg diagonal
            else:
                ax = inner(np.conj(Df), dx, axC) + rho_x * Xf
            xrrs = rrs(ax, b)
        # ---- relaxation (admm/admm.py:877-885)
        if rlx == 1.0:
            AX = X
        else:
            AX = rlx * X + (1 - rlx) * Y
        # ---- ystep (admm/cbpdn.py:614-620, 785-794, 297-311)
        if joint:
            Y = prox_sl1l2(AX + U, (lmbda / rho) * wl1, (mu_ / rho) * wl21, axis=axC)
        else:
            Y = prox_l1(AX + U, (lmbda / rho) * wl1)
        if o['NonNegCoef']:
            Y[Y < 0.0] = 0.0
        if o['NoBndryCross']:
            Y[1 - hD[0]:, :] = 0.0
            Y[:, 1 - hD[1]:] = 0.0
        if ams is not None:                  # AddMaskSim.ystep: the impulse maps bypass the prox
            Yi = AX[..., -ams[1]:] + U[..., -ams[1]:]
            Yi[np.where(ams[0].astype(bool))] = 0.0
            Y[..., -ams[1]:] = Yi
        # ---- ustep (admm/admm.py:434-437)
        U = U + (AX - Y)
        # ---- residuals (admm/admm.py:462-486, 959-983)
        need_rsdl = ar['Enabled'] or not o['FastSolve']
        if need_rsdl:
            if norm_reduce is None:
                nX, nY, nU = norm(X), norm(Y), norm(U)
                nR = norm(X - Y)
                nS = norm(rho * (Yprev - Y))
            else:
                loc = np.array([np.sum(X.astype(np.float64) ** 2),
                                np.sum(Y.astype(np.float64) ** 2),
                                np.sum(U.astype(np.float64) ** 2),
                                np.sum((X - Y).astype(np.float64) ** 2),
                                np.sum((Yprev - Y).astype(np.float64) ** 2)])
                g = norm_reduce(loc)
                nX, nY, nU, nR = [rdt.type(np.sqrt(v)) for v in g[0:4]]
                nS = rho * rdt.type(np.sqrt(g[4]))
            if ar['StdResiduals']:
                r = nR
                s = nS
                epri = np.sqrt(Nc) * o['AbsStopTol'] + max(nX, nY) * o['RelStopTol']
                edua = np.sqrt(Nx) * o['AbsStopTol'] + rho * nU * o['RelStopTol']
            else:
                rn = max(nX, nY)
                if rn == 0.0:
                    rn = 1.0
                sn = rho * nU
                if sn == 0.0:
                    sn = 1.0
                r = nR / rn
                s = nS / sn
                epri = np.sqrt(Nc) * o['AbsStopTol'] / rn + o['RelStopTol']
                edua = np.sqrt(Nx) * o['AbsStopTol'] / sn + o['RelStopTol']
        # ---- objective (admm/cbpdn.py:325-344, 624-630, 798-807)
        if not o['FastSolve']:
            fvar = fft.rfftn(Y, None, axN) if o['AuxVarObj'] else Xf
            gvar = Y if o['AuxVarObj'] else X
            if ams is not None:              # AddMaskSim.obfn_gvar: impulse maps do not count
                gvar = gvar.copy()
                gvar[..., -ams[1]:] = 0
            Ef = inner(Df, fvar, axM) - Sf
            dfd = rfl2norm2(Ef, Sm.shape, axis=axN) / 2.0
            rl1 = np.linalg.norm((wl1 * gvar).ravel(), 1)
            if norm_reduce is not None:
                g = norm_reduce(np.array([dfd, rl1], dtype=np.float64))
                dfd, rl1 = g[0], g[1]
            if joint:
                rl21 = np.sum(wl21 * np.sqrt(np.sum(gvar ** 2, axis=axC)))
                if norm_reduce is not None:
                    rl21 = norm_reduce(np.array([rl21], dtype=np.float64))[0]
                obj = dfd + (lmbda * rl1 + mu_ * rl21)
                row = (k, obj, dfd, rl1, rl21, r, s, epri, edua, rho, xrrs,
                       time.perf_counter() - t_start)
            elif grd:                                   # admm/cbpdn.py:1205-1216
                rgr = rfl2norm2(np.sqrt(GHGf * np.conj(fvar) * fvar), dims.Nv, axis=axN) / 2.0
                if norm_reduce is not None:
                    rgr = norm_reduce(np.array([rgr], dtype=np.float64))[0]
                obj = dfd + (lmbda * rl1 + gmu * rgr)
                row = (k, obj, dfd, rl1, rgr, r, s, epri, edua, rho, xrrs,
                       time.perf_counter() - t_start)
            elif enet:                                  # admm/cbpdn.py:978-986
                rl2 = 0.5 * np.linalg.norm(gvar) ** 2
                if norm_reduce is not None:
                    rl2 = norm_reduce(np.array([rl2], dtype=np.float64))[0]
                obj = dfd + (lmbda * rl1 + emu * rl2)
                row = (k, obj, dfd, rl1, rl2, r, s, epri, edua, rho, xrrs,
                       time.perf_counter() - t_start)
            else:
                obj = dfd + lmbda * rl1
                row = (k, obj, dfd, rl1, r, s, epri, edua, rho, xrrs,
                       time.perf_counter() - t_start)
            itstat.append(row)
        if record:
            trace.append({'rho': float(rho), 'r': float(r) if need_rsdl else None,
                          's': float(s) if need_rsdl else None})
        # ---- rho update (admm/admm.py:549-575)
        if ar['Enabled'] and need_rsdl:
            if k != 0 and np.mod(k + 1, ar['Period']) == 0:
                if ar['AutoScaling']:
                    if s == 0.0 or r == 0.0:
                        rhomlt = tau
                    else:
                        rhomlt = np.sqrt(r / (s * xi) if r > s * xi else (s * xi) / r)
                        if rhomlt > tau:
                            rhomlt = tau
                else:
                    rhomlt = tau
                rsf = 1.0
                if r > xi * mur * s:
                    rsf = rhomlt
                elif s > (mur / xi) * r:
                    rsf = 1.0 / rhomlt
                rho = rho * rdt.type(rsf)
                U = U / rsf
                U = U.astype(dtype, copy=False)
        if timing is not None:
            timing.setdefault('iter_end', []).append(time.perf_counter() - t_start)
        if need_rsdl and r < epri and s < edua:
            break
    if timing is not None:
        timing['solve'] = time.perf_counter() - t_start
        timing['iters'] = k + 1
    res.X, res.Y, res.U, res.Xf = X, Y, U, Xf
    res.Df, res.Sf, res.rho, res.lmbda, res.xi = Df, Sf, rho, lmbda, xi
    res.itstat, res.trace, res.k, res.dims = itstat, trace, k + 1, dims
    res.wl1 = wl1
    return res


def reconstruct(Df, X, dims, fft=None):
    """irfftn(sum_m Df * rfftn(X)) (admm/cbpdn.py:373-380)."""
    fft = fft or FFTBackend()
    Xf = fft.rfftn(X, None, dims.axisN)
    return fft.irfftn(np.sum(Df * Xf, axis=dims.axisM), dims.Nv, dims.axisN)


# ----------------------------------------------------------------------------------
# PGM / FISTA ConvBPDN
# ----------------------------------------------------------------------------------
PGM_DEFAULTS = {
    # pgm/pgm.py:157-166 merged with pgm/cbpdn.py:115-118
    'MaxMainIter': 1000, 'RelStopTol': 1e-3, 'L': 500.0, 'FastSolve': False,
    'DataType': None, 'NonNegCoef': False, 'NoBndryCross': False, 'L1Weight': 1.0,
    'X0': None,
    'Backtrack': None,      # None | {'gamma_u': 1.2, 'maxiter': 50}  (BacktrackStandard)
                            #      | {'kind': 'robust', 'gamma_d': 0.9, 'gamma_u': 2.0, 'maxiter': 50}
    'StepSizePolicy': None,  # None | 'cauchy' | 'bb'   (pgm/stepsize.py; ignored with Backtrack)
    'Monotone': False,       # pgm/pgm.py:802-831 (without backtracking)
    'AutoStop': {'Enabled': False, 'Tau0': 1e-2},
}


def pgm_convbpdn(D, S, lmbda=None, opt=None, dimK=None, fft=None, timing=None, W=None):
    """FISTA ConvBPDN with Nesterov momentum and optional standard backtracking; with `W` the
    masked data fidelity of pgm.cbpdn.ConvBPDNMask (pgm/cbpdn.py:387-508)."""
    fft = fft or FFTBackend()
    o = _merge(PGM_DEFAULTS, opt)
    dims = Dims(D, S, dimK=dimK)
    dtype = np.dtype(o['DataType']) if o['DataType'] is not None else np.dtype(S.dtype)
    axN, axC, axM = dims.axisN, dims.axisC, dims.axisM

    Dm = np.asarray(D.reshape(dims.shpD), dtype=dtype)
    Sm = np.asarray(S.reshape(dims.shpS), dtype=dtype)
    if lmbda is None:
        Df0 = fft.rfftn(D.reshape(dims.shpD), dims.Nv, axN)
        Sf0 = fft.rfftn(S.reshape(dims.shpS), None, axN)
        lmbda = 0.1 * abs(np.conj(Df0) * Sf0).max()
    lmbda = dtype.type(lmbda)
    wl1 = np.asarray(o['L1Weight'], dtype=dtype)
    L = dtype.type(o['L']) if o['L'] is not None else dtype.type(1.0)

    X = np.zeros(dims.shpX, dtype) if o['X0'] is None else o['X0'].astype(dtype, copy=True)
    Sf = fft.rfftn(Sm, None, axN)
    Df = fft.rfftn(Dm, dims.Nv, axN)
    Xf = fft.rfftn(X, None, axN)
    Yf = Xf.copy()
    Yfprv = Yf.copy() + 1e5
    hD = Dm.shape[0:2]
    t = 1
    bt = o['Backtrack']

    def eval_Rf(Vf):
        return inner(Df, Vf, axM) - Sf

    if W is not None:
        W5 = np.asarray(W.reshape(msk_shape(W, dims)), dtype=dtype)

    def grad_f(Vf):
        if W is None:
            g = np.conj(Df) * eval_Rf(Vf)
        else:                                            # pgm/cbpdn.py:461-474
            Ry = fft.irfftn(eval_Rf(Vf), dims.Nv, axN)
            WRyf = fft.rfftn((W5 ** 2) * Ry, dims.Nv, axN)
            g = np.conj(Df) * WRyf
        if dims.Cd > 1:
            g = np.sum(g, axis=axC, keepdims=True)
        return g

    def obfn_f(Vf):
        if W is None:
            return 0.5 * np.linalg.norm(eval_Rf(Vf).flatten(), 2) ** 2
        R = fft.irfftn(eval_Rf(Vf), dims.Nv, axN)        # pgm/cbpdn.py:490-506
        WRf = fft.rfftn(W5 * R, dims.Nv, axN)
        return 0.5 * np.linalg.norm(WRf.flatten(), 2) ** 2

    def obfn_dfd(Vf):
        if W is None:
            return rfl2norm2(eval_Rf(Vf), Sm.shape, axis=axN) / 2.0
        E = fft.irfftn(eval_Rf(Vf), dims.Nv, axN)        # pgm/cbpdn.py:478-486
        return (np.linalg.norm(W5 * E) ** 2) / 2.0

    def prox_g(V, Lc):
        Uo = prox_l1(V, (lmbda / Lc) * wl1)
        if o['NonNegCoef']:
            Uo[Uo < 0.0] = 0.0
        if o['NoBndryCross']:
            Uo[1 - hD[0]:, :] = 0.0
            Uo[:, 1 - hD[1]:] = 0.0
        return Uo

    def xstep(gradf, Lc):
        Vf = Yf - (1. / Lc) * gradf
        V = fft.irfftn(Vf, dims.Nv, axN)
        Xn = prox_g(V, Lc)
        return Xn, fft.rfftn(Xn, None, axN)

    def hessian_f(V):                                    # pgm/cbpdn.py:302-310
        h = np.conj(Df) * inner(Df, V, axM)
        if dims.Cd > 1:
            h = np.sum(h, axis=axC, keepdims=True)
        return h

    def eval_objfn(Xf_, X_):                             # pgm/cbpdn.py:320-356
        dfd_ = obfn_dfd(Xf_)
        rl1_ = np.linalg.norm((wl1 * X_).ravel(), 1)
        return (dfd_ + lmbda * rl1_, dfd_, rl1_)

    robust = bt is not None and bt.get('kind') == 'robust'
    policy = o['StepSizePolicy'] if bt is None else None   # pgm/pgm.py:236-239
    mono = bool(o['Monotone'])
    if mono and bt is not None:
        raise NotImplementedError('oracle: Monotone together with backtracking')
    Tk, Zrb = 0., None                                   # BacktrackRobust state (pgm/backtrack.py:150-151)
    bb_xprv, bb_gprv = 0.0, 0.0                          # StepSizePolicyBB state (pgm/stepsize.py:108-109)
    objfn = objfn_prev = None
    ZZf = None

    res = ADMMResult()
    itstat = []
    F = Q = itbt = None
    k = 0
    t_start = time.perf_counter()
    for k in range(0, o['MaxMainIter']):
        # on_iteration_start (pgm/pgm.py:835-846)
        Xfprv = Xf.copy()
        if not o['FastSolve'] or robust:
            Yfprv = Yf.copy()
        if mono:
            if k == 0:
                objfn = eval_objfn(Xf, X)
            objfn_prev = objfn
        if robust:                                       # pgm/backtrack.py:153-210
            if Zrb is None:
                Zrb = Xf.copy()
            L = L * bt.get('gamma_d', 0.9)
            itbt = 0
            search = True
            while search and itbt < bt.get('maxiter', 50):
                tt = float(1. + np.sqrt(1. + 4. * L * Tk)) / (2. * L)
                T = Tk + tt
                Yf = (Tk * Xfprv + tt * Zrb) / T
                gradY = grad_f(Yf)
                X, Xf = xstep(gradY, L)
                F = obfn_f(Xf)
                Dxy = Xf - Yf
                Q = obfn_f(Yf) + np.sum(np.real(np.conj(Dxy) * gradY)) + \
                    (L / 2.) * np.linalg.norm(Dxy.flatten(), 2) ** 2
                if F <= Q:
                    search = False
                else:
                    L = L * bt.get('gamma_u', 2.0)
                itbt += 1
            Tk = T
            Zrb = Zrb + (tt * L * (Xf - Yf))
        elif bt is not None:                             # pgm/backtrack.py:74-107
            gradY = grad_f(Yf)
            itbt = 0
            search = True
            while search and itbt < bt.get('maxiter', 50):
                X, Xf = xstep(gradY, L)
                F = obfn_f(Xf)
                Dxy = Xf - Yf
                Q = obfn_f(Yf) + np.sum(np.real(np.conj(Dxy) * gradY)) + \
                    (L / 2.) * np.linalg.norm(Dxy.flatten(), 2) ** 2
                if F <= Q:
                    search = False
                else:
                    L = L * dtype.type(bt.get('gamma_u', 1.2))
                itbt += 1
        else:                                            # PGMDFT.xstep (pgm/pgm.py:779-811)
            gradf = grad_f(Yf)
            if policy is not None:
                if k > 1:
                    if policy == 'cauchy':               # pgm/stepsize.py:68-87
                        den = np.sum(np.real(np.conj(gradf) * gradf))
                        num = np.sum(np.real(np.conj(gradf) * hessian_f(gradf)))
                        L = num / den
                    else:                                # Barzilai-Borwein, pgm/stepsize.py:125-145
                        dx = Xf - bb_xprv
                        dg = gradf - bb_gprv
                        den = np.sum(np.real(np.conj(dx) * dg))
                        num = np.sum(np.real(np.conj(dg) * dg))
                        Lbb = num / den
                        if not Lbb < 0.:
                            L = Lbb
                if policy == 'bb':
                    bb_xprv, bb_gprv = Xf, gradf
            X, Xf = xstep(gradf, L)
            if mono and k > 0:
                ZZf = Xf.copy()
                objfn = eval_objfn(Xf, X)
                if objfn_prev[0] < objfn[0]:
                    Xf = Xfprv.copy()
                    objfn = objfn_prev
        if not robust:
            # momentum (pgm/pgm.py:815-831, pgm/momentum.py:45-48)
            tprv = t
            t = 0.5 * float(1. + np.sqrt(1. + 4. * t ** 2))
            if mono and k > 0:
                Yf = Xf + (tprv / t) * (ZZf - Xf) + ((tprv - 1.) / t) * (Xf - Xfprv)
            else:
                Yf = Xf + ((tprv - 1.) / t) * (Xf - Xfprv)
        if not o['FastSolve']:
            frcxd = rfl2norm2(Xf - Yfprv, X.shape, axis=axN)
            tol = o['RelStopTol']
            if o['AutoStop']['Enabled']:
                tol = o['AutoStop']['Tau0'] / (1. + k)
            if mono:                                     # pgm/pgm.py:546-549: the tracked objective
                obj, dfd, rl1 = objfn
            else:
                dfd = obfn_dfd(Xf)
                rl1 = np.linalg.norm((wl1 * X).ravel(), 1)
                obj = dfd + lmbda * rl1
            itstat.append((k, obj, dfd, rl1, frcxd, F, Q, itbt, L,
                           time.perf_counter() - t_start))
            if frcxd < tol:
                break
    if timing is not None:
        timing['solve'] = time.perf_counter() - t_start
        timing['iters'] = k + 1
    res.X, res.Xf, res.Yf, res.L, res.lmbda = X, Xf, Yf, L, lmbda
    res.Df, res.Sf, res.itstat, res.k, res.dims, res.t = Df, Sf, itstat, k + 1, dims, t
    return res

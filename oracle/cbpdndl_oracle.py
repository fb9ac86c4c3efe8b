"""CPU oracle for convolutional dictionary learning -- TEST INFRASTRUCTURE, not product code.

Restates, in numpy and in the reference's layout, ``sporco.dictlrn.cbpdndl.ConvBPDNDictLearn``
with its default solvers: X-step = one ADMM ConvBPDN iteration per outer iteration
(sporco/dictlrn/cbpdndl.py:49-55, admm/cbpdn.py), D-step = one PGM iteration of the
constrained convolutional MOD problem (sporco/pgm/ccmod.py:28-404, pgm/pgm.py:328-370), the
alternation of sporco/dictlrn/dictlrn.py:327-363 and the constraint-set projection
``Pcn`` (sporco/cnvrep.py:868-1074: crop to the filter support, optional zero mean, unit
norm).  Greyscale, single-channel dictionary with multi-channel signals (the channels then act as
further images in the D step, pgm/ccmod.py:232-237, 264-281) and multi-channel dictionaries; one
support size.  Signals are given as (N0, N1, K) or (N0, N1, C, K).
Pinned bit-for-bit to the live reference by oracle/make_golden.py (fixture tests/golden/cdl_*.npz).
"""

import time

import numpy as np

from . import cbpdn_oracle as co


def dsz_blocks(dsz):
    """Blocks [(h, w, [Cd,] Mb), ...] of a single- or multi-scale dictionary size (cnvrep.py:277-360)."""
    return [tuple(b) for b in dsz] if isinstance(dsz[0], (tuple, list)) else [tuple(dsz)]


def dsz_max(dsz):
    """(max h, max w, [Cd,] M): the support that holds every filter."""
    bl = dsz_blocks(dsz)
    return (max(b[0] for b in bl), max(b[1] for b in bl)) + tuple(bl[0][2:-1]) + (sum(b[-1] for b in bl),)


def pcn(x, dsz, Nv, zm=False):
    """normalise(zeromean(zpad(bcrop(x)))) for x of shape (N0, N1, Cd, 1, M)  (cnvrep.py:953-1033); every block of
    a multi-scale dictionary over its own support (cnvrep.py:609-668, 894-950)."""
    p = np.zeros(x.shape, dtype=x.dtype)
    m0 = 0
    for b in dsz_blocks(dsz):
        h, w, mb = b[0], b[1], b[-1]
        p[0:h, 0:w, ..., m0:m0 + mb] = x[0:h, 0:w, ..., m0:m0 + mb]
        if zm:
            p[0:h, 0:w, ..., m0:m0 + mb] -= np.mean(p[0:h, 0:w, ..., m0:m0 + mb], (0, 1))
        m0 += mb
    vn = np.sqrt(np.sum(p ** 2, (0, 1, 2), keepdims=True))
    vn[vn == 0] = 1.0
    return np.asarray(p / vn, dtype=x.dtype)


CDL_DEFAULTS = {
    'MaxMainIter': 10,
    'CBPDN': {'rho': None, 'RelaxParam': 1.8, 'NonNegCoef': False, 'NoBndryCross': False,
              'AuxVarObj': False,
              'AutoRho': {'Enabled': True, 'Period': 10, 'AutoScaling': False, 'RsdlRatio': 10.0,
                          'Scaling': 2.0, 'RsdlTarget': 1.0, 'StdResiduals': False},
              'RelStopTol': 1e-3, 'AbsStopTol': 0.0},
    'CCMOD': {'L': None, 'ZeroMean': False},
}


def cbpdndl(D0, S, lmbda, opt=None, fft=None, reduce=None):
    """Run ConvBPDNDictLearn(D0, S, lmbda, opt) with xmethod='admm', dmethod='pgm'.
    Returns a dict with the learned dictionary (cropped), the coefficient maps and the
    iteration statistics columns.  `reduce`, if given, maps a float64 array of local sums to
    global sums: the form of the algorithm with the training images sharded over ranks (every
    rank codes its own images; squared norms, objective terms and the dictionary gradient are
    summed), which must reproduce the unsharded run."""
    fft = fft or co.FFTBackend()
    o = {'MaxMainIter': 10, 'CBPDN': dict(CDL_DEFAULTS['CBPDN']), 'CCMOD': dict(CDL_DEFAULTS['CCMOD'])}
    o['CBPDN']['AutoRho'] = dict(CDL_DEFAULTS['CBPDN']['AutoRho'])
    for k, v in (opt or {}).items():
        if k in ('CBPDN', 'CCMOD'):
            for kk, vv in v.items():
                if kk == 'AutoRho':
                    o[k][kk].update(vv)
                else:
                    o[k][kk] = vv
        else:
            o[k] = v
    xo, do = o['CBPDN'], o['CCMOD']
    ar = xo['AutoRho']
    dtype = np.dtype(S.dtype)
    rdt = co._rdt(dtype)
    dsz = o.get('DictSize') or D0.shape                     # possibly multi-scale; D0 holds the largest support
    mxd = dsz_max(dsz)
    Cd = D0.shape[2] if D0.ndim == 4 else 1
    C = S.shape[2] if S.ndim == 4 else 1
    N0, N1, K = S.shape[0], S.shape[1], S.shape[-1]
    M = D0.shape[-1]
    Cx = C - Cd + 1
    Nv = (N0, N1)
    axN, axC, axK, axM = (0, 1), 2, 3, 4

    # ---- initial dictionary: cropped + normalised (cbpdndl.py:448-454)
    Dn = pcn(np.asarray(D0.reshape(mxd[0], mxd[1], Cd, 1, M)), dsz, (mxd[0], mxd[1]), zm=do['ZeroMean'])
    X0d = np.zeros((N0, N1, Cd, 1, M), dtype=D0.dtype)
    X0d[0:mxd[0], 0:mxd[1]] = Dn

    # ---- X-step state (admm/cbpdn.py ctor)
    Sm = np.asarray(S.reshape(N0, N1, C, K, 1), dtype=dtype)
    Sf = fft.rfftn(Sm, None, axN)
    # D step: with a single-channel dictionary the channels of the signal count as images
    Sd = Sm.reshape(N0, N1, 1, C * K, 1) if (Cd == 1 and C > 1) else Sm
    Sdf = fft.rfftn(Sd, None, axN)
    lm = rdt.type(lmbda)
    rho = rdt.type(xo['rho']) if xo['rho'] is not None else rdt.type(50.0 * lm + 1.0)
    tau, mur = rdt.type(ar['Scaling']), rdt.type(ar['RsdlRatio'])
    xi = rdt.type(ar['RsdlTarget'])
    rlx = rdt.type(xo['RelaxParam'])
    Y = np.zeros((N0, N1, Cx, K, M), dtype)
    U = np.zeros((N0, N1, Cx, K, M), dtype)
    Dcur = np.asarray(Dn, dtype=dtype)
    Nx = np.prod(np.array(Y.shape))
    kx = 0

    # ---- D-step state (pgm/ccmod.py ctor)
    # NB: the reference means K*14 as default (pgm/ccmod.py:218) but PGM.__init__ has already set
    # L = 1.0 (pgm/pgm.py:242) and set_attr does not overwrite it: the effective default is 1.0
    L = dtype.type(do['L']) if do['L'] is not None else dtype.type(1.0)
    Xd = X0d.astype(dtype, copy=True)
    Xdf = fft.rfftn(Xd, None, axN)
    Ydf = Xdf
    t = 1

    cols = {n: [] for n in ('ObjFun', 'DFid', 'RegL1', 'Cnstr', 'XPrRsdl', 'XDlRsdl', 'XRho',
                            'D_L', 'D_Rsdl')}
    t0 = time.perf_counter()
    for j in range(o['MaxMainIter']):
        # ================= X step: one ADMM iteration =================
        Df = fft.rfftn(Dcur, Nv, axN)
        DSf = np.conj(Df) * Sf
        if Cd > 1:
            DSf = np.sum(DSf, axis=axC, keepdims=True)
        Yprev = Y.copy()
        b = DSf + rho * fft.rfftn(Y - U, None, axN)
        if Cd == 1:
            Xf = co.solvedbi_sm(Df, rho, b, axM)
        else:
            Xf = co.solvemdbi_ism(Df, rho, b, axM, axC)
        X = fft.irfftn(Xf, Nv, axN)
        AX = X if rlx == 1.0 else rlx * X + (1 - rlx) * Y
        Y = co.prox_l1(AX + U, (lm / rho))
        if xo['NonNegCoef']:
            Y[Y < 0.0] = 0.0
        U = U + (AX - Y)
        if reduce is None:
            nX, nY, nU = np.linalg.norm(X), np.linalg.norm(Y), np.linalg.norm(U)
            nR, nS = np.linalg.norm(X - Y), np.linalg.norm(rho * (Yprev - Y))
        else:
            g = reduce(np.array([np.sum(a.astype(np.float64) ** 2) for a in (X, Y, U, X - Y, Yprev - Y)]))
            nX, nY, nU, nR = [rdt.type(np.sqrt(v)) for v in g[0:4]]
            nS = rho * rdt.type(np.sqrt(g[4]))
        rn = max(nX, nY)
        rn = 1.0 if rn == 0.0 else rn
        sn = rho * nU
        sn = 1.0 if sn == 0.0 else sn
        r = nR / rn
        s = nS / sn
        Ef = co.inner(Df, Xf, axM) - Sf
        dfd = co.rfl2norm2(Ef, Sm.shape, axis=axN) / 2.0
        rl1 = np.linalg.norm(X.ravel(), 1)
        if reduce is not None:
            dfd, rl1 = reduce(np.array([dfd, rl1], dtype=np.float64))
        xrho = rho
        if ar['Enabled'] and kx != 0 and np.mod(kx + 1, ar['Period']) == 0:
            if ar['AutoScaling']:
                if s == 0.0 or r == 0.0:
                    mlt = tau
                else:
                    mlt = np.sqrt(r / (s * xi) if r > s * xi else (s * xi) / r)
                    if mlt > tau:
                        mlt = tau
            else:
                mlt = tau
            rsf = 1.0
            if r > xi * mur * s:
                rsf = mlt
            elif s > (mur / xi) * r:
                rsf = 1.0 / mlt
            rho = rho * rdt.type(rsf)
            U = (U / rsf).astype(dtype, copy=False)
        kx += 1
        # ================= D step: one PGM iteration on the dictionary =================
        Zc = Y.reshape(N0, N1, 1, Cx * K, M) if (Cd == 1 and C > 1) else Y
        Zf = fft.rfftn(Zc, Nv, axN)                             # setcoef (pgm/ccmod.py:264-281)
        Xdfprv = Xdf.copy()
        Ydfprv = Ydf.copy()
        Ryf = co.inner(Zf, Ydf, axM) - Sdf
        gradf = co.inner(np.conj(Zf), Ryf, axK)
        if reduce is not None:                                   # sum of the shards' gradients
            gradf = (reduce(gradf.real.astype(np.float64)) + 1j * reduce(gradf.imag.astype(np.float64))
                     ).astype(gradf.dtype)
        Vf = Ydf - (1. / L) * gradf
        V = fft.irfftn(Vf, Nv, axN)
        Xd = pcn(V, dsz, Nv, zm=do['ZeroMean'])
        Xdf = fft.rfftn(Xd, None, axN)
        tprv = t
        t = 0.5 * float(1. + np.sqrt(1. + 4. * t ** 2))
        Ydf = Xdf + ((tprv - 1.) / t) * (Xdf - Xdfprv)
        drsdl = co.rfl2norm2(Xdf - Ydfprv, Xd.shape, axis=axN)
        cns = np.linalg.norm((pcn(Xd, dsz, Nv, zm=do['ZeroMean']) - Xd))
        # ================= book-keeping (dictlrn/dictlrn.py:327-363) =================
        Dcur = np.asarray(Xd[0:mxd[0], 0:mxd[1]], dtype=dtype)
        for name, val in (('ObjFun', dfd + lm * rl1), ('DFid', dfd), ('RegL1', rl1), ('Cnstr', cns),
                          ('XPrRsdl', r), ('XDlRsdl', s), ('XRho', xrho), ('D_L', L),
                          ('D_Rsdl', drsdl)):
            cols[name].append(float(val))
    out = {k: np.array(v, dtype=np.float64) for k, v in cols.items()}
    out['D'] = Dcur.reshape(mxd)
    out['X'] = Y
    out['time'] = time.perf_counter() - t0
    return out


# =====================================================================================
# Consensus dictionary update: sporco.admm.ccmod.ConvCnstrMOD_Consensus (sporco/admm/ccmod.py:613-911)
# on ADMMConsensus / ADMM (sporco/admm/admm.py:293-389, 434-486, 549-575, 1419-1707).
# =====================================================================================
CNS_DEFAULTS = {
    'MaxMainIter': 1000, 'rho': None, 'RelaxParam': 1.8, 'ZeroMean': False, 'Y0': None,
    'RelStopTol': 1e-3, 'AbsStopTol': 0.0, 'AuxVarObj': True,
    'AutoRho': {'Enabled': False, 'Period': 10, 'Scaling': 2.0, 'RsdlRatio': 10.0, 'RsdlTarget': 1.0,
                'AutoScaling': False, 'StdResiduals': False},
}


class ConsensusCCMOD(object):
    """State and iteration of ConvCnstrMOD_Consensus (objective on the consensus variable: AuxVarObj True,
    the class default).  S: (N0, N1, K) or (N0, N1, C, K); coefficient maps Z: (N0, N1, Cx, K, M); dsz =
    (h, w, M) or (h, w, Cd, M).  `reduce`, if given, maps a float64 array of local sums to global sums (blocks
    sharded over ranks: the supports of the block mean and the squared norms are summed)."""

    def __init__(self, S, dsz, opt=None, fft=None, reduce=None, nb_global=None):
        self.fft = fft or co.FFTBackend()
        o = {k: (dict(v) if isinstance(v, dict) else v) for k, v in CNS_DEFAULTS.items()}
        for k, v in (opt or {}).items():
            if k == 'AutoRho':
                o[k].update(v)
            else:
                o[k] = v
        self.o = o
        self.reduce = reduce
        self.dtype = np.dtype(S.dtype)
        self.rdt = co._rdt(self.dtype)
        self.dsz = dsz
        self.mxd = dsz_max(dsz)
        self.Cd = self.mxd[2] if len(self.mxd) == 4 else 1
        self.C = S.shape[2] if S.ndim == 4 else 1
        self.N0, self.N1, self.K = S.shape[0], S.shape[1], S.shape[-1]
        self.M = self.mxd[-1]
        self.Nv = (self.N0, self.N1)
        N0, N1, C, K, Cd, M = self.N0, self.N1, self.C, self.K, self.Cd, self.M
        self.Nb = (K if C == Cd else C * K)                      # local blocks (ccmod.py:684-686)
        self.NbG = self.Nb if nb_global is None else int(nb_global)     # blocks over all ranks
        Sm = np.asarray(S.reshape(N0, N1, C, K, 1), dtype=self.dtype)
        if Cd == 1 and C > 1:
            Sm = Sm.reshape(N0, N1, 1, C * K, 1)
        self.S = Sm
        self.Sf = self.fft.rfftn(Sm, None, (0, 1))
        # NB the reference means the number of images as default (ccmod.py:691-692) but ADMM.__init__ has already
        # set rho = 1 (admm.py:247) and set_attr keeps a value that is set: the effective default is 1
        self.rho = self.rdt.type(1.0 if o['rho'] is None else o['rho'])
        self.rlx = self.rdt.type(o['RelaxParam'])
        yshape = (N0, N1, Cd, 1, M)
        self.Nx = self.NbG * int(np.prod(yshape))
        self.Nc = self.Nx
        if o['Y0'] is None:
            self.Y = np.zeros(yshape, self.dtype)
            self.U = np.zeros(yshape + (self.Nb,), self.dtype)
        else:
            self.Y = np.asarray(o['Y0'], dtype=self.dtype).reshape(yshape)
            self.U = (np.repeat(self.Y[..., np.newaxis], self.Nb, axis=-1) / self.rho).astype(self.dtype)
        self.k = 0
        self.itstat = []

    def setcoef(self, Z):
        N0, N1, C, K, Cd, M = self.N0, self.N1, self.C, self.K, self.Cd, self.M
        Cx = C - Cd + 1
        Z = np.asarray(Z, dtype=self.dtype).reshape(N0, N1, Cx, K, M)
        if Cd == 1 and C > 1:
            Z = Z.reshape(N0, N1, 1, Cx * K, M)
        self.Zf = self.fft.rfftn(Z, self.Nv, (0, 1))
        self.ZSf = np.conj(self.Zf) * self.Sf

    def _pcn(self, x):
        return pcn(x, self.dsz, self.Nv, zm=self.o['ZeroMean'])

    def step(self):
        """One pass of the loop body of ADMM.solve (admm.py:331-377).  Returns True when the stopping
        test fires."""
        fft, o, ar = self.fft, self.o, self.o['AutoRho']
        rdt, Nb = self.rdt, self.Nb
        axN, axK, axM = (0, 1), 3, 4
        Yprev = self.Y.copy()
        # xstep (ccmod.py:787-813; the per-block form of xistep :825-838 is the same arithmetic)
        YU = self.Y[..., np.newaxis] - self.U
        X = np.empty_like(self.U)
        Xfs = []
        for i in range(Nb):
            b = np.take(self.ZSf, [i], axis=axK) + self.rho * fft.rfftn(YU[..., i], None, axN)
            Xf = co.solvedbi_sm(np.take(self.Zf, [i], axis=axK), self.rho, b, axM)
            Xfs.append(Xf)
            X[..., i] = fft.irfftn(Xf, self.Nv, axN)
        # relax_AX (admm.py:1608-1616)
        AX = X if self.rlx == 1.0 else self.rlx * X + (1 - self.rlx) * self.Y[..., np.newaxis]
        # ystep (admm.py:1585-1591) with prox_g = Pcn
        if self.reduce is None:
            mAXU = np.mean(AX + self.U, axis=-1)
        else:       # blocks sharded over ranks: only the filter supports of the mean are exchanged
            loc = np.sum((AX + self.U).astype(np.float64), axis=-1) / self.NbG
            h, w = self.mxd[0], self.mxd[1]
            supp = self.reduce(loc[0:h, 0:w].copy())
            mAXU = np.zeros(self.Y.shape, self.dtype)
            mAXU[0:h, 0:w] = supp.astype(self.dtype)
        self.Y = self._pcn(np.asarray(mAXU, dtype=self.dtype))
        # ustep (admm.py:434-437)
        self.U = self.U + (AX - self.Y[..., np.newaxis])
        # compute_residuals (admm.py:462-486 with ADMMConsensus.rsdl_* :1673-1707)
        if self.reduce is None:
            nX, nU = np.linalg.norm(X), np.linalg.norm(self.U)
            nR = np.linalg.norm(X - self.Y[..., np.newaxis])
        else:
            g = self.reduce(np.array([np.sum(a.astype(np.float64) ** 2)
                                      for a in (X, self.U, X - self.Y[..., np.newaxis])]))
            nX, nU, nR = [rdt.type(np.sqrt(v)) for v in g]
        nY = np.linalg.norm(self.Y)
        sNb = np.sqrt(self.NbG)
        r = nR
        s = np.linalg.norm(sNb * self.rho * (Yprev - self.Y))
        rn = max(nX, sNb * nY)
        sn = self.rho * nU
        if ar['StdResiduals']:
            epri = np.sqrt(self.Nc) * o['AbsStopTol'] + rn * o['RelStopTol']
            edua = np.sqrt(self.Nx) * o['AbsStopTol'] + sn * o['RelStopTol']
        else:
            rn = 1.0 if rn == 0.0 else rn
            sn = 1.0 if sn == 0.0 else sn
            r = r / rn
            s = s / sn
            epri = np.sqrt(self.Nc) * o['AbsStopTol'] / rn + o['RelStopTol']
            edua = np.sqrt(self.Nx) * o['AbsStopTol'] / sn + o['RelStopTol']
        if o['AuxVarObj']:
            # objective on Y (ccmod.py:861-902 with fEvalX False, gEvalY True)
            Yf = fft.rfftn(self.Y, None, axN)
            Ef = co.inner(self.Zf, Yf, axM) - self.Sf
            dfd = co.rfl2norm2(Ef, self.S.shape, axis=axN) / 2.0
            if self.reduce is not None:
                dfd = self.reduce(np.array([dfd], dtype=np.float64))[0]
            cns = np.linalg.norm(self._pcn(self.Y) - self.Y)
        else:
            # objective on the block variables (fEvalX True, gEvalY False): data fidelity of every block with its own
            # X_i (obfn_fvarf = swapaxes(Xf), ccmod.py:872-892), constraint violation of the block mean (admm.py:1641-1646)
            assert self.reduce is None
            Xfb = np.concatenate(Xfs, axis=axK)                  # (N0, N1f, Cd, Nb, M): block i on the image axis
            Ef = co.inner(self.Zf, Xfb, axM) - self.Sf
            dfd = co.rfl2norm2(Ef, self.S.shape, axis=axN) / 2.0
            Yg = np.mean(X, axis=-1)
            cns = np.linalg.norm(self._pcn(Yg) - Yg)
        self.itstat.append((self.k, float(dfd), float(cns), float(r), float(s), float(epri), float(edua),
                            float(self.rho)))
        # update_rho (admm.py:549-575)
        if ar['Enabled'] and self.k != 0 and np.mod(self.k + 1, ar['Period']) == 0:
            tau, mu, xi = rdt.type(ar['Scaling']), rdt.type(ar['RsdlRatio']), rdt.type(ar['RsdlTarget'])
            if ar['AutoScaling']:
                if s == 0.0 or r == 0.0:
                    mlt = tau
                else:
                    mlt = np.sqrt(r / (s * xi) if r > s * xi else (s * xi) / r)
                    if mlt > tau:
                        mlt = tau
            else:
                mlt = tau
            rsf = 1.0
            if r > xi * mu * s:
                rsf = mlt
            elif s > (mu / xi) * r:
                rsf = 1.0 / mlt
            self.rho = self.rho * rdt.type(rsf)
            self.U = (self.U / rsf).astype(self.dtype, copy=False)
        stop = bool(r < epri and s < edua)
        self.k += 1
        return stop

    def solve(self):
        for _ in range(self.o['MaxMainIter']):
            if self.step():
                break
        return self.Y

    def getdict(self):
        return self.Y[0:self.mxd[0], 0:self.mxd[1]]

"""Device-time measurements of the other BASELINE.json configurations (not the headline metric).

    python tools/bench_configs.py [cfg2 cfg3a cfg3b cfg4 cfg5 f64 ...]

One JSON line per configuration: ms per iteration from CUDA events on the library's stream (ADMM) or
wall clock around synchronised calls (PGM, dictionary learning), the algorithmic bytes of one iteration
(every logical array a kernel of the schedule reads or writes, counted once per
kernel) and the fraction of the HBM peak that corresponds to.  `bench.py` imports `measure()` for the
"configs" block of its JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def unit(D):
    return D / np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))


def admm_bytes(N0, N1, M, K, Cx, C, Cd, esz, fused):
    """Algorithmic bytes of one ADMM iteration.  B_r = one real X-shaped array; Zt = its half spectrum."""
    n1f = N1 // 2 + 1
    b_r = float(esz) * N0 * N1 * Cx * K * M
    zt = 2.0 * esz * N0 * n1f * Cx * K * M
    small = 2.0 * esz * N0 * n1f * (Cd * M + C * K)          # dictionary + signal spectra
    col = 2 * zt + small
    if fused:       # column kernel + prox kernel that also emits the next row spectra
        return col + (zt + 4 * b_r + zt)
    return (2 * b_r + zt) + col + (zt + 4 * b_r)              # row forward, column, row inverse + prox


def _admm(b, iters, warm, shape):
    h = b._h
    h.admm_configure(**b._admm_config())
    h.admm_iterate(warm, False)
    _, done, _ = h.admm_iterate(iters, False)
    ms, _ = h.admm_last_timing()
    prof_n = min(20, iters)
    kms = [x / prof_n for x in h.admm_profile(prof_n)]
    sched = h.admm_schedule_info()
    gb = admm_bytes(fused=bool(sched['fused']), **shape) / 1e9
    return {'ms': ms / iters, 'it_per_s': iters / ms * 1e3, 'algorithmic_GB': gb,
            'kernel_ms': {'row_fwd': kms[0], 'col': kms[1], 'row_inv_prox': kms[2], 'scalars': kms[3]},
            'schedule': sched}


def measure(name, peak_gbs, quick=True):
    """Run configuration `name`; returns a dict with ms, algorithmic_GB and frac (of peak_gbs)."""
    from sporco_b200.admm import cbpdn
    rng = np.random.default_rng(12345)
    f32 = np.float32
    out = None
    if name == 'cfg2':
        D = unit(rng.standard_normal((8, 8, 32))).astype(f32)
        S = rng.standard_normal((256, 256)).astype(f32)
        o = cbpdn.ConvBPDN.Options({'RelStopTol': 0.0, 'FastSolve': True})
        out = _admm(cbpdn.ConvBPDN(D, S, 0.1, o), 200, 20,
                    dict(N0=256, N1=256, M=32, K=1, Cx=1, C=1, Cd=1, esz=4))
        out['workload'] = 'admm.cbpdn.ConvBPDN 256x256, 8x8x32, one image, f32'
        out['note'] = 'working set 25 MB: lives in L2, not graded against HBM'
    elif name == 'cfg3a':
        D = unit(rng.standard_normal((8, 8, 64))).astype(f32)
        S = rng.standard_normal((256, 256, 3, 32)).astype(f32)
        o = cbpdn.ConvBPDNJoint.Options({'RelStopTol': 0.0, 'FastSolve': True})
        out = _admm(cbpdn.ConvBPDNJoint(D, S, 0.1, 0.01, o), 30 if quick else 50, 10,
                    dict(N0=256, N1=256, M=64, K=32, Cx=3, C=3, Cd=1, esz=4))
        out['workload'] = 'admm.cbpdn.ConvBPDNJoint 256x256x3, 8x8x64 (greyscale dictionary), 32 images, f32'
    elif name == 'cfg3b':
        D = unit(rng.standard_normal((8, 8, 3, 64))).astype(f32)
        S = rng.standard_normal((256, 256, 3, 32)).astype(f32)
        o = cbpdn.ConvBPDNJoint.Options({'RelStopTol': 0.0, 'FastSolve': True})
        out = _admm(cbpdn.ConvBPDNJoint(D, S, 0.1, 0.01, o), 30 if quick else 50, 10,
                    dict(N0=256, N1=256, M=64, K=32, Cx=1, C=3, Cd=3, esz=4))
        out['workload'] = 'admm.cbpdn.ConvBPDNJoint 256x256x3, 8x8x3x64 (colour dictionary), 32 images, f32'
    elif name == 'f64':
        D = unit(rng.standard_normal((8, 8, 64)))
        S = rng.standard_normal((256, 256, 8))
        o = cbpdn.ConvBPDN.Options({'RelStopTol': 0.0, 'FastSolve': True})
        out = _admm(cbpdn.ConvBPDN(D, S, 0.1, o, dimK=1), 50, 10,
                    dict(N0=256, N1=256, M=64, K=8, Cx=1, C=1, Cd=1, esz=8))
        out['workload'] = 'admm.cbpdn.ConvBPDN 256x256, 8x8x64, 8 images, float64'
    elif name == 'cfg4':
        from sporco_b200.pgm import cbpdn as pcbpdn
        from sporco_b200.pgm.backtrack import BacktrackStandard
        D = unit(rng.standard_normal((12, 12, 128))).astype(f32)
        S = rng.standard_normal((512, 512)).astype(f32)
        o = pcbpdn.ConvBPDN.Options({'MaxMainIter': 10, 'RelStopTol': 0.0, 'L': 1.0,
                                     'Backtrack': BacktrackStandard(maxiter=15)})
        b = pcbpdn.ConvBPDN(D, S, 0.05, o)
        b.solve()
        n = 30 if quick else 50
        b.opt['MaxMainIter'] = n
        b._h.synchronize()
        t0 = time.perf_counter()
        b.solve()
        b._h.synchronize()
        dt = time.perf_counter() - t0
        its = b.getitstat()
        trials = float(np.sum(its.IterBTrack[-n:]))
        # per trial: gradient step R Yf W Vt (column kernel) + R Vt W X W Xt (row kernel) + R Xt R Yf W Xf
        # (evaluation); per iteration additionally the momentum step R Xf R Xfprv W Yf
        zt = 8.0 * 512 * 257 * 128
        b_r = 4.0 * 512 * 512 * 128
        per_trial = (2 * zt) + (2 * zt + b_r) + (3 * zt)
        per_iter = per_trial * trials / n + 3 * zt
        out = {'ms': dt / n * 1e3, 'it_per_s': n / dt, 'trials_per_iter': trials / n,
               'ms_per_trial': dt / trials * 1e3, 'algorithmic_GB': per_iter / 1e9,
               'workload': 'pgm.cbpdn.ConvBPDN 512x512, 12x12x128, BacktrackStandard, f32',
               'final_L': float(its.L[-1])}
    elif name == 'cfg5':
        from sporco_b200.dictlrn import cbpdndl
        D0 = rng.standard_normal((8, 8, 64)).astype(f32)
        S = rng.standard_normal((256, 256, 16)).astype(f32)
        o = cbpdndl.ConvBPDNDictLearn.Options({'MaxMainIter': 10})
        b = cbpdndl.ConvBPDNDictLearn(D0, S, 0.1, o)
        b.solve()
        n = 40 if quick else 100
        b.opt['MaxMainIter'] = n
        b.xstep._h.synchronize()
        t0 = time.perf_counter()
        b.solve()
        b.xstep._h.synchronize()
        dt = time.perf_counter() - t0
        its = b.getitstat()
        # X step: one unfused ADMM iteration with statistics (10.03 B_r) ; hand-over: R Y W Zf (row) + R/W Zf
        # (columns); D step: gradient R Zf, evaluation R Zf; dictionary-sized arrays are negligible
        b_r = 4.0 * 256 * 256 * 16 * 64
        zt = 8.0 * 256 * 129 * 16 * 64
        per_iter = (2 * b_r + zt) + 2 * zt + (zt + 4 * b_r) + (b_r + zt) + 2 * zt + 2 * zt
        out = {'ms': dt / n * 1e3, 'it_per_s': n / dt, 'algorithmic_GB': per_iter / 1e9,
               'workload': 'dictlrn.cbpdndl.ConvBPDNDictLearn 16 images 256x256, 8x8x64, ADMM X step / PGM D step, f32',
               'ObjFun_first_last': [float(its.ObjFun[0]), float(its.ObjFun[-1])]}
    elif name == 'cfg5_cns':
        # the same problem with the consensus ADMM dictionary update (dmethod 'cns'): per outer iteration the D step
        # runs Y - U_i (R U W W: 2 B), forward rows (R W W Z: B + Zt), block solves in place against the block's
        # coefficient spectra (R W Z + R Zf: 3 Zt), inverse rows (Zt + B), dual update and norms (R X U W U: 3 B),
        # data fidelity on Y (R Zf: Zt); the X step and the hand-over as in cfg5
        from sporco_b200.dictlrn import cbpdndl
        D0 = rng.standard_normal((8, 8, 64)).astype(f32)
        S = rng.standard_normal((256, 256, 16)).astype(f32)
        o = cbpdndl.ConvBPDNDictLearn.Options({'MaxMainIter': 10, 'CCMOD': {'rho': 16.0}}, dmethod='cns')
        b = cbpdndl.ConvBPDNDictLearn(D0, S, 0.1, o, dmethod='cns')
        b.solve()
        n = 40 if quick else 100
        b.opt['MaxMainIter'] = n
        b.xstep._h.synchronize()
        t0 = time.perf_counter()
        b.solve()
        b.xstep._h.synchronize()
        dt = time.perf_counter() - t0
        its = b.getitstat()
        b_r = 4.0 * 256 * 256 * 16 * 64
        zt = 8.0 * 256 * 129 * 16 * 64
        xstep = (2 * b_r + zt) + 2 * zt + (zt + 4 * b_r) + (b_r + zt) + 2 * zt
        dstep = 2 * b_r + (b_r + zt) + 3 * zt + (zt + b_r) + 3 * b_r + zt
        out = {'ms': dt / n * 1e3, 'it_per_s': n / dt, 'algorithmic_GB': (xstep + dstep) / 1e9,
               'workload': 'dictlrn.cbpdndl.ConvBPDNDictLearn 16 images 256x256, 8x8x64, ADMM X step / consensus ADMM D step '
                           '(dmethod cns), f32',
               'ObjFun_first_last': [float(its.ObjFun[0]), float(its.ObjFun[-1])],
               'DPrRsdl_last': float(its.DPrRsdl[-1])}
    else:
        raise ValueError(name)
    out['GBps'] = out['algorithmic_GB'] / (out['ms'] / 1e3)
    out['frac'] = out['GBps'] / peak_gbs
    return out


def main():
    which = sys.argv[1:] or ['cfg2', 'cfg3a', 'cfg3b', 'cfg4', 'cfg5', 'f64']
    peak = 6567.4
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peak = float(json.load(open(p))['hbm_gbs'])
    for name in which:
        r = measure(name, peak, quick=False)
        r['config'] = name
        print(json.dumps(r), flush=True)


if __name__ == '__main__':
    main()

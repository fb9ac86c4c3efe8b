"""Device-time measurements of the other BASELINE.json configurations (not the headline metric).
    python tools/bench_configs.py [cfg2 cfg3 cfg4 f64 ...]
Prints one JSON line per configuration: ms per iteration from CUDA events on the library's stream
(ADMM) or wall clock around synchronised trial calls (PGM)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sporco_b200.admm import cbpdn            # noqa: E402
from sporco_b200.pgm import cbpdn as pcbpdn   # noqa: E402
from sporco_b200.pgm.backtrack import BacktrackStandard   # noqa: E402


def unit(D):
    return D / np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))


def admm(name, b, iters, warm):
    h = b._h
    h.admm_configure(**b._admm_config())
    h.admm_iterate(warm, False)
    _, done, _ = h.admm_iterate(iters, False)
    ms, launches = h.admm_last_timing()
    kms = [x / 20 for x in h.admm_profile(20)]
    print(json.dumps({'config': name, 'ms_per_iter': ms / iters, 'it_per_s': iters / ms * 1e3,
                      'kernel_ms': kms, 'schedule': h.admm_schedule_info()}))


def main():
    which = sys.argv[1:] or ['cfg2', 'cfg3', 'cfg3b', 'cfg4', 'f64']
    rng = np.random.default_rng(12345)
    if 'cfg2' in which:
        D = unit(rng.standard_normal((8, 8, 32))).astype(np.float32)
        S = rng.standard_normal((256, 256)).astype(np.float32)
        o = cbpdn.ConvBPDN.Options({'RelStopTol': 0.0, 'FastSolve': True})
        admm('cfg2: ConvBPDN 256x256, 8x8x32, K=1, f32', cbpdn.ConvBPDN(D, S, 0.1, o), 200, 20)
    if 'cfg3' in which:
        D = unit(rng.standard_normal((8, 8, 64))).astype(np.float32)
        S = rng.standard_normal((256, 256, 3, 32)).astype(np.float32)
        o = cbpdn.ConvBPDNJoint.Options({'RelStopTol': 0.0, 'FastSolve': True})
        admm('cfg3a: ConvBPDNJoint 256x256x3, 8x8x64 (Cd=1), K=32, f32',
             cbpdn.ConvBPDNJoint(D, S, 0.1, 0.01, o), 50, 10)
    if 'cfg3b' in which:
        D = unit(rng.standard_normal((8, 8, 3, 64))).astype(np.float32)
        S = rng.standard_normal((256, 256, 3, 32)).astype(np.float32)
        o = cbpdn.ConvBPDNJoint.Options({'RelStopTol': 0.0, 'FastSolve': True})
        admm('cfg3b: ConvBPDNJoint 256x256x3, 8x8x3x64 (Cd=3), K=32, f32',
             cbpdn.ConvBPDNJoint(D, S, 0.1, 0.01, o), 50, 10)
    if 'f64' in which:
        D = unit(rng.standard_normal((8, 8, 64)))
        S = rng.standard_normal((256, 256, 8))
        o = cbpdn.ConvBPDN.Options({'RelStopTol': 0.0, 'FastSolve': True})
        admm('f64: ConvBPDN 256x256, 8x8x64, K=8, float64', cbpdn.ConvBPDN(D, S, 0.1, o, dimK=1), 50, 10)
    if 'cfg5' in which:
        from sporco_b200.dictlrn import cbpdndl
        D0 = rng.standard_normal((8, 8, 64)).astype(np.float32)
        S = rng.standard_normal((256, 256, 16)).astype(np.float32)
        o = cbpdndl.ConvBPDNDictLearn.Options({'MaxMainIter': 20})
        b = cbpdndl.ConvBPDNDictLearn(D0, S, 0.1, o)
        b.solve()
        b.opt['MaxMainIter'] = 100
        b.xstep._h.synchronize()
        t0 = time.perf_counter()
        b.solve()
        b.xstep._h.synchronize()
        dt = time.perf_counter() - t0
        # split: X step alone / D step alone, same sizes
        h = b.xstep._h
        t1 = time.perf_counter()
        for _ in range(50):
            b.run_xstep()
        h.synchronize()
        tx = (time.perf_counter() - t1) / 50
        t1 = time.perf_counter()
        for _ in range(50):
            b.post_xstep()
            b.run_dstep()
            b.post_dstep()
        h.synchronize()
        td = (time.perf_counter() - t1) / 50
        its = b.getitstat()
        print(json.dumps({'config': 'cfg5: ConvBPDNDictLearn 256x256x16, 8x8x64, admm X / pgm D, f32',
                          'ms_per_outer_iter': dt / 100 * 1e3, 'outer_it_per_s': 100 / dt,
                          'ms_xstep': tx * 1e3, 'ms_dstep_incl_handover': td * 1e3,
                          'ObjFun_first_last': [float(its.ObjFun[0]), float(its.ObjFun[-1])]}))
        if 'cpu' in which:
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from oracle import cbpdn_oracle as co, cbpdndl_oracle as ocdl
            r = ocdl.cbpdndl(D0, S, 0.1, {'MaxMainIter': 2}, fft=co.FFTBackend('scipy', workers=os.cpu_count()))
            print(json.dumps({'config': 'cfg5 CPU oracle (scipy.fft workers=all), 2 outer iterations',
                              's_per_outer_iter': r['time'] / 2}))
    if 'cfg4' in which:
        D = unit(rng.standard_normal((12, 12, 128))).astype(np.float32)
        S = rng.standard_normal((512, 512)).astype(np.float32)
        o = pcbpdn.ConvBPDN.Options({'MaxMainIter': 10, 'RelStopTol': 0.0, 'L': 1.0,
                                     'Backtrack': BacktrackStandard(maxiter=15)})
        b = pcbpdn.ConvBPDN(D, S, 0.05, o)
        b.solve()
        b.opt['MaxMainIter'] = 50
        b._h.synchronize()
        t0 = time.perf_counter()
        b.solve()
        b._h.synchronize()
        dt = time.perf_counter() - t0
        its = b.getitstat()
        trials = float(np.sum(its.IterBTrack[-50:]))
        print(json.dumps({'config': 'cfg4: PGM ConvBPDN 512x512, 12x12x128, backtracking, f32',
                          'ms_per_iter': dt / 50 * 1e3, 'it_per_s': 50 / dt,
                          'trials_per_iter': trials / 50, 'ms_per_trial': dt / trials * 1e3,
                          'final_L': float(its.L[-1])}))


if __name__ == '__main__':
    main()

"""Schedule sweep at the metric configuration: the cross-iteration-fused schedule against the
wavefront schedule (SPCSC_WAVE=g,s) for several group sizes / stream counts.

    python tools/wave_sweep.py [--k 32] [--variants "fused;2,1;2,2;..."] [--keep 0,1]

For each variant: driver-style window (5 warm-up + 20 timed iterations from a cold start, rho
unsettled) and steady state (iterations 200..400), plus the final rho as a parity fingerprint."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sporco_b200.admm import cbpdn            # noqa: E402


def run(D, S, env):
    for k in [k for k in os.environ if k.startswith('SPCSC_')]:
        os.environ.pop(k, None)
    os.environ.update(env)
    o = cbpdn.ConvBPDN.Options({'RelStopTol': 0.0, 'FastSolve': True, 'AutoRho': {'Enabled': True}})
    b = cbpdn.ConvBPDN(D, S, 0.1, o, dimK=1)
    h = b._h
    h.admm_configure(**b._admm_config())
    h.admm_iterate(5, False)
    h.admm_iterate(20, False)
    ms_drv, _ = h.admm_last_timing()
    h.admm_iterate(175, False)
    h.admm_iterate(200, False)
    ms_st, launches = h.admm_last_timing()
    rho = float(b.rho)
    ysum = float(np.abs(b.Y[:, :, 0, 0, :4]).sum()) if hasattr(b, 'Y') else 0.0
    out = {'env': env, 'sched': h.admm_schedule_info(), 'driver_ms': ms_drv / 20, 'driver_its': 20e3 / ms_drv, 'steady_ms': ms_st / 200,
           'steady_its': 200e3 / ms_st, 'rho': rho, 'ysum': ysum, 'launches_per_iter': launches / 200}
    if os.environ.get('WAVE_SWEEP_PROFILE'):
        out['kernel_ms'] = [round(x / 20, 4) for x in h.admm_profile(20)]
    del b
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--k', type=int, default=32)
    ap.add_argument('--variants', default='fused;nofuse;1,1;2,1;4,1;8,1;1,2;2,2;4,2;1,3;2,3;1,4;2,4')
    ap.add_argument('--keep', default='0,1')
    ap.add_argument('--persist', default='0')
    a = ap.parse_args()
    rng = np.random.default_rng(12345)
    D = rng.standard_normal((8, 8, 64)).astype(np.float32)
    D /= np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))
    S = rng.standard_normal((256, 256, a.k)).astype(np.float32)
    for v in a.variants.split(';'):
        if v == 'fused':
            print(json.dumps(run(D, S, {})), flush=True)
        elif v == 'nofuse':
            print(json.dumps(run(D, S, {'SPCSC_FUSE': '0'})), flush=True)
        elif v.startswith('env:'):          # env:KEY=VAL,KEY=VAL
            print(json.dumps(run(D, S, dict(kv.split('=') for kv in v[4:].split(',')))), flush=True)
        elif v.startswith('f:'):
            print(json.dumps(run(D, S, {'SPCSC_WAVE': v[2:], 'SPCSC_WAVE_FUSED': '1'})), flush=True)
        else:
            for keep in a.keep.split(','):
                for pers in a.persist.split(','):
                    env = {'SPCSC_WAVE': v, 'SPCSC_WAVE_KEEP': keep}
                    if pers != '0':
                        env['SPCSC_WAVE_PERSIST'] = pers
                    print(json.dumps(run(D, S, env)), flush=True)


if __name__ == '__main__':
    main()

"""Time signal.tikhonov_filter on the device (any-size transforms: 512 + 2*16 = 544 = 2^5 * 17 per axis).
    python tools/tikhonov_time.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sporco_b200 import _lib                     # noqa: E402


def main():
    rng = np.random.default_rng(1)
    out = []
    for n, npd, batch in ((512, 16, 8), (256, 16, 32)):
        x = rng.standard_normal((batch, n, n)).astype(np.float32)
        _lib.tikhonov_filter(x, 5.0, npd)
        t0 = time.perf_counter()
        for _ in range(3):
            sl, sh = _lib.tikhonov_filter(x, 5.0, npd)
        dt = (time.perf_counter() - t0) / 3
        out.append({'image': n, 'npd': npd, 'padded': n + 2 * npd, 'batch': batch, 'ms_per_call_incl_copies': dt * 1e3,
                    'ms_per_image': dt * 1e3 / batch})
    print(json.dumps({'tikhonov_filter': out}))


if __name__ == '__main__':
    main()

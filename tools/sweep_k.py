"""Per-image cost of one ADMM iteration as a function of the batch size (L2 residency effects).
    python tools/sweep_k.py [K ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sporco_b200.admm import cbpdn            # noqa: E402


def main():
    ks = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32]
    rng = np.random.default_rng(12345)
    D = rng.standard_normal((8, 8, 64)).astype(np.float32)
    D /= np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))
    for K in ks:
        S = rng.standard_normal((256, 256, K)).astype(np.float32)
        o = cbpdn.ConvBPDN.Options({'RelStopTol': 0.0, 'FastSolve': True})
        b = cbpdn.ConvBPDN(D, S, 0.1, o, dimK=1)
        h = b._h
        h.admm_configure(**b._admm_config())
        h.admm_iterate(40, False)
        h.admm_iterate(200, False)
        ms, _ = h.admm_last_timing()
        kms = [x / 40 for x in h.admm_profile(40)]
        print(json.dumps({'K': K, 'ms_per_iter': ms / 200, 'us_per_image': ms / 200 / K * 1e3,
                          'kernel_ms': [round(x, 4) for x in kms]}))
        del b


if __name__ == '__main__':
    main()

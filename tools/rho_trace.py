"""Trace of rho / r / s over a long run of the metric configuration (how often does AutoRho fire?)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sporco_b200 import _lib                  # noqa: E402
if os.environ.get('SPCSC_LIBRARY'):
    import ctypes
    _lib.use_library(_lib._declare(ctypes.CDLL(os.environ['SPCSC_LIBRARY'])))
from sporco_b200.admm import cbpdn            # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1050
    rng = np.random.default_rng(12345)
    D = rng.standard_normal((8, 8, 64)).astype(np.float32)
    D /= np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))
    S = rng.standard_normal((256, 256, K)).astype(np.float32)
    o = cbpdn.ConvBPDN.Options({'RelStopTol': 0.0, 'MaxMainIter': n})
    b = cbpdn.ConvBPDN(D, S, 0.1, o, dimK=1)
    b.solve()
    its = b.getitstat()
    rho = np.array(its.Rho, dtype=np.float64)
    r = np.array(its.PrimalRsdl, dtype=np.float64)
    s = np.array(its.DualRsdl, dtype=np.float64)
    chg = (np.diff(rho) != 0).astype(int)
    win = [int(chg[i:i + 50].sum()) for i in range(0, len(chg), 50)]
    print(json.dumps({'K': K, 'changes_per_50': win}))
    for i in list(range(40, n, 50)):
        print(i, '%.6f' % rho[i], '%.4e %.4e ratio %.3f' % (r[i], s[i], r[i] / s[i]))


if __name__ == '__main__':
    main()

"""Long trajectories: device vs the numpy oracle (float32 and float64) on a mid-size problem."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cbpdn_oracle as orc          # noqa: E402
from sporco_b200.admm import cbpdn            # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    rng = np.random.default_rng(12345)
    D = rng.standard_normal((8, 8, 64)).astype(np.float32)
    D /= np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))
    S = rng.standard_normal((64, N, 2)).astype(np.float32)
    opt = {'RelStopTol': 0.0, 'MaxMainIter': n}
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(opt), dimK=1)
    b.solve()
    its = b.getitstat()
    r32 = orc.admm_convbpdn(D, S, 0.1, opt=opt, dimK=1)
    r64 = orc.admm_convbpdn(D.astype(np.float64), S.astype(np.float64), 0.1, opt=opt, dimK=1)
    for i in list(range(0, n, 40)) + [n - 1]:
        print('%4d dev rho %.5f r %.3e s %.3e | f32 rho %.5f r %.3e s %.3e | f64 rho %.5f r %.3e s %.3e' % (
            i, its.Rho[i], its.PrimalRsdl[i], its.DualRsdl[i],
            r32.itstat[i][8], r32.itstat[i][4], r32.itstat[i][5],
            r64.itstat[i][8], r64.itstat[i][4], r64.itstat[i][5]))


if __name__ == '__main__':
    main()

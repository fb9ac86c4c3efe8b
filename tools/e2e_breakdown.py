"""Where the end-to-end time of the bench goes: construct (host D, S -> device, transforms), solve, coefficient
maps back to the host.    python tools/e2e_breakdown.py [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                    # noqa: E402
from sporco_b200.admm import cbpdn              # noqa: E402


def main():
    import torch
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    D, S = bench.make_inputs(bench.K_PER_GPU)
    opt = dict(bench.OPT_BENCH)
    opt['MaxMainIter'] = steps
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = cbpdn.ConvBPDN(D, S, bench.LMBDA, cbpdn.ConvBPDN.Options(opt), dimK=1)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        b.run()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        Y = b.getmin()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print('rep %d: construct %.2f ms, run(%d) %.2f ms, getmin (Y to host, %d MB) %.2f ms, total %.2f ms -> %.1f it/s'
              % (rep, 1e3 * (t1 - t0), steps, 1e3 * (t2 - t1), Y.nbytes >> 20, 1e3 * (t3 - t2), 1e3 * (t3 - t0),
                 steps / (t3 - t0)), flush=True)
        del b, Y


if __name__ == '__main__':
    main()

"""Accuracy at the literal metric configuration (256x256, 8x8x64, 32 images, AutoRho, float32), 10 iterations:
the device result and the float32 oracle (= the reference's arithmetic) against the float64 oracle.
    python tools/k32_accuracy.py [iters]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cbpdn_oracle as orc          # noqa: E402
from sporco_b200.admm import cbpdn              # noqa: E402


def rel(a, b):
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rng = np.random.default_rng(12345)
    D = rng.standard_normal((8, 8, 64))
    D /= np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))
    D = D.astype(np.float32)
    S = rng.standard_normal((256, 256, 32)).astype(np.float32)
    opt = {'MaxMainIter': iters, 'RelStopTol': 0.0, 'AutoRho': {'Enabled': True}}
    fft = orc.FFTBackend('scipy', os.cpu_count() or 8)
    b = cbpdn.ConvBPDN(D, S, 0.1, cbpdn.ConvBPDN.Options(opt), dimK=1)
    Yg = b.solve().copy()
    rho_g = np.asarray(b.getitstat().Rho, dtype=np.float64)
    r32 = orc.admm_convbpdn(D, S, 0.1, opt=opt, dimK=1, fft=fft)
    Y32 = r32.Y
    rho32 = np.array([row[8] for row in r32.itstat])
    del r32
    r64 = orc.admm_convbpdn(D.astype(np.float64), S.astype(np.float64), 0.1, opt=opt, dimK=1, fft=fft)
    Y64 = r64.Y
    rho64 = np.array([row[8] for row in r64.itstat])
    print(json.dumps({'iters': iters, 'gpu32_vs_orc32': rel(Yg, Y32), 'gpu32_vs_orc64': rel(Yg, Y64),
                      'orc32_vs_orc64': rel(Y32, Y64),
                      'rho_gpu': rho_g.tolist(), 'rho_orc32': rho32.tolist(), 'rho_orc64': rho64.tolist()}))


if __name__ == '__main__':
    main()

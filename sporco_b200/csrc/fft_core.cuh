// fft_core.cuh -- radix-2/4/8/16 butterflies and the shared-memory Stockham stage used by
// every transform in the library.  Replaces, on the device, what the reference delegates
// to FFTW / pocketfft through sporco/fft.py:257-314 (rfftn / irfftn over axes (0,1)).
//
// Conventions: forward = exp(-2*pi*i*jk/N), unnormalised; inverse = exp(+...), also
// unnormalised here (callers fold the 1/N into their own epilogue).
#pragma once

#include "platform.h"

namespace spcsc {

template <typename T, bool INV>
SPCSC_HD void dft2(C2<T>& a, C2<T>& b) {
    C2<T> t = a;
    a = t + b;
    b = t - b;
}

template <typename T, bool INV>
SPCSC_HD void dft4(C2<T>& a0, C2<T>& a1, C2<T>& a2, C2<T>& a3) {
    C2<T> t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = a1 - a3;
    a0 = t0 + t2;
    a2 = t0 - t2;
    if (INV) {
        a1 = t1 + mul_i(t3);
        a3 = t1 - mul_i(t3);
    } else {
        a1 = t1 - mul_i(t3);
        a3 = t1 + mul_i(t3);
    }
}

// exp(-+ 2 pi i e / 16) for compile-time e
template <typename T, bool INV, int E16>
SPCSC_HD C2<T> w16() {
    constexpr double c[16] = {1.0, 0.92387953251128675613, 0.70710678118654752440,
                              0.38268343236508977173, 0.0, -0.38268343236508977173,
                              -0.70710678118654752440, -0.92387953251128675613, -1.0,
                              -0.92387953251128675613, -0.70710678118654752440,
                              -0.38268343236508977173, 0.0, 0.38268343236508977173,
                              0.70710678118654752440, 0.92387953251128675613};
    constexpr int e = E16 & 15;
    constexpr double co = c[e];
    constexpr double si = c[(e + 12) & 15];   // sin(x) = cos(x - pi/2)
    return mk<T>((T)co, INV ? (T)si : (T)(-si));
}

template <typename T, int R, bool INV>
struct SmallDFT;

template <typename T, bool INV>
struct SmallDFT<T, 1, INV> {
    static SPCSC_HD void run(C2<T>*) {}
};
template <typename T, bool INV>
struct SmallDFT<T, 2, INV> {
    static SPCSC_HD void run(C2<T>* v) { dft2<T, INV>(v[0], v[1]); }
};
template <typename T, bool INV>
struct SmallDFT<T, 4, INV> {
    static SPCSC_HD void run(C2<T>* v) { dft4<T, INV>(v[0], v[1], v[2], v[3]); }
};
template <typename T, bool INV>
struct SmallDFT<T, 8, INV> {
    static SPCSC_HD void run(C2<T>* v) {
        // even / odd split, then one radix-2 level with w8^k
        dft4<T, INV>(v[0], v[2], v[4], v[6]);
        dft4<T, INV>(v[1], v[3], v[5], v[7]);
        C2<T> o1 = v[3] * w16<T, INV, 2>();
        C2<T> o2 = INV ? mul_i(v[5]) : mul_mi(v[5]);
        C2<T> o3 = v[7] * w16<T, INV, 6>();
        C2<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
        v[0] = e0 + o0;  v[4] = e0 - o0;
        v[1] = e1 + o1;  v[5] = e1 - o1;
        v[2] = e2 + o2;  v[6] = e2 - o2;
        v[3] = e3 + o3;  v[7] = e3 - o3;
    }
};
template <typename T, bool INV>
struct SmallDFT<T, 16, INV> {
    static SPCSC_HD void run(C2<T>* v) {
        // 16 = 4 x 4:  A_r = DFT4(v[r], v[r+4], v[r+8], v[r+12]);
        //              X[q + 4p] = DFT4 over r of (w16^(r q) A_r[q]) at p
        dft4<T, INV>(v[0], v[4], v[8], v[12]);
        dft4<T, INV>(v[1], v[5], v[9], v[13]);
        dft4<T, INV>(v[2], v[6], v[10], v[14]);
        dft4<T, INV>(v[3], v[7], v[11], v[15]);
        // A_r[q] now sits in v[r + 4q]
        v[5] = v[5] * w16<T, INV, 1>();
        v[9] = v[9] * w16<T, INV, 2>();
        v[13] = v[13] * w16<T, INV, 3>();
        v[6] = v[6] * w16<T, INV, 2>();
        v[10] = v[10] * w16<T, INV, 4>();
        v[14] = v[14] * w16<T, INV, 6>();
        v[7] = v[7] * w16<T, INV, 3>();
        v[11] = v[11] * w16<T, INV, 6>();
        v[15] = v[15] * w16<T, INV, 9>();
        // for each q: DFT4 over r of v[r + 4q]  -> results p land in v[p + 4q] = X[q + 4p]
        dft4<T, INV>(v[0], v[1], v[2], v[3]);
        dft4<T, INV>(v[4], v[5], v[6], v[7]);
        dft4<T, INV>(v[8], v[9], v[10], v[11]);
        dft4<T, INV>(v[12], v[13], v[14], v[15]);
        // transpose 4x4 so that v[k] = X[k]
        C2<T> t;
        t = v[1];  v[1] = v[4];   v[4] = t;
        t = v[2];  v[2] = v[8];   v[8] = t;
        t = v[3];  v[3] = v[12];  v[12] = t;
        t = v[6];  v[6] = v[9];   v[9] = t;
        t = v[7];  v[7] = v[13];  v[13] = t;
        t = v[11]; v[11] = v[14]; v[14] = t;
    }
};

// Elements handled per thread for a length-N transform.
template <typename T>
constexpr int fft_elems(int N) {
    return N >= 64 ? 8 : (N >= 16 ? 4 : 2);
}

// ------------------------------------------------------------------------------------
// One Stockham pass over a length-N sequence that lives in shared memory, executed by
// TPF = N/E cooperating threads (`t` is the thread's index inside that group).  All
// threads of the block must call it (it contains block-wide barriers); threads with
// active == false only take part in the barriers.
//   tw[j*TWS] must hold exp(-2 pi i j / N) for j < N.
// ------------------------------------------------------------------------------------
template <typename T, int N, int E, bool INV, int TWS, int Ns>
struct StockhamStage {
    static constexpr int REM = N / Ns;
    static constexpr int R = REM >= E ? E : REM;
    static constexpr int NB = E / R;
    static constexpr int TPF = N / E;

    static SPCSC_DEV void run(C2<T>* buf, int t, const C2<T>* SPCSC_RESTRICT tw,
                              bool active) {
        C2<T> v[NB][R];
        if (active) {
            SPCSC_UNROLL
            for (int i = 0; i < NB; ++i) {
                const int j = t + i * TPF;
                const int k = j & (Ns - 1);
                SPCSC_UNROLL
                for (int r = 0; r < R; ++r) v[i][r] = buf[j + r * (N / R)];
                if (Ns > 1) {
                    SPCSC_UNROLL
                    for (int r = 1; r < R; ++r) {
                        C2<T> w = tw[(size_t)(r * k) * (N / (Ns * R)) * TWS];
                        v[i][r] = INV ? mulc(v[i][r], w) : v[i][r] * w;
                    }
                }
                SmallDFT<T, R, INV>::run(v[i]);
            }
        }
        __syncthreads();
        if (active) {
            SPCSC_UNROLL
            for (int i = 0; i < NB; ++i) {
                const int j = t + i * TPF;
                const int k = j & (Ns - 1);
                const int j0 = (j - k) * R + k;
                SPCSC_UNROLL
                for (int r = 0; r < R; ++r) buf[j0 + r * Ns] = v[i][r];
            }
        }
        __syncthreads();
        if constexpr (Ns * R < N) {
            StockhamStage<T, N, E, INV, TWS, Ns * R>::run(buf, t, tw, active);
        }
    }
};

// Full in-place transform of a length-N sequence in shared memory (natural order in/out).
template <typename T, int N, bool INV, int TWS>
SPCSC_DEV void fft_smem(C2<T>* buf, int t, const C2<T>* SPCSC_RESTRICT tw, bool active) {
    constexpr int E = fft_elems<T>(N);
    if (N > 1) StockhamStage<T, N, E, INV, TWS, 1>::run(buf, t, tw, active);
}

template <typename T, int N>
constexpr int fft_tpf() { return N / fft_elems<T>(N); }

}  // namespace spcsc

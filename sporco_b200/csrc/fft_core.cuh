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


// =====================================================================================
// Register-resident transform (kernel set v2).
//
// A length-N transform is done by TPF = N/E lanes of ONE warp (TPF <= 32), each holding E
// elements in registers in the "strided" layout: slot p <-> element t + TPF*p.  The first
// stage (radix E) works straight on the registers; later stages exchange data through a
// per-transform shared-memory region of N elements with an XOR swizzle (conflict-free for
// the radix-16 x 16 plan, at most 2-way otherwise) and only warp-level barriers.  Input and
// output use the same strided layout, so forward -> pointwise -> inverse needs no shuffling.
// Stage twiddles come from a small table laid out [r][k] per stage (consecutive lanes read
// consecutive entries); make_stage_twiddles() on the host builds it with the same plan.
// =====================================================================================
// Exchange regions are padded by one element per 16 (index i lives at i + i/16): the radix-16
// first-stage writes of a lane group (stride 16) then fall on distinct banks, and -- unlike an
// XOR swizzle -- every access of a stage is "lane base + compile-time offset", so the exchanges
// cost no integer arithmetic per element.  A region therefore needs fft_region(N) elements.
SPCSC_HD int fft_pad(int i) { return i + (i >> 4); }
constexpr int fft_region(int N) { return N + N / 16; }

// total number of stage-twiddle entries of the (N, E) plan
constexpr int stage_tw_len(int N, int E) {
    int len = 0;
    int Ns = (E < N ? E : N);
    while (Ns < N) {
        int R = (N / Ns >= E) ? E : N / Ns;
        len += R * Ns;
        Ns *= R;
    }
    return len;
}

// TWR: the (single) twiddled stage of a two-stage plan takes its factors from a thread-local array laid out
// [i][r] (see load_stage_tw_regs) instead of the shared-memory table: they depend only on the lane, so a
// persistent kernel loads them once and keeps them in registers.
template <typename T, int N, int E, bool INV, int Ns, int TWOFF, bool TWR = false>
struct RegStage {
    static constexpr int REM = N / Ns;
    static constexpr int R = REM >= E ? E : REM;
    static constexpr int NB = E / R;
    static constexpr int TPF = N / E;
    static constexpr bool LAST = (Ns * R >= N);
    static constexpr int RS = N / R;                         // read stride
    static constexpr bool RD_FAST = (RS % 16 == 0);          // r*RS never carries into the pad term
    static constexpr bool WR_FAST16 = (Ns % 16 == 0);
    static constexpr bool WR_BLOCK = (!WR_FAST16 && Ns * R <= 16 && 16 % (Ns * R) == 0);

    // v: E registers.  Ns == 1: strided input layout.  On return from the last stage v is
    // again in strided layout (slot p <-> X[t + TPF*p]).
    static SPCSC_DEV void run(C2<T>* v, C2<T>* buf, const C2<T>* SPCSC_RESTRICT stw, int t) {
        if (Ns > 1) {
            SPCSC_UNROLL
            for (int i = 0; i < NB; ++i) {
                const int j = t + i * TPF;
                const int k = j & (Ns - 1);
                const C2<T>* rd = buf + fft_pad(j);
                SPCSC_UNROLL
                for (int r = 0; r < R; ++r) {
                    C2<T> x = RD_FAST ? rd[r * (RS + RS / 16)] : buf[fft_pad(j + r * RS)];
                    if (r > 0) {
                        const C2<T> w = TWR ? stw[i * R + r] : stw[TWOFF + r * Ns + k];
                        x = INV ? mulc(x, w) : x * w;
                    }
                    v[i * R + r] = x;
                }
            }
        }
        SPCSC_UNROLL
        for (int i = 0; i < NB; ++i) SmallDFT<T, R, INV>::run(v + i * R);
        if constexpr (!LAST) {
            if (Ns > 1) __syncwarp();          // everyone has read before anyone overwrites
            SPCSC_UNROLL
            for (int i = 0; i < NB; ++i) {
                const int j = t + i * TPF;
                const int k = j & (Ns - 1);
                const int j0 = (j - k) * R + k;
                if (WR_FAST16) {
                    C2<T>* wr = buf + fft_pad(j0);
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) wr[r * (Ns + Ns / 16)] = v[i * R + r];
                } else if (WR_BLOCK) {
                    // j0 + r*Ns stays inside the 16-aligned block that contains j0 - k
                    C2<T>* wr = buf + j0 + ((j0 - k) >> 4);
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) wr[r * Ns] = v[i * R + r];
                } else {
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) buf[fft_pad(j0 + r * Ns)] = v[i * R + r];
                }
            }
            __syncwarp();
            static_assert(!TWR || Ns == 1, "register twiddles: two-stage plans only");
            RegStage<T, N, E, INV, Ns * R, TWOFF + (Ns > 1 ? R * Ns : 0), TWR>::run(v, buf, stw, t);
        } else {
            // slot i*R + r holds X[t + TPF*(i + NB*r)]: rename registers into strided order
            if (NB > 1) {
                C2<T> o[E];
                SPCSC_UNROLL
                for (int i = 0; i < NB; ++i) {
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) o[i + NB * r] = v[i * R + r];
                }
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) v[p] = o[p];
            }
        }
    }
};

// Two transforms at once (the two columns a lane group of the column kernels holds): same plan, the exchange
// moves both columns' values as one 16-byte element (half the shared-memory instructions of two separate
// transforms, one stage-twiddle fetch for both), and the butterflies of the two columns interleave.
template <typename T>
struct alignas(4 * sizeof(T)) C4 {
    C2<T> a, b;
};
template <typename T> SPCSC_HD C4<T> mk4(C2<T> a, C2<T> b) { C4<T> r; r.a = a; r.b = b; return r; }

template <typename T, int N, int E, bool INV, int Ns, int TWOFF>
struct RegStage2 {
    static constexpr int REM = N / Ns;
    static constexpr int R = REM >= E ? E : REM;
    static constexpr int NB = E / R;
    static constexpr int TPF = N / E;
    static constexpr bool LAST = (Ns * R >= N);
    static constexpr int RS = N / R;
    static constexpr bool RD_FAST = (RS % 16 == 0);
    static constexpr bool WR_FAST16 = (Ns % 16 == 0);
    static constexpr bool WR_BLOCK = (!WR_FAST16 && Ns * R <= 16 && 16 % (Ns * R) == 0);

    static SPCSC_DEV void run(C2<T>* v0, C2<T>* v1, C4<T>* buf, const C2<T>* SPCSC_RESTRICT stw, int t) {
        if (Ns > 1) {
            SPCSC_UNROLL
            for (int i = 0; i < NB; ++i) {
                const int j = t + i * TPF;
                const int k = j & (Ns - 1);
                const C4<T>* rd = buf + fft_pad(j);
                SPCSC_UNROLL
                for (int r = 0; r < R; ++r) {
                    C4<T> x = RD_FAST ? rd[r * (RS + RS / 16)] : buf[fft_pad(j + r * RS)];
                    if (r > 0) {
                        const C2<T> w = stw[TWOFF + r * Ns + k];
                        x.a = INV ? mulc(x.a, w) : x.a * w;
                        x.b = INV ? mulc(x.b, w) : x.b * w;
                    }
                    v0[i * R + r] = x.a;
                    v1[i * R + r] = x.b;
                }
            }
        }
        SPCSC_UNROLL
        for (int i = 0; i < NB; ++i) {
            SmallDFT<T, R, INV>::run(v0 + i * R);
            SmallDFT<T, R, INV>::run(v1 + i * R);
        }
        if constexpr (!LAST) {
            if (Ns > 1) __syncwarp();
            SPCSC_UNROLL
            for (int i = 0; i < NB; ++i) {
                const int j = t + i * TPF;
                const int k = j & (Ns - 1);
                const int j0 = (j - k) * R + k;
                if (WR_FAST16) {
                    C4<T>* wr = buf + fft_pad(j0);
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) wr[r * (Ns + Ns / 16)] = mk4<T>(v0[i * R + r], v1[i * R + r]);
                } else if (WR_BLOCK) {
                    C4<T>* wr = buf + j0 + ((j0 - k) >> 4);
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) wr[r * Ns] = mk4<T>(v0[i * R + r], v1[i * R + r]);
                } else {
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) buf[fft_pad(j0 + r * Ns)] = mk4<T>(v0[i * R + r], v1[i * R + r]);
                }
            }
            __syncwarp();
            RegStage2<T, N, E, INV, Ns * R, TWOFF + (Ns > 1 ? R * Ns : 0)>::run(v0, v1, buf, stw, t);
        } else {
            if (NB > 1) {
                C2<T> o[E];
                SPCSC_UNROLL
                for (int i = 0; i < NB; ++i) {
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) o[i + NB * r] = v0[i * R + r];
                }
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) v0[p] = o[p];
                SPCSC_UNROLL
                for (int i = 0; i < NB; ++i) {
                    SPCSC_UNROLL
                    for (int r = 0; r < R; ++r) o[i + NB * r] = v1[i * R + r];
                }
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) v1[p] = o[p];
            }
        }
    }
};
template <typename T, int N, int E, bool INV>
SPCSC_DEV void fft_regs2(C2<T>* v0, C2<T>* v1, C4<T>* buf, const C2<T>* SPCSC_RESTRICT stw, int t) {
    static_assert(E <= N && (N / E) <= 32, "plan must fit in one warp");
    RegStage2<T, N, E, INV, 1, 0>::run(v0, v1, buf, stw, t);
}

// v in strided layout -> transform -> v in strided layout.  `buf`: this transform's N-element
// shared-memory region; `stw`: stage twiddle table of the (N, E) plan (forward sign).
// All lanes of the warp must call it together.
template <typename T, int N, int E, bool INV>
SPCSC_DEV void fft_regs(C2<T>* v, C2<T>* buf, const C2<T>* SPCSC_RESTRICT stw, int t) {
    static_assert(E <= N && (N / E) <= 32, "plan must fit in one warp");
    RegStage<T, N, E, INV, 1, 0>::run(v, buf, stw, t);
}

// Two-stage plans (E < N <= E*E) with the second stage's factors in registers: twr[E] is filled once per
// thread by load_stage_tw_regs from the (N, E) table and then passed to fft_regs_twr for every transform.
template <int N, int E>
constexpr bool fft_two_stage() { return N > E && N <= E * E; }
template <typename T, int N, int E>
SPCSC_DEV void load_stage_tw_regs(C2<T>* twr, const C2<T>* SPCSC_RESTRICT stw, int t) {
    constexpr int Ns = E, R = N / E, NB = E / R, TPF = N / E;
    SPCSC_UNROLL
    for (int i = 0; i < NB; ++i) {
        const int k = (t + i * TPF) & (Ns - 1);
        SPCSC_UNROLL
        for (int r = 0; r < R; ++r) twr[i * R + r] = stw[r * Ns + k];
    }
}
template <typename T, int N, int E, bool INV>
SPCSC_DEV void fft_regs_twr(C2<T>* v, C2<T>* buf, const C2<T>* twr, int t) {
    static_assert(fft_two_stage<N, E>() && (N / E) <= 32, "two-stage plan in one warp");
    RegStage<T, N, E, INV, 1, 0, true>::run(v, buf, twr, t);
}

}  // namespace spcsc

#ifndef SPCSC_FFT_HOST_ONLY
#include <vector>
namespace spcsc {
// Host: stage twiddle table of the (N, E) plan, forward sign, layout per stage [r][k].
template <typename T>
inline std::vector<C2<T>> make_stage_twiddles(int N, int E) {
    std::vector<C2<T>> tab;
    const double two_pi = 6.283185307179586476925286766559;
    int Ns = (E < N ? E : N);
    while (Ns < N) {
        const int R = (N / Ns >= E) ? E : N / Ns;
        for (int r = 0; r < R; ++r)
            for (int k = 0; k < Ns; ++k) {
                const double a = -two_pi * (double)(r * k) / (double)(Ns * R);
                tab.push_back(mk<T>((T)std::cos(a), (T)std::sin(a)));
            }
        Ns *= R;
    }
    if (tab.empty()) tab.push_back(mk<T>(1, 0));
    return tab;
}
}  // namespace spcsc
#endif

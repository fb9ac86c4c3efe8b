// kernels.cuh -- device kernels of the ConvBPDN hot path.
//
// Device data layout (all arrays dense, C order, spatial axis fastest):
//   coefficient-shaped real arrays  Y, U           [K][Cx][M][N0][N1]
//   row-spectrum / column slabs     Zt             [K][Cx][N1f][M][N0]   complex, N1f = N1/2+1
//   dictionary spectrum             Df             [Cd][N1f][M][N0]      complex
//   signal spectrum                 Sf             [K][C][N1f][N0]       complex
//   Gram of the dictionary          G              [N1f][N0][Cd][Cd]     complex (Cd==1: real part = sum_m |Df|^2)
// The reference keeps (N0,N1,C,K,M) with M fastest (sporco/cnvrep.py:104-198); conversion
// happens once on upload / download (k_to_internal / k_from_internal).
//
// One ADMM iteration (sporco/admm/admm.py:331-377) is three data kernels + one scalar kernel:
//   k_row_fwd       Y-U, real row FFT, transposed store         (admm/cbpdn.py:271-273, first half of rfftn)
//   k_col           column FFT + Sherman-Morrison solve + column IFFT  (admm/cbpdn.py:273-281, linalg.py:232-297 / 370-444)
//   k_row_inv_prox  row c2r IFFT + relax + prox + dual update + residual sums
//                   (admm/admm.py:877-885, admm/cbpdn.py:614-620 / 785-794 / 297-311, admm/admm.py:434-437, 462-486)
//   k_admm_scalars  residuals, stopping test, rho update  (admm/admm.py:462-486, 549-575)
#pragma once

#include "fft_core.cuh"

namespace spcsc {

// ------------------------------------------------------------------------------------
// Device-resident solver scalars.
// ------------------------------------------------------------------------------------
// Accumulator block on the device: ACC_N doubles (floating-point atomics: objective / diagnostic
// sums) followed by kAccDet rows of kDetBins 64-bit integer bins (the sums that steer the
// algorithm -- residual norms and regularisation terms -- accumulated order-independently).
constexpr int kAccDet = 7;
constexpr size_t kAccBytes = 16 * sizeof(double) + (size_t)kAccDet * kDetBins * sizeof(unsigned long long);
enum { ACC_X2 = 0, ACC_Y2, ACC_U2, ACC_R2, ACC_S2, ACC_L1, ACC_L21, ACC_DFID,
       ACC_AX2, ACC_B2, ACC_AXB2, ACC_PGM_FY, ACC_PGM_F, ACC_PGM_LIN, ACC_PGM_DXY2,
       ACC_PGM_RSDL, ACC_N = 16 };
// ConvBPDNGradReg (ADMM only, so the PGM slots are free): sum_m w_m GHG |x_m|^2, Hermitian-weighted
enum { ACC_RGRAD = ACC_PGM_FY };   // (SOLVE 4 writes ACC_PGM_F / _LIN itself)

template <typename T>
struct AdmmState {
    T rho;          // penalty parameter used by the next x-step
    T udiv;         // pending dual rescale: U_effective = U_stored / udiv
    int k;          // iterations completed
    int stopped;    // 1 once the residual stopping test has fired
    int zt_stale;   // 1: the pre-computed row spectra of (Y - U) do not match the current U scaling (or were not made)
    int emit;       // 1: this iteration's prox kernel also produces the next x-step's row spectra (fused schedule)
};

template <typename T>
struct AdmmParams {
    T lmbda, mu, rlx, tau, mur, xi;
    double abs_tol, rel_tol, n_x;        // n_x = number of elements of X (== Nc)
    double inv_n;                        // 1 / (N0*N1)
    int autorho, period, autoscaling, stdres;
    int need_rsdl, need_obj, joint, linsolve_check;
    int dfid_direct;                     // 1: ACC_DFID already holds the weighted |sum_m Df Yf - Sf|^2 (AuxVarObj)
    int enet;                            // 1: ConvElasticNet: rows carry RegL2 = ||x||^2 / 2 in the regl21 column
    T enet_mu;                           // its l2 weight: the x-step diagonal is enet_mu + rho
    int ams_m0;                          // AddMaskSim: filters m >= ams_m0 are the appended impulse maps (M: none)
    int gradreg;                         // 1: ConvBPDNGradReg: DFid = |q|^2 sums without the rho^2 factor, rows carry RegGrad
    int emit_policy;                     // cross-iteration fusion: 0 always emit the next row spectra, 1 skip it in the
                                         // iteration after a change of rho (changes come in runs; a skipped emission
                                         // costs one row-forward pass, a wasted one costs that pass AND the emission)
};

struct StatRow {
    double k, obj, dfid, regl1, regl21, r, s, epri, edua, rho, xrrs, pad;
};

// l1 / l2,1 weight with numpy-style broadcasting over the internal index (k, c, m, n0, n1)
template <typename T>
struct WeightView {
    const T* p;
    long long sk, sc, sm, s0, s1;
    int spatial_uniform;     // s0 == s1 == 0
};

// ------------------------------------------------------------------------------------
// k_row_fwd:  Zt[b][wf][m][h] = rfft_row( A[b][m][h][:] - B[b][m][h][:] / udiv )
//   A, B real [nb][M][N0][N1]; B may be null.  One CTA = TR consecutive rows of one (b,m).
//   Real transform of length N1 = 2H done as a complex transform of length H on
//   (even, odd) pairs followed by the usual split.
// ------------------------------------------------------------------------------------
template <typename T, int H>
SPCSC_GLOBAL void k_row_fwd(const T* SPCSC_RESTRICT A, const T* SPCSC_RESTRICT B,
                            const AdmmState<T>* SPCSC_RESTRICT st, C2<T>* SPCSC_RESTRICT Zt,
                            const C2<T>* SPCSC_RESTRICT tw, int N0, int M, int TR) {
    if (st && st->stopped) return;
    SPCSC_DYN_SMEM(smem_raw);
    C2<T>* buf = reinterpret_cast<C2<T>*>(smem_raw);
    constexpr int P = H + 1;
    constexpr int TPF = fft_tpf<T, H>();
    const int tid = threadIdx.x, nt = blockDim.x;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, b = blockIdx.z;
    const T udiv = (st && B) ? st->udiv : (T)1;

    const size_t rowbase = (((size_t)b * M + m) * N0 + h0) * H;   // in C2 units
    const C2<T>* A2 = reinterpret_cast<const C2<T>*>(A) + rowbase;
    const C2<T>* B2 = B ? reinterpret_cast<const C2<T>*>(B) + rowbase : nullptr;
    for (int e = tid; e < TR * H; e += nt) {
        const int r = e / H, j = e - r * H;
        C2<T> z = A2[e];
        if (B2) {
            C2<T> u = B2[e];
            z.re -= u.re / udiv;
            z.im -= u.im / udiv;
        }
        buf[r * P + j] = z;
    }
    __syncthreads();
    {
        const int row = tid / TPF, t = tid - row * TPF;
        const bool active = row < TR;
        fft_smem<T, H, false, 2>(buf + (active ? row : 0) * P, t, tw, active);
    }
    const int N1f = H + 1;
    C2<T>* out = Zt + (((size_t)b * N1f) * M + m) * N0 + h0;
    const size_t wstride = (size_t)M * N0;
    for (int e = tid; e < TR * N1f; e += nt) {
        const int wf = e / TR, r = e - wf * TR;
        const C2<T> a = buf[r * P + (wf == H ? 0 : wf)];
        const C2<T> bb = conj(buf[r * P + (wf == 0 ? 0 : H - wf)]);
        const C2<T> w = tw[wf];
        const C2<T> sum = a + bb, dif = mul_mi((a - bb) * w);
        out[wf * wstride + r] = mk<T>((T)0.5 * (sum.re + dif.re), (T)0.5 * (sum.im + dif.im));
    }
}

// ------------------------------------------------------------------------------------
// Shared pieces of the row-inverse kernels: gather TR rows of one (b,m) from the slab
// layout, undo the real-transform split, inverse complex FFT of length H.  Afterwards
// buf[r*P + j] * scale = (x[h0+r][2j], x[h0+r][2j+1]).
// ------------------------------------------------------------------------------------
template <typename T, int H>
SPCSC_DEV void row_inverse_to_smem(C2<T>* buf, const C2<T>* SPCSC_RESTRICT Zt,
                                   const C2<T>* SPCSC_RESTRICT tw, int b, int m, int h0,
                                   int N0, int M, int TR) {
    constexpr int P = H + 1;
    constexpr int TPF = fft_tpf<T, H>();
    constexpr int N1f = H + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const C2<T>* in = Zt + (((size_t)b * N1f) * M + m) * N0 + h0;
    const size_t wstride = (size_t)M * N0;
    for (int e = tid; e < TR * N1f; e += nt) {
        const int wf = e / TR, r = e - wf * TR;
        buf[r * P + wf] = in[wf * wstride + r];
    }
    __syncthreads();
    constexpr int NP = H / 2 + 1;           // pairs (kk, H-kk), kk = 0..H/2
    for (int e = tid; e < TR * NP; e += nt) {
        const int r = e / NP, kk = e - r * NP;
        C2<T>* row = buf + r * P;
        if (kk == 0) {
            const T a = row[0].re, c = row[H].re;    // c2r ignores the imaginary parts
            row[0] = mk<T>(a + c, a - c);
        } else {
            const C2<T> Xa = row[kk], Xb = row[H - kk];
            const C2<T> w = tw[kk];
            // Z[kk]   = (Xa + conj Xb) + i conj(w) (Xa - conj Xb)
            // Z[H-kk] = (Xb + conj Xa) - i w      (Xb - conj Xa)
            const C2<T> s1 = Xa + conj(Xb), d1 = Xa - conj(Xb);
            const C2<T> s2 = Xb + conj(Xa), d2 = Xb - conj(Xa);
            row[kk] = s1 + mul_i(mulc(d1, w));
            if (H - kk != kk) row[H - kk] = s2 - mul_i(d2 * w);
        }
    }
    __syncthreads();
    const int rowi = tid / TPF, t = tid - rowi * TPF;
    const bool active = rowi < TR;
    fft_smem<T, H, true, 2>(buf + (active ? rowi : 0) * P, t, tw, active);
}

// k_row_inv: X[b][m][h][:] = irfft_row(Zt) * scale   (plain inverse; used for get X,
// reconstruction and the unit irfft2 entry point)
template <typename T, int H>
SPCSC_GLOBAL void k_row_inv(const C2<T>* SPCSC_RESTRICT Zt, T* SPCSC_RESTRICT X,
                            const C2<T>* SPCSC_RESTRICT tw, int N0, int M, int TR, T scale) {
    SPCSC_DYN_SMEM(smem_raw);
    C2<T>* buf = reinterpret_cast<C2<T>*>(smem_raw);
    constexpr int P = H + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, b = blockIdx.z;
    row_inverse_to_smem<T, H>(buf, Zt, tw, b, m, h0, N0, M, TR);
    C2<T>* X2 = reinterpret_cast<C2<T>*>(X) + (((size_t)b * M + m) * N0 + h0) * H;
    for (int e = tid; e < TR * H; e += nt) {
        const int r = e / H, j = e - r * H;
        C2<T> z = buf[r * P + j];
        X2[e] = mk<T>(z.re * scale, z.im * scale);
    }
}

template <typename T>
SPCSC_DEV T soft_threshold(T v, T thr) {
    const T t = fabs(v) - thr;
    return t > (T)0 ? copysign(t, v) : (T)0;
}

// ------------------------------------------------------------------------------------
// k_row_inv_prox: for TR rows of one (k, m), all Cx channels:
//   X = irfft_row(Zt)*scale ; AX = rlx X + (1-rlx) Y ; V = AX + U/udiv
//   Y' = prox(V) (+NonNeg, +NoBndryCross) ; U' = U/udiv + (AX - Y') ; residual sums.
// ------------------------------------------------------------------------------------
template <typename T, int H, int CX>
SPCSC_GLOBAL void k_row_inv_prox(const C2<T>* SPCSC_RESTRICT Zt, T* SPCSC_RESTRICT Y,
                                 T* SPCSC_RESTRICT U, const AdmmState<T>* SPCSC_RESTRICT st,
                                 AdmmParams<T> prm, WeightView<T> wl1, WeightView<T> wl21,
                                 double* SPCSC_RESTRICT acc, const C2<T>* SPCSC_RESTRICT tw,
                                 int N0, int M, int TR, T scale, int nonneg,
                                 int bnd0, int bnd1, int reg_on_y) {
    if (st->stopped) return;
    constexpr int Cx = CX;
    SPCSC_DYN_SMEM(smem_raw);
    C2<T>* buf = reinterpret_cast<C2<T>*>(smem_raw);
    constexpr int P = H + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, k = blockIdx.z;
    const bool ams = m >= prm.ams_m0;   // additive-mask-simulation map: no clipping, not part of RegL1
    const int N1 = 2 * H;
    const size_t cbuf = (size_t)TR * P;

    for (int c = 0; c < Cx; ++c)
        row_inverse_to_smem<T, H>(buf + c * cbuf, Zt, tw, k * Cx + c, m, h0, N0, M, TR);
    __syncthreads();

    const T rho = st->rho, udiv = st->udiv;
    const T lr = prm.lmbda / rho;
    const T mr = prm.joint ? prm.mu / rho : (T)0;
    const T rlx = prm.rlx;
    double sums[7] = {0, 0, 0, 0, 0, 0, 0};

    for (int e = tid; e < TR * H; e += nt) {
        const int r = e / H, j = e - r * H;
        const int h = h0 + r;
        T a2[2] = {0, 0}, g2[2] = {0, 0};
        T wv[CX][2], ax[CX][2], ue[CX][2], xv[CX][2], yp[CX][2];
        SPCSC_UNROLL
        for (int c = 0; c < Cx; ++c) {
            const size_t off = ((((size_t)(k * Cx + c) * M + m) * N0 + h) * H + j);
            const C2<T> z = buf[c * cbuf + r * P + j];
            const C2<T> y2 = reinterpret_cast<const C2<T>*>(Y)[off];
            const C2<T> u2 = reinterpret_cast<const C2<T>*>(U)[off];
            const T xs[2] = {z.re * scale, z.im * scale};
            const T ys[2] = {y2.re, y2.im};
            const T us[2] = {u2.re / udiv, u2.im / udiv};
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                const T w1 = wl1.p[(size_t)k * wl1.sk + (size_t)c * wl1.sc + (size_t)m * wl1.sm +
                                  (size_t)h * wl1.s0 + (size_t)(2 * j + q) * wl1.s1];
                const T axv = (rlx == (T)1) ? xs[q] : rlx * xs[q] + ((T)1 - rlx) * ys[q];
                const T v = axv + us[q];
                const T w = soft_threshold(v, lr * w1);
                xv[c][q] = xs[q];
                yp[c][q] = ys[q];
                ax[c][q] = axv;
                ue[c][q] = us[q];
                wv[c][q] = w;
                a2[q] += w * w;
                if (!reg_on_y) {
                    sums[ACC_L1] += ams ? 0 : (double)fabs(w1 * xs[q]);
                    g2[q] += xs[q] * xs[q];
                }
            }
        }
        T fac[2] = {1, 1};
        T w21[2] = {1, 1};
        if (prm.joint) {
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                w21[q] = wl21.p[(size_t)k * wl21.sk + (size_t)m * wl21.sm + (size_t)h * wl21.s0 +
                                (size_t)(2 * j + q) * wl21.s1];
                const T a = sqrt(a2[q]);
                const T bq = fmax((T)0, a - mr * w21[q]);
                fac[q] = (a != (T)0) ? bq / a : (T)0;
            }
        }
        SPCSC_UNROLL
        for (int c = 0; c < Cx; ++c) {
            const size_t off = ((((size_t)(k * Cx + c) * M + m) * N0 + h) * H + j);
            T yn[2], un[2];
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                T y = prm.joint ? fac[q] * wv[c][q] : wv[c][q];
                if (nonneg && !ams && y < (T)0) y = (T)0;
                if (!ams && (h >= bnd0 || (2 * j + q) >= bnd1)) y = (T)0;
                const T u = ue[c][q] + (ax[c][q] - y);
                yn[q] = y;
                un[q] = u;
                const T x = xv[c][q];
                const T dr = x - y, ds = yp[c][q] - y;
                sums[ACC_X2] += (double)x * x;
                sums[ACC_Y2] += (double)y * y;
                sums[ACC_U2] += (double)u * u;
                sums[ACC_R2] += (double)dr * dr;
                sums[ACC_S2] += (double)ds * ds;
                if (reg_on_y) {
                    const T w1 = wl1.p[(size_t)k * wl1.sk + (size_t)c * wl1.sc +
                                      (size_t)m * wl1.sm + (size_t)h * wl1.s0 +
                                      (size_t)(2 * j + q) * wl1.s1];
                    sums[ACC_L1] += ams ? 0 : (double)fabs(w1 * y);
                    g2[q] += y * y;
                }
            }
            reinterpret_cast<C2<T>*>(Y)[off] = mk<T>(yn[0], yn[1]);
            reinterpret_cast<C2<T>*>(U)[off] = mk<T>(un[0], un[1]);
        }
        if (prm.joint) {
            sums[ACC_L21] += (double)(w21[0] * sqrt(g2[0])) + (double)(w21[1] * sqrt(g2[1]));
        }
    }
    if (prm.need_rsdl || prm.need_obj) {
        double* red = reinterpret_cast<double*>(smem_raw);
        block_accumulate_det<7>(sums, red, reinterpret_cast<unsigned long long*>(acc + ACC_N));
    }
}

// ------------------------------------------------------------------------------------
// k_col: one CTA per (wf, b) slab of M columns x N0.
//   DO_FWD   : forward FFT along the column before the solve
//   SOLVE    : 0 none, 1 ADMM  q = (rho I + G)^-1 (Sf - s),  2 gradient  q = (Sf - s)/L
//              (PGM: also stores s in `sumout` and accumulates sum|Sf - s|^2),
//              3 sum only: write s_c = sum_m Df_c * col to `sumout` ([nb][Cd][N1f][N0]),
//              4 PGM evaluation of a candidate Xf: no update; accumulates sum|s - Sf|^2 (plain and
//                Hermitian-weighted), sum Re(conj(s - sY)(sY - Sf)) with sY read from `sumin`,
//                and sum|Xf - Yf|^2 (plain and weighted) against the slabs in `ref`
//   DO_INV   : inverse FFT (unnormalised) along the column after the update
//   out = in + sum_c conj(Df_c) q_c       with s_c = sum_m Df_c[m] in[m]
// The slab is processed in chunks of MC columns that fit shared memory; when the whole
// slab fits (nchunk == 1) it stays resident between the reduction and the update.
// ------------------------------------------------------------------------------------
struct ColArgs {
    int N0, M, Cd, Cs;     // Cs: channel count of Sf (C); b -> (k, c) = (b / Cx, b % Cx)
    int Cx, N1f, MC, nchunk, parts;
    int dfid_on, even_n1, check_on;
    int ntiles;            // k_col2: number of (wf, b) slabs = N1f * nb
    int gradreg;           // k_col SOLVE 1 / 4: ConvBPDNGradReg (G.im = GHG, sumin[m] = (mu w_m, w_m))
    int pgm_mask;          // k_col SOLVE 2 / 4: masked data fidelity (pgm.ConvBPDNMask): the residual spectra
                           // W^2-filtered in the signal domain arrive ready made (SOLVE 2: sumin, SOLVE 4: G)
    // consensus dictionary update (admm.ccmod.ConvCnstrMOD_Consensus): every slab b has its OWN "dictionary" --
    // the coefficient spectra of block b / df_bdiv -- and Gram row: Df + (b / df_bdiv) df_bstride, likewise G.
    // 0: one dictionary shared by all slabs (ConvBPDN)
    long long df_bstride, g_bstride;
    int df_bdiv;
};

template <typename T, int MAXCD>
SPCSC_DEV void hpd_solve(C2<T> (&A)[MAXCD][MAXCD], C2<T> (&bvec)[MAXCD], int n) {
    // Gaussian elimination without pivoting on a Hermitian positive definite system.
    for (int i = 0; i < n; ++i) {
        const T inv = (T)1 / A[i][i].re;
        for (int r = i + 1; r < n; ++r) {
            const C2<T> f = inv * A[r][i];
            for (int c = i; c < n; ++c) A[r][c] = A[r][c] - f * A[i][c];
            bvec[r] = bvec[r] - f * bvec[i];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        C2<T> s = bvec[i];
        for (int c = i + 1; c < n; ++c) s = s - A[i][c] * bvec[c];
        const T d = abs2(A[i][i]);
        bvec[i] = mk<T>((s.re * A[i][i].re + s.im * A[i][i].im) / d,
                        (s.im * A[i][i].re - s.re * A[i][i].im) / d);
    }
}

template <typename T>
SPCSC_DEV void col_load_chunk(C2<T>* buf, const C2<T>* SPCSC_RESTRICT src, int m0, int mc, int N0) {
    for (int e = threadIdx.x; e < mc * N0; e += blockDim.x) buf[e] = src[(size_t)m0 * N0 + e];
    __syncthreads();
}
// ---- any-length transforms: mixed-radix Stockham passes with run-time radices ----------------------------
// A length N whose prime factors are all <= kGenMaxRadix is transformed in O(N sum(radices)) instead of the
// O(N^2) direct DFT: per pass every butterfly gathers R inputs, applies the inter-stage twiddles, evaluates
// the R-point DFT directly from the twiddle table (tw[j] = exp(-2 pi i j / N)) and scatters in autosort
// order; passes ping-pong between two buffers.  Used by the any-size row and column kernels (image sizes
// the reference accepts that are not powers of two, e.g. the padded sizes of signal.tikhonov_filter).
constexpr int kGenMaxRadix = 32;
SPCSC_HD int gen_factor(int n, int* rad) {
    int c = 0;
    while (n % 4 == 0) { rad[c++] = 4; n /= 4; }
    if (n % 2 == 0) { rad[c++] = 2; n /= 2; }
    for (int p = 3; p <= kGenMaxRadix && n > 1; p += 2)
        while (n % p == 0) { rad[c++] = p; n /= p; }
    return n == 1 ? c : 0;
}
// nseq sequences of length N stored back to back in `a`; `b` is scratch of the same size.  All threads of the
// block take part.  Returns the buffer that holds the result (natural order).
template <typename T, bool INV>
SPCSC_DEV C2<T>* gen_fft_batch(C2<T>* a, C2<T>* b, const C2<T>* SPCSC_RESTRICT tw, int nseq, int N,
                               const int* rad, int nrad) {
    int Ns = 1;
    for (int stg = 0; stg < nrad; ++stg) {
        const int R = rad[stg], nb = N / R;
        const int qs = N / R;                                   // table step of exp(-2 pi i / R)
        const int ts = N / (Ns * R);                            // table step of the inter-stage twiddle
        for (int e = threadIdx.x; e < nseq * nb; e += blockDim.x) {
            const int sq = e / nb, j = e - sq * nb;
            const int k = j % Ns;
            const C2<T>* x = a + (size_t)sq * N;
            C2<T> v[kGenMaxRadix];
            int ti = 0;
            for (int r = 0; r < R; ++r) {
                C2<T> xv = x[j + r * nb];
                if (ti != 0) {
                    const C2<T> w = tw[ti];
                    xv = INV ? mulc(xv, w) : xv * w;
                }
                v[r] = xv;
                ti += ts * k;                                   // r k N / (Ns R) < N
            }
            C2<T>* y = b + (size_t)sq * N + (size_t)(j - k) * R + k;
            for (int q = 0; q < R; ++q) {
                C2<T> s = v[0];
                int wi = 0;
                for (int r = 1; r < R; ++r) {
                    wi += q * qs;                               // (q r mod R) N / R
                    if (wi >= N) wi -= N;
                    const C2<T> w = tw[wi];
                    s = s + (INV ? mulc(v[r], w) : v[r] * w);
                }
                y[(size_t)q * Ns] = s;
            }
        }
        __syncthreads();
        C2<T>* t2 = a; a = b; b = t2;
        Ns *= R;
    }
    return a;
}
// Any-length column transform in place in `buf` (`buf2`: scratch of the same size): mixed radix when the
// length factors into small primes, else the direct DFT from the twiddle table.
template <typename T, bool INV>
SPCSC_DEV void col_dft_chunk(C2<T>* buf, C2<T>* buf2, const C2<T>* SPCSC_RESTRICT tw, int mc, int N0) {
    __shared__ int rad[16];
    __shared__ int nrad;
    if (threadIdx.x == 0) nrad = gen_factor(N0, rad);
    __syncthreads();
    if (nrad > 0) {
        C2<T>* res = gen_fft_batch<T, INV>(buf, buf2, tw, mc, N0, rad, nrad);
        if (res != buf) {
            for (int e = threadIdx.x; e < mc * N0; e += blockDim.x) buf[e] = buf2[e];
            __syncthreads();
        }
        return;
    }
    for (int e = threadIdx.x; e < mc * N0; e += blockDim.x) {
        const int col = e / N0, k = e - col * N0;
        const C2<T>* x = buf + (size_t)col * N0;
        C2<T> s = mk<T>(0, 0);
        int idx = 0;
        for (int n = 0; n < N0; ++n) {
            const C2<T> w = tw[idx];
            s = s + (INV ? mulc(x[n], w) : x[n] * w);
            idx += k;
            if (idx >= N0) idx -= N0;
        }
        buf2[e] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < mc * N0; e += blockDim.x) buf[e] = buf2[e];
    __syncthreads();
}
template <typename T, int N0, bool INV>
SPCSC_DEV void col_fft_chunk(C2<T>* buf, const C2<T>* SPCSC_RESTRICT tw, int mc, int MC) {
    constexpr int TPF = fft_tpf<T, N0>();
    const int tid = threadIdx.x, nt = blockDim.x;
    const int cols_per_pass = nt / TPF > 0 ? nt / TPF : 1;
    for (int c0 = 0; c0 < MC; c0 += cols_per_pass) {
        const int col = c0 + tid / TPF, t = tid % TPF;
        const bool active = (tid / TPF) < cols_per_pass && col < mc;
        fft_smem<T, N0, INV, 1>(buf + (size_t)(active ? col : 0) * N0, t, tw, active);
    }
}

template <typename T, int N0T, bool DO_FWD, int SOLVE, bool DO_INV>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS(sizeof(T) == 4 ? 1024 : 512) k_col(const C2<T>* SPCSC_RESTRICT in, C2<T>* SPCSC_RESTRICT out,
                        const C2<T>* SPCSC_RESTRICT Df, const C2<T>* SPCSC_RESTRICT Sf,
                        const C2<T>* SPCSC_RESTRICT G, C2<T>* SPCSC_RESTRICT sumout,
                        const C2<T>* SPCSC_RESTRICT sumin, const C2<T>* SPCSC_RESTRICT ref,
                        const AdmmState<T>* SPCSC_RESTRICT st, T Lstep,
                        double* SPCSC_RESTRICT acc, const C2<T>* SPCSC_RESTRICT tw,
                        ColArgs a) {
    if (st && st->stopped) return;
    SPCSC_DYN_SMEM(smem_raw);
    constexpr bool GEN = (N0T == 0);             // 0: run-time length, direct DFT (any size)
    const int N0 = GEN ? a.N0 : N0T;
    constexpr int MAXCD = 4;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int wf = blockIdx.x, b = blockIdx.y;
    const int M = a.M, Cd = a.Cd, MC = a.MC;
    if (a.df_bstride) {                      // per-slab dictionary (consensus dictionary update)
        const size_t ib = (size_t)(b / a.df_bdiv);
        Df += ib * (size_t)a.df_bstride;
        G += ib * (size_t)a.g_bstride;
    }
    C2<T>* buf = reinterpret_cast<C2<T>*>(smem_raw);                 // [MC][N0]
    C2<T>* buf2 = buf + (size_t)MC * N0;                              // [MC][N0], direct-DFT path only
    C2<T>* spart = buf + (size_t)MC * N0 * (GEN ? 2 : 1);             // [Cd][parts][N0]
    const size_t slab = (((size_t)b * a.N1f + wf) * M) * N0;
    const C2<T>* src = in + slab;
    C2<T>* dst = out ? out + slab : nullptr;
    const bool resident = (a.nchunk == 1);
    // SOLVE 1 with an l2 term (ConvElasticNet, admm/cbpdn.py:948-955; its weight arrives in Lstep):
    // (D^H D + rho_x I) x = D^H s + rho z, rho_x = mu + rho, is the plain system for z' = (rho/rho_x) z
    const T rho0 = (SOLVE == 1) ? st->rho : (T)0;
    const T rho = (SOLVE == 1) ? rho0 + Lstep : (T)0;
    const T zeta = (SOLVE == 1 && Lstep != (T)0) ? rho0 / rho : (T)1;

    if (SOLVE != 0) {
        // ---- phase A: s_c[h] = sum_m Df_c[m][h] * col[m][h]
        const int parts = a.parts;
        const int CdS = Cd + ((a.gradreg && SOLVE == 1) ? 1 : 0);     // + gamma pseudo-channel
        for (int e = tid; e < CdS * parts * N0; e += nt) spart[e] = mk<T>(0, 0);
        __syncthreads();
        for (int ch = 0; ch < a.nchunk; ++ch) {
            const int m0 = ch * MC, mc = (M - m0 < MC) ? M - m0 : MC;
            col_load_chunk<T>(buf, src, m0, mc, N0);
            if (DO_FWD) {
                if constexpr (GEN) col_dft_chunk<T, false>(buf, buf2, tw, mc, N0);
                else col_fft_chunk<T, N0T, false>(buf, tw, mc, MC);
            }
            if (zeta != (T)1) {
                for (int e = tid; e < mc * N0; e += nt) buf[e] = zeta * buf[e];
                __syncthreads();
            }
            if (a.gradreg && SOLVE == 1) {
                // diagonal d_m = mu w_m GHG + rho (admm/cbpdn.py:1173-1201, linalg.solvedbd_sm):
                // sigma = sum_m Df z / d, gamma = sum_m |Df|^2 / d (kept as pseudo-channel 1)
                for (int e = tid; e < parts * N0; e += nt) {
                    const int part = e / N0, h = e - part * N0;
                    const T ghg = G[(size_t)wf * N0 + h].im;
                    const C2<T>* dfc = Df + ((size_t)wf * M + m0) * N0 + h;
                    C2<T> s = mk<T>(0, 0);
                    T gam = 0;
                    for (int mm = part; mm < mc; mm += parts) {
                        const C2<T> df = dfc[(size_t)mm * N0];
                        const T inv = (T)1 / (sumin[m0 + mm].re * ghg + rho);
                        s = s + inv * (df * buf[(size_t)mm * N0 + h]);
                        gam += inv * abs2(df);
                    }
                    C2<T>* sp = spart + ((size_t)part) * N0 + h;
                    *sp = *sp + s;
                    C2<T>* gp = spart + ((size_t)parts + part) * N0 + h;
                    gp->re += gam;
                }
                __syncthreads();
                continue;
            }
            if (a.gradreg && SOLVE == 4) {   // RegGrad of the evaluated spectra (AuxVarObj)
                const double wg = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
                double rg[1] = {0.0};
                for (int e = tid; e < mc * N0; e += nt) {
                    const int mm = e / N0, h = e - mm * N0;
                    rg[0] += wg * (double)(sumin[m0 + mm].im * G[(size_t)wf * N0 + h].im * abs2(buf[e]));
                }
                double* red = reinterpret_cast<double*>(spart + (size_t)Cd * parts * N0);
                block_accumulate<1>(rg, red, acc + ACC_RGRAD);
            }
            for (int e = tid; e < parts * N0; e += nt) {
                const int part = e / N0, h = e - part * N0;
                for (int c = 0; c < Cd; ++c) {
                    const C2<T>* dfc = Df + (((size_t)c * a.N1f + wf) * M + m0) * N0 + h;
                    C2<T> s = mk<T>(0, 0);
                    for (int mm = part; mm < mc; mm += parts)
                        s = s + dfc[(size_t)mm * N0] * buf[(size_t)mm * N0 + h];
                    C2<T>* sp = spart + ((size_t)c * parts + part) * N0 + h;
                    *sp = *sp + s;
                }
            }
            __syncthreads();
        }
        // ---- q_c[h]
        double dsum[1] = {0.0};
        double psum[3] = {0.0, 0.0, 0.0};            // PGM: plain |s-Sf|^2, weighted, linear term
        const double wgt = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
        const int k = b / a.Cx, cx = b - k * a.Cx;
        for (int h = tid; h < N0; h += nt) {
            C2<T> d[MAXCD];
            for (int c = 0; c < Cd; ++c) {
                C2<T> s = mk<T>(0, 0);
                for (int p = 0; p < parts; ++p) s = s + spart[((size_t)c * parts + p) * N0 + h];
                if (SOLVE == 3) {
                    d[c] = s;
                } else {
                    const int cs = (Cd > 1) ? c : cx;
                    const C2<T> sf = Sf[(((size_t)k * a.Cs + cs) * a.N1f + wf) * N0 + h];
                    d[c] = sf - s;
                    if (a.pgm_mask) {
                        const size_t si = (((size_t)b * Cd + c) * a.N1f + wf) * N0 + h;
                        if (SOLVE == 2) {
                            d[c] = mk<T>(-sumin[si].re, -sumin[si].im);        // - rfft(W^2 irfft(s_Y - Sf))
                        } else if (SOLVE == 4) {
                            sumout[si] = s;                                     // s_X, for the masked F / DFid
                            const C2<T> dx = s - sumin[si], gy = G[si];
                            psum[2] += (double)(dx.re * gy.re + dx.im * gy.im);
                        }
                        continue;
                    }
                    if (SOLVE == 2 && sumout) {
                        sumout[(((size_t)b * Cd + c) * a.N1f + wf) * N0 + h] = s;
                        psum[0] += (double)abs2(d[c]);
                    }
                    if (SOLVE == 4) {
                        const double e2 = (double)abs2(d[c]);
                        psum[0] += e2;
                        psum[1] += wgt * e2;
                        if (sumin && !a.gradreg) {
                            const C2<T> sy = sumin[(((size_t)b * Cd + c) * a.N1f + wf) * N0 + h];
                            const C2<T> dx = s - sy, gy = sy - sf;   // Re(conj(dx) * gy)
                            psum[2] += (double)(dx.re * gy.re + dx.im * gy.im);
                        }
                    }
                }
            }
            if (SOLVE == 1) {
                if (a.gradreg) {
                    // q = (Sf - rho sigma) / (1 + gamma); then x_m = (rho z_m + conj(Df_m) q) / d_m
                    // and sum_m Df_m x_m - Sf = -q
                    T gam = 0;
                    C2<T> sg = mk<T>(0, 0);
                    for (int p = 0; p < parts; ++p) {
                        sg = sg + spart[((size_t)p) * N0 + h];
                        gam += spart[((size_t)parts + p) * N0 + h].re;
                    }
                    const C2<T> sf = Sf[(((size_t)k * a.Cs + cx) * a.N1f + wf) * N0 + h];
                    const T den = (T)1 + gam;
                    const C2<T> num = sf - rho * sg;
                    d[0] = mk<T>(num.re / den, num.im / den);
                } else if (Cd == 1) {
                    const T g = G[(size_t)wf * N0 + h].re;
                    const T den = g + rho;
                    d[0] = mk<T>(d[0].re / den, d[0].im / den);
                } else {
                    C2<T> A[MAXCD][MAXCD];
                    const C2<T>* Gp = G + ((size_t)wf * N0 + h) * Cd * Cd;
                    for (int i = 0; i < Cd; ++i)
                        for (int j = 0; j < Cd; ++j) {
                            A[i][j] = Gp[i * Cd + j];
                            if (i == j) A[i][j].re += rho;
                        }
                    hpd_solve<T, MAXCD>(A, d, Cd);
                }
                if (a.dfid_on) {
                    double q2 = 0.0;
                    for (int c = 0; c < Cd; ++c) q2 += (double)abs2(d[c]);
                    dsum[0] += wgt * q2;
                }
            } else if (SOLVE == 2) {
                for (int c = 0; c < Cd; ++c) d[c] = mk<T>(d[c].re / Lstep, d[c].im / Lstep);
            }
            if (SOLVE == 3) {
                for (int c = 0; c < Cd; ++c)
                    sumout[(((size_t)b * Cd + c) * a.N1f + wf) * N0 + h] = d[c];
            } else if (SOLVE == 4) {
                // nothing to keep: the evaluation does not update the slab
            } else {
                for (int c = 0; c < Cd; ++c) spart[((size_t)c * parts) * N0 + h] = d[c];
            }
        }
        __syncthreads();
        if (SOLVE == 1 && a.dfid_on) {
            double* red = reinterpret_cast<double*>(spart + (size_t)CdS * parts * N0);
            block_accumulate<1>(dsum, red, acc + ACC_DFID);
        }
        if (SOLVE == 2 && sumout) {
            double* red = reinterpret_cast<double*>(spart + (size_t)Cd * parts * N0);
            double one[1] = {psum[0]};
            block_accumulate<1>(one, red, acc + ACC_PGM_FY);
        }
        if (SOLVE == 4) {
            double* red = reinterpret_cast<double*>(spart + (size_t)Cd * parts * N0);
            double a1[1] = {psum[0]}, a2[1] = {psum[1]}, a3[1] = {psum[2]};
            block_accumulate<1>(a1, red, acc + ACC_PGM_F);
            block_accumulate<1>(a2, red, acc + ACC_DFID);
            block_accumulate<1>(a3, red, acc + ACC_PGM_LIN);
        }
        if (SOLVE == 3) return;
    }

    // ---- phase B: update, inverse transform, store
    for (int ch = 0; ch < a.nchunk; ++ch) {
        const int m0 = ch * MC, mc = (M - m0 < MC) ? M - m0 : MC;
        if (!(resident && SOLVE != 0)) {
            col_load_chunk<T>(buf, src, m0, mc, N0);
            if (DO_FWD) {
                if constexpr (GEN) col_dft_chunk<T, false>(buf, buf2, tw, mc, N0);
                else col_fft_chunk<T, N0T, false>(buf, tw, mc, MC);
            }
            if (zeta != (T)1) {
                for (int e = tid; e < mc * N0; e += nt) buf[e] = zeta * buf[e];
                __syncthreads();
            }
        }
        if (SOLVE == 1 && a.gradreg) {
            const int parts = a.parts;
            const double wg = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
            double rg[1] = {0.0};
            for (int e = tid; e < mc * N0; e += nt) {
                const int mm = e / N0, h = e - mm * N0;
                const T ghg = G[(size_t)wf * N0 + h].im;
                const C2<T> gwm = sumin[m0 + mm];
                const T inv = (T)1 / (gwm.re * ghg + rho);
                const C2<T> df = Df[((size_t)wf * M + m0 + mm) * N0 + h];
                const C2<T> q = spart[(size_t)h];
                const C2<T> x = inv * (rho * buf[e] + mulc(q, df));
                if (a.dfid_on) rg[0] += wg * (double)(gwm.im * ghg * abs2(x));
                buf[e] = x;
            }
            __syncthreads();
            if (a.dfid_on) {
                double* red = reinterpret_cast<double*>(spart + (size_t)(Cd + 1) * parts * N0);
                block_accumulate<1>(rg, red, acc + ACC_RGRAD);
            }
        } else if (SOLVE == 1 || SOLVE == 2) {
            const int parts = a.parts;
            for (int e = tid; e < mc * N0; e += nt) {
                const int mm = e / N0, h = e - mm * N0;
                C2<T> v = buf[e];
                for (int c = 0; c < Cd; ++c) {
                    const C2<T> df = Df[(((size_t)c * a.N1f + wf) * M + m0 + mm) * N0 + h];
                    const C2<T> q = spart[((size_t)c * parts) * N0 + h];
                    v = v + mulc(q, df);
                }
                buf[e] = v;
            }
            __syncthreads();
        }
        if (SOLVE == 4 && ref) {
            const double wgt = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
            double dd[1] = {0.0};
            const C2<T>* rs = ref + slab;
            for (int e = tid; e < mc * N0; e += nt)
                dd[0] += (double)abs2(buf[e] - rs[(size_t)m0 * N0 + e]);
            double* red = reinterpret_cast<double*>(spart + (size_t)Cd * a.parts * N0);
            block_accumulate<1>(dd, red, acc + ACC_PGM_DXY2);
            dd[0] *= wgt;
            block_accumulate<1>(dd, red, acc + ACC_PGM_RSDL);
        }
        if (DO_INV) {
            if constexpr (GEN) col_dft_chunk<T, true>(buf, buf2, tw, mc, N0);
            else col_fft_chunk<T, N0T, true>(buf, tw, mc, MC);
        }
        if (dst)
            for (int e = tid; e < mc * N0; e += nt) dst[(size_t)m0 * N0 + e] = buf[e];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// PGM proximal step in the row domain (pgm/pgm.py:796-800, pgm/cbpdn.py:288-298):
//   V = irfft_row(Vt)*scale ; X = prox_l1(V, (lmbda/L) wl1) [+NonNeg, NoBndryCross] ;
//   Xt = rfft_row(X)   written over Vt;  X stored; RegL1 accumulated.
// ------------------------------------------------------------------------------------
template <typename T, int H>
SPCSC_GLOBAL void k_row_inv_prox_fwd(C2<T>* SPCSC_RESTRICT Vt, T* SPCSC_RESTRICT X, T thr_scale,
                                     WeightView<T> wl1, double* SPCSC_RESTRICT acc,
                                     const C2<T>* SPCSC_RESTRICT tw, int N0, int M, int Cx, int TR,
                                     T scale, int nonneg, int bnd0, int bnd1) {
    SPCSC_DYN_SMEM(smem_raw);
    C2<T>* buf = reinterpret_cast<C2<T>*>(smem_raw);
    constexpr int P = H + 1;
    constexpr int TPF = fft_tpf<T, H>();
    constexpr int N1f = H + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, b = blockIdx.z;
    const int k = b / Cx, c = b - k * Cx;
    row_inverse_to_smem<T, H>(buf, Vt, tw, b, m, h0, N0, M, TR);
    __syncthreads();
    double sums[1] = {0.0};
    C2<T>* X2 = reinterpret_cast<C2<T>*>(X) + (((size_t)b * M + m) * N0 + h0) * H;
    for (int e = tid; e < TR * H; e += nt) {
        const int r = e / H, j = e - r * H;
        const int h = h0 + r;
        const C2<T> z = buf[r * P + j];
        T xs[2] = {z.re * scale, z.im * scale};
        SPCSC_UNROLL
        for (int q = 0; q < 2; ++q) {
            const T w1 = wl1.p[(size_t)k * wl1.sk + (size_t)c * wl1.sc + (size_t)m * wl1.sm +
                              (size_t)h * wl1.s0 + (size_t)(2 * j + q) * wl1.s1];
            T x = soft_threshold(xs[q], thr_scale * w1);
            if (nonneg && x < (T)0) x = (T)0;
            if (h >= bnd0 || (2 * j + q) >= bnd1) x = (T)0;
            xs[q] = x;
            sums[0] += (double)fabs(w1 * x);
        }
        X2[e] = mk<T>(xs[0], xs[1]);
        buf[r * P + j] = mk<T>(xs[0], xs[1]);
    }
    __syncthreads();
    {
        const int row = tid / TPF, t = tid - row * TPF;
        const bool active = row < TR;
        fft_smem<T, H, false, 2>(buf + (active ? row : 0) * P, t, tw, active);
    }
    C2<T>* out = Vt + (((size_t)b * N1f) * M + m) * N0 + h0;
    const size_t wstride = (size_t)M * N0;
    for (int e = tid; e < TR * N1f; e += nt) {
        const int wf = e / TR, r = e - wf * TR;
        const C2<T> aa = buf[r * P + (wf == H ? 0 : wf)];
        const C2<T> bb = conj(buf[r * P + (wf == 0 ? 0 : H - wf)]);
        const C2<T> w = tw[wf];
        const C2<T> sum = aa + bb, dif = mul_mi((aa - bb) * w);
        out[wf * wstride + r] = mk<T>((T)0.5 * (sum.re + dif.re), (T)0.5 * (sum.im + dif.im));
    }
    double* red = reinterpret_cast<double*>(smem_raw);
    block_accumulate<1>(sums, red, acc + ACC_L1);
}

// Momentum step in the frequency domain (pgm/pgm.py:815-831): Yf = Xf + coef (Xf - Xfprv).
template <typename T>
SPCSC_GLOBAL void k_pgm_momentum(const C2<T>* SPCSC_RESTRICT Xf, const C2<T>* SPCSC_RESTRICT Xfprv,
                                 C2<T>* SPCSC_RESTRICT Yf, T coef, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const C2<T> x = Xf[i], p = Xfprv[i];
        Yf[i] = mk<T>(x.re + coef * (x.re - p.re), x.im + coef * (x.im - p.im));
    }
}

// out = a X + b Y  (spectra; robust backtracking: y = (T_k x_prev + t z) / T, pgm/backtrack.py:176-178;
// monotone FISTA after a rejected step: y = x + (t_prev / t)(z - x), pgm/pgm.py:826-828)
template <typename T>
SPCSC_GLOBAL void k_spec_axpby(const C2<T>* SPCSC_RESTRICT X, const C2<T>* SPCSC_RESTRICT Y,
                               C2<T>* SPCSC_RESTRICT out, T a, T b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const C2<T> x = X[i], y = Y[i];
        out[i] = mk<T>(a * x.re + b * y.re, a * x.im + b * y.im);
    }
}
// Z += c (X - Y)   (robust backtracking: z += t L (x - y), pgm/backtrack.py:203)
template <typename T>
SPCSC_GLOBAL void k_spec_add_diff(C2<T>* SPCSC_RESTRICT Z, const C2<T>* SPCSC_RESTRICT X,
                                  const C2<T>* SPCSC_RESTRICT Y, T c, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const C2<T> z = Z[i], x = X[i], y = Y[i];
        Z[i] = mk<T>(z.re + c * (x.re - y.re), z.im + c * (x.im - y.im));
    }
}
// Hermitian-weighted squared distance of two spectra in slab order [nb][N1f][per_wf], i.e. N times
// rfl2norm2(A - B) (fft.py:449-484), added to *slot.
template <typename T>
SPCSC_GLOBAL void k_spec_wdist2(const C2<T>* SPCSC_RESTRICT A, const C2<T>* SPCSC_RESTRICT B,
                                double* SPCSC_RESTRICT slot, int nb, int N1f, size_t per_wf, int even_n1) {
    __shared__ double red[32];
    const size_t n = (size_t)nb * N1f * per_wf;
    double s[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int wf = (int)((i / per_wf) % N1f);
        const double wgt = (wf == 0 || (even_n1 && wf == N1f - 1)) ? 1.0 : 2.0;
        s[0] += wgt * (double)abs2(A[i] - B[i]);
    }
    block_accumulate<1>(s, red, slot);
}
// sum |w x| over a real array in device order [K][Cx][M][N0][N1] with the broadcast weight view
template <typename T>
SPCSC_GLOBAL void k_l1_sum(const T* SPCSC_RESTRICT X, WeightView<T> w, double* SPCSC_RESTRICT slot, int K,
                           int Cx, int M, int N0, int N1) {
    __shared__ double red[32];
    const size_t n = (size_t)K * Cx * M * N0 * N1;
    double s[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t t = i;
        const int n1 = (int)(t % N1); t /= N1;
        const int n0 = (int)(t % N0); t /= N0;
        const int m = (int)(t % M); t /= M;
        const int c = (int)(t % Cx);
        const int k = (int)(t / Cx);
        const T wv = w.p[(size_t)k * w.sk + (size_t)c * w.sc + (size_t)m * w.sm + (size_t)n0 * w.s0 +
                         (size_t)n1 * w.s1];
        s[0] += (double)fabs(wv * X[i]);
    }
    block_accumulate<1>(s, red, slot);
}
// Scalars of the step-size policies (pgm/stepsize.py:50-145) and of the objective, from the per-frequency
// sums only.  With R = sY - Sf (residual spectrum at the auxiliary point, Cd-vector per frequency; gradient
// A R with A = [conj(Df_c,m)]) and G = A^H A:
//   out[0] = sum R^H G R = ||grad||^2                 out[1] = sum |G R|^2 = <grad, Hess grad>      (Cauchy)
//   out[2] = sum dR^H G dR = ||grad - grad_prev||^2   out[3] = sum Re(dsX^H dR) = <x - x_prev, dgrad>   (BB)
//   out[4] = Hermitian-weighted sum |sX - Sf|^2  (N times 2 DFid of the accepted iterate)
// Plain sums over the stored half spectrum for [0..3], as np.sum over rfftn output does.  With `store` the
// current R and sX replace the remembered ones afterwards (StepSizePolicyBB.store_prev_state).
// Layout of sY, sX, Sf, Rprev, sXprev: [nb][CD][N1f][N0]; G: [N1f][N0][CD][CD].
template <typename T, int CD>
SPCSC_GLOBAL void k_pgm_policy(const C2<T>* SPCSC_RESTRICT sY, const C2<T>* SPCSC_RESTRICT sX,
                               const C2<T>* SPCSC_RESTRICT Sf, const C2<T>* SPCSC_RESTRICT G,
                               C2<T>* SPCSC_RESTRICT Rprev, C2<T>* SPCSC_RESTRICT sXprev,
                               double* SPCSC_RESTRICT out, int nb, int N1f, int N0, int even_n1, int store) {
    __shared__ double red[5 * 32];
    const size_t nf = (size_t)N1f * N0, n = (size_t)nb * nf;
    double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / nf, f = i - b * nf;
        const int wf = (int)(f / N0);
        const double wgt = (wf == 0 || (even_n1 && wf == N1f - 1)) ? 1.0 : 2.0;
        C2<T> R[CD], dR[CD], dS[CD];
        SPCSC_UNROLL
        for (int d = 0; d < CD; ++d) {
            const size_t j = (b * CD + d) * nf + f;
            const C2<T> sf = Sf[j], sy = sY[j], sx = sX[j];
            R[d] = sy - sf;
            dR[d] = R[d] - Rprev[j];
            dS[d] = sx - sXprev[j];
            s[4] += wgt * (double)abs2(sx - sf);
            if (store) {
                Rprev[j] = R[d];
                sXprev[j] = sx;
            }
        }
        const C2<T>* Gp = G + f * CD * CD;
        SPCSC_UNROLL
        for (int r = 0; r < CD; ++r) {
            C2<T> y = mk<T>(0, 0), dy = mk<T>(0, 0);
            SPCSC_UNROLL
            for (int c = 0; c < CD; ++c) {
                C2<T> g = Gp[r * CD + c];
                if (CD == 1) g.im = 0;           // the imaginary part may carry another table (GHG)
                y = y + g * R[c];
                dy = dy + g * dR[c];
            }
            s[0] += (double)(R[r].re * y.re + R[r].im * y.im);
            s[1] += (double)abs2(y);
            s[2] += (double)(dR[r].re * dy.re + dR[r].im * dy.im);
            s[3] += (double)(dS[r].re * dR[r].re + dS[r].im * dR[r].im);
        }
    }
    block_accumulate<5>(s, red, out);
}

// ------------------------------------------------------------------------------------
// LinSolveCheck (admm/cbpdn.py:283-293): relative residual of the x-step system, taken
// honestly from the stored solution Xf (column-spectrum slab, before the inverse column
// transform) -- a separate diagnostic kernel, only launched when the option is on.
//   Zt holds Xf slabs; Zin holds the right-hand-side slabs rho*Z is formed from.
// ------------------------------------------------------------------------------------
template <typename T>
SPCSC_GLOBAL void k_linsolve_check(const C2<T>* SPCSC_RESTRICT Xf, const C2<T>* SPCSC_RESTRICT Zf,
                                   const C2<T>* SPCSC_RESTRICT Df, const C2<T>* SPCSC_RESTRICT Sf,
                                   const AdmmState<T>* SPCSC_RESTRICT st,
                                   double* SPCSC_RESTRICT acc, ColArgs a, T l2w,
                                   const C2<T>* SPCSC_RESTRICT G, const C2<T>* SPCSC_RESTRICT gw) {
    // one CTA per (wf, b); thread per h (strided); loops over m twice.  Slow, diagnostic only.
    if (st->stopped) return;
    SPCSC_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int wf = blockIdx.x, b = blockIdx.y, N0 = a.N0, M = a.M, Cd = a.Cd;
    const int k = b / a.Cx, cx = b - k * a.Cx;
    const T rho = st->rho;
    const size_t slab = (((size_t)b * a.N1f + wf) * M) * N0;
    double sums[3] = {0, 0, 0};
    for (int h = tid; h < N0; h += nt) {
        C2<T> e[4], sf[4];
        for (int c = 0; c < Cd; ++c) {
            C2<T> s = mk<T>(0, 0);
            for (int m = 0; m < M; ++m)
                s = s + Df[(((size_t)c * a.N1f + wf) * M + m) * N0 + h] * Xf[slab + (size_t)m * N0 + h];
            e[c] = s;
            const int cs = (Cd > 1) ? c : cx;
            sf[c] = Sf[(((size_t)k * a.Cs + cs) * a.N1f + wf) * N0 + h];
        }
        for (int m = 0; m < M; ++m) {
            const C2<T> x = Xf[slab + (size_t)m * N0 + h], z = Zf[slab + (size_t)m * N0 + h];
            T dg = rho + l2w;
            if (a.gradreg) dg = rho + gw[m].re * G[(size_t)wf * N0 + h].im;
            C2<T> ax = dg * x, bb = rho * z;
            for (int c = 0; c < Cd; ++c) {
                const C2<T> df = Df[(((size_t)c * a.N1f + wf) * M + m) * N0 + h];
                ax = ax + mulc(e[c], df);
                bb = bb + mulc(sf[c], df);
            }
            sums[0] += (double)abs2(ax);
            sums[1] += (double)abs2(bb);
            sums[2] += (double)abs2(ax - bb);
        }
    }
    double* red = reinterpret_cast<double*>(smem_raw);
    block_accumulate<3>(sums, red, acc + ACC_AX2);
}

// ------------------------------------------------------------------------------------
// k_admm_scalars: one thread.  Replays admm/admm.py:462-486 (residuals, stopping
// tolerances) and admm/admm.py:549-575 (rho update) in the working precision T from the
// double-precision sums, writes one StatRow, advances k, clears the accumulators.
// ------------------------------------------------------------------------------------
// Fold the integer bins into acc[0..kAccDet) and clear them: warp i handles sum i, lane l the bins
// l and l+32; the butterfly reduction has a fixed order, so the result is reproducible.  Needs
// blockDim.x >= 32*kAccDet; ends with a block barrier.
SPCSC_DEV void fold_det_bins(double* acc) {
    unsigned long long* bins = reinterpret_cast<unsigned long long*>(acc + ACC_N);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < kAccDet) {
        unsigned long long* row = bins + warp * kDetBins;
        double x = det_bin_term(row, lane) + det_bin_term(row, lane + 32);
        x = warp_sum(x);
        row[lane] = 0ull;
        row[lane + 32] = 0ull;
        if (lane == 0) acc[warp] += x;
    }
    __syncthreads();
}
// ---- all-reduce of the ACC_N accumulators over peer memory (NVLink / NVSwitch) --------------
// Every rank owns one P2pSlots block; peers hold it mapped (CUDA IPC).  Rank r writes its ACC_N
// doubles into slot [parity][r] of EVERY rank's block, fences, then raises flag [parity][r] there to
// the sequence number of this exchange; each rank waits for all flags of its own block and adds the
// slots in rank order -- the same order everywhere, so all ranks get bit-identical sums (they must:
// the rho update and the stopping test branch on them).  Two parities: a rank can be at most one
// exchange ahead of the slowest one (it cannot pass exchange n+1 without that rank's flag n+1, which
// is raised only after exchange n was read), so buffer n+2 never overwrites unread data.
constexpr int kP2pMaxRanks = 8;
constexpr int kP2pVecMax = 16384;      // largest vector (elements) the peer-memory vector all-reduce takes
struct P2pSlots {
    double vals[2][kP2pMaxRanks][ACC_N];
    unsigned long long flag[2][kP2pMaxRanks];
};
// The block every rank allocates and its peers map: the accumulator slots, this rank's exchange counters
// (advanced only by exchanges that execute), and the slots of the vector all-reduce (dictionary learning:
// the cropped dictionary gradient, h*w*Cd*M values, pgm/ccmod.py + cnvrep.py:953-981).
struct P2pBlock {
    P2pSlots slots;
    unsigned long long seq, vseq;
    unsigned long long vflag[2][kP2pMaxRanks];
    double vec[2][kP2pMaxRanks][kP2pVecMax];
};
struct P2pView {
    P2pBlock* peer[kP2pMaxRanks];      // peer[r]: rank r's block as mapped here (peer[rank] = own block)
    int nranks, rank;
};
// All threads of the (single) block; blockDim.x >= ACC_N * nranks.  Returns false on time-out.
SPCSC_DEV bool p2p_allreduce(const P2pView& pv, double* acc) {
#ifdef SPCSC_EMU
    (void)pv; (void)acc;
    return true;
#else
    __shared__ int timed_out;
    __shared__ unsigned long long seq_s;
    const int tid = threadIdx.x;
    if (tid == 0) {
        timed_out = 0;
        seq_s = ++pv.peer[pv.rank]->seq;           // > 0; advanced only by exchanges that execute
    }
    __syncthreads();
    const unsigned long long seq = seq_s;
    const int par = (int)(seq & 1ull);
    if (tid < ACC_N * pv.nranks) {
        const int r = tid / ACC_N, i = tid % ACC_N;
        pv.peer[r]->slots.vals[par][pv.rank][i] = acc[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pv.nranks) {
        volatile unsigned long long* f = &pv.peer[tid]->slots.flag[par][pv.rank];
        *f = seq;
    }
    if (tid < pv.nranks) {
        volatile unsigned long long* f = &pv.peer[pv.rank]->slots.flag[par][tid];
        const long long t0 = clock64();
        while (*f != seq) {
            if (clock64() - t0 > 60000000000LL) { timed_out = 1; break; }     // ~30 s: ranks may start far apart
        }
    }
    __threadfence_system();
    __syncthreads();
    if (timed_out) return false;
    if (tid < ACC_N) {
        const volatile double* v = &pv.peer[pv.rank]->slots.vals[par][0][tid];
        double s = 0.0;
        for (int r = 0; r < pv.nranks; ++r) s += v[(size_t)r * ACC_N];
        acc[tid] = s;
    }
    __syncthreads();
    return true;
#endif
}

// All-reduce (sum) of a vector of n <= kP2pVecMax values over the ranks, same protocol as p2p_allreduce with
// its own counter, flags and slots; sums in double, in rank order, so every rank ends with identical values.
// One block; any number of threads.  On time-out *status is set to 2.
template <typename T>
SPCSC_GLOBAL void k_p2p_allreduce_vec(P2pView pv, T* SPCSC_RESTRICT data, int n, int* SPCSC_RESTRICT status) {
#ifndef SPCSC_EMU
    __shared__ int timed_out;
    __shared__ unsigned long long seq_s;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) {
        timed_out = 0;
        seq_s = ++pv.peer[pv.rank]->vseq;
    }
    __syncthreads();
    const unsigned long long seq = seq_s;
    const int par = (int)(seq & 1ull);
    for (int r = 0; r < pv.nranks; ++r) {
        double* dst = pv.peer[r]->vec[par][pv.rank];
        for (int i = tid; i < n; i += nt) dst[i] = (double)data[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pv.nranks) {
        volatile unsigned long long* f = &pv.peer[tid]->vflag[par][pv.rank];
        *f = seq;
    }
    if (tid < pv.nranks) {
        volatile unsigned long long* f = &pv.peer[pv.rank]->vflag[par][tid];
        const long long t0 = clock64();
        while (*f != seq) {
            if (clock64() - t0 > 60000000000LL) { timed_out = 1; break; }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (timed_out) {
        if (tid == 0 && status) *status = 2;
        return;
    }
    for (int i = tid; i < n; i += nt) {
        double s = 0.0;
        for (int r = 0; r < pv.nranks; ++r) {
            const volatile double* v = pv.peer[pv.rank]->vec[par][r];
            s += v[i];
        }
        data[i] = (T)s;
    }
#else
    (void)pv; (void)data; (void)n; (void)status;
#endif
}
// the accumulator all-reduce as a kernel of its own (dictionary learning: the data fidelity of the new iterate)
template <int DUMMY>
SPCSC_GLOBAL void k_p2p_allreduce_acc(P2pView pv, double* acc, int* status) {
    if (!p2p_allreduce(pv, acc) && threadIdx.x == 0 && status) *status = 2;
}
// gather / scatter of the filter supports (top-left hd x wd of every [Cd*M] plane of N0 x N1)
template <typename T>
SPCSC_GLOBAL void k_support_copy(T* SPCSC_RESTRICT full, T* SPCSC_RESTRICT compact, int planes, int N0, int N1,
                                 int hd, int wd, int scatter) {
    const size_t n = (size_t)planes * hd * wd;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % wd), y = (int)((i / wd) % hd);
        const size_t pl = i / ((size_t)hd * wd);
        const size_t o = (pl * N0 + y) * N1 + x;
        if (scatter) full[o] = compact[i];
        else compact[i] = full[o];
    }
}

// Used before a multi-rank all-reduce (which then sums plain doubles in NCCL's fixed order).
template <int DUMMY>
SPCSC_GLOBAL void k_fold_bins(double* acc) {
    fold_det_bins(acc);
}

template <typename T>
SPCSC_GLOBAL void k_admm_scalars(AdmmState<T>* st, AdmmParams<T> p, double* acc,
                                 StatRow* rows, int k_base, int row_cap, P2pView pv) {
    if (st->stopped) return;                       // uniform over the (single) block
    fold_det_bins(acc);
    if (pv.nranks > 1 && !p2p_allreduce(pv, acc)) {
        if (threadIdx.x == 0) st->stopped = 2;     // a peer never arrived: stop, the host reports it
        return;
    }
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int k = st->k;
    T rho = st->rho;
    T r = 0, s = 0;
    double epri = 0, edua = 0;
    if (p.need_rsdl) {
        const T nX = (T)sqrt(acc[ACC_X2]), nY = (T)sqrt(acc[ACC_Y2]), nU = (T)sqrt(acc[ACC_U2]);
        const T nR = (T)sqrt(acc[ACC_R2]), nS = rho * (T)sqrt(acc[ACC_S2]);
        const T mx = nX > nY ? nX : nY;
        if (p.stdres) {
            r = nR;
            s = nS;
            epri = sqrt(p.n_x) * p.abs_tol + (double)mx * p.rel_tol;
            edua = sqrt(p.n_x) * p.abs_tol + (double)(rho * nU) * p.rel_tol;
        } else {
            T rn = mx, sn = rho * nU;
            if (rn == (T)0) rn = 1;
            if (sn == (T)0) sn = 1;
            r = nR / rn;
            s = nS / sn;
            epri = sqrt(p.n_x) * p.abs_tol / (double)rn + p.rel_tol;
            edua = sqrt(p.n_x) * p.abs_tol / (double)sn + p.rel_tol;
        }
    }
    if (rows && k - k_base >= 0 && k - k_base < row_cap) {
        StatRow& o = rows[k - k_base];
        o.k = k;
        o.r = r; o.s = s; o.epri = epri; o.edua = edua; o.rho = rho;
        o.dfid = o.regl1 = o.regl21 = o.obj = 0;
        o.xrrs = -1.0;
        if (p.need_obj) {
            const double rho_x = (double)(T)(rho + (p.enet ? p.enet_mu : (T)0));
            o.dfid = (p.dfid_direct || p.gradreg) ? 0.5 * acc[ACC_DFID] * p.inv_n
                                                  : 0.5 * rho_x * rho_x * acc[ACC_DFID] * p.inv_n;
            o.regl1 = acc[ACC_L1];
            o.regl21 = acc[ACC_L21];
            o.obj = o.dfid + (double)p.lmbda * o.regl1 + (p.joint ? (double)p.mu * o.regl21 : 0.0);
            if (p.gradreg) {               // (mu/2) sum_m w_m ||G x_m||^2, taken in the DFT domain
                o.regl21 = 0.5 * acc[ACC_RGRAD] * p.inv_n;
                o.obj = o.dfid + ((double)p.lmbda * o.regl1 + (double)p.mu * o.regl21);
            }
            if (p.enet) {                  // (mu/2)||x||^2 on the objective's variable (X, or Y with AuxVarObj)
                o.regl21 = 0.5 * acc[p.dfid_direct ? ACC_Y2 : ACC_X2];
                o.obj = o.dfid + ((double)p.lmbda * o.regl1 + (double)p.enet_mu * o.regl21);
            }
        }
        if (p.linsolve_check) {
            const double na = sqrt(acc[ACC_AX2]), nb = sqrt(acc[ACC_B2]);
            const double nm = na > nb ? na : nb;
            o.xrrs = nm == 0.0 ? 0.0 : sqrt(acc[ACC_AXB2]) / nm;
        }
    }
    T udiv = 1;
    if (p.autorho && p.need_rsdl) {
        if (k != 0 && ((k + 1) % p.period) == 0) {
            T mlt;
            if (p.autoscaling) {
                if (s == (T)0 || r == (T)0) {
                    mlt = p.tau;
                } else {
                    const T sx = s * p.xi;
                    mlt = (T)sqrt(r > sx ? r / sx : sx / r);
                    if (mlt > p.tau) mlt = p.tau;
                }
            } else {
                mlt = p.tau;
            }
            T rsf = 1;
            if (r > p.xi * p.mur * s) rsf = mlt;
            else if (s > (p.mur / p.xi) * r) rsf = (T)1 / mlt;
            rho = rho * rsf;
            udiv = rsf;
        }
    }
    st->rho = rho;
    st->udiv = udiv;
    // the spectra the prox kernel may have written for the next x-step are unusable when rho changed (U is
    // rescaled), and absent when the kernel was told not to write them
    st->zt_stale = (udiv != (T)1 || !st->emit) ? 1 : 0;
    st->emit = (p.emit_policy == 0 || udiv == (T)1) ? 1 : 0;
    st->k = k + 1;
    if (p.need_rsdl && (double)r < epri && (double)s < edua) st->stopped = 1;
    for (int i = 0; i < ACC_N; ++i) acc[i] = 0.0;
}

// Fold a pending dual rescale into the stored U (used before U is read back).
template <typename T>
SPCSC_GLOBAL void k_apply_udiv(T* U, AdmmState<T>* st, size_t n) {
    const T d = st->udiv;
    if (d == (T)1) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        U[i] = U[i] / d;
}
template <typename T>
SPCSC_GLOBAL void k_reset_udiv(AdmmState<T>* st) { st->udiv = 1; st->zt_stale = 1; }
template <typename T>
SPCSC_GLOBAL void k_mark_stale(AdmmState<T>* st) { st->zt_stale = 1; }

// ------------------------------------------------------------------------------------
// Layout conversion between the reference's (N0,N1,C,K,M) order and the device order
// [K][C][M][N0*N1]: a tiled transpose of a (N0*N1) x (C*K*M) matrix with a row permutation.
// ------------------------------------------------------------------------------------
template <typename T>
SPCSC_GLOBAL void k_to_internal(const T* SPCSC_RESTRICT ext, T* SPCSC_RESTRICT inr, int NP,
                                int C, int K, int M) {
    __shared__ T tile[32][33];
    const int J = C * K * M;
    const int p0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, j = j0 + tx;
        if (p < NP && j < J) tile[i][tx] = ext[(size_t)p * J + j];
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int j = j0 + i, p = p0 + tx;
        if (p < NP && j < J) {
            const int m = j % M, kk = (j / M) % K, c = j / (M * K);
            inr[(((size_t)kk * C + c) * M + m) * NP + p] = tile[tx][i];
        }
    }
}

template <typename T>
SPCSC_GLOBAL void k_from_internal(const T* SPCSC_RESTRICT inr, T* SPCSC_RESTRICT ext, int NP,
                                  int C, int K, int M) {
    __shared__ T tile[32][33];
    const int J = C * K * M;
    const int p0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int j = j0 + i, p = p0 + tx;
        if (p < NP && j < J) {
            const int m = j % M, kk = (j / M) % K, c = j / (M * K);
            tile[i][tx] = inr[(((size_t)kk * C + c) * M + m) * NP + p];
        }
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, j = j0 + tx;
        if (p < NP && j < J) ext[(size_t)p * J + j] = tile[tx][i];
    }
}

// Zero-padded dictionary in device order [Cd][M][N0][N1] from the reference's (hd,wd,Cd,M).
template <typename T>
SPCSC_GLOBAL void k_pad_dict(const T* SPCSC_RESTRICT D, T* SPCSC_RESTRICT Dp, int hd, int wd,
                             int Cd, int M, int N0, int N1) {
    const size_t n = (size_t)Cd * M * N0 * N1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % N1), y = (int)((i / N1) % N0);
        const int m = (int)((i / ((size_t)N1 * N0)) % M), c = (int)(i / ((size_t)N1 * N0 * M));
        Dp[i] = (y < hd && x < wd) ? D[(((size_t)y * wd + x) * Cd + c) * M + m] : (T)0;
    }
}

// ---- masked data fidelity of pgm.cbpdn.ConvBPDNMask (pgm/cbpdn.py:461-506) -----------------
template <typename T>
SPCSC_GLOBAL void k_spec_sub(const C2<T>* SPCSC_RESTRICT a, const C2<T>* SPCSC_RESTRICT b,
                             C2<T>* SPCSC_RESTRICT out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = a[i] - b[i];
}
// o1 = W r, o2 = W^2 r (either may be null); optionally sum (W r)^2 * scale2 into acc_slot
template <typename T>
SPCSC_GLOBAL void k_mask_mul(const T* SPCSC_RESTRICT r, const T* SPCSC_RESTRICT W, T* SPCSC_RESTRICT o1,
                             T* SPCSC_RESTRICT o2, double* SPCSC_RESTRICT acc_slot, double scale2, size_t n) {
    __shared__ double red[32];
    double s[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const T w = W[i], wr = w * r[i];
        if (o1) o1[i] = wr;
        if (o2) o2[i] = w * wr;
        s[0] += (double)wr * (double)wr;
    }
    if (acc_slot) {
        s[0] *= scale2;
        block_accumulate<1>(s, red, acc_slot);
    }
}
// plain sum of |z|^2 over a spectrum buffer (the half-spectrum norm the backtracking test uses)
template <typename T>
SPCSC_GLOBAL void k_spec_sumsq(const C2<T>* SPCSC_RESTRICT z, double* SPCSC_RESTRICT acc_slot, size_t n) {
    __shared__ double red[32];
    double s[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        s[0] += (double)abs2(z[i]);
    block_accumulate<1>(s, red, acc_slot);
}

// ---- level-1 entry points (stand-alone launches of the arithmetic the fused kernels use) ---------------
// linalg.solvedbi_sm / solvemdbi_ism (linalg.py:232-297, 370-444): (rho I + sum_c a_c a_c^H) x = b with
// a_c = conj(ah_c), per position f and right-hand side k, reduction over the M axis.
//   ah [nf][CD][M], b and x [nf][nk][M].  One warp per (f, k): lanes stride over M.
template <typename T, int CD>
SPCSC_GLOBAL void k_solvedbi(const C2<T>* SPCSC_RESTRICT ah, const C2<T>* SPCSC_RESTRICT b,
                             C2<T>* SPCSC_RESTRICT x, long long nf, int nk, int M, T rho) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = warp; w < nf * nk; w += nwarp) {
        const long long f = w / nk;
        const C2<T>* a = ah + (size_t)f * CD * M;
        const C2<T>* bb = b + (size_t)w * M;
        // s_c = sum_m ah_c,m b_m ;  G_ij = sum_m ah_i,m conj(ah_j,m)
        C2<T> sv[CD], A[CD][CD];
        SPCSC_UNROLL
        for (int i = 0; i < CD; ++i) {
            sv[i] = mk<T>(0, 0);
            SPCSC_UNROLL
            for (int j = 0; j < CD; ++j) A[i][j] = mk<T>(0, 0);
        }
        for (int m = lane; m < M; m += 32) {
            const C2<T> bm = bb[m];
            C2<T> am[CD];
            SPCSC_UNROLL
            for (int i = 0; i < CD; ++i) am[i] = a[(size_t)i * M + m];
            SPCSC_UNROLL
            for (int i = 0; i < CD; ++i) {
                sv[i] = sv[i] + am[i] * bm;
                SPCSC_UNROLL
                for (int j = 0; j < CD; ++j) A[i][j] = A[i][j] + mulc(am[i], am[j]);
            }
        }
        SPCSC_UNROLL
        for (int i = 0; i < CD; ++i) {
            SPCSC_UNROLL
            for (int o = 16; o > 0; o >>= 1) {
                sv[i].re += __shfl_xor_sync(0xffffffffu, sv[i].re, o);
                sv[i].im += __shfl_xor_sync(0xffffffffu, sv[i].im, o);
            }
            SPCSC_UNROLL
            for (int j = 0; j < CD; ++j) {
                SPCSC_UNROLL
                for (int o = 16; o > 0; o >>= 1) {
                    A[i][j].re += __shfl_xor_sync(0xffffffffu, A[i][j].re, o);
                    A[i][j].im += __shfl_xor_sync(0xffffffffu, A[i][j].im, o);
                }
            }
            A[i][i].re += rho;
        }
        if (CD == 1) {
            sv[0] = mk<T>(sv[0].re / A[0][0].re, sv[0].im / A[0][0].re);
        } else {
            hpd_solve<T, CD>(A, sv, CD);
        }
        for (int m = lane; m < M; m += 32) {
            C2<T> v = bb[m];
            SPCSC_UNROLL
            for (int i = 0; i < CD; ++i) v = v - mulc(sv[i], a[(size_t)i * M + m]);   // conj(ah) * y
            x[(size_t)w * M + m] = mk<T>(v.re / rho, v.im / rho);
        }
    }
}
// prox_l1 (prox/_lp.py:144-183) and prox_sl1l2 (prox/_l21.py:51-88, the l2 shrinkage over the middle axis of
// [n_outer][C][n_inner]); w: optional weights of the l1 term, same shape as v
template <typename T>
SPCSC_GLOBAL void k_prox_l1l2(const T* SPCSC_RESTRICT v, const T* SPCSC_RESTRICT w, T* SPCSC_RESTRICT out,
                              long long n_outer, int C, long long n_inner, T alpha, T beta, int joint) {
    const long long n = n_outer * n_inner;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long o = i / n_inner, r = i - o * n_inner;
        const size_t base = (size_t)o * C * n_inner + r;
        T a2 = 0;
        for (int c = 0; c < C; ++c) {
            const size_t j = base + (size_t)c * n_inner;
            const T t = soft_threshold(v[j], alpha * (w ? w[j] : (T)1));
            out[j] = t;
            a2 += t * t;
        }
        if (joint) {
            const T a = sqrt(a2);
            const T fac = (a != (T)0) ? fmax((T)0, a - beta) / a : (T)0;
            for (int c = 0; c < C; ++c) out[base + (size_t)c * n_inner] *= fac;
        }
    }
}

// ---- sporco.signal.tikhonov_filter (signal.py:244-303) ------------------------------------
// symmetric padding by npd on every side of each image: index -1 -> 0, N -> N-1 (numpy 'symmetric')
template <typename T>
SPCSC_GLOBAL void k_pad_symmetric(const T* SPCSC_RESTRICT in, T* SPCSC_RESTRICT out, int batch, int N0,
                                  int N1, int npd) {
    const int P0 = N0 + 2 * npd, P1 = N1 + 2 * npd;
    const size_t n = (size_t)batch * P0 * P1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % P1), y = (int)((i / P1) % P0), b = (int)(i / ((size_t)P0 * P1));
        int sy = y - npd, sx = x - npd;
        sy = sy < 0 ? -sy - 1 : (sy >= N0 ? 2 * N0 - 1 - sy : sy);
        sx = sx < 0 ? -sx - 1 : (sx >= N1 ? 2 * N1 - 1 - sx : sx);
        out[i] = in[((size_t)b * N0 + sy) * N1 + sx];
    }
}
// slab spectra [batch][N1f][N0] divided by A = 1 + lmbda |Gr|^2 + lmbda |Gc|^2, the transfer function of
// I + lmbda (Gr^T Gr + Gc^T Gc) for the periodic forward differences: |G(f)|^2 = 2 - 2 cos(2 pi f / N)
template <typename T>
SPCSC_GLOBAL void k_tikhonov_divide(C2<T>* SPCSC_RESTRICT Z, int batch, int N1f, int N0, int N1, T lmbda) {
    const size_t n = (size_t)batch * N1f * N0;
    const double two_pi = 6.283185307179586476925286766559;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int h = (int)(i % N0), wf = (int)((i / N0) % N1f);
        const double gr = 2.0 - 2.0 * cos(two_pi * (double)h / (double)N0);
        const double gc = 2.0 - 2.0 * cos(two_pi * (double)wf / (double)N1);
        const T a = (T)1 + lmbda * (T)gr + lmbda * (T)gc;
        Z[i] = mk<T>(Z[i].re / a, Z[i].im / a);
    }
}
// sl = centre crop of the filtered padded image; sh = s - sl
template <typename T>
SPCSC_GLOBAL void k_crop_split(const T* SPCSC_RESTRICT padded, const T* SPCSC_RESTRICT s,
                               T* SPCSC_RESTRICT sl, T* SPCSC_RESTRICT sh, int batch, int N0, int N1,
                               int npd) {
    const int P0 = N0 + 2 * npd, P1 = N1 + 2 * npd;
    const size_t n = (size_t)batch * N0 * N1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % N1), y = (int)((i / N1) % N0), b = (int)(i / ((size_t)N0 * N1));
        const T l = padded[((size_t)b * P0 + y + npd) * P1 + x + npd];
        sl[i] = l;
        sh[i] = s[i] - l;
    }
}

// ConvBPDNGradReg: GHG[wf][h] rides in the (otherwise zero) imaginary part of the Cd = 1 Gram table.
template <typename T>
SPCSC_GLOBAL void k_set_ghg(C2<T>* SPCSC_RESTRICT G, const T* SPCSC_RESTRICT ghg, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        G[i].im = ghg[i];
}

// ConvBPDNGradReg with a multi-channel dictionary (sporco/admm/cbpdn.py:1181-1184: solvemdbi_ism with the diagonal
// d_m = mu w_m GHG + rho in place of rho; linalg.py:370-444).  Closed form per frequency, A = [conj(Df_c,m)] (M x C):
//   sigma_c = sum_m Df_c,m z_m / d_m ,  Gamma_cc' = sum_m Df_c,m conj(Df_c',m) / d_m ,
//   q = (I + Gamma)^-1 (Sf - rho sigma) ,  x_m = (rho z_m + sum_c conj(Df_c,m) q_c) / d_m ,  sum_m Df_c,m x_m - Sf_c = -q_c
// (the Cd = 1 case of this is in k_col).  Works in place on slabs that are ALREADY in the 2-D frequency domain (the engine
// brackets it with the forward / inverse column passes of k_col).  One CTA per (wf, 32 frequencies h, slab): 32 h-lanes x
// 8 filter groups; general path, not tuned.
template <typename T, int CD>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS(256)
k_gradreg_mc(C2<T>* SPCSC_RESTRICT Zf, const C2<T>* SPCSC_RESTRICT Df, const C2<T>* SPCSC_RESTRICT Sf,
             const T* SPCSC_RESTRICT ghg, const C2<T>* SPCSC_RESTRICT gw, const AdmmState<T>* SPCSC_RESTRICT st,
             double* SPCSC_RESTRICT acc, int N1f, int M, int N0, int Cs, int even_n1, int stats) {
    if (st->stopped) return;
    __shared__ C2<T> red[8][33];
    __shared__ C2<T> qs[CD][32];
    __shared__ double dred[2 * 32];
    const int lane = threadIdx.x & 31, mg = threadIdx.x >> 5;
    const int wf = blockIdx.x, h = blockIdx.y * 32 + lane, b = blockIdx.z;
    const bool hv = h < N0;
    const T rho = st->rho;
    const T g = hv ? ghg[(size_t)wf * N0 + h] : (T)0;
    const size_t slab = (((size_t)b * N1f + wf) * M) * N0;
    const size_t dfc = (size_t)N1f * M * N0;
    C2<T> sig[CD], gam[CD][CD];
    SPCSC_UNROLL
    for (int c = 0; c < CD; ++c) {
        sig[c] = mk<T>(0, 0);
        SPCSC_UNROLL
        for (int d = 0; d < CD; ++d) gam[c][d] = mk<T>(0, 0);
    }
    if (hv) {
        for (int m = mg; m < M; m += 8) {
            const C2<T> z = Zf[slab + (size_t)m * N0 + h];
            const T inv = (T)1 / (gw[m].re * g + rho);
            C2<T> df[CD];
            SPCSC_UNROLL
            for (int c = 0; c < CD; ++c) df[c] = Df[(size_t)c * dfc + ((size_t)wf * M + m) * N0 + h];
            SPCSC_UNROLL
            for (int c = 0; c < CD; ++c) {
                sig[c] = sig[c] + inv * (df[c] * z);
                SPCSC_UNROLL
                for (int d = 0; d < CD; ++d) gam[c][d] = gam[c][d] + inv * mulc(df[c], df[d]);
            }
        }
    }
    // sums over the 8 filter groups, one quantity at a time through a small shared array
    auto group_sum = [&](C2<T> v) -> C2<T> {
        red[mg][lane] = v;
        __syncthreads();
        C2<T> s = mk<T>(0, 0);
        SPCSC_UNROLL
        for (int q = 0; q < 8; ++q) s = s + red[q][lane];
        __syncthreads();
        return s;
    };
    SPCSC_UNROLL
    for (int c = 0; c < CD; ++c) {
        sig[c] = group_sum(sig[c]);
        SPCSC_UNROLL
        for (int d = 0; d < CD; ++d) gam[c][d] = group_sum(gam[c][d]);
    }
    double dsum[1] = {0.0};
    if (mg == 0) {
        C2<T> A[CD][CD], rhs[CD];
        SPCSC_UNROLL
        for (int c = 0; c < CD; ++c) {
            const C2<T> sf = hv ? Sf[(((size_t)b * Cs + c) * N1f + wf) * N0 + h] : mk<T>(0, 0);
            rhs[c] = sf - rho * sig[c];
            SPCSC_UNROLL
            for (int d = 0; d < CD; ++d) {
                A[c][d] = gam[c][d];
                if (c == d) A[c][d].re += (T)1;
            }
        }
        hpd_solve<T, CD>(A, rhs, CD);
        const double wg = (wf == 0 || (even_n1 && wf == N1f - 1)) ? 1.0 : 2.0;
        SPCSC_UNROLL
        for (int c = 0; c < CD; ++c) {
            qs[c][lane] = rhs[c];
            if (hv) dsum[0] += wg * (double)abs2(rhs[c]);
        }
    }
    __syncthreads();
    double rg[1] = {0.0};
    if (hv) {
        const double wg = (wf == 0 || (even_n1 && wf == N1f - 1)) ? 1.0 : 2.0;
        for (int m = mg; m < M; m += 8) {
            const C2<T> z = Zf[slab + (size_t)m * N0 + h];
            const C2<T> gwm = gw[m];
            const T inv = (T)1 / (gwm.re * g + rho);
            C2<T> x = rho * z;
            SPCSC_UNROLL
            for (int c = 0; c < CD; ++c)
                x = x + mulc(qs[c][lane], Df[(size_t)c * dfc + ((size_t)wf * M + m) * N0 + h]);
            x = inv * x;
            Zf[slab + (size_t)m * N0 + h] = x;
            rg[0] += wg * (double)(gwm.im * g * abs2(x));
        }
    }
    if (stats) {
        block_accumulate<1>(dsum, dred, acc + ACC_DFID);
        block_accumulate<1>(rg, dred, acc + ACC_RGRAD);
    }
}

// Gram of the dictionary per frequency: G[wf][h][c][c'] = sum_m Df_c[m] conj(Df_c'[m]).
template <typename T>
SPCSC_GLOBAL void k_gram(const C2<T>* SPCSC_RESTRICT Df, C2<T>* SPCSC_RESTRICT G, int N1f, int N0,
                         int M, int Cd) {
    const size_t n = (size_t)N1f * N0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int h = (int)(i % N0), wf = (int)(i / N0);
        for (int c = 0; c < Cd; ++c)
            for (int d = 0; d < Cd; ++d) {
                C2<T> s = mk<T>(0, 0);
                for (int m = 0; m < M; ++m)
                    s = s + mulc(Df[(((size_t)c * N1f + wf) * M + m) * N0 + h],
                                 Df[(((size_t)d * N1f + wf) * M + m) * N0 + h]);
                G[(i * Cd + c) * Cd + d] = s;
            }
    }
}

// out[b][h][wf] <-> in[b][wf][h]  (complex; unit-test entry points and Xf/Df/Sf read-back)
template <typename T>
SPCSC_GLOBAL void k_swap_last2(const C2<T>* SPCSC_RESTRICT in, C2<T>* SPCSC_RESTRICT out, int A, int B) {
    // in [nb][A][B] -> out [nb][B][A]
    const size_t n = (size_t)A * B;
    const size_t base = (size_t)blockIdx.y * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int bb = (int)(i % B), aa = (int)(i / B);
        out[base + (size_t)bb * A + aa] = in[base + i];
    }
}

}  // namespace spcsc

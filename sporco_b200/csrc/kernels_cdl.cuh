// kernels_cdl.cuh -- dictionary update (constrained convolutional MOD by PGM) on the device.
//
// Replaces the array work of sporco.pgm.ccmod.ConvCnstrMOD (sporco/pgm/ccmod.py:264-404) inside
// the alternation of sporco.dictlrn.dictlrn.DictLearn.solve (dictlrn/dictlrn.py:327-363):
//   setcoef     Zf = rfftn(Z)                                    -> forward2d of the ADMM Y
//   grad_f      g_m = sum_k conj(Zf_km) (sum_m' Zf_km' Yf_m' - Sf_k)   -> k_ccmod_grad
//   xstep       Vf = Yf - g/L ; V = irfftn(Vf) ; X = Pcn(V) ; Xf = rfftn(X)
//   ystep       Yf = Xf + coef (Xf - Xfprv)
//   rsdl / obfn rfl2norm2(Xf - Yfprv), rfl2norm2(sum_m Zf Xf - Sf)/2, ||Pcn(X) - X||
// The dictionary iterate lives in the layout of Df ([Cd][N1f][M][N0] spectra, [Cd][M][N0][N1]
// real), so handing the new dictionary to the X step is a device copy.  The coefficient spectra
// Zf use the slab layout of the X step ([K][N1f][M][N0]); the only large traffic of a D step is
// one read of Zf for the gradient and one for the data-fidelity value.
#pragma once

#include "kernels.cuh"

namespace spcsc {

enum { ACC_CDL_F = 8, ACC_CDL_DFID = 9, ACC_CDL_RSDL = 10, ACC_CDL_CNS = 11 };

// One CTA per (wf, tile of HT=32 frequencies h): threads = 32 h-lanes x 8 filter groups.
// For every image k: R_k[h] = sum_m Zf[k][m][h] Yf[m][h] - Sf[k][h];  g[m][h] += conj(Zf[k][m][h]) R_k[h].
// GRAD == false: only the (plain and Hermitian-weighted) sums of |R_k|^2 are accumulated.
// Sf holds sfs channel planes per image; this launch uses plane sfc (multi-channel dictionaries
// run one launch per channel; with a single-channel dictionary the K "images" are all
// (image, channel) pairs and sfs = 1).
template <typename T, int MI, bool GRAD>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS(256)
k_ccmod_grad(const C2<T>* SPCSC_RESTRICT Zf, const C2<T>* SPCSC_RESTRICT Yf,
             const C2<T>* SPCSC_RESTRICT Sf, C2<T>* SPCSC_RESTRICT gout, double* SPCSC_RESTRICT acc,
             int K, int N1f, int M, int N0, int even_n1, int sfs, int sfc) {
    __shared__ C2<T> red[8][33];
    __shared__ double dred[2 * 32];
    const int lane = threadIdx.x & 31, mg = threadIdx.x >> 5;
    const int wf = blockIdx.x, h = blockIdx.y * 32 + lane;
    const bool hv = h < N0;
    C2<T> y[MI], g[MI];
    SPCSC_UNROLL
    for (int i = 0; i < MI; ++i) {
        const int m = mg + 8 * i;
        y[i] = (hv && m < M) ? Yf[((size_t)wf * M + m) * N0 + h] : mk<T>(0, 0);
        g[i] = mk<T>(0, 0);
    }
    double fsum = 0.0;
    for (int k = 0; k < K; ++k) {
        const C2<T>* zk = Zf + (((size_t)k * N1f + wf) * M) * N0;
        C2<T> z[MI];
        C2<T> part = mk<T>(0, 0);
        SPCSC_UNROLL
        for (int i = 0; i < MI; ++i) {
            const int m = mg + 8 * i;
            z[i] = (hv && m < M) ? zk[(size_t)m * N0 + h] : mk<T>(0, 0);
            part = part + z[i] * y[i];
        }
        red[mg][lane] = part;
        __syncthreads();
        C2<T> R = mk<T>(0, 0);
        SPCSC_UNROLL
        for (int q = 0; q < 8; ++q) R = R + red[q][lane];
        if (hv) R = R - Sf[(((size_t)k * sfs + sfc) * N1f + wf) * N0 + h];      // image k, channel sfc of sfs
        __syncthreads();
        if (GRAD) {
            SPCSC_UNROLL
            for (int i = 0; i < MI; ++i) g[i] = g[i] + mulc(R, z[i]);      // R * conj(z)
        }
        if (mg == 0 && hv) fsum += (double)abs2(R);
    }
    if (GRAD) {
        SPCSC_UNROLL
        for (int i = 0; i < MI; ++i) {
            const int m = mg + 8 * i;
            if (hv && m < M) gout[((size_t)wf * M + m) * N0 + h] = g[i];
        }
    }
    const double wgt = (wf == 0 || (even_n1 && wf == N1f - 1)) ? 1.0 : 2.0;
    double s1[1] = {fsum}, s2[1] = {fsum * wgt};
    block_accumulate<1>(s1, dred, acc + ACC_CDL_F);
    block_accumulate<1>(s2, dred, acc + ACC_CDL_DFID);
}

// Vf = ys Yf - g / L   (ys = 1; with images sharded over R ranks ys = 1/R, so that the rank sum of the
// inverse transforms is irfftn(Yf - sum_r g_r / L))
template <typename T>
SPCSC_GLOBAL void k_ccmod_step(const C2<T>* SPCSC_RESTRICT Yf, const C2<T>* SPCSC_RESTRICT g,
                               C2<T>* SPCSC_RESTRICT Vf, T L, T ys, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const C2<T> a = Yf[i], b = g[i];
        const T inv = (T)1 / L;
        if (ys == (T)1)
            Vf[i] = mk<T>(a.re - inv * b.re, a.im - inv * b.im);
        else
            Vf[i] = mk<T>(ys * a.re - inv * b.re, ys * a.im - inv * b.im);
    }
}

// Constraint-set projection of one filter per CTA (sporco/cnvrep.py:953-1033):
//   crop to the hd x wd support (all Cd channels), optional zero mean over the support of each
//   channel, normalise to unit l2 norm (a zero filter stays zero), zero elsewhere.
// V, X: real [Cd][M][N0][N1].  With `check` the squared distance ||Pcn(V) - V||^2 (over the
// support) is accumulated instead of writing X.
template <typename T>
SPCSC_GLOBAL void k_pcn(const T* SPCSC_RESTRICT V, T* SPCSC_RESTRICT X, double* SPCSC_RESTRICT acc,
                        int Cd, int M, int N0, int N1, int hd, int wd, int zero_mean, int check,
                        const int* SPCSC_RESTRICT supp = nullptr) {
    __shared__ double red[4 * 32];
    __shared__ double bc[8];
    const int m = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    // multi-scale dictionaries (cnvrep.py:277-360): filter m has its own support hm x wm <= hd x wd
    const int hm = supp ? supp[2 * m] : hd, wm = supp ? supp[2 * m + 1] : wd;
    const int ns = hm * wm;
    // per-channel means over the support, then the norm of the (mean-free) support
    double nrm2 = 0.0;
    for (int c = 0; c < Cd; ++c) {
        const T* v = V + ((size_t)c * M + m) * N0 * N1;
        double mean = 0.0;
        if (zero_mean) {
            double s[1] = {0.0};
            for (int e = tid; e < ns; e += nt) s[0] += (double)v[(size_t)(e / wm) * N1 + (e % wm)];
            if (tid == 0) bc[0] = 0.0;
            __syncthreads();
            block_accumulate<1>(s, red, bc);
            __syncthreads();
            mean = bc[0] / (double)ns;
            __syncthreads();
        }
        double s2[1] = {0.0};
        for (int e = tid; e < ns; e += nt) {
            const double x = (double)((T)((double)v[(size_t)(e / wm) * N1 + (e % wm)] - (T)mean));
            s2[0] += x * x;
        }
        if (tid == 0) bc[1] = 0.0;
        __syncthreads();
        block_accumulate<1>(s2, red, bc + 1);
        __syncthreads();
        nrm2 += bc[1];
        if (zero_mean && tid == 0) bc[2 + c] = mean;
        __syncthreads();
    }
    T vn = (T)sqrt(nrm2);
    if (vn == (T)0) vn = (T)1;
    // Outside the support X is zero and stays zero (the buffer is cleared once, at reset), so only
    // the support is written -- or, with `check`, compared: Pcn(V) - V vanishes elsewhere when V
    // is itself a projected iterate, which is the only way the reference evaluates it.
    double d2[1] = {0.0};
    for (int c = 0; c < Cd; ++c) {
        const T mean = zero_mean ? (T)bc[2 + c] : (T)0;
        const T* v = V + ((size_t)c * M + m) * N0 * N1;
        T* x = X ? X + ((size_t)c * M + m) * N0 * N1 : nullptr;
        for (int e = tid; e < ns; e += nt) {
            const size_t o = (size_t)(e / wm) * N1 + (e % wm);
            const T p = (v[o] - mean) / vn;
            if (check) {
                const double d = (double)(p - v[o]);
                d2[0] += d * d;
            } else {
                x[o] = p;
            }
        }
        if (supp) {       // the part of the largest support that lies outside this filter's own
            for (int e = tid; e < hd * wd; e += nt) {
                const int y = e / wd, xx = e % wd;
                if (y < hm && xx < wm) continue;
                const size_t o = (size_t)y * N1 + xx;
                if (check) d2[0] += (double)v[o] * (double)v[o];
                else x[o] = (T)0;
            }
        }
    }
    if (check) {
        __syncthreads();
        block_accumulate<1>(d2, red, acc + ACC_CDL_CNS);
    }
}

// Hermitian-weighted sum of |A - B|^2 over spectra in [nb][N1f][M][N0] order (rfl2norm2 weights).
template <typename T>
SPCSC_GLOBAL void k_spec_diffnorm(const C2<T>* SPCSC_RESTRICT A, const C2<T>* SPCSC_RESTRICT B,
                                  double* SPCSC_RESTRICT acc, int nb, int N1f, size_t per_wf,
                                  int even_n1) {
    __shared__ double red[32];
    const size_t n = (size_t)nb * N1f * per_wf;
    double s[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int wf = (int)((i / per_wf) % N1f);
        const double wgt = (wf == 0 || (even_n1 && wf == N1f - 1)) ? 1.0 : 2.0;
        s[0] += wgt * (double)abs2(A[i] - B[i]);
    }
    block_accumulate<1>(s, red, acc + ACC_CDL_RSDL);
}

// Backtracking terms of the dictionary update over spectra in slab order: sum Re(conj(X - Y) g)  (eval_linear_approx,
// pgm/pgm.py:886-894) and sum |X - Y|^2, plain sums over the stored half spectrum.
template <typename T>
SPCSC_GLOBAL void k_spec_lin(const C2<T>* SPCSC_RESTRICT X, const C2<T>* SPCSC_RESTRICT Y, const C2<T>* SPCSC_RESTRICT G,
                             double* SPCSC_RESTRICT acc, size_t n) {
    __shared__ double red[2 * 32];
    double s[2] = {0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const C2<T> d = X[i] - Y[i], g = G[i];
        s[0] += (double)d.re * (double)g.re + (double)d.im * (double)g.im;
        s[1] += (double)abs2(d);
    }
    block_accumulate<2>(s, red, acc + ACC_CDL_RSDL);
}

// Cropped dictionary in the reference's order (hd, wd, Cd, M) from the device order [Cd][M][N0][N1].
template <typename T>
SPCSC_GLOBAL void k_crop_dict(const T* SPCSC_RESTRICT X, T* SPCSC_RESTRICT D, int hd, int wd, int Cd,
                              int M, int N0, int N1) {
    const size_t n = (size_t)hd * wd * Cd * M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t t = i;
        const int m = (int)(t % M); t /= M;
        const int c = (int)(t % Cd); t /= Cd;
        const int x = (int)(t % wd); t /= wd;
        const int y = (int)t;
        D[i] = X[(((size_t)c * M + m) * N0 + y) * N1 + x];
    }
}

// =====================================================================================
// Consensus dictionary update: sporco.admm.ccmod.ConvCnstrMOD_Consensus (sporco/admm/ccmod.py:613-911 over
// sporco/admm/admm.py:1419-1707).  One block i per (image, coefficient channel) with its own copy X_i of the
// dictionary, dual U_i, and the consensus variable Y = the constrained dictionary:
//   x step   Xf_i = solvedbi_sm(Zf_i, rho, conj(Zf_i) Sf_i + rho rfftn(Y - U_i))      ccmod.py:787-813
//            = the ConvBPDN column kernel with the coefficient spectra of block i as its "dictionary"
//              (ColArgs::df_bstride), i.e. the roles of the images and the filters swapped
//   relax    AX_i = alpha X_i + (1 - alpha) Y                                          admm.py:1608-1616
//   y step   Y = Pcn(mean_i (AX_i + U_i))                                              admm.py:1585-1591
//            Pcn = normalise . zpad . bcrop (cnvrep.py:953-1033) and bcrop is linear: only the filter SUPPORTS of
//            the mean are formed (k_cns_support_mean) -- with images sharded over GPUs that is also all that is
//            exchanged
//   u step   U_i += AX_i - Y                                                           admm.py:434-437
//   norms    ||X||, ||X - Y||, ||U||, ||Y||, ||Yprev - Y|| for rsdl_r / rsdl_s / rsdl_rn / rsdl_sn  admm.py:1673-1707
// Device layout: X, U [NB][M][N0][N1] with batch index bb = (image, channel); Y [Cd][M][N0][N1] (bb % Cd is the
// dictionary channel of batch bb; Cd == 1: the signal channels count as further blocks, ccmod.py:697-705).
// =====================================================================================
enum { ACC_CNS_X2 = 8, ACC_CNS_R2 = 9, ACC_CNS_U2 = 10, ACC_CNS_Y2 = 11, ACC_CNS_S2 = 12 };

// W[bb] = Y[bb % Cd] - U[bb] * uinv     (U /= rsf of a change of rho is applied lazily: uinv = 1 / udiv)
// grid (x, NB): blockIdx.y is the batch index, so no per-element division; 4 values per thread and step where the
// plane size allows (the first version divided per element and ran at 2.4 TB/s, profiles/r02_configs_ncu.md)
template <typename T>
SPCSC_GLOBAL void k_cns_yu(const T* SPCSC_RESTRICT Y, const T* SPCSC_RESTRICT U, T* SPCSC_RESTRICT W, int NB,
                           int Cd, size_t plane, T uinv) {
    const size_t bb = blockIdx.y;
    if (bb >= (size_t)NB) return;
    const T* y = Y + (bb % Cd) * plane;
    const T* u = U + bb * plane;
    T* w = W + bb * plane;
    const size_t start = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    if (plane % 4 == 0) {
        for (size_t i = start * 4; i < plane; i += step * 4) {
            T a[4], b[4];
            SPCSC_UNROLL
            for (int q = 0; q < 4; ++q) { a[q] = y[i + q]; b[q] = u[i + q]; }
            SPCSC_UNROLL
            for (int q = 0; q < 4; ++q) w[i + q] = a[q] - b[q] * uinv;
        }
    } else {
        for (size_t i = start; i < plane; i += step) w[i] = y[i] - u[i] * uinv;
    }
}

// U[bb] = Y[bb % Cd] * s   (uinit with a given Y0: U = Y / rho for every block, ccmod.py:739-750)
template <typename T>
SPCSC_GLOBAL void k_cns_uinit(const T* SPCSC_RESTRICT Y, T* SPCSC_RESTRICT U, int NB, int Cd, size_t plane, T s) {
    const size_t n = (size_t)NB * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t bb = i / plane, o = i - bb * plane;
        U[i] = Y[(bb % Cd) * plane + o] * s;
    }
}

// Filter supports of  sum_{bb = c mod Cd} (alpha X[bb] + U[bb] uinv) * wsum  +  ymul (1 - alpha) Y[c]  written into the
// supports of V [Cd][M][N0][N1]; wsum = 1 / (number of blocks over all ranks), ymul = 1 / (number of ranks)
// so that the rank sum of the supports is the global mean.  One CTA per (m, c).
template <typename T>
SPCSC_GLOBAL void k_cns_support_mean(const T* SPCSC_RESTRICT X, const T* SPCSC_RESTRICT U, const T* SPCSC_RESTRICT Y,
                                     T* SPCSC_RESTRICT V, int NB, int Cd, int M, int N0, int N1, int hd, int wd,
                                     T alpha, T uinv, T wsum, T ymul) {
    const int m = blockIdx.x, c = blockIdx.y;
    const size_t img = (size_t)N0 * N1, plane = (size_t)M * img;
    for (int e = threadIdx.x; e < hd * wd; e += blockDim.x) {
        const size_t o = (size_t)m * img + (size_t)(e / wd) * N1 + (e % wd);
        T acc = 0;
        for (int bb = c; bb < NB; bb += Cd) acc += alpha * X[(size_t)bb * plane + o] + U[(size_t)bb * plane + o] * uinv;
        V[(size_t)c * plane + o] = acc * wsum + ymul * ((T)1 - alpha) * Y[(size_t)c * plane + o];
    }
}

// U[bb] <- U[bb] uinv + alpha X[bb] + (1 - alpha) Yold[c] - Ynew[c], and the sums of X^2, (X - Ynew)^2, Unew^2
// grid (x, NB) as k_cns_yu
template <typename T>
SPCSC_GLOBAL void k_cns_update(const T* SPCSC_RESTRICT X, T* SPCSC_RESTRICT U, const T* SPCSC_RESTRICT Yold,
                               const T* SPCSC_RESTRICT Ynew, double* SPCSC_RESTRICT acc, int NB, int Cd,
                               size_t plane, T alpha, T uinv) {
    __shared__ double red[3 * 32];
    const size_t bb = blockIdx.y;
    double s[3] = {0.0, 0.0, 0.0};
    if (bb < (size_t)NB) {
        const T* x = X + bb * plane;
        T* u = U + bb * plane;
        const T* y0 = Yold + (bb % Cd) * plane;
        const T* y1 = Ynew + (bb % Cd) * plane;
        const size_t start = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
        const int V = (plane % 4 == 0) ? 4 : 1;
        float fs[3] = {0.f, 0.f, 0.f};                       // per-thread partial sums of a few hundred terms at most
        for (size_t i = start * V; i < plane; i += step * V) {
            for (int q = 0; q < V; ++q) {
                const T xv = x[i + q], a = y0[i + q], b = y1[i + q];
                const T ax = alpha * xv + ((T)1 - alpha) * a;
                const T un = u[i + q] * uinv + (ax - b);
                u[i + q] = un;
                const T r = xv - b;
                if (sizeof(T) == 4) {
                    fs[0] += (float)(xv * xv); fs[1] += (float)(r * r); fs[2] += (float)(un * un);
                } else {
                    s[0] += (double)xv * (double)xv; s[1] += (double)r * (double)r; s[2] += (double)un * (double)un;
                }
            }
        }
        if (sizeof(T) == 4) { s[0] = fs[0]; s[1] = fs[1]; s[2] = fs[2]; }
    }
    block_accumulate<3>(s, red, acc + ACC_CNS_X2);
}

// sums of Ynew^2 and (Yold - Ynew)^2 over the dictionary planes
template <typename T>
SPCSC_GLOBAL void k_cns_ynorms(const T* SPCSC_RESTRICT Yold, const T* SPCSC_RESTRICT Ynew, double* SPCSC_RESTRICT acc,
                               size_t n) {
    __shared__ double red[2 * 32];
    double s[2] = {0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const T a = Yold[i], b = Ynew[i];
        s[0] += (double)b * (double)b;
        s[1] += (double)(a - b) * (double)(a - b);
    }
    block_accumulate<2>(s, red, acc + ACC_CNS_Y2);
}

// Objective on the block variables (AuxVarObj False: gEvalY False, admm.py:1641-1646): G[c] = mean over the blocks bb = c
// mod Cd of X[bb], full planes; and the part of ||Pcn(G) - G||^2 that lies outside the largest filter support (there
// Pcn(G) = 0), added to the slot k_pcn's check mode uses for the part inside.
template <typename T>
SPCSC_GLOBAL void k_cns_mean_x(const T* SPCSC_RESTRICT X, T* SPCSC_RESTRICT Gm, double* SPCSC_RESTRICT acc, int NB, int Cd,
                               int M, int N0, int N1, int hd, int wd, T winv) {
    __shared__ double red[32];
    const size_t img = (size_t)N0 * N1, plane = (size_t)M * img, n = (size_t)Cd * plane;
    double s[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / plane, o = i - c * plane;
        T a = 0;
        for (int bb = (int)c; bb < NB; bb += Cd) a += X[(size_t)bb * plane + o];
        a *= winv;
        Gm[i] = a;
        const size_t px = o % img;
        const int y = (int)(px / N1), x = (int)(px % N1);
        if (y >= hd || x >= wd) s[0] += (double)a * (double)a;
    }
    block_accumulate<1>(s, red, acc + ACC_CDL_CNS);
}

// LinSolveCheck of the consensus x step (ccmod.py:815-824): relative residual of the block solves SUMMED over the
// blocks,  ax = sum_i [conj(Zf_i) (sum_m Zf_i,m Xf_i,m) + rho Xf_i],  b = sum_i [conj(Zf_i) Sf_i + rho rfftn(Y - U_i)],
// XSlvRelRes = ||ax - b|| / ||b|| (plain norms over the stored half spectra).  One CTA per (wf, tile of 32 frequencies):
// 32 h-lanes x 8 filter groups, loop over the batches bb = c, c + Cd, ... of dictionary channel c with ax, b in registers.
enum { ACC_CNS_LSR = 13, ACC_CNS_LSB = 14 };
template <typename T, int MI>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS(256)
k_cns_linsolve(const C2<T>* SPCSC_RESTRICT Zf, const C2<T>* SPCSC_RESTRICT Xf, const C2<T>* SPCSC_RESTRICT Zin,
               const C2<T>* SPCSC_RESTRICT Sf, double* SPCSC_RESTRICT acc, int NB, int Cd, int c, int N1f, int M,
               int N0, T rho) {
    __shared__ C2<T> red[8][33];
    __shared__ double dred[2 * 32];
    const int lane = threadIdx.x & 31, mg = threadIdx.x >> 5;
    const int wf = blockIdx.x, h = blockIdx.y * 32 + lane;
    const bool hv = h < N0;
    C2<T> ax[MI], bv[MI];
    SPCSC_UNROLL
    for (int i = 0; i < MI; ++i) ax[i] = bv[i] = mk<T>(0, 0);
    for (int bb = c; bb < NB; bb += Cd) {
        const C2<T>* zk = Zf + (((size_t)(bb / Cd) * N1f + wf) * M) * N0;
        const size_t slab = (((size_t)bb * N1f + wf) * M) * N0;
        C2<T> z[MI], x[MI];
        C2<T> part = mk<T>(0, 0);
        SPCSC_UNROLL
        for (int i = 0; i < MI; ++i) {
            const int m = mg + 8 * i;
            const bool ok = hv && m < M;
            z[i] = ok ? zk[(size_t)m * N0 + h] : mk<T>(0, 0);
            x[i] = ok ? Xf[slab + (size_t)m * N0 + h] : mk<T>(0, 0);
            part = part + z[i] * x[i];
        }
        red[mg][lane] = part;
        __syncthreads();
        C2<T> sx = mk<T>(0, 0);
        SPCSC_UNROLL
        for (int q = 0; q < 8; ++q) sx = sx + red[q][lane];
        const C2<T> sf = hv ? Sf[((size_t)bb * N1f + wf) * N0 + h] : mk<T>(0, 0);
        __syncthreads();
        SPCSC_UNROLL
        for (int i = 0; i < MI; ++i) {
            const int m = mg + 8 * i;
            if (hv && m < M) {
                const C2<T> zi = Zin[slab + (size_t)m * N0 + h];
                ax[i] = ax[i] + mulc(sx, z[i]) + mk<T>(rho * x[i].re, rho * x[i].im);
                bv[i] = bv[i] + mulc(sf, z[i]) + mk<T>(rho * zi.re, rho * zi.im);
            }
        }
    }
    double s1[1] = {0.0}, s2[1] = {0.0};
    SPCSC_UNROLL
    for (int i = 0; i < MI; ++i) {
        s1[0] += (double)abs2(ax[i] - bv[i]);
        s2[0] += (double)abs2(bv[i]);
    }
    block_accumulate<1>(s1, dred, acc + ACC_CNS_LSR);
    block_accumulate<1>(s2, dred, acc + ACC_CNS_LSB);
}

}  // namespace spcsc

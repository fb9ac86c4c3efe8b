// kernels2.cuh -- kernel set v2 for the ADMM iteration (single-channel dictionary, power-of-two
// sizes that fit the one-warp register plans).  Same arithmetic as kernels.cuh, restructured
// around what the v1 profile showed (profiles/r01_v1_ncu_summary.md):
//   * transforms live in registers (fft_regs): first stage straight from global memory, later
//     stages through a swizzled per-transform shared-memory region with warp-level barriers only;
//   * the column kernel splits the M columns of a slab over a thread-block CLUSTER; each CTA keeps
//     its columns' spectra in registers between the forward transform, the Sherman-Morrison
//     update and the inverse transform; the per-frequency sums over M are combined through
//     distributed shared memory.  Small CTAs (~40 KB smem) let several be resident per SM, so
//     loads, transforms and stores of different slabs overlap without a hand-written pipeline;
//   * per-thread float partial sums, no integer divisions by run-time values, U/udiv as a
//     multiplication by a reciprocal that is exactly 1 when rho did not change.
// kernels.cuh stays the general path (any supported size, Cd > 1, double) and the set-up path.
#pragma once

#include "kernels.cuh"

#ifdef SPCSC_EMU
#define SPCSC_LAUNCH_BOUNDS2(t, b)
#else
#define SPCSC_LAUNCH_BOUNDS2(t, b) __launch_bounds__(t, b)
#endif

namespace spcsc {

// ------------------------------------------------------------------------------------
// k_row_fwd2: TR = NT/TPF rows of one (b, m) per CTA; each row by TPF = H/E lanes.
// ------------------------------------------------------------------------------------
template <typename T, int H, int E, int NT>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, 2)
k_row_fwd2(const T* SPCSC_RESTRICT A, const T* SPCSC_RESTRICT B,
           const AdmmState<T>* SPCSC_RESTRICT st, C2<T>* SPCSC_RESTRICT Zt,
           const C2<T>* SPCSC_RESTRICT tw, const C2<T>* SPCSC_RESTRICT stw, int N0, int M,
           int nb, int gated) {
    if (st && st->stopped) return;
    // `gated`: the previous iteration's prox kernel already produced these spectra; they are
    // only stale (and this kernel only has work) when rho -- hence the scaling of U -- changed.
    if (gated && st && !st->zt_stale) return;
    SPCSC_DYN_SMEM(smem_raw);
    constexpr int TPF = H / E, TR = NT / TPF, P = H + H / 16 + 1, N1f = H + 1;
    constexpr int TWLEN = stage_tw_len(H, E);
    C2<T>* reg = reinterpret_cast<C2<T>*>(smem_raw);          // [TR][P]
    C2<T>* stw_s = reg + TR * P;                               // [TWLEN]
    const int tid = threadIdx.x;
    for (int i = tid; i < TWLEN; i += NT) stw_s[i] = stw[i];
    const int g = tid / TPF, t = tid % TPF;
    T uinv = 1;
    if (st && B) {
        const T ud = st->udiv;
        if (ud != (T)1) uinv = (T)1 / ud;
    }
    const int tiles_h = N0 / TR;
    const long long ntiles = (long long)tiles_h * M * nb;
    __syncthreads();                                         // stage twiddles are in place
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int h0 = (int)(tile % tiles_h) * TR;
        const int m = (int)((tile / tiles_h) % M), b = (int)(tile / ((long long)tiles_h * M));
        const size_t rowbase = ((((size_t)b * M + m) * N0 + h0 + g) * H);
        const C2<T>* A2 = reinterpret_cast<const C2<T>*>(A) + rowbase;
        C2<T> v[E];
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) v[p] = A2[t + TPF * p];
        if (B) {
            const C2<T>* B2 = reinterpret_cast<const C2<T>*>(B) + rowbase;
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) {
                const C2<T> u = B2[t + TPF * p];
                v[p].re -= u.re * uinv;
                v[p].im -= u.im * uinv;
            }
        }
        fft_regs<T, H, E, false>(v, reg + g * P, stw_s, t);
        __syncwarp();
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) reg[g * P + t + TPF * p] = v[p];
        __syncthreads();
        C2<T>* out = Zt + (((size_t)b * N1f) * M + m) * N0 + h0;
        const size_t wstride = (size_t)M * N0;
        for (int e = tid; e < TR * N1f; e += NT) {
            const int wf = e / TR, r = e % TR;
            const C2<T> a = reg[r * P + (wf == H ? 0 : wf)];
            const C2<T> bb = conj(reg[r * P + (wf == 0 ? 0 : H - wf)]);
            const C2<T> w = tw[wf];
            const C2<T> sum = a + bb, dif = mul_mi((a - bb) * w);
            out[wf * wstride + r] = mk<T>((T)0.5 * (sum.re + dif.re), (T)0.5 * (sum.im + dif.im));
        }
        __syncthreads();                                     // reg is reused by the next tile
    }
}

// ------------------------------------------------------------------------------------
// k_row_inv_prox2: inverse row transform + relaxation + prox + dual update + residual sums.
// ------------------------------------------------------------------------------------
template <typename T, int H, int E, int CX, int NT>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, (CX == 1 ? 2 : 1))
k_row_inv_prox2(const C2<T>* SPCSC_RESTRICT Zt, T* SPCSC_RESTRICT Y, T* SPCSC_RESTRICT U,
                const AdmmState<T>* SPCSC_RESTRICT st, AdmmParams<T> prm, WeightView<T> wl1,
                WeightView<T> wl21, double* SPCSC_RESTRICT acc, const C2<T>* SPCSC_RESTRICT tw,
                const C2<T>* SPCSC_RESTRICT stw, int N0, int M, T scale, int nonneg, int bnd0,
                int bnd1, int reg_on_y) {
    if (st->stopped) return;
    SPCSC_DYN_SMEM(smem_raw);
    constexpr int TPF = H / E, TR = NT / TPF, P = H + H / 16 + 1, N1f = H + 1;
    constexpr int TWLEN = stage_tw_len(H, E);
    C2<T>* reg = reinterpret_cast<C2<T>*>(smem_raw);          // [CX][TR][P]
    C2<T>* stw_s = reg + CX * TR * P;
    const int tid = threadIdx.x;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, k = blockIdx.z;
    const bool ams = m >= prm.ams_m0;   // additive-mask-simulation map: no clipping, not part of RegL1
    for (int i = tid; i < TWLEN; i += NT) stw_s[i] = stw[i];
    const size_t wstride = (size_t)M * N0;
    SPCSC_UNROLL
    for (int c = 0; c < CX; ++c) {
        const C2<T>* in = Zt + (((size_t)(k * CX + c) * N1f) * M + m) * N0 + h0;
        for (int e = tid; e < TR * N1f; e += NT) {
            const int wf = e / TR, r = e % TR;
            reg[(c * TR + r) * P + wf] = in[wf * wstride + r];
        }
    }
    __syncthreads();
    const int g = tid / TPF, t = tid % TPF;
    const int h = h0 + g;
    C2<T> v[CX][E];
    SPCSC_UNROLL
    for (int c = 0; c < CX; ++c) {
        C2<T>* row = reg + (c * TR + g) * P;
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) {
            const int kk = t + TPF * p;
            if (kk == 0) {
                const T a = row[0].re, cc = row[H].re;       // c2r ignores the imaginary parts
                v[c][p] = mk<T>(a + cc, a - cc);
            } else {
                const C2<T> Xa = row[kk], Xb = row[H - kk];
                const C2<T> w = tw[kk];
                const C2<T> s1 = Xa + conj(Xb), d1 = Xa - conj(Xb);
                v[c][p] = s1 + mul_i(mulc(d1, w));
            }
        }
        __syncwarp();
        fft_regs<T, H, E, true>(v[c], row, stw_s, t);
    }

    const T rho = st->rho;
    T uinv = 1;
    {
        const T ud = st->udiv;
        if (ud != (T)1) uinv = (T)1 / ud;
    }
    const T lr = prm.lmbda / rho;
    const T mr = prm.joint ? prm.mu / rho : (T)0;
    const T rlx = prm.rlx;
    const bool relax = rlx != (T)1;
    const T rl1 = (T)1 - rlx;
    T sums[7] = {0, 0, 0, 0, 0, 0, 0};
    T w1u[CX];
    const size_t wbase = (size_t)k * wl1.sk + (size_t)m * wl1.sm + (size_t)h * wl1.s0;
    SPCSC_UNROLL
    for (int c = 0; c < CX; ++c) w1u[c] = wl1.p[(size_t)k * wl1.sk + (size_t)c * wl1.sc + (size_t)m * wl1.sm];
    const size_t w21base = (size_t)k * wl21.sk + (size_t)m * wl21.sm + (size_t)h * wl21.s0;

    SPCSC_UNROLL
    for (int p = 0; p < E; ++p) {
        const int j = t + TPF * p;
        T wv[CX][2], ax[CX][2], ue[CX][2], yp[CX][2];
        T a2[2] = {0, 0}, g2[2] = {0, 0};
        SPCSC_UNROLL
        for (int c = 0; c < CX; ++c) {
            const size_t off = ((((size_t)(k * CX + c) * M + m) * N0 + h) * H + j);
            const C2<T> y2 = reinterpret_cast<const C2<T>*>(Y)[off];
            const C2<T> u2 = reinterpret_cast<const C2<T>*>(U)[off];
            const T xs[2] = {v[c][p].re * scale, v[c][p].im * scale};
            const T ys[2] = {y2.re, y2.im};
            const T us[2] = {u2.re * uinv, u2.im * uinv};
            v[c][p] = mk<T>(xs[0], xs[1]);
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                const T w1 = wl1.spatial_uniform
                                 ? w1u[c]
                                 : wl1.p[wbase + (size_t)c * wl1.sc + (size_t)(2 * j + q) * wl1.s1];
                const T axv = relax ? rlx * xs[q] + rl1 * ys[q] : xs[q];
                const T vv = axv + us[q];
                const T w = soft_threshold(vv, lr * w1);
                yp[c][q] = ys[q];
                ax[c][q] = axv;
                ue[c][q] = us[q];
                wv[c][q] = w;
                a2[q] += w * w;
                if (!reg_on_y) {
                    sums[ACC_L1] += ams ? 0 : fabs(w1 * xs[q]);
                    g2[q] += xs[q] * xs[q];
                }
            }
        }
        T fac[2] = {1, 1}, w21[2] = {1, 1};
        if (prm.joint) {
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                w21[q] = wl21.p[w21base + (size_t)(2 * j + q) * wl21.s1];
                const T a = sqrt(a2[q]);
                const T bq = fmax((T)0, a - mr * w21[q]);
                fac[q] = (a != (T)0) ? bq / a : (T)0;
            }
        }
        SPCSC_UNROLL
        for (int c = 0; c < CX; ++c) {
            const size_t off = ((((size_t)(k * CX + c) * M + m) * N0 + h) * H + j);
            T yn[2], un[2];
            const T xs[2] = {v[c][p].re, v[c][p].im};
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                T y = prm.joint ? fac[q] * wv[c][q] : wv[c][q];
                if (nonneg && !ams && y < (T)0) y = (T)0;
                if (!ams && (h >= bnd0 || (2 * j + q) >= bnd1)) y = (T)0;
                const T u = ue[c][q] + (ax[c][q] - y);
                yn[q] = y;
                un[q] = u;
                const T x = xs[q];
                const T dr = x - y, ds = yp[c][q] - y;
                sums[ACC_X2] += x * x;
                sums[ACC_Y2] += y * y;
                sums[ACC_U2] += u * u;
                sums[ACC_R2] += dr * dr;
                sums[ACC_S2] += ds * ds;
                if (reg_on_y) {
                    const T w1 = wl1.spatial_uniform
                                     ? w1u[c]
                                     : wl1.p[wbase + (size_t)c * wl1.sc + (size_t)(2 * j + q) * wl1.s1];
                    sums[ACC_L1] += ams ? 0 : fabs(w1 * y);
                    g2[q] += y * y;
                }
            }
            reinterpret_cast<C2<T>*>(Y)[off] = mk<T>(yn[0], yn[1]);
            reinterpret_cast<C2<T>*>(U)[off] = mk<T>(un[0], un[1]);
        }
        if (prm.joint) sums[ACC_L21] += w21[0] * sqrt(g2[0]) + w21[1] * sqrt(g2[1]);
    }
    if (prm.need_rsdl || prm.need_obj) {
        double d[7];
        SPCSC_UNROLL
        for (int i = 0; i < 7; ++i) d[i] = (double)sums[i];
        double* red = reinterpret_cast<double*>(smem_raw);
        block_accumulate_det<7>(d, red, reinterpret_cast<unsigned long long*>(acc + ACC_N));
    }
}

// ------------------------------------------------------------------------------------
// k_row_inv_prox3: inverse row transform + relaxation + prox (l1, or l1 + l2,1 over the CX
// coefficient channels) + dual update + residual sums + (optionally) the NEXT iteration's
// Y - U row spectra.  Every global read of the CTA -- the transposed Zt tiles and the contiguous
// Y and U tiles (TR rows x N1 reals per channel) -- is issued up front as asynchronous
// global->shared copies (cp.async), so the memory system is kept busy independently of the
// register budget; the transforms and the prox then run out of shared memory and registers.
// (The v2 profile showed the synchronous-load version waiting on its own loads 60 % of the time.)
// PLAIN: no NonNegCoef / NoBndryCross, spatially uniform weights, regulariser evaluated on X --
// the common configuration; the flag tests and per-element weight fetches then compile away.
// ------------------------------------------------------------------------------------
// Shared-memory plan of k_row_inv_prox3.  A warp holds 32/TPF rows; 64-bit shared accesses are
// served per half-warp, so with TPF = 8 two rows share a pass and must fall on complementary
// halves of the 32 banks.  REMAP: lane groups 2i and 2i+1 take rows i and i+8 -- 8 rows of odd
// stride P apart is 16 banks -- and rows >= 8 of the Y/U tiles are shifted by 8 elements for the
// same reason (the profile of the unshifted layout showed every row access 2-way conflicted).
template <typename T, int H, int E, int CX, int NT>
struct Prox3Plan {
    static constexpr int TPF = H / E, TR = NT / TPF, P = H + H / 16 + 1, N1f = H + 1;
    static constexpr bool REMAP = (sizeof(C2<T>) == 8 && TPF == 8 && E == 16 && TR % 16 == 0);
    static constexpr int YS = TR * H + (REMAP ? TR : 0);       // one channel of a Y / U tile
    static constexpr int TWLEN = stage_tw_len(H, E);
    // TWG: with three or more coefficient channels the tiles alone are 75 KB; the two twiddle tables (2.5 KB) are then
    // read from global memory through L1 instead of being staged, which is what lets a third CTA fit on the SM
    // (profiles/r02_configs_ncu.md: the joint prox of cfg3a ran at 12 % warp occupancy with two)
    static constexpr bool TWG = (CX >= 3);
    static constexpr size_t smem_bytes =
        ((size_t)CX * (2 * YS + TR * P) + (TWG ? 0 : TWLEN + N1f)) * sizeof(C2<T>);
    static SPCSC_HD int row_of_group(int gi) {
        return REMAP ? ((gi & ~15) | ((gi & 1) << 3) | ((gi & 15) >> 1)) : gi;
    }
    static SPCSC_HD int yoff(int g) { return g * H + (REMAP ? 8 * (g >> 3) : 0); }
};

template <typename T, int H, int E, int CX, int NT, bool PLAIN>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, (NT <= 128 ? (CX == 1 ? 4 : 3) : 2))
k_row_inv_prox3(const C2<T>* SPCSC_RESTRICT Zt, C2<T>* SPCSC_RESTRICT Znext, T* SPCSC_RESTRICT Y,
                T* SPCSC_RESTRICT U,
                const AdmmState<T>* SPCSC_RESTRICT st, AdmmParams<T> prm, WeightView<T> wl1,
                WeightView<T> wl21, double* SPCSC_RESTRICT acc, const C2<T>* SPCSC_RESTRICT tw,
                const C2<T>* SPCSC_RESTRICT stw, int N0, int M, T scale, int nonneg, int bnd0,
                int bnd1, int reg_on_y) {
    if (st->stopped) return;
    SPCSC_DYN_SMEM(smem_raw);
    using PL = Prox3Plan<T, H, E, CX, NT>;
    constexpr int TPF = PL::TPF, TR = PL::TR, P = PL::P, N1f = PL::N1f, YS = PL::YS;
    constexpr int TWLEN = PL::TWLEN;
    constexpr int VEC = 16 / sizeof(C2<T>);                    // complex values per 16-byte copy
    constexpr int WSTEP = NT / TR;
    constexpr int WIT = (N1f + WSTEP - 1) / WSTEP;             // wf iterations of the tile loops
    C2<T>* ybuf = reinterpret_cast<C2<T>*>(smem_raw);          // [CX][YS]  (rows of H complex pairs)
    C2<T>* ubuf = ybuf + CX * YS;                              // [CX][YS]
    C2<T>* reg = ubuf + CX * YS;                               // [CX][TR][P]
    C2<T>* tab_s = reg + CX * TR * P;                          // [TWLEN] stage twiddles, [N1f] split twiddles
    const C2<T>* stw_s = PL::TWG ? stw : tab_s;                // (TWG: read from global memory, not staged)
    const C2<T>* tw_s = PL::TWG ? tw : tab_s + TWLEN;
    const int tid = threadIdx.x;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, k = blockIdx.z;
    const bool ams = m >= prm.ams_m0;   // additive-mask-simulation map: no clipping, not part of RegL1
    const size_t wstride = (size_t)M * N0;
    const int gr = tid % TR, wf0 = tid / TR;                   // this thread's row / first wf in tile loops
    // group 0: the twiddle tables (asynchronous as well: the profile of the synchronous copy showed 16 % of
    // the kernel's stall samples on the stores that waited for these loads) ...
    if constexpr (!PL::TWG) {
        for (int i = tid; i < TWLEN; i += NT) cp_async<sizeof(C2<T>)>(tab_s + i, stw + i);
        for (int i = tid; i < N1f; i += NT) cp_async<sizeof(C2<T>)>(tab_s + TWLEN + i, tw + i);
    }
    SPCSC_UNROLL
    for (int c = 0; c < CX; ++c) {   // ... and the Zt tiles, transposed on the fly
        const C2<T>* src = Zt + (((size_t)(k * CX + c) * N1f) * M + m) * N0 + h0 +
                           (size_t)wf0 * wstride + gr;
        C2<T>* dst = reg + (c * TR + gr) * P + wf0;
        SPCSC_UNROLL
        for (int it = 0; it < WIT; ++it)
            if (wf0 + it * WSTEP < N1f)
                cp_async<sizeof(C2<T>)>(dst + it * WSTEP, src + (size_t)it * WSTEP * wstride);
    }
    cp_async_commit();
    SPCSC_UNROLL
    for (int c = 0; c < CX; ++c) {   // group 1: the Y and U tiles, contiguous
        const size_t tile = ((((size_t)(k * CX + c) * M + m) * N0 + h0) * H);
        const C2<T>* y2 = reinterpret_cast<const C2<T>*>(Y) + tile;
        const C2<T>* u2 = reinterpret_cast<const C2<T>*>(U) + tile;
        SPCSC_UNROLL
        for (int e = tid * VEC; e < TR * H; e += NT * VEC) {
            const int d = c * YS + e + (PL::REMAP ? 8 * (e / (8 * H)) : 0);
            cp_async<16>(ybuf + d, y2 + e);
            cp_async<16>(ubuf + d, u2 + e);
        }
    }
    cp_async_commit();
    const int g = PL::row_of_group(tid / TPF), t = tid % TPF;
    const int h = h0 + g;
    const int yrow = PL::yoff(g);
    cp_async_wait<1>();
    __syncthreads();

    C2<T> v[CX][E];
    SPCSC_UNROLL
    for (int c = 0; c < CX; ++c) {
        C2<T>* row = reg + (c * TR + g) * P;
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) {
            const int kk = t + TPF * p;
            if (kk == 0) {
                const T a = row[0].re, cc = row[H].re;       // c2r ignores the imaginary parts
                v[c][p] = mk<T>(a + cc, a - cc);
            } else {
                const C2<T> Xa = row[kk], Xb = row[H - kk];
                const C2<T> w = tw_s[kk];
                const C2<T> s1 = Xa + conj(Xb), d1 = Xa - conj(Xb);
                v[c][p] = s1 + mul_i(mulc(d1, w));
            }
        }
        __syncwarp();
        fft_regs<T, H, E, true>(v[c], row, stw_s, t);
    }
    cp_async_wait<0>();
    __syncthreads();

    const T rho = st->rho;
    T uinv = 1;
    {
        const T ud = st->udiv;
        if (ud != (T)1) uinv = (T)1 / ud;
    }
    const T lr = prm.lmbda / rho;
    const T mr = prm.joint ? prm.mu / rho : (T)0;
    const T rlx = prm.rlx;
    const bool relax = rlx != (T)1;
    const T rl1 = (T)1 - rlx;
    T sums[7] = {0, 0, 0, 0, 0, 0, 0};
    T w1u[CX];
    SPCSC_UNROLL
    for (int c = 0; c < CX; ++c)
        w1u[c] = wl1.p[(size_t)k * wl1.sk + (size_t)c * wl1.sc + (size_t)m * wl1.sm];
    const T w21u = prm.joint ? wl21.p[(size_t)k * wl21.sk + (size_t)m * wl21.sm] : (T)0;
    // with a single coefficient channel the l2 norm over the channel axis is |.|, so
    // S_2,beta(S_1,alpha(v)) = S_1,alpha+beta(v)  (prox/_l21.py:88 with a 1-element axis)
    const T joint1 = (CX == 1) ? mr * w21u : (T)0;
    const size_t wbase = (size_t)k * wl1.sk + (size_t)m * wl1.sm + (size_t)h * wl1.s0;
    const size_t w21base = (size_t)k * wl21.sk + (size_t)m * wl21.sm + (size_t)h * wl21.s0;
    if constexpr (PLAIN && CX == 1) {
        // The common configuration, on (re, im) pairs with packed arithmetic: threshold as
        // v - clamp(v, -t, t); weights are uniform, so sum |w x| = |w| sum |x| (likewise l2,1).
        const T thr = lr * w1u[0] + joint1;
        C2<T> sx2 = mk<T>(0, 0), sy2 = sx2, su2 = sx2, sr2 = sx2, ss2 = sx2;
        T sabs = 0;
        const size_t tile = ((((size_t)k * M + m) * N0 + h) * H);
        C2<T>* yg = reinterpret_cast<C2<T>*>(Y) + tile;
        C2<T>* ug = reinterpret_cast<C2<T>*>(U) + tile;
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) {
            const int j = t + TPF * p;
            const C2<T> y2 = ybuf[yrow + j], u2 = ubuf[yrow + j];
            const C2<T> xs = pmul(v[0][p], scale);
            const C2<T> us = pmul(u2, uinv);
            const C2<T> ax = relax ? pfma(xs, rlx, pmul(y2, rl1)) : xs;
            const C2<T> vv = ax + us;
            const C2<T> cl = mk<T>(fmin(fmax(vv.re, -thr), thr), fmin(fmax(vv.im, -thr), thr));
            const C2<T> y = vv - cl;
            const C2<T> u = us + (ax - y);
            const C2<T> dr = xs - y, ds = y2 - y;
            sx2 = pfma(xs, xs, sx2);
            sy2 = pfma(y, y, sy2);
            su2 = pfma(u, u, su2);
            sr2 = pfma(dr, dr, sr2);
            ss2 = pfma(ds, ds, ss2);
            sabs += fabs(xs.re) + fabs(xs.im);
            yg[j] = y;
            ug[j] = u;
            v[0][p] = y - u;                                   // next x-step input, if rho stays
        }
        sums[ACC_X2] = sx2.re + sx2.im;
        sums[ACC_Y2] = sy2.re + sy2.im;
        sums[ACC_U2] = su2.re + su2.im;
        sums[ACC_R2] = sr2.re + sr2.im;
        sums[ACC_S2] = ss2.re + ss2.im;
        sums[ACC_L1] = fabs(w1u[0]) * sabs;
        if (prm.joint) sums[ACC_L21] = w21u * sabs;
    } else if constexpr (PLAIN && CX > 1) {
        // Several coefficient channels, common configuration: the same packed (re, im)-pair arithmetic;
        // the l2 shrinkage over the channel axis (prox_sl1l2) is one factor per pixel.
        C2<T> sx2 = mk<T>(0, 0), sy2 = sx2, su2 = sx2, sr2 = sx2, ss2 = sx2;
        T sl1 = 0, sl21 = 0;
        const T mrw = mr * w21u;
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) {
            const int j = t + TPF * p;
            C2<T> xs[CX], ax[CX], us[CX], yp[CX], w[CX];
            C2<T> a2 = mk<T>(0, 0), g2 = mk<T>(0, 0);
            SPCSC_UNROLL
            for (int c = 0; c < CX; ++c) {
                yp[c] = ybuf[c * YS + yrow + j];
                xs[c] = pmul(v[c][p], scale);
                us[c] = pmul(ubuf[c * YS + yrow + j], uinv);
                ax[c] = relax ? pfma(xs[c], rlx, pmul(yp[c], rl1)) : xs[c];
                const C2<T> vv = ax[c] + us[c];
                const T thr = lr * w1u[c];
                w[c] = vv - mk<T>(fmin(fmax(vv.re, -thr), thr), fmin(fmax(vv.im, -thr), thr));
                a2 = pfma(w[c], w[c], a2);
                g2 = pfma(xs[c], xs[c], g2);
                sl1 += fabs(w1u[c]) * (fabs(xs[c].re) + fabs(xs[c].im));
            }
            C2<T> fac = mk<T>(1, 1);
            if (prm.joint) {
                const T a0 = sqrt(a2.re), a1 = sqrt(a2.im);
                fac = mk<T>(a0 != (T)0 ? fmax((T)0, a0 - mrw) / a0 : (T)0,
                            a1 != (T)0 ? fmax((T)0, a1 - mrw) / a1 : (T)0);
                sl21 += w21u * (sqrt(g2.re) + sqrt(g2.im));
            }
            SPCSC_UNROLL
            for (int c = 0; c < CX; ++c) {
                const C2<T> y = prm.joint ? pfma(w[c], fac, mk<T>(0, 0)) : w[c];
                const C2<T> u = us[c] + (ax[c] - y);
                const C2<T> dr = xs[c] - y, ds = yp[c] - y;
                sx2 = pfma(xs[c], xs[c], sx2);
                sy2 = pfma(y, y, sy2);
                su2 = pfma(u, u, su2);
                sr2 = pfma(dr, dr, sr2);
                ss2 = pfma(ds, ds, ss2);
                const size_t off = ((((size_t)(k * CX + c) * M + m) * N0 + h) * H + j);
                reinterpret_cast<C2<T>*>(Y)[off] = y;
                reinterpret_cast<C2<T>*>(U)[off] = u;
                v[c][p] = y - u;                               // next x-step input, if rho stays
            }
        }
        sums[ACC_X2] = sx2.re + sx2.im;
        sums[ACC_Y2] = sy2.re + sy2.im;
        sums[ACC_U2] = su2.re + su2.im;
        sums[ACC_R2] = sr2.re + sr2.im;
        sums[ACC_S2] = ss2.re + ss2.im;
        sums[ACC_L1] = sl1;
        sums[ACC_L21] = sl21;
    } else {
    SPCSC_UNROLL
    for (int p = 0; p < E; ++p) {
        const int j = t + TPF * p;
        T wv[CX][2], ax[CX][2], ue[CX][2], yp[CX][2], w1s[CX][2];
        T a2[2] = {0, 0};
        SPCSC_UNROLL
        for (int c = 0; c < CX; ++c) {
            const C2<T> y2 = ybuf[c * YS + yrow + j], u2 = ubuf[c * YS + yrow + j];
            const T xs[2] = {v[c][p].re * scale, v[c][p].im * scale};
            const T ys[2] = {y2.re, y2.im};
            const T us[2] = {u2.re * uinv, u2.im * uinv};
            v[c][p] = mk<T>(xs[0], xs[1]);
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                const T w1 = (PLAIN || wl1.spatial_uniform)
                                 ? w1u[c]
                                 : wl1.p[wbase + (size_t)c * wl1.sc + (size_t)(2 * j + q) * wl1.s1];
                const T axv = relax ? rlx * xs[q] + rl1 * ys[q] : xs[q];
                const T w = soft_threshold(axv + us[q], lr * w1 + joint1);
                yp[c][q] = ys[q];
                ax[c][q] = axv;
                ue[c][q] = us[q];
                wv[c][q] = w;
                w1s[c][q] = w1;
                a2[q] += w * w;
            }
        }
        T fac[2] = {1, 1}, w21[2] = {w21u, w21u};
        if (CX > 1 && prm.joint) {
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                if (!PLAIN && !wl21.spatial_uniform) w21[q] = wl21.p[w21base + (size_t)(2 * j + q) * wl21.s1];
                const T a = sqrt(a2[q]);
                const T bq = fmax((T)0, a - mr * w21[q]);
                fac[q] = (a != (T)0) ? bq / a : (T)0;
            }
        }
        T g2[2] = {0, 0};
        SPCSC_UNROLL
        for (int c = 0; c < CX; ++c) {
            T yn[2], un[2];
            SPCSC_UNROLL
            for (int q = 0; q < 2; ++q) {
                T y = (CX > 1 && prm.joint) ? fac[q] * wv[c][q] : wv[c][q];
                if (!PLAIN) {
                    if (nonneg && !ams && y < (T)0) y = (T)0;
                    if (!ams && (h >= bnd0 || (2 * j + q) >= bnd1)) y = (T)0;
                }
                const T u = ue[c][q] + (ax[c][q] - y);
                yn[q] = y;
                un[q] = u;
                const T x = (q == 0) ? v[c][p].re : v[c][p].im;
                const T dr = x - y, ds = yp[c][q] - y;
                sums[ACC_X2] += x * x;
                sums[ACC_Y2] += y * y;
                sums[ACC_U2] += u * u;
                sums[ACC_R2] += dr * dr;
                sums[ACC_S2] += ds * ds;
                const T gq = (!PLAIN && reg_on_y) ? y : x;
                sums[ACC_L1] += ams ? 0 : fabs(w1s[c][q] * gq);
                g2[q] += gq * gq;
            }
            const size_t off = ((((size_t)(k * CX + c) * M + m) * N0 + h) * H + j);
            reinterpret_cast<C2<T>*>(Y)[off] = mk<T>(yn[0], yn[1]);
            reinterpret_cast<C2<T>*>(U)[off] = mk<T>(un[0], un[1]);
            v[c][p] = mk<T>(yn[0] - un[0], yn[1] - un[1]);   // next x-step input, if rho stays
        }
        if (prm.joint) sums[ACC_L21] += w21[0] * sqrt(g2[0]) + w21[1] * sqrt(g2[1]);
    }
    }
    if (prm.need_rsdl || prm.need_obj) {
#ifdef SPCSC_OLD_REDUCE
        if constexpr (false) {
#else
        if constexpr (sizeof(T) == 4) {
#endif
            const float s8[8] = {(float)sums[0], (float)sums[1], (float)sums[2], (float)sums[3],
                                 (float)sums[4], (float)sums[5], (float)sums[6], 0.f};
            block_accumulate_det_f<7>(s8, reinterpret_cast<float*>(smem_raw),
                                      reinterpret_cast<unsigned long long*>(acc + ACC_N));
        } else {
            double d[7];
            SPCSC_UNROLL
            for (int i = 0; i < 7; ++i) d[i] = (double)sums[i];
            double* red = reinterpret_cast<double*>(smem_raw);
            block_accumulate_det<7>(d, red, reinterpret_cast<unsigned long long*>(acc + ACC_N));
        }
    }
    if (Znext && st->emit) {
        // Cross-iteration fusion: the row spectra of Y - U for the next iteration, valid as long
        // as rho (hence the scaling of U) does not change; otherwise k_row_fwd3 redoes them.  Skipped
        // (st->emit, set by the scalar kernel) in the iteration after a change of rho.
        SPCSC_UNROLL
        for (int c = 0; c < CX; ++c) {
            C2<T>* row = reg + (c * TR + g) * P;
            fft_regs<T, H, E, false>(v[c], row, stw_s, t);
            __syncwarp();
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) row[t + TPF * p] = v[c][p];
        }
        __syncthreads();
        SPCSC_UNROLL
        for (int c = 0; c < CX; ++c) {
            const C2<T>* row = reg + (c * TR + gr) * P;
            C2<T>* out = Znext + (((size_t)(k * CX + c) * N1f) * M + m) * N0 + h0 +
                         (size_t)wf0 * wstride + gr;
            SPCSC_UNROLL
            for (int it = 0; it < WIT; ++it) {
                const int wf = wf0 + it * WSTEP;
                if (wf < N1f) {
                    const C2<T> a = row[wf == H ? 0 : wf];
                    const C2<T> bb = conj(row[wf == 0 ? 0 : H - wf]);
                    const C2<T> w = tw_s[wf];
                    const C2<T> sum = a + bb, dif = mul_mi((a - bb) * w);
                    out[(size_t)it * WSTEP * wstride] =
                        mk<T>((T)0.5 * (sum.re + dif.re), (T)0.5 * (sum.im + dif.im));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// k_row_fwd3: Zt = row spectra of A - B / udiv, as k_row_fwd2, with the tile flow of k_row_inv_prox3: the
// contiguous A and B tiles (TR rows x N1 reals) arrive by 16-byte asynchronous copies in the conflict-free
// row layout of Prox3Plan, and -- the CTAs walk over tiles with a grid stride -- the copies of the NEXT tile are
// issued as soon as every thread holds its values of the current one, so they overlap the transform and the
// transposed stores.  This is the kernel that redoes the row spectra when rho changed (`gated`), i.e. in
// every other iteration of the first ~25 of an AutoRho run.
// ------------------------------------------------------------------------------------
template <typename T, int H, int E, int NT>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, (NT <= 128 ? 4 : 2))
k_row_fwd3(const T* SPCSC_RESTRICT A, const T* SPCSC_RESTRICT B,
           const AdmmState<T>* SPCSC_RESTRICT st, C2<T>* SPCSC_RESTRICT Zt,
           const C2<T>* SPCSC_RESTRICT tw, const C2<T>* SPCSC_RESTRICT stw, int N0, int M,
           int nb, int gated) {
    if (st && st->stopped) return;
    if (gated && st && !st->zt_stale) return;
    SPCSC_DYN_SMEM(smem_raw);
    using PL = Prox3Plan<T, H, E, 1, NT>;
    constexpr int TPF = PL::TPF, TR = PL::TR, P = PL::P, N1f = PL::N1f, YS = PL::YS, TWLEN = PL::TWLEN;
    constexpr int VEC = 16 / sizeof(C2<T>);
    constexpr int WSTEP = NT / TR, WIT = (N1f + WSTEP - 1) / WSTEP;
    C2<T>* abuf = reinterpret_cast<C2<T>*>(smem_raw);          // [YS]
    C2<T>* bbuf = abuf + YS;                                   // [YS]
    C2<T>* reg = bbuf + YS;                                    // [TR][P]
    C2<T>* stw_s = reg + TR * P;                               // [TWLEN]
    C2<T>* tw_s = stw_s + TWLEN;                               // [N1f]
    const int tid = threadIdx.x;
    T uinv = 1;
    if (st && B) {
        const T ud = st->udiv;
        if (ud != (T)1) uinv = (T)1 / ud;
    }
    const int tiles_h = N0 / TR;
    const long long ntiles = (long long)tiles_h * M * nb;
    const int g = PL::row_of_group(tid / TPF), t = tid % TPF;
    const int yrow = PL::yoff(g);
    const int gr = tid % TR, wf0 = tid / TR;
    const size_t wstride = (size_t)M * N0;
    auto issue = [&](long long tile) {
        const int h0 = (int)(tile % tiles_h) * TR;
        const int m = (int)((tile / tiles_h) % M), b = (int)(tile / ((long long)tiles_h * M));
        const size_t base = ((((size_t)b * M + m) * N0 + h0) * H);
        const C2<T>* a2 = reinterpret_cast<const C2<T>*>(A) + base;
        const C2<T>* b2 = B ? reinterpret_cast<const C2<T>*>(B) + base : nullptr;
        SPCSC_UNROLL
        for (int e = tid * VEC; e < TR * H; e += NT * VEC) {
            const int d = e + (PL::REMAP ? 8 * (e / (8 * H)) : 0);
            cp_async<16>(abuf + d, a2 + e);
            if (b2) cp_async<16>(bbuf + d, b2 + e);
        }
    };
    for (int i = tid; i < TWLEN; i += NT) cp_async<sizeof(C2<T>)>(stw_s + i, stw + i);
    for (int i = tid; i < N1f; i += NT) cp_async<sizeof(C2<T>)>(tw_s + i, tw + i);
    if ((long long)blockIdx.x < ntiles) issue(blockIdx.x);
    cp_async_commit();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int h0 = (int)(tile % tiles_h) * TR;
        const int m = (int)((tile / tiles_h) % M), b = (int)(tile / ((long long)tiles_h * M));
        cp_async_wait<0>();
        __syncthreads();
        C2<T> v[E];
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) {
            const int j = t + TPF * p;
            C2<T> a = abuf[yrow + j];
            if (B) a = a - pmul(bbuf[yrow + j], uinv);
            v[p] = a;
        }
        __syncthreads();                                     // the tile buffers are free: fetch the next tile now
        if (tile + gridDim.x < ntiles) issue(tile + gridDim.x);
        cp_async_commit();
        C2<T>* row = reg + g * P;
        fft_regs<T, H, E, false>(v, row, stw_s, t);
        __syncwarp();
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) row[t + TPF * p] = v[p];
        __syncthreads();
        {
            const C2<T>* rrow = reg + gr * P;
            C2<T>* out = Zt + (((size_t)b * N1f) * M + m) * N0 + h0 + (size_t)wf0 * wstride + gr;
            SPCSC_UNROLL
            for (int it = 0; it < WIT; ++it) {
                const int wf = wf0 + it * WSTEP;
                if (wf < N1f) {
                    const C2<T> aa = rrow[wf == H ? 0 : wf];
                    const C2<T> bb = conj(rrow[wf == 0 ? 0 : H - wf]);
                    const C2<T> sum = aa + bb, dif = mul_mi((aa - bb) * tw_s[wf]);
                    out[(size_t)it * WSTEP * wstride] =
                        mk<T>((T)0.5 * (sum.re + dif.re), (T)0.5 * (sum.im + dif.im));
                }
            }
        }
        // reg is next written after the barrier that follows the wait at the top of the loop
    }
}

// ------------------------------------------------------------------------------------
// k_row_prox_fwd3: the PGM proximal step in the row domain (pgm/pgm.py:796-800, pgm/cbpdn.py:288-298)
// on the register plans: V = irfft_row(Vt)*scale ; X = prox_l1(V, (lmbda/L) wl1) [+NonNeg,
// NoBndryCross] stored; Xt = rfft_row(X) written over Vt (the CTA owns its TR rows of every wf);
// RegL1 accumulated.  Same tile flow and conflict-free row mapping as k_row_inv_prox3.
// ------------------------------------------------------------------------------------
template <typename T, int H, int E, int NT>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, (NT <= 128 ? 4 : 2))
k_row_prox_fwd3(C2<T>* SPCSC_RESTRICT Vt, T* SPCSC_RESTRICT X, T thr_scale, WeightView<T> wl1,
                double* SPCSC_RESTRICT acc, const C2<T>* SPCSC_RESTRICT tw,
                const C2<T>* SPCSC_RESTRICT stw, int N0, int M, int Cx, T scale, int nonneg, int bnd0,
                int bnd1) {
    SPCSC_DYN_SMEM(smem_raw);
    using PL = Prox3Plan<T, H, E, 1, NT>;
    constexpr int TPF = PL::TPF, TR = PL::TR, P = PL::P, N1f = PL::N1f, TWLEN = PL::TWLEN;
    constexpr int WSTEP = NT / TR, WIT = (N1f + WSTEP - 1) / WSTEP;
    C2<T>* reg = reinterpret_cast<C2<T>*>(smem_raw);           // [TR][P]
    C2<T>* stw_s = reg + TR * P;                               // [TWLEN]
    C2<T>* tw_s = stw_s + TWLEN;                               // [N1f]
    double* red = reinterpret_cast<double*>(tw_s + N1f + (N1f & 1));   // [32]
    const int tid = threadIdx.x;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, b = blockIdx.z;
    const int k = b / Cx, c = b - k * Cx;
    const size_t wstride = (size_t)M * N0;
    const int gr = tid % TR, wf0 = tid / TR;
    C2<T>* tile = Vt + (((size_t)b * N1f) * M + m) * N0 + h0 + (size_t)wf0 * wstride + gr;
    {
        C2<T>* dst = reg + gr * P + wf0;
        SPCSC_UNROLL
        for (int it = 0; it < WIT; ++it)
            if (wf0 + it * WSTEP < N1f)
                cp_async<sizeof(C2<T>)>(dst + it * WSTEP, tile + (size_t)it * WSTEP * wstride);
    }
    for (int i = tid; i < TWLEN; i += NT) cp_async<sizeof(C2<T>)>(stw_s + i, stw + i);
    for (int i = tid; i < N1f; i += NT) cp_async<sizeof(C2<T>)>(tw_s + i, tw + i);
    cp_async_commit();
    const int g = PL::row_of_group(tid / TPF), t = tid % TPF;
    const int h = h0 + g;
    cp_async_wait<0>();
    __syncthreads();

    C2<T> v[E];
    C2<T>* row = reg + g * P;
    SPCSC_UNROLL
    for (int p = 0; p < E; ++p) {
        const int kk = t + TPF * p;
        if (kk == 0) {
            const T a0 = row[0].re, cc = row[H].re;            // c2r ignores the imaginary parts
            v[p] = mk<T>(a0 + cc, a0 - cc);
        } else {
            const C2<T> Xa = row[kk], Xb = row[H - kk];
            const C2<T> s1 = Xa + conj(Xb), d1 = Xa - conj(Xb);
            v[p] = s1 + mul_i(mulc(d1, tw_s[kk]));
        }
    }
    __syncwarp();
    fft_regs<T, H, E, true>(v, row, stw_s, t);

    const size_t wbase = (size_t)k * wl1.sk + (size_t)c * wl1.sc + (size_t)m * wl1.sm + (size_t)h * wl1.s0;
    C2<T>* xg = reinterpret_cast<C2<T>*>(X) + (((size_t)b * M + m) * N0 + h) * H;
    double sums[1] = {0.0};
    T sabs = 0;
    SPCSC_UNROLL
    for (int p = 0; p < E; ++p) {
        const int j = t + TPF * p;
        T xs[2] = {v[p].re * scale, v[p].im * scale};
        SPCSC_UNROLL
        for (int q = 0; q < 2; ++q) {
            const T w1 = wl1.spatial_uniform ? wl1.p[wbase] : wl1.p[wbase + (size_t)(2 * j + q) * wl1.s1];
            T x = soft_threshold(xs[q], thr_scale * w1);
            if (nonneg && x < (T)0) x = (T)0;
            if (h >= bnd0 || (2 * j + q) >= bnd1) x = (T)0;
            xs[q] = x;
            sabs += fabs(w1 * x);
        }
        v[p] = mk<T>(xs[0], xs[1]);
        xg[j] = v[p];
    }
    sums[0] = (double)sabs;
    fft_regs<T, H, E, false>(v, row, stw_s, t);
    __syncwarp();
    SPCSC_UNROLL
    for (int p = 0; p < E; ++p) row[t + TPF * p] = v[p];
    __syncthreads();
    {
        const C2<T>* rrow = reg + gr * P;
        SPCSC_UNROLL
        for (int it = 0; it < WIT; ++it) {
            const int wf = wf0 + it * WSTEP;
            if (wf < N1f) {
                const C2<T> aa = rrow[wf == H ? 0 : wf];
                const C2<T> bb = conj(rrow[wf == 0 ? 0 : H - wf]);
                const C2<T> sum = aa + bb, dif = mul_mi((aa - bb) * tw_s[wf]);
                tile[(size_t)it * WSTEP * wstride] =
                    mk<T>((T)0.5 * (sum.re + dif.re), (T)0.5 * (sum.im + dif.im));
            }
        }
    }
    block_accumulate<1>(sums, red, acc + ACC_L1);
}

// ------------------------------------------------------------------------------------
// k_col2: cluster of CS CTAs per (wf, b) slab; CTA `cr` owns columns
//   m = (cr*G + g)*CPG + c,  g = group (TPF lanes) index, c < CPG, kept in registers.
//   SOLVE 1: q = (Sf - s)/(g + rho)   (ADMM, Cd == 1)      SOLVE 2: q = (Sf - s)/L  (PGM gradient step;
//   also stores s and sums |Sf - s|^2)      SOLVE 4: no update: PGM evaluation of the transformed
//   slab against Sf, the sums `sumin` of the momentum point and its slabs `ref` (F, DFid, linear
//   term, |X - Y|^2)
// ------------------------------------------------------------------------------------
// BULK: persistent clusters (the grid is the number of clusters that fit on the GPU; each walks
// over slabs with that stride) whose NEXT slab -- the CTA's NG*CPG columns are contiguous in
// memory -- is fetched into shared memory by one bulk asynchronous copy (TMA) while the current
// one is transformed, so the loads of a CTA are in flight during its whole lifetime.
template <typename T, int N0, int E, int CPG, int NT, int CD, bool DO_FWD, int SOLVE, bool DO_INV,
          bool BULK>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, 2)
k_col2(const C2<T>* SPCSC_RESTRICT in, C2<T>* SPCSC_RESTRICT out, const C2<T>* SPCSC_RESTRICT Df,
       const C2<T>* SPCSC_RESTRICT Sf, const C2<T>* SPCSC_RESTRICT G,
       const AdmmState<T>* SPCSC_RESTRICT st, T Lstep, double* SPCSC_RESTRICT acc,
       const C2<T>* SPCSC_RESTRICT stw, ColArgs a, C2<T>* SPCSC_RESTRICT sumout,
       const C2<T>* SPCSC_RESTRICT sumin, const C2<T>* SPCSC_RESTRICT ref) {
    if (st && st->stopped) return;
    SPCSC_DYN_SMEM(smem_raw);
    constexpr int TPF = N0 / E, NG = NT / TPF;
    constexpr int TWLEN = stage_tw_len(N0, E);
    constexpr int XP = fft_region(N0);                          // padded exchange region per lane group
    C2<T>* xbuf = reinterpret_cast<C2<T>*>(smem_raw);         // [NG][XP] exchange / partial sums
    C2<T>* sloc = xbuf + NG * XP;                              // [CD][N0] this CTA's sums over its columns
    C2<T>* qbuf = sloc + CD * N0;                              // [CD][N0]
    C2<T>* stw_s = qbuf + CD * N0;                             // [TWLEN]
    C2<T>* pre = stw_s + TWLEN;                                // [2][N0] CD == 1: Sf and G rows of this slab,
                                                               //   fetched asynchronously at slab entry
    double* red = reinterpret_cast<double*>(pre + 2 * N0);     // [32]
    mbar_t* bar = reinterpret_cast<mbar_t*>(red + 32);         // BULK: arrival of the prefetched slab
    C2<T>* tbuf = reinterpret_cast<C2<T>*>(bar + 2);           // BULK: [NG*CPG][N0] prefetched columns
    const int tid = threadIdx.x;
    const unsigned cr = cluster_rank(), cs = cluster_size();
    const int M = a.M;
    const int g = tid / TPF, t = tid % TPF;
    for (int i = tid; i < TWLEN; i += NT) stw_s[i] = stw[i];
    const size_t dfc = (size_t)a.N1f * M * N0;                 // stride between dictionary channels
    int mcol[CPG];
    SPCSC_UNROLL
    for (int c = 0; c < CPG; ++c) mcol[c] = ((int)cr * NG + g) * CPG + c;
    // slabs of this cluster: tile, tile + tstep, ...   (BULK off: exactly one, from the grid)
    const int tstep = BULK ? (int)(gridDim.x / cs) : a.ntiles;
    int tile = BULK ? (int)(blockIdx.x / cs) : (int)(blockIdx.y * a.N1f + blockIdx.x / cs);
    const int col0 = (int)cr * NG * CPG;                       // first column of this CTA
    const int ncols = (M - col0 < NG * CPG) ? (M - col0) : NG * CPG;
    const unsigned tbytes = ncols > 0 ? (unsigned)ncols * N0 * (unsigned)sizeof(C2<T>) : 0u;
    unsigned parity = 0;
    if (BULK && tid == 0) {
        mbar_init(bar, 1);
        if (tbytes && tile < a.ntiles) {
            const int wf0 = tile % a.N1f, b0 = tile / a.N1f;
            bulk_load(tbuf, in + (((size_t)b0 * a.N1f + wf0) * M + col0) * N0, tbytes, bar);
        }
    }
    __syncthreads();                                         // stage twiddles / barrier are in place
    for (; tile < a.ntiles; tile += tstep) {
    const int wf = tile % a.N1f, b = tile / a.N1f;
    const size_t slab = (((size_t)b * a.N1f + wf) * M) * N0;
    // per-slab dictionary and Gram row (consensus dictionary update), else shared by all slabs
    const C2<T>* Dfb = a.df_bstride ? Df + (size_t)(b / a.df_bdiv) * (size_t)a.df_bstride : Df;
    const C2<T>* Gb = a.df_bstride ? G + (size_t)(b / a.df_bdiv) * (size_t)a.g_bstride : G;
    const C2<T>* dfw = Dfb + ((size_t)wf * M) * N0;
    if constexpr (SOLVE != 0 && CD == 1) {
        // the solve needs one signal and one Gram value per frequency: start fetching them now, so
        // that their L2 latency is not exposed between the two cluster barriers
        const int kk = b / a.Cx, cxx = b - kk * a.Cx;
        for (int h = tid; h < N0; h += NT) {
            cp_async<sizeof(C2<T>)>(pre + h, Sf + (((size_t)kk * a.Cs + cxx) * a.N1f + wf) * N0 + h);
            if (SOLVE == 1) cp_async<sizeof(C2<T>)>(pre + N0 + h, Gb + (size_t)wf * N0 + h);
        }
        cp_async_commit();
    }

    C2<T> v[CPG][E];
    if (BULK) {
        if (tbytes) mbar_wait(bar, parity);
        parity ^= 1u;
        SPCSC_UNROLL
        for (int c = 0; c < CPG; ++c) {
            const C2<T>* src = tbuf + (size_t)(g * CPG + c) * N0;
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) v[c][p] = (mcol[c] < M) ? src[t + TPF * p] : mk<T>(0, 0);
        }
    } else {
        SPCSC_UNROLL
        for (int c = 0; c < CPG; ++c) {
            if (mcol[c] < M) {
                const C2<T>* src = in + slab + (size_t)mcol[c] * N0;
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) v[c][p] = ld_stream(src + t + TPF * p);
            } else {
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) v[c][p] = mk<T>(0, 0);
            }
        }
    }
    if (DO_FWD) {
        SPCSC_UNROLL
        for (int c = 0; c < CPG; ++c) {
            fft_regs<T, N0, E, false>(v[c], xbuf + g * XP, stw_s, t);
            __syncwarp();
        }
    }
    if constexpr (SOLVE != 0) {
    // s_d[h] = sum over this CTA's columns of Df_d[m][h] * col[m][h], one dictionary channel at a time
    SPCSC_UNROLL
    for (int d = 0; d < CD; ++d) {
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) {
            const int h = t + TPF * p;
            C2<T> s = mk<T>(0, 0);
            SPCSC_UNROLL
            for (int c = 0; c < CPG; ++c)
                if (mcol[c] < M) s = s + ld_keep(dfw + d * dfc + (size_t)mcol[c] * N0 + h) * v[c][p];
            xbuf[g * XP + h] = s;
        }
        __syncthreads();
        // Every thread has by now USED what it read from the prefetch buffer (a barrier alone
        // does not wait for shared-memory loads still in flight, and the bulk copy engine is
        // not ordered behind them), so the buffer can be refilled with the cluster's next slab.
        if (BULK && d == 0 && tid == 0 && tbytes && tile + tstep < a.ntiles) {
            const int nt = tile + tstep, wf1 = nt % a.N1f, b1 = nt / a.N1f;
            bulk_load(tbuf, in + (((size_t)b1 * a.N1f + wf1) * M + col0) * N0, tbytes, bar);
        }
        for (int h = tid; h < N0; h += NT) {
            C2<T> s = mk<T>(0, 0);
            for (int gg = 0; gg < NG; ++gg) s = s + xbuf[gg * XP + h];
            sloc[d * N0 + h] = s;
        }
        if (d + 1 < CD) __syncthreads();
    }
    cluster_arrive();
    cluster_wait();
    const int k = b / a.Cx, cx = b - k * a.Cx;
    const T rho = (SOLVE == 1) ? st->rho : (T)0;
    double dsum[1] = {0.0};
    if constexpr (SOLVE != 0 && CD == 1) cp_async_wait<0>();   // own copies only: same h as below
    double psum[3] = {0.0, 0.0, 0.0};                          // PGM: plain |s - Sf|^2, weighted, linear term
    const double wgt_wf = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
    for (int h = tid; h < N0; h += NT) {
        C2<T> dv[CD];
        SPCSC_UNROLL
        for (int d = 0; d < CD; ++d) {
            C2<T> s = mk<T>(0, 0);
            for (unsigned rk = 0; rk < cs; ++rk) {
                const C2<T>* ps = (rk == cr) ? sloc : cluster_peer(sloc, rk);
                s = s + ps[d * N0 + h];
            }
            const int csig = (CD > 1) ? d : cx;
            C2<T> sfv;
            if constexpr (SOLVE != 0 && CD == 1)
                sfv = pre[h];
            else
                sfv = Sf[(((size_t)k * a.Cs + csig) * a.N1f + wf) * N0 + h];
            dv[d] = sfv - s;
            if ((SOLVE == 2 || SOLVE == 4) && cr == 0) {
                const size_t si = (((size_t)b * CD + d) * a.N1f + wf) * N0 + h;
                const double e2 = (double)abs2(dv[d]);
                psum[0] += e2;
                if (SOLVE == 2 && sumout) sumout[si] = s;
                if (SOLVE == 4) {
                    psum[1] += wgt_wf * e2;
                    if (sumin) {
                        const C2<T> sy = sumin[si];
                        const C2<T> dx = s - sy, gy = sy - sfv;          // Re(conj(dx) * gy)
                        psum[2] += (double)(dx.re * gy.re + dx.im * gy.im);
                    }
                }
            }
        }
        if (SOLVE == 1) {
            if (CD == 1) {
                const T den = pre[N0 + h].re + rho;
                dv[0] = mk<T>(dv[0].re / den, dv[0].im / den);
            } else {
                C2<T> A[CD][CD];
                const C2<T>* Gp = Gb + ((size_t)wf * N0 + h) * CD * CD;
                SPCSC_UNROLL
                for (int i = 0; i < CD; ++i) {
                    SPCSC_UNROLL
                    for (int j = 0; j < CD; ++j) {
                        A[i][j] = Gp[i * CD + j];
                        if (i == j) A[i][j].re += rho;
                    }
                }
                hpd_solve<T, CD>(A, dv, CD);
            }
            if (a.dfid_on && cr == 0) {
                const double wgt = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
                double q2 = 0.0;
                SPCSC_UNROLL
                for (int d = 0; d < CD; ++d) q2 += (double)abs2(dv[d]);
                dsum[0] += wgt * q2;
            }
        } else {
            SPCSC_UNROLL
            for (int d = 0; d < CD; ++d) dv[d] = mk<T>(dv[d].re / Lstep, dv[d].im / Lstep);
        }
        SPCSC_UNROLL
        for (int d = 0; d < CD; ++d) qbuf[d * N0 + h] = dv[d];
    }
    cluster_arrive_relaxed();                                // done reading the peers' sums
    __syncthreads();
    if (SOLVE == 1 && a.dfid_on) block_accumulate<1>(dsum, red, acc + ACC_DFID);
    if (SOLVE == 2 && sumout) {
        double one[1] = {psum[0]};
        block_accumulate<1>(one, red, acc + ACC_PGM_FY);
    }
    if (SOLVE == 4) {
        double a1[1] = {psum[0]}, a2[1] = {psum[1]}, a3[1] = {psum[2]};
        block_accumulate<1>(a1, red, acc + ACC_PGM_F);
        block_accumulate<1>(a2, red, acc + ACC_DFID);
        block_accumulate<1>(a3, red, acc + ACC_PGM_LIN);
    }
    }
    double dxy[1] = {0.0};
    SPCSC_UNROLL
    for (int c = 0; c < CPG; ++c) {
        if (mcol[c] < M) {
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) {
                const int h = t + TPF * p;
                C2<T> x = v[c][p];
                if constexpr (SOLVE == 1 || SOLVE == 2) {
                    SPCSC_UNROLL
                    for (int d = 0; d < CD; ++d)
                        x = x + mulc(qbuf[d * N0 + h], ld_keep(dfw + d * dfc + (size_t)mcol[c] * N0 + h));
                }
                if constexpr (SOLVE == 4) {
                    if (ref) dxy[0] += (double)abs2(x - ld_stream(ref + slab + (size_t)mcol[c] * N0 + h));
                }
                v[c][p] = x;
            }
        }
        if (DO_INV) {
            fft_regs<T, N0, E, true>(v[c], xbuf + g * XP, stw_s, t);
            __syncwarp();
        }
        if (mcol[c] < M) {
            C2<T>* dst = out + slab + (size_t)mcol[c] * N0;
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) dst[t + TPF * p] = v[c][p];
        }
    }
    if constexpr (SOLVE == 4) {
        if (ref) {
            __syncthreads();
            block_accumulate<1>(dxy, red, acc + ACC_PGM_DXY2);
            dxy[0] *= (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
            block_accumulate<1>(dxy, red, acc + ACC_PGM_RSDL);
        }
    }
    if constexpr (SOLVE != 0) cluster_wait();                // peers are done with my shared memory
    }
}

}  // namespace spcsc

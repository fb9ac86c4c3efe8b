// kernels3.cuh -- k_col3: the ADMM column kernel (column FFT + Sherman-Morrison / Woodbury solve + column
// IFFT, sporco/linalg.py:232-297 / 370-444 inside sporco/admm/cbpdn.py:271-281) as PERSISTENT clusters.
//
// Same register plan and arithmetic as k_col2<..., DO_FWD, SOLVE=1, DO_INV> (kernels2.cuh); what changes is
// how the CTAs of a cluster meet and what they keep between slabs -- both answers to the round-1 profile of
// k_col2 (profiles/r01_v7_ncu_summary.md: 25 % of the stall samples waiting on L2 reads of the dictionary,
// 24 % warp occupancy, two cluster barriers per slab):
//   * the per-frequency sums  s[h] = sum_m Df[m][h] z[m][h]  of a CTA's columns are PUSHED into the peers'
//     shared memory (st.async, completing on the receiver's mbarrier) instead of being read by the peers
//     after a cluster barrier.  barrier.cluster.{arrive.release, wait.acquire} carry cluster-scope fences --
//     ptxas emits CCTL.IVALL for them, which empties L1 -- so in k_col2 the second read of the dictionary slice
//     (the correction x = z + conj(Df) q) never hit L1.  Here nothing invalidates L1 inside the slab loop;
//   * a cluster walks over work items (wf, run of images) instead of one slab, so the dictionary slice of its
//     frequency column stays in L1 across the images of the run and the Gram row is fetched once per item;
//   * receive buffers and barriers are double-buffered by slab parity: a CTA cannot be more than one slab
//     ahead of a peer (it needs that peer's sums for every slab), so buffer n+2 never overwrites unread data.
// The sums are added in cluster-rank order on every CTA, so all CTAs of a cluster compute the same q.
#pragma once

#include "kernels2.cuh"

namespace spcsc {

// shared-memory bytes of k_col3 for a cluster of `cs` CTAs
template <typename T, int N0, int E, int NT, int CD, bool PAIR, int DFCOLS = 0>
constexpr size_t col3_smem_bytes(int cs) {
    return ((size_t)(NT / (N0 / E)) * fft_region(N0) * (PAIR ? 2 : 1)   // xbuf (16-byte elements with PAIR)
            + (size_t)DFCOLS * N0                            // staged dictionary columns of this CTA
            + (size_t)CD * N0                                // qbuf
            + (size_t)stage_tw_len(N0, E)                    // stage twiddles
            + (size_t)2 * N0                                 // Sf row, Gram row
            + (size_t)2 * cs * CD * N0)                      // receive buffers [2][cs][CD][N0]
               * sizeof(C2<T>) +
           32 * sizeof(double) + 2 * sizeof(mbar_t);
}

// DFS: the CTA's columns of the dictionary slice are copied into shared memory once per work item and read from
// there in both the sum and the correction phase (single-channel dictionaries).
template <typename T, int N0, int E, int CPG, int NT, int CD, bool PAIR, bool DFS = false>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, (CPG == 1 && sizeof(T) == 4 ? 3 : 2))
k_col3(const C2<T>* SPCSC_RESTRICT in, C2<T>* SPCSC_RESTRICT out, const C2<T>* SPCSC_RESTRICT Df,
       const C2<T>* SPCSC_RESTRICT Sf, const C2<T>* SPCSC_RESTRICT G,
       const AdmmState<T>* SPCSC_RESTRICT st, double* SPCSC_RESTRICT acc,
       const C2<T>* SPCSC_RESTRICT stw, ColArgs a, int nb, int chunk, int pf) {
    if (st->stopped) return;                                   // same value in every CTA of the cluster
    SPCSC_DYN_SMEM(smem_raw);
    constexpr int TPF = N0 / E, NG = NT / TPF;
    constexpr int TWLEN = stage_tw_len(N0, E);
    constexpr int XP = fft_region(N0) * (PAIR ? 2 : 1);        // per lane group, in C2 units
    constexpr int HPT = (N0 + NT - 1) / NT;                    // frequencies per thread in the solve
    static_assert(!PAIR || CPG == 2, "the paired transform takes the two columns of a lane group");
    const unsigned cr = cluster_rank(), cs = cluster_size();
    C2<T>* xbuf = reinterpret_cast<C2<T>*>(smem_raw);         // [NG][XP] FFT exchange / partial sums
    C2<T>* qbuf = xbuf + NG * XP;                              // [CD][N0]
    C2<T>* stw_s = qbuf + CD * N0;                             // [TWLEN]
    C2<T>* pre = stw_s + TWLEN;                                // [2][N0] Sf row of the slab, Gram row of the item
    C2<T>* recv = pre + 2 * N0;                                // [2][cs][CD][N0] sums pushed by the peers
    C2<T>* dfs = recv + (size_t)2 * cs * CD * N0;              // DFS: [NG*CPG][N0] this CTA's dictionary columns
    double* red = reinterpret_cast<double*>(dfs + (DFS ? (size_t)NG * CPG * N0 : 0));   // [32]
    static_assert(!DFS || CD == 1, "dictionary staging is for single-channel dictionaries");
    mbar_t* bar = reinterpret_cast<mbar_t*>(red + 32);         // [2]
    const int tid = threadIdx.x;
    const int M = a.M;
    const int g = tid / TPF, t = tid % TPF;
    for (int i = tid; i < TWLEN; i += NT) stw_s[i] = stw[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
    }
    __syncthreads();
    if (cs > 1) {                                              // the peers' barriers exist before anyone pushes
        cluster_arrive();
        cluster_wait();
    }
    const size_t dfc = (size_t)a.N1f * M * N0;                 // stride between dictionary channels
    int mcol[CPG];
    SPCSC_UNROLL
    for (int c = 0; c < CPG; ++c) mcol[c] = ((int)cr * NG + g) * CPG + c;
    const int nchunks = (nb + chunk - 1) / chunk;
    const int nitems = a.N1f * nchunks;
    const int ncl = (int)(gridDim.x / cs);
    const unsigned rbytes = (unsigned)((cs - 1) * CD * N0 * sizeof(C2<T>));
    unsigned slab_no = 0;
    for (int item = (int)(blockIdx.x / cs); item < nitems; item += ncl) {
        const int wf = item / nchunks;
        const int b0 = (item - wf * nchunks) * chunk;
        const int b1 = (b0 + chunk < nb) ? b0 + chunk : nb;
        const C2<T>* dfw = Df + ((size_t)wf * M) * N0;
        const double wgt_wf = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
        if constexpr (DFS) {
            __syncthreads();                                   // everyone is done with the previous item's columns
            const int col0 = (int)cr * NG * CPG;
            const int ncol = (M - col0 < NG * CPG) ? (M - col0) : NG * CPG;
            constexpr int VEC = 16 / (int)sizeof(C2<T>);
            const C2<T>* srcd = dfw + (size_t)col0 * N0;
            for (int e = tid * VEC; e < ncol * N0; e += NT * VEC) cp_async<16>(dfs + e, srcd + e);
            cp_async_commit();
        }
        for (int b = b0; b < b1; ++b, ++slab_no) {
            const unsigned par = slab_no & 1u, ph = (slab_no >> 1) & 1u;
            const size_t slab = (((size_t)b * a.N1f + wf) * M) * N0;
            const int k = b / a.Cx, cx = b - k * a.Cx;
            if constexpr (CD == 1) {
                // one signal value per frequency (and, once per item, the Gram row): fetched asynchronously
                // now, consumed after the exchange
                for (int h = tid; h < N0; h += NT) {
                    cp_async<sizeof(C2<T>)>(pre + h, Sf + (((size_t)k * a.Cs + cx) * a.N1f + wf) * N0 + h);
                    if (b == b0) cp_async<sizeof(C2<T>)>(pre + N0 + h, G + (size_t)wf * N0 + h);
                }
                cp_async_commit();
            }
            if (cs > 1 && tid == 0) mbar_expect_tx(bar + par, rbytes);

            C2<T> v[CPG][E];
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
            if (pf) {
                // the cluster's next slab: ask L2 for its lines now (lane t takes line t of each column; TPF
                // lanes x sizeof(C2) = one line), so that the loads at the top of the next pass find them there
                int bn = b + 1, wfn = wf;
                bool have = true;
                if (bn >= b1) {
                    const int itn = item + ncl;
                    have = itn < nitems;
                    wfn = itn / nchunks;
                    bn = (itn - wfn * nchunks) * chunk;
                }
                if (have && t < E) {
                    const C2<T>* nxt = in + (((size_t)bn * a.N1f + wfn) * M) * N0 + TPF * t;
                    SPCSC_UNROLL
                    for (int c = 0; c < CPG; ++c)
                        if (mcol[c] < M) prefetch_l2(nxt + (size_t)mcol[c] * N0);
                }
            }
            if constexpr (PAIR) {
                fft_regs2<T, N0, E, false>(v[0], v[1], reinterpret_cast<C4<T>*>(xbuf + g * XP), stw_s, t);
                __syncwarp();
            } else {
                SPCSC_UNROLL
                for (int c = 0; c < CPG; ++c) {
                    fft_regs<T, N0, E, false>(v[c], xbuf + g * XP, stw_s, t);
                    __syncwarp();
                }
            }
            if constexpr (DFS) {
                if (b == b0) {                                 // the item's dictionary columns have landed
                    cp_async_wait<0>();
                    __syncthreads();
                }
            }
            // s_d[h] over this CTA's columns, pushed to every peer
            C2<T> mine[CD][HPT];
            SPCSC_UNROLL
            for (int d = 0; d < CD; ++d) {
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) {
                    const int h = t + TPF * p;
                    C2<T> s = mk<T>(0, 0);
                    SPCSC_UNROLL
                    for (int c = 0; c < CPG; ++c) {
                        if (mcol[c] < M) {
                            const C2<T> dv = DFS ? dfs[(size_t)(g * CPG + c) * N0 + h]
                                                 : ld_keep(dfw + d * dfc + (size_t)mcol[c] * N0 + h);
                            s = s + dv * v[c][p];
                        }
                    }
                    xbuf[g * XP + h] = s;
                }
                __syncthreads();
                SPCSC_UNROLL
                for (int i = 0; i < HPT; ++i) {
                    const int h = tid + NT * i;
                    C2<T> s = mk<T>(0, 0);
                    if (h < N0) {
                        for (int gg = 0; gg < NG; ++gg) s = s + xbuf[gg * XP + h];
                        for (unsigned rk = 0; rk < cs; ++rk) {
                            if (rk == cr) continue;
                            C2<T>* slot = recv + (((size_t)par * cs + cr) * CD + d) * N0 + h;
                            push_remote(cluster_remote(slot, rk), s, cluster_remote(bar + par, rk));
                        }
                    }
                    mine[d][i] = s;
                }
                if (d + 1 < CD) __syncthreads();               // xbuf is reused for the next channel
            }
            if (cs > 1) mbar_wait(bar + par, ph);
            if constexpr (CD == 1) cp_async_wait<0>();        // own copies only: same h as below
            const T rho = st->rho;
            double dsum[1] = {0.0};
            SPCSC_UNROLL
            for (int i = 0; i < HPT; ++i) {
                const int h = tid + NT * i;
                if (h < N0) {
                    C2<T> dv[CD];
                    SPCSC_UNROLL
                    for (int d = 0; d < CD; ++d) {
                        C2<T> s = mk<T>(0, 0);
                        for (unsigned rk = 0; rk < cs; ++rk)
                            s = s + ((rk == cr) ? mine[d][i]
                                                : recv[(((size_t)par * cs + rk) * CD + d) * N0 + h]);
                        const int csig = (CD > 1) ? d : cx;
                        C2<T> sfv;
                        if constexpr (CD == 1)
                            sfv = pre[h];
                        else
                            sfv = Sf[(((size_t)k * a.Cs + csig) * a.N1f + wf) * N0 + h];
                        dv[d] = sfv - s;
                    }
                    if constexpr (CD == 1) {
                        const T den = pre[N0 + h].re + rho;
                        dv[0] = mk<T>(dv[0].re / den, dv[0].im / den);
                    } else {
                        C2<T> A[CD][CD];
                        const C2<T>* Gp = G + ((size_t)wf * N0 + h) * CD * CD;
                        SPCSC_UNROLL
                        for (int i2 = 0; i2 < CD; ++i2) {
                            SPCSC_UNROLL
                            for (int j2 = 0; j2 < CD; ++j2) {
                                A[i2][j2] = Gp[i2 * CD + j2];
                                if (i2 == j2) A[i2][j2].re += rho;
                            }
                        }
                        hpd_solve<T, CD>(A, dv, CD);
                    }
                    if (a.dfid_on && cr == 0) {
                        double q2 = 0.0;
                        SPCSC_UNROLL
                        for (int d = 0; d < CD; ++d) q2 += (double)abs2(dv[d]);
                        dsum[0] += wgt_wf * q2;
                    }
                    SPCSC_UNROLL
                    for (int d = 0; d < CD; ++d) qbuf[d * N0 + h] = dv[d];
                }
            }
            __syncthreads();
            if (a.dfid_on) block_accumulate<1>(dsum, red, acc + ACC_DFID);
            SPCSC_UNROLL
            for (int c = 0; c < CPG; ++c) {
                if (mcol[c] < M) {
                    SPCSC_UNROLL
                    for (int p = 0; p < E; ++p) {
                        const int h = t + TPF * p;
                        C2<T> x = v[c][p];
                        SPCSC_UNROLL
                        for (int d = 0; d < CD; ++d) {
                            const C2<T> dv = DFS ? dfs[(size_t)(g * CPG + c) * N0 + h]
                                                 : ld_keep(dfw + d * dfc + (size_t)mcol[c] * N0 + h);
                            x = x + mulc(qbuf[d * N0 + h], dv);
                        }
                        v[c][p] = x;
                    }
                }
                if constexpr (!PAIR) {
                    fft_regs<T, N0, E, true>(v[c], xbuf + g * XP, stw_s, t);
                    __syncwarp();
                    if (mcol[c] < M) {
                        C2<T>* dst = out + slab + (size_t)mcol[c] * N0;
                        SPCSC_UNROLL
                        for (int p = 0; p < E; ++p) dst[t + TPF * p] = v[c][p];
                    }
                }
            }
            if constexpr (PAIR) {
                fft_regs2<T, N0, E, true>(v[0], v[1], reinterpret_cast<C4<T>*>(xbuf + g * XP), stw_s, t);
                __syncwarp();
                SPCSC_UNROLL
                for (int c = 0; c < CPG; ++c) {
                    if (mcol[c] < M) {
                        C2<T>* dst = out + slab + (size_t)mcol[c] * N0;
                        SPCSC_UNROLL
                        for (int p = 0; p < E; ++p) dst[t + TPF * p] = v[c][p];
                    }
                }
            }
            // qbuf, xbuf and pre are next written after the next slab's first block barrier or by the thread
            // that read them; recv[par] is next written by a peer that has received this CTA's NEXT sums
        }
    }
}

// ------------------------------------------------------------------------------------
// k_col4: the ADMM column kernel with everything a slab needs ALREADY IN SHARED MEMORY when the arithmetic
// starts.  The round-2 profile of k_col3 (profiles/r02a_ncu_summary.md) still showed 29 % of the stall samples
// waiting on global / L2 loads (the slab itself at the top of a pass, the dictionary slice twice per slab) with
// 16 warps per SM and the whole register file spent on payload.  Here
//   * one CTA of NT = 512 threads per SM, ONE column per lane group (half the registers per thread, the same 16
//     warps), clusters of M / (NT / TPF) CTAs (2 at the metric configuration);
//   * a cluster owns a CONTIGUOUS range of the (frequency column, image) slabs in frequency-major order, so it
//     changes frequency column at most a few times: its columns of the dictionary slice (64 KB) and the Gram row
//     are fetched by ONE bulk copy (TMA) per frequency column and read from shared memory in both the sum and the
//     correction phase -- the 2 x 541 MB of L2 reads per launch become ~45 MB;
//   * the CTA's columns of the NEXT slab (64 KB, contiguous) and its signal row arrive by bulk copy while the
//     current slab is transformed: the stage is refilled as soon as every thread holds its column in registers
//     (after the first block barrier of the pass -- by then every thread has USED its loads of the stage, which is
//     what the refill must wait for, see DESIGN.md on the bulk-copy race);
//   * the sums travel between the CTAs of a cluster as in k_col3 (st.async pushes, double-buffered by parity).
// Single-channel dictionaries (the staged slice must fit); float32 and float64.
// ------------------------------------------------------------------------------------
template <typename T, int N0, int E, int NT>
constexpr size_t col4_smem_bytes(int cs) {
    return ((size_t)2 * (NT / (N0 / E)) * N0                 // stage, staged dictionary columns
            + (size_t)3 * N0                                 // signal rows [2], Gram row
            + (size_t)(NT / (N0 / E)) * fft_region(N0)       // xbuf
            + (size_t)N0                                     // qbuf
            + (size_t)stage_tw_len(N0, E)                    // stage twiddles
            + (size_t)2 * cs * N0)                           // receive buffers [2][cs][N0]
               * sizeof(C2<T>) +
           32 * sizeof(double) + 4 * sizeof(mbar_t);
}

template <typename T, int N0, int E, int NT, bool DBG = false>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, 1)
k_col4(const C2<T>* SPCSC_RESTRICT in, C2<T>* SPCSC_RESTRICT out, const C2<T>* SPCSC_RESTRICT Df,
       const C2<T>* SPCSC_RESTRICT Sf, const C2<T>* SPCSC_RESTRICT G,
       const AdmmState<T>* SPCSC_RESTRICT st, double* SPCSC_RESTRICT acc,
       const C2<T>* SPCSC_RESTRICT stw, ColArgs a, int nb, unsigned* dbg = nullptr) {
    if (st->stopped) return;                                   // same value in every CTA of the cluster
    SPCSC_DYN_SMEM(smem_raw);
    constexpr int TPF = N0 / E, NG = NT / TPF;
    constexpr int TWLEN = stage_tw_len(N0, E);
    constexpr int XP = fft_region(N0);
    constexpr int HPT = (N0 + NT - 1) / NT;
    const unsigned cr = cluster_rank(), cs = cluster_size();
    C2<T>* stage = reinterpret_cast<C2<T>*>(smem_raw);        // [NG][N0] this CTA's columns of the slab (bulk copy)
    C2<T>* dfs = stage + NG * N0;                              // [NG][N0] its columns of the dictionary slice
    C2<T>* pre = dfs + NG * N0;                                // [2][N0] signal rows by parity, [N0] Gram row
    C2<T>* xbuf = pre + 3 * N0;                                // [NG][XP] FFT exchange / partial sums
    C2<T>* qbuf = xbuf + NG * XP;                              // [N0]
    C2<T>* stw_s = qbuf + N0;                                  // [TWLEN]
    C2<T>* recv = stw_s + TWLEN;                               // [2][cs][N0] sums pushed by the peers
    double* red = reinterpret_cast<double*>(recv + (size_t)2 * cs * N0);   // [32]
    mbar_t* bar = reinterpret_cast<mbar_t*>(red + 32);         // [0..1] receive, [2] stage full, [3] dictionary
    const int tid = threadIdx.x;
    const int M = a.M;
    const int g = tid / TPF, t = tid % TPF;
    const int col0 = (int)cr * NG;
    const int ncol = (M - col0 < NG) ? (M - col0) : NG;
    const bool have = g < ncol;
    for (int i = tid; i < TWLEN; i += NT) stw_s[i] = stw[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_init(bar + 2, 1);
        mbar_init(bar + 3, 1);
    }
    __syncthreads();
    if (cs > 1) {                                              // the peers' barriers exist before anyone pushes
        cluster_arrive();
        cluster_wait();
    }
    // this cluster's slabs: [lo, hi) of the index L = wf * nb + b
    const int ncl = (int)(gridDim.x / cs), ci = (int)(blockIdx.x / cs);
    const long long total = (long long)a.N1f * nb;
    const int lo = (int)(total * ci / ncl), hi = (int)(total * (ci + 1) / ncl);
    const unsigned zbytes = (unsigned)((size_t)ncol * N0 * sizeof(C2<T>));
    const unsigned rowbytes = (unsigned)(N0 * sizeof(C2<T>));
    const unsigned rbytes = (unsigned)((cs - 1) * N0 * sizeof(C2<T>));
    if (tid == 0 && lo < hi) {
        const int wf = lo / nb, b = lo - wf * nb;
        const int k = b / a.Cx, cx = b - k * a.Cx;
        mbar_expect_tx(bar + 2, zbytes + rowbytes);
        bulk_copy(stage, in + (((size_t)b * a.N1f + wf) * M + col0) * N0, zbytes, bar + 2);
        bulk_copy(pre, Sf + (((size_t)k * a.Cs + cx) * a.N1f + wf) * N0, rowbytes, bar + 2);
    }
    int cur_wf = -1;
    unsigned nslab = 0, ndf = 0;
    // DBG: cycles per phase, summed over the slabs, for two probe threads (a reducer and a non-reducer)
    unsigned ph_c[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned tk = 0;
#ifndef SPCSC_EMU
#define SPCSC_MARK(i) if constexpr (DBG) { const unsigned now_ = (unsigned)clock64(); ph_c[i] += now_ - tk; tk = now_; }
#else
#define SPCSC_MARK(i)
#endif
    for (int L = lo; L < hi; ++L, ++nslab) {
        const int wf = L / nb, b = L - wf * nb;
        const unsigned par = nslab & 1u, ph = (nslab >> 1) & 1u;
        const size_t slab = (((size_t)b * a.N1f + wf) * M) * N0;
        const double wgt_wf = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
        bool df_wait = false;
#ifndef SPCSC_EMU
        if constexpr (DBG) tk = (unsigned)clock64();
#endif
        if (wf != cur_wf) {                                    // the same decision in every CTA of the cluster
            cur_wf = wf;
            __syncthreads();                                   // everyone has used the previous column's slice
            if (tid == 0) {
                mbar_expect_tx(bar + 3, zbytes + rowbytes);
                bulk_copy(dfs, Df + ((size_t)wf * M + col0) * N0, zbytes, bar + 3);
                bulk_copy(pre + 2 * N0, G + (size_t)wf * N0, rowbytes, bar + 3);
            }
            df_wait = true;
        }
        if (cs > 1 && tid == 0) mbar_expect_tx(bar + par, rbytes);
        SPCSC_MARK(0)
        mbar_wait(bar + 2, par);                               // the slab's columns and signal row have landed
        SPCSC_MARK(1)
        C2<T> v[E];
        if (have) {
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) v[p] = stage[g * N0 + t + TPF * p];
        } else {
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) v[p] = mk<T>(0, 0);
        }
        fft_regs<T, N0, E, false>(v, xbuf + g * XP, stw_s, t);
        __syncwarp();
        SPCSC_MARK(2)
        if (df_wait) {
            mbar_wait(bar + 3, ndf & 1u);
            ++ndf;
        }
        // this column's products with its dictionary column; summed over the CTA's columns after the barrier
        SPCSC_UNROLL
        for (int p = 0; p < E; ++p) {
            const int h = t + TPF * p;
            xbuf[g * XP + h] = have ? dfs[g * N0 + h] * v[p] : mk<T>(0, 0);
        }
        SPCSC_MARK(3)
        __syncthreads();
        SPCSC_MARK(4)
        if (tid == 0 && L + 1 < hi) {                          // the stage is free: fetch the next slab now
            const int wfn = (L + 1) / nb, bn = (L + 1) - wfn * nb;
            const int kn = bn / a.Cx, cxn = bn - kn * a.Cx;
            mbar_expect_tx(bar + 2, zbytes + rowbytes);
            bulk_copy(stage, in + (((size_t)bn * a.N1f + wfn) * M + col0) * N0, zbytes, bar + 2);
            bulk_copy(pre + (par ^ 1u) * N0, Sf + (((size_t)kn * a.Cs + cxn) * a.N1f + wfn) * N0, rowbytes,
                      bar + 2);
        }
        C2<T> mine[HPT];
        SPCSC_UNROLL
        for (int i = 0; i < HPT; ++i) {
            const int h = tid + NT * i;
            C2<T> s0 = mk<T>(0, 0), s1 = mk<T>(0, 0);
            if (h < N0) {
                SPCSC_UNROLL
                for (int gg = 0; gg < NG; gg += 2) {
                    s0 = s0 + xbuf[gg * XP + h];
                    if (gg + 1 < NG) s1 = s1 + xbuf[(gg + 1) * XP + h];
                }
                s0 = s0 + s1;
                for (unsigned rk = 0; rk < cs; ++rk) {
                    if (rk == cr) continue;
                    C2<T>* slot = recv + ((size_t)par * cs + cr) * N0 + h;
                    push_remote(cluster_remote(slot, rk), s0, cluster_remote(bar + par, rk));
                }
            }
            mine[i] = s0;
        }
        SPCSC_MARK(5)
        if (cs > 1) mbar_wait(bar + par, ph);
        SPCSC_MARK(6)
        const T rho = st->rho;
        double dsum[1] = {0.0};
        SPCSC_UNROLL
        for (int i = 0; i < HPT; ++i) {
            const int h = tid + NT * i;
            if (h < N0) {
                C2<T> s = mk<T>(0, 0);
                for (unsigned rk = 0; rk < cs; ++rk)
                    s = s + ((rk == cr) ? mine[i] : recv[((size_t)par * cs + rk) * N0 + h]);
                const C2<T> d = pre[par * N0 + h] - s;
                const T den = pre[2 * N0 + h].re + rho;
                const C2<T> q = mk<T>(d.re / den, d.im / den);
                if (a.dfid_on && cr == 0) dsum[0] += wgt_wf * (double)abs2(q);
                qbuf[h] = q;
            }
        }
        SPCSC_MARK(7)
        __syncthreads();
        SPCSC_MARK(8)
        if (a.dfid_on) block_accumulate<1>(dsum, red, acc + ACC_DFID);
        if (have) {
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) {
                const int h = t + TPF * p;
                v[p] = v[p] + mulc(qbuf[h], dfs[g * N0 + h]);
            }
        }
        SPCSC_MARK(9)
        fft_regs<T, N0, E, true>(v, xbuf + g * XP, stw_s, t);
        __syncwarp();
        SPCSC_MARK(10)
        if (have) {
            C2<T>* dst = out + slab + (size_t)(col0 + g) * N0;
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) dst[t + TPF * p] = v[p];
        }
        SPCSC_MARK(11)
        // xbuf rows, qbuf and recv[par] are next written after barriers every thread passes only once it is
        // done with them here (see k_col3); the stage and pre[par ^ 1] are being refilled meanwhile
    }
    if constexpr (DBG) {
        if (dbg && (tid == 0 || tid == NT / 2)) {
            unsigned* o = dbg + ((size_t)blockIdx.x * 2 + (tid ? 1 : 0)) * 13;
            for (int i = 0; i < 12; ++i) o[i] = ph_c[i];
            o[12] = nslab;
        }
    }
#undef SPCSC_MARK
}

// ------------------------------------------------------------------------------------
// k_col5: k_col4 with the CTA split into NGRP INDEPENDENT thread groups that work on different slabs of the
// same frequency column.  The profile of k_col4 (profiles/r02b_ncu_summary.md) showed a kernel whose phases
// run in lock step over the whole SM: the transforms saturate the FMA pipe while shared memory idles, the
// pointwise phases saturate shared memory (stall reason mio 17 %) while the FMA pipe idles, and 8 % of the
// samples sit at the two block barriers -- with one CTA per SM nothing fills those gaps.  Here
//   * a group of NT / NGRP threads owns its columns of ONE slab from the bulk copy to the store, with its own
//     stage, exchange regions, receive buffers, mbarriers and a NAMED barrier (bar.sync id) -- the groups drift
//     apart, so one group's transforms overlap the other's pointwise and exchange phases, as two CTAs per SM
//     would, but
//   * the groups SHARE the staged dictionary columns (which two CTAs could not: 2 x 64 KB), so each slab is
//     spread over twice as many CTAs (clusters of 4 at the metric configuration);
//   * the second stage's twiddle factors of the 16 x 16 transform plan live in registers (they depend only on
//     the lane): 16 % fewer shared-memory wavefronts than k_col4.
// Work split: the cluster walks over the segments of constant frequency column of its slab range; within a
// segment group j takes slabs j, j + NGRP, ...  The data term (DFid) is summed per thread over all slabs and
// reduced once at the end.
// ------------------------------------------------------------------------------------
template <typename T, int N0, int E, int NT, int NGRP>
constexpr size_t col5_smem_bytes(int cs) {
    constexpr int NGG = (NT / NGRP) / (N0 / E);
    return ((size_t)NGRP * NGG * N0 + (size_t)NGG * N0       // stages, staged dictionary columns
            + (size_t)NGRP * 2 * N0 + N0                     // signal rows [NGRP][2], Gram row
            + (size_t)NGRP * NGG * fft_region(N0)            // xbuf
            + (size_t)NGRP * N0                              // qbuf
            + (size_t)stage_tw_len(N0, E)                    // stage twiddles
            + (size_t)NGRP * 2 * cs * N0)                    // receive buffers [NGRP][2][cs][N0]
               * sizeof(C2<T>) +
           32 * sizeof(double) + (3 * NGRP + 1) * sizeof(mbar_t);
}

// TST: the result leaves through the exchange regions and bulk copies shared -> global instead of 16 stores per
// thread (measured 3 % slower at the metric configuration -- the extra pass through shared memory costs more than the
// store queue it relieves -- so it stays an experiment).
template <typename T, int N0, int E, int NT, int NGRP, bool TST = false>
SPCSC_GLOBAL void SPCSC_LAUNCH_BOUNDS2(NT, 1)
k_col5(const C2<T>* SPCSC_RESTRICT in, C2<T>* SPCSC_RESTRICT out, const C2<T>* SPCSC_RESTRICT Df,
       const C2<T>* SPCSC_RESTRICT Sf, const C2<T>* SPCSC_RESTRICT G,
       const AdmmState<T>* SPCSC_RESTRICT st, double* SPCSC_RESTRICT acc,
       const C2<T>* SPCSC_RESTRICT stw, ColArgs a, int nb, int stagger) {
    if (st->stopped) return;                                   // same value in every CTA of the cluster
    SPCSC_DYN_SMEM(smem_raw);
    constexpr int TPF = N0 / E, NTG = NT / NGRP, NGG = NTG / TPF;
    constexpr int TWLEN = stage_tw_len(N0, E);
    constexpr int XP = fft_region(N0);
    constexpr int HPT = (N0 + NTG - 1) / NTG;
    constexpr bool TWR = fft_two_stage<N0, E>();
    const unsigned cr = cluster_rank(), cs = cluster_size();
    const int tid = threadIdx.x;
    const int gi = tid / NTG, tg = tid % NTG;                  // thread group, thread within it
    const int g = tg / TPF, t = tg % TPF;                      // column within the group, lane of the column
    C2<T>* stage = reinterpret_cast<C2<T>*>(smem_raw) + (size_t)gi * NGG * N0;     // [NGG][N0] (bulk copy)
    C2<T>* dfs = reinterpret_cast<C2<T>*>(smem_raw) + (size_t)NGRP * NGG * N0;     // [NGG][N0] shared
    C2<T>* pre0 = dfs + NGG * N0;                              // [NGRP][2][N0], then the Gram row
    C2<T>* pre = pre0 + (size_t)gi * 2 * N0;
    C2<T>* gram = pre0 + (size_t)NGRP * 2 * N0;
    C2<T>* xbuf = gram + N0 + (size_t)gi * NGG * XP;           // [NGG][XP] of this group
    C2<T>* qbuf = gram + N0 + (size_t)NGRP * NGG * XP + (size_t)gi * N0;
    C2<T>* stw_s = gram + N0 + (size_t)NGRP * NGG * XP + (size_t)NGRP * N0;
    C2<T>* recv = stw_s + TWLEN + (size_t)gi * 2 * cs * N0;    // [2][cs][N0] of this group
    double* red = reinterpret_cast<double*>(stw_s + TWLEN + (size_t)NGRP * 2 * cs * N0);   // [32]
    mbar_t* bar0 = reinterpret_cast<mbar_t*>(red + 32);
    mbar_t* bar = bar0 + 3 * gi;                               // [0..1] receive by parity, [2] stage full
    mbar_t* bar_df = bar0 + 3 * NGRP;                          // dictionary columns + Gram row
    const int M = a.M;
    const int col0 = (int)cr * NGG;
    const int ncol = (M - col0 < NGG) ? (M - col0 > 0 ? M - col0 : 0) : NGG;
    const bool have = g < ncol;
    for (int i = tid; i < TWLEN; i += NT) stw_s[i] = stw[i];
    if (tid == 0)
        for (int i = 0; i < 3 * NGRP + 1; ++i) mbar_init(bar0 + i, 1);
    __syncthreads();
    C2<T> twr[E];
    if constexpr (TWR) load_stage_tw_regs<T, N0, E>(twr, stw_s, t);
    if (cs > 1) {                                              // the peers' barriers exist before anyone pushes
        cluster_arrive();
        cluster_wait();
    }
    if (stagger > 0 && gi > 0) nap((unsigned)(stagger * gi));  // start the groups out of phase
    const int ncl = (int)(gridDim.x / cs), ci = (int)(blockIdx.x / cs);
    const long long total = (long long)a.N1f * nb;
    const int lo = (int)(total * ci / ncl), hi = (int)(total * (ci + 1) / ncl);
    const unsigned zbytes = (unsigned)((size_t)ncol * N0 * sizeof(C2<T>));
    const unsigned rowbytes = (unsigned)(N0 * sizeof(C2<T>));
    const unsigned rbytes = (unsigned)((cs - 1) * N0 * sizeof(C2<T>));
    double dsum[1] = {0.0};
    unsigned nsl = 0, ndf = 0;                                 // slabs of this group, segments so far
    int L = lo;
    while (L < hi) {                                           // one segment = one frequency column
        const int wf = L / nb;
        const int segend = ((long long)(wf + 1) * nb < hi) ? (wf + 1) * nb : hi;
        const double wgt_wf = (wf == 0 || (a.even_n1 && wf == a.N1f - 1)) ? 1.0 : 2.0;
        __syncthreads();                                       // every group has used the previous column's slice
        if (tid == 0) {
            mbar_expect_tx(bar_df, zbytes + rowbytes);
            if (zbytes) bulk_copy(dfs, Df + ((size_t)wf * M + col0) * N0, zbytes, bar_df);
            bulk_copy(gram, G + (size_t)wf * N0, rowbytes, bar_df);
        }
        int Lg = L + gi;
        if (tg == 0 && Lg < segend) {                          // this group's first slab of the segment
            const int b = Lg - wf * nb, k = b / a.Cx, cx = b - k * a.Cx;
            mbar_expect_tx(bar + 2, zbytes + rowbytes);
            if (zbytes) bulk_copy(stage, in + (((size_t)b * a.N1f + wf) * M + col0) * N0, zbytes, bar + 2);
            bulk_copy(pre + (nsl & 1u) * N0, Sf + (((size_t)k * a.Cs + cx) * a.N1f + wf) * N0, rowbytes, bar + 2);
        }
        bool df_wait = true;
        for (; Lg < segend; Lg += NGRP, ++nsl) {
            const int b = Lg - wf * nb;
            const unsigned par = nsl & 1u, ph = (nsl >> 1) & 1u;
            const size_t slab = (((size_t)b * a.N1f + wf) * M) * N0;
            if (cs > 1 && tg == 0) mbar_expect_tx(bar + par, rbytes);
            mbar_wait(bar + 2, par);                           // the slab's columns and signal row have landed
            C2<T> v[E];
            if (have) {
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) v[p] = stage[g * N0 + t + TPF * p];
            } else {
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) v[p] = mk<T>(0, 0);
            }
            if constexpr (TST) {                               // the previous slab's result has left this region
                if (t == 0) bulk_store_wait_read();
                __syncwarp();
            }
            if constexpr (TWR)
                fft_regs_twr<T, N0, E, false>(v, xbuf + g * XP, twr, t);
            else
                fft_regs<T, N0, E, false>(v, xbuf + g * XP, stw_s, t);
            __syncwarp();
            if (df_wait) {
                mbar_wait(bar_df, ndf & 1u);
                df_wait = false;
            }
            SPCSC_UNROLL
            for (int p = 0; p < E; ++p) {
                const int h = t + TPF * p;
                xbuf[g * XP + h] = have ? dfs[g * N0 + h] * v[p] : mk<T>(0, 0);
            }
            group_barrier(1 + gi, NTG);
            if (tg == 0 && Lg + NGRP < segend) {               // the stage is free: fetch this group's next slab
                const int bn = b + NGRP, kn = bn / a.Cx, cxn = bn - kn * a.Cx;
                mbar_expect_tx(bar + 2, zbytes + rowbytes);
                if (zbytes) bulk_copy(stage, in + (((size_t)bn * a.N1f + wf) * M + col0) * N0, zbytes, bar + 2);
                bulk_copy(pre + (par ^ 1u) * N0, Sf + (((size_t)kn * a.Cs + cxn) * a.N1f + wf) * N0, rowbytes,
                          bar + 2);
            }
            const T rho = st->rho;
            C2<T> mine[HPT];
            SPCSC_UNROLL
            for (int i = 0; i < HPT; ++i) {
                const int h = tg + NTG * i;
                C2<T> s0 = mk<T>(0, 0), s1 = mk<T>(0, 0);
                if (h < N0) {
                    SPCSC_UNROLL
                    for (int gg = 0; gg < NGG; gg += 2) {
                        s0 = s0 + xbuf[gg * XP + h];
                        if (gg + 1 < NGG) s1 = s1 + xbuf[(gg + 1) * XP + h];
                    }
                    s0 = s0 + s1;
                    for (unsigned rk = 0; rk < cs; ++rk) {
                        if (rk == cr) continue;
                        C2<T>* slot = recv + ((size_t)par * cs + cr) * N0 + h;
                        push_remote(cluster_remote(slot, rk), s0, cluster_remote(bar + par, rk));
                    }
                }
                mine[i] = s0;
            }
            if (cs > 1) mbar_wait(bar + par, ph);
            SPCSC_UNROLL
            for (int i = 0; i < HPT; ++i) {
                const int h = tg + NTG * i;
                if (h < N0) {
                    C2<T> s = mk<T>(0, 0);
                    for (unsigned rk = 0; rk < cs; ++rk)
                        s = s + ((rk == cr) ? mine[i] : recv[((size_t)par * cs + rk) * N0 + h]);
                    const C2<T> d = pre[par * N0 + h] - s;
                    const T inv = (T)1 / (gram[h].re + rho);
                    const C2<T> q = mk<T>(d.re * inv, d.im * inv);
                    if (a.dfid_on && cr == 0) dsum[0] += wgt_wf * (double)abs2(q);
                    qbuf[h] = q;
                }
            }
            group_barrier(1 + gi, NTG);
            if (have) {
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) {
                    const int h = t + TPF * p;
                    v[p] = v[p] + mulc(qbuf[h], dfs[g * N0 + h]);
                }
            }
            if constexpr (TWR)
                fft_regs_twr<T, N0, E, true>(v, xbuf + g * XP, twr, t);
            else
                fft_regs<T, N0, E, true>(v, xbuf + g * XP, stw_s, t);
            __syncwarp();
            if constexpr (TST) {
                if (have) {
                    SPCSC_UNROLL
                    for (int p = 0; p < E; ++p) xbuf[g * XP + t + TPF * p] = v[p];
                }
                fence_async_smem();
                __syncwarp();
                if (have && t == 0) {
                    bulk_store(out + slab + (size_t)(col0 + g) * N0, xbuf + g * XP, (unsigned)(N0 * sizeof(C2<T>)));
                    bulk_store_commit();
                }
            } else if (have) {
                C2<T>* dst = out + slab + (size_t)(col0 + g) * N0;
                SPCSC_UNROLL
                for (int p = 0; p < E; ++p) dst[t + TPF * p] = v[p];
            }
        }
        if (df_wait) mbar_wait(bar_df, ndf & 1u);              // a group without a slab in this segment
        ++ndf;
        L = segend;
    }
    if constexpr (TST) bulk_store_wait_all();
    if (a.dfid_on) block_accumulate<1>(dsum, red, acc + ACC_DFID);
}

}  // namespace spcsc

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

}  // namespace spcsc

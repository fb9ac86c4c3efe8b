// launchers_impl.cuh -- definitions of the size-templated launchers (see launchers.cuh).
#pragma once

#include "launchers.cuh"

#include <cstdlib>
#include <string>
#include <vector>
#include <cstdio>

namespace spcsc {

inline int round_up32(int n) { return (n + 31) / 32 * 32; }
extern int g_col_variant;      // spcsc.cu: which cluster column kernel the last COL_ADMM launch used

template <typename T, int H>
cudaError_t row_fwd_launch(const RowArgs<T>& r, const T* A, const T* B, const AdmmState<T>* st,
                           C2<T>* Zt) {
    constexpr int TPF = fft_tpf<T, H>();
    const int nt = round_up32(r.TR * TPF);
    size_t smem = (size_t)r.TR * (H + 1) * sizeof(C2<T>);
    dim3 grid(r.N0 / r.TR, r.M, r.nb);
    return launch(k_row_fwd<T, H>, grid, dim3(nt), smem, r.stream, A, B, st, Zt, r.tw, r.N0, r.M,
                  r.TR);
}

template <typename T, int H>
cudaError_t row_inv_launch(const RowArgs<T>& r, const C2<T>* Zt, T* X, T scale) {
    constexpr int TPF = fft_tpf<T, H>();
    const int nt = round_up32(r.TR * TPF);
    size_t smem = (size_t)r.TR * (H + 1) * sizeof(C2<T>);
    dim3 grid(r.N0 / r.TR, r.M, r.nb);
    return launch(k_row_inv<T, H>, grid, dim3(nt), smem, r.stream, Zt, X, r.tw, r.N0, r.M, r.TR,
                  scale);
}

template <typename T, int H, int CX>
static cudaError_t row_inv_prox_cx(const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt, T* Y,
                                   T* U, const AdmmState<T>* st) {
    constexpr int TPF = fft_tpf<T, H>();
    const int nt = round_up32(r.TR * TPF);
    size_t smem = (size_t)CX * r.TR * (H + 1) * sizeof(C2<T>);
    if (smem < 7 * 32 * sizeof(double)) smem = 7 * 32 * sizeof(double);
    dim3 grid(r.N0 / r.TR, r.M, r.nb / CX);
    return launch(k_row_inv_prox<T, H, CX>, grid, dim3(nt), smem, r.stream, Zt, Y, U, st, p.prm,
                  p.wl1, p.wl21, p.acc, r.tw, r.N0, r.M, r.TR, p.scale, p.nonneg, p.bnd0, p.bnd1,
                  p.reg_on_y);
}

template <typename T, int H>
cudaError_t row_inv_prox_launch(const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt, T* Y,
                                T* U, const AdmmState<T>* st) {
    switch (r.Cx) {
        case 1: return row_inv_prox_cx<T, H, 1>(r, p, Zt, Y, U, st);
        case 2: return row_inv_prox_cx<T, H, 2>(r, p, Zt, Y, U, st);
        case 3: return row_inv_prox_cx<T, H, 3>(r, p, Zt, Y, U, st);
        case 4: return row_inv_prox_cx<T, H, 4>(r, p, Zt, Y, U, st);
        default: return cudaErrorInvalidValue;
    }
}

template <typename T, int H>
cudaError_t row_inv_prox_fwd_launch(const RowArgs<T>& r, const PgmRowArgs<T>& p, C2<T>* Vt, T* X) {
    if constexpr (row2_elems(H, 1, (int)sizeof(T)) != 0) {
        // register-plan kernel: float32, power-of-two row length, 128-thread CTAs
        constexpr int E = row2_elems(H, 1, (int)sizeof(T)), TPF2 = H / E, NT2 = 128, TR2 = NT2 / TPF2;
        if (p.stw && TPF2 <= 16 && r.N0 % TR2 == 0) {
            using PL = Prox3Plan<T, H, E, 1, NT2>;
            const size_t smem2 = ((size_t)TR2 * PL::P + PL::TWLEN + PL::N1f + (PL::N1f & 1)) * sizeof(C2<T>) +
                                 32 * sizeof(double);
            dim3 grid2(r.N0 / TR2, r.M, r.nb);
            return launch(k_row_prox_fwd3<T, H, E, NT2>, grid2, dim3(NT2), smem2, r.stream, Vt, X, p.thr_scale,
                          p.wl1, p.acc, r.tw, p.stw, r.N0, r.M, r.Cx, p.scale, p.nonneg, p.bnd0, p.bnd1);
        }
    }
    constexpr int TPF = fft_tpf<T, H>();
    const int nt = round_up32(r.TR * TPF);
    size_t smem = (size_t)r.TR * (H + 1) * sizeof(C2<T>);
    if (smem < 32 * sizeof(double)) smem = 32 * sizeof(double);
    dim3 grid(r.N0 / r.TR, r.M, r.nb);
    return launch(k_row_inv_prox_fwd<T, H>, grid, dim3(nt), smem, r.stream, Vt, X, p.thr_scale, p.wl1,
                  p.acc, r.tw, r.N0, r.M, r.Cx, r.TR, p.scale, p.nonneg, p.bnd0, p.bnd1);
}

// Choose threads / chunking for the column kernel.
template <typename T, int N0T>
inline void col_plan(ColArgs& a, int& nthreads, size_t& smem) {
    constexpr bool GEN = (N0T == 0);
    const int N0 = GEN ? a.N0 : N0T;
    int TPF = 1;
    if constexpr (!GEN) TPF = fft_tpf<T, N0T>();
    const int maxt = sizeof(T) == 4 ? 1024 : 512;
    const size_t sz = sizeof(C2<T>);
    int want = round_up32(a.M * TPF);
    nthreads = want < maxt ? want : maxt;
    if (nthreads < 32) nthreads = 32;
    int parts = nthreads / N0;
    if (parts < 1) parts = 1;
    if (parts > 8) parts = 8;
    if (parts > a.M) parts = a.M;
    a.parts = parts;
    const size_t fixed = (size_t)(a.Cd + (a.gradreg ? 1 : 0)) * parts * N0 * sz + 64 * sizeof(double);
    size_t room = kSmemLimit - fixed;
    int mc = (int)(room / ((size_t)N0 * sz * (GEN ? 2 : 1)));
    if (mc > a.M) mc = a.M;
    if (mc < 1) mc = 1;
    a.MC = mc;
    a.nchunk = (a.M + mc - 1) / mc;
    smem = (size_t)mc * N0 * sz * (GEN ? 2 : 1) + fixed;
}

template <typename T, int N0, bool F, int S, bool I>
static cudaError_t col_go(ColLaunch<T>& c) {
    int nt;
    size_t smem;
    col_plan<T, N0>(c.a, nt, smem);
    dim3 grid(c.a.N1f, c.nb);
    return launch(k_col<T, N0, F, S, I>, grid, dim3(nt), smem, c.stream, c.in, c.out, c.Df, c.Sf,
                  c.G, c.sumout, c.sumin, c.ref, c.st, c.Lstep, c.acc, c.tw, c.a);
}

template <typename T, int N0>
cudaError_t col_launch(int mode, ColLaunch<T> c) {
    if (N0 != 0) c.a.N0 = N0;
    switch (mode) {
        case COL_FWD: return col_go<T, N0, true, 0, false>(c);
        case COL_INV: return col_go<T, N0, false, 0, true>(c);
        case COL_ADMM: return col_go<T, N0, true, 1, true>(c);
        case COL_ADMM_NOFFT: return col_go<T, N0, false, 1, false>(c);
        case COL_GRAD_INV: return col_go<T, N0, false, 2, true>(c);
        case COL_FWD_SUM: return col_go<T, N0, true, 3, false>(c);
        case COL_SUM: return col_go<T, N0, false, 3, false>(c);
        case COL_FWD_EVAL: return col_go<T, N0, true, 4, false>(c);
        default: return cudaErrorInvalidValue;
    }
}


// ---- any-size path --------------------------------------------------------------------
template <typename T>
inline int gen_threads(int work) {
    int nt = round_up32(work);
    return nt > 256 ? 256 : (nt < 32 ? 32 : nt);
}
// Mixed-radix path of the any-size row kernels: taken when the row length factors into primes <= kGenMaxRadix
// and the two complex work buffers (TR rows each) fit into shared memory next to what the kernel needs anyway.
template <typename T>
inline int gen_rows_fast(const GenRowArgs<T>& r, size_t base_bytes, size_t& smem) {
    int rad[16];
    smem = base_bytes;
    if (gen_factor(r.N1, rad) == 0) return 0;
    const size_t extra = gen_align16(base_bytes) + (size_t)2 * r.TR * r.N1 * sizeof(C2<T>);
    if (extra > kSmemLimit) return 0;
    smem = extra;
    return 1;
}
template <typename T>
cudaError_t row_fwd_gen_launch(const GenRowArgs<T>& r, const T* A, const T* B, const AdmmState<T>* st,
                               C2<T>* Zt) {
    size_t smem;
    const int fast = gen_rows_fast<T>(r, (size_t)r.TR * r.N1 * sizeof(T), smem);
    dim3 grid((r.N0 + r.TR - 1) / r.TR, r.M, r.nb);
    return launch(k_row_fwd_gen<T>, grid, dim3(gen_threads<T>(r.TR * r.N1)), smem, r.stream, A, B, st,
                  Zt, r.tw, r.N0, r.N1, r.M, r.TR, fast);
}
template <typename T>
cudaError_t row_inv_gen_launch(const GenRowArgs<T>& r, const C2<T>* Zt, T* X, T scale) {
    const int N1f = r.N1 / 2 + 1;
    size_t smem;
    const int fast = gen_rows_fast<T>(r, (size_t)r.TR * N1f * sizeof(C2<T>) + (size_t)r.TR * r.N1 * sizeof(T), smem);
    dim3 grid((r.N0 + r.TR - 1) / r.TR, r.M, r.nb);
    return launch(k_row_inv_gen<T>, grid, dim3(gen_threads<T>(r.TR * r.N1)), smem, r.stream, Zt, X,
                  r.tw, r.N0, r.N1, r.M, r.TR, scale, fast);
}
template <typename T, int CX>
static cudaError_t row_inv_prox_gen_cx(const GenRowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt,
                                       T* Y, T* U, const AdmmState<T>* st) {
    const int N1f = r.N1 / 2 + 1;
    size_t smem;
    const int fast = gen_rows_fast<T>(r, (size_t)r.TR * N1f * sizeof(C2<T>) + (size_t)CX * r.TR * r.N1 * sizeof(T), smem);
    if (smem < 7 * 32 * sizeof(double)) smem = 7 * 32 * sizeof(double);
    dim3 grid((r.N0 + r.TR - 1) / r.TR, r.M, r.nb / CX);
    return launch(k_row_inv_prox_gen<T, CX>, grid, dim3(gen_threads<T>(r.TR * r.N1)), smem, r.stream,
                  Zt, Y, U, st, p.prm, p.wl1, p.wl21, p.acc, r.tw, r.N0, r.N1, r.M, r.TR, p.scale,
                  p.nonneg, p.bnd0, p.bnd1, p.reg_on_y, fast);
}
template <typename T>
cudaError_t row_inv_prox_gen_launch(const GenRowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt,
                                    T* Y, T* U, const AdmmState<T>* st) {
    switch (r.Cx) {
        case 1: return row_inv_prox_gen_cx<T, 1>(r, p, Zt, Y, U, st);
        case 2: return row_inv_prox_gen_cx<T, 2>(r, p, Zt, Y, U, st);
        case 3: return row_inv_prox_gen_cx<T, 3>(r, p, Zt, Y, U, st);
        case 4: return row_inv_prox_gen_cx<T, 4>(r, p, Zt, Y, U, st);
        default: return cudaErrorInvalidValue;
    }
}
template <typename T>
cudaError_t row_inv_prox_fwd_gen_launch(const GenRowArgs<T>& r, const PgmRowArgs<T>& p, C2<T>* Vt,
                                        T* X) {
    const int N1f = r.N1 / 2 + 1;
    size_t smem;
    const int fast = gen_rows_fast<T>(r, (size_t)r.TR * N1f * sizeof(C2<T>) + (size_t)r.TR * r.N1 * sizeof(T), smem);
    if (smem < 32 * sizeof(double)) smem = 32 * sizeof(double);
    dim3 grid((r.N0 + r.TR - 1) / r.TR, r.M, r.nb);
    return launch(k_row_inv_prox_fwd_gen<T>, grid, dim3(gen_threads<T>(r.TR * r.N1)), smem, r.stream,
                  Vt, X, p.thr_scale, p.wl1, p.acc, r.tw, r.N0, r.N1, r.M, r.Cx, r.TR, p.scale,
                  p.nonneg, p.bnd0, p.bnd1, fast);
}

// ---- kernel set v2 -----------------------------------------------------------------
template <typename T, int H>
cudaError_t row_fwd2_launch(const RowArgs<T>& r, const T* A, const T* B, const AdmmState<T>* st,
                            C2<T>* Zt, const C2<T>* stw, int gated) {
    if constexpr (row2_elems(H, 1, (int)sizeof(T)) != 0) {
        constexpr int E = row2_elems(H, 1, (int)sizeof(T)), NT = kRow2Threads, TR = row2_tile(H, 1, (int)sizeof(T));
        if constexpr (H / E <= 16) {
            // asynchronous-copy tile flow (as the prox kernel): 128-thread CTAs, next tile prefetched
            constexpr int NT3 = 128, TR3 = NT3 / (H / E);
            const bool use3 = !(getenv("SPCSC_ROWFWD") && std::string(getenv("SPCSC_ROWFWD")) == "2");
            if (use3 && r.N0 % TR3 == 0) {
                using PL = Prox3Plan<T, H, E, 1, NT3>;
                const size_t smem3 = ((size_t)2 * PL::YS + (size_t)TR3 * PL::P + PL::TWLEN + PL::N1f) * sizeof(C2<T>);
                const long long ntiles3 = (long long)(r.N0 / TR3) * r.M * r.nb;
                const long long cap3 = 148LL * 4 * 4;
                dim3 grid3((unsigned)(ntiles3 < cap3 ? ntiles3 : cap3));
                return launch(k_row_fwd3<T, H, E, NT3>, grid3, dim3(NT3), smem3, r.stream, A, B, st, Zt, r.tw,
                              stw, r.N0, r.M, r.nb, gated);
            }
        }
        const size_t smem = ((size_t)TR * (H + H / 16 + 1) + stage_tw_len(H, E)) * sizeof(C2<T>);
        const long long ntiles = (long long)(r.N0 / TR) * r.M * r.nb;
        // grid-stride over tiles: a gated launch that finds nothing to do retires in microseconds
        const long long cap = 148LL * 24;
        dim3 grid((unsigned)(ntiles < cap ? ntiles : cap));
        return launch(k_row_fwd2<T, H, E, NT>, grid, dim3(NT), smem, r.stream, A, B, st, Zt, r.tw,
                      stw, r.N0, r.M, r.nb, gated);
    } else {
        return cudaErrorInvalidValue;
    }
}

template <typename T, int H, int CX>
static cudaError_t row_inv_prox2_cx(const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt,
                                    T* Y, T* U, const AdmmState<T>* st, const C2<T>* stw) {
    if constexpr (row2_elems(H, CX, (int)sizeof(T)) != 0) {
        constexpr int E = row2_elems(H, CX, (int)sizeof(T)), NT = kRow2Threads, TR = row2_tile(H, CX, (int)sizeof(T));
        const size_t smem = ((size_t)CX * TR * (H + H / 16 + 1) + stage_tw_len(H, E)) * sizeof(C2<T>);
        dim3 grid(r.N0 / TR, r.M, r.nb / CX);
        return launch(k_row_inv_prox2<T, H, E, CX, NT>, grid, dim3(NT), smem, r.stream, Zt, Y, U, st,
                      p.prm, p.wl1, p.wl21, p.acc, r.tw, stw, r.N0, r.M, p.scale, p.nonneg, p.bnd0,
                      p.bnd1, p.reg_on_y);
    } else {
        return cudaErrorInvalidValue;
    }
}

template <typename T, int H, int CX, int NT>
static cudaError_t row_inv_prox3_nt(const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt,
                                    T* Y, T* U, const AdmmState<T>* st, const C2<T>* stw) {
    constexpr int E = row2_elems(H, CX, (int)sizeof(T)), TR = NT / (H / E);
    if (r.N0 % TR != 0) return cudaErrorInvalidValue;
    const size_t smem = Prox3Plan<T, H, E, CX, NT>::smem_bytes;
    dim3 grid(r.N0 / TR, r.M, r.nb / CX);
    const bool plain = !p.nonneg && p.bnd0 >= r.N0 && p.bnd1 >= 2 * H && !p.reg_on_y &&
                       p.wl1.spatial_uniform && (!p.prm.joint || p.wl21.spatial_uniform);
    if (plain)
        return launch(k_row_inv_prox3<T, H, E, CX, NT, true>, grid, dim3(NT), smem, r.stream, Zt,
                      reinterpret_cast<C2<T>*>(p.znext), Y, U, st,
                      p.prm, p.wl1, p.wl21, p.acc, r.tw, stw, r.N0, r.M, p.scale, p.nonneg, p.bnd0, p.bnd1,
                      p.reg_on_y);
    return launch(k_row_inv_prox3<T, H, E, CX, NT, false>, grid, dim3(NT), smem, r.stream, Zt,
                  reinterpret_cast<C2<T>*>(p.znext), Y, U, st,
                  p.prm, p.wl1, p.wl21, p.acc, r.tw, stw, r.N0, r.M, p.scale, p.nonneg, p.bnd0, p.bnd1,
                  p.reg_on_y);
}

// threads per CTA of the cp.async prox kernel: small CTAs (8-16 rows) give the best overlap
template <typename T, int H, int CX>
static cudaError_t row_inv_prox3_go(const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt,
                                    T* Y, T* U, const AdmmState<T>* st, const C2<T>* stw) {
    if constexpr (row2_elems(H, CX, (int)sizeof(T)) != 0) {
        constexpr int E = row2_elems(H, CX, (int)sizeof(T)), TPF = H / E;
        if constexpr (TPF <= 16) {
            if (p.prox_threads == 128 && r.N0 % (128 / TPF) == 0)
                return row_inv_prox3_nt<T, H, CX, 128>(r, p, Zt, Y, U, st, stw);
        }
        return row_inv_prox3_nt<T, H, CX, kRow2Threads>(r, p, Zt, Y, U, st, stw);
    } else {
        return cudaErrorInvalidValue;
    }
}

template <typename T, int H>
cudaError_t row_inv_prox2_launch(const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt, T* Y,
                                 T* U, const AdmmState<T>* st, const C2<T>* stw) {
    const bool async_ok = !p.use_v2_sync &&
                          ((!p.prm.joint) || p.wl21.spatial_uniform || r.Cx > 1);
    switch (r.Cx) {
        case 1:
            if (async_ok) return row_inv_prox3_go<T, H, 1>(r, p, Zt, Y, U, st, stw);
            return row_inv_prox2_cx<T, H, 1>(r, p, Zt, Y, U, st, stw);
        case 2:
            if (async_ok) return row_inv_prox3_go<T, H, 2>(r, p, Zt, Y, U, st, stw);
            return row_inv_prox2_cx<T, H, 2>(r, p, Zt, Y, U, st, stw);
        case 3:
            if (async_ok) return row_inv_prox3_go<T, H, 3>(r, p, Zt, Y, U, st, stw);
            return row_inv_prox2_cx<T, H, 3>(r, p, Zt, Y, U, st, stw);
        case 4:
            if (async_ok) return row_inv_prox3_go<T, H, 4>(r, p, Zt, Y, U, st, stw);
            return row_inv_prox2_cx<T, H, 4>(r, p, Zt, Y, U, st, stw);
        default: return cudaErrorInvalidValue;
    }
}

// k_col3 launch: returns false when the variant does not fit (shared memory, no resident cluster)
template <typename T, int N0, int E, int CPG, int NT, int CD, bool PAIR, bool DFS = false>
static bool col3_go(ColLaunch<T>& c, const C2<T>* stw, unsigned cs, cudaError_t& result) {
    auto kern = k_col3<T, N0, E, CPG, NT, CD, PAIR, DFS>;
    const size_t smem3 = col3_smem_bytes<T, N0, E, NT, CD, PAIR, (DFS ? (NT / (N0 / E)) * CPG : 0)>((int)cs);
    if (smem3 > kSmemLimit) return false;
    static int resident3[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // per cluster size
    if (resident3[cs] == 0) {
#ifndef SPCSC_EMU
        // leave what two CTAs per SM do not need of the unified array to L1: it holds the dictionary slice
        const int per_sm = (CPG == 1 && sizeof(T) == 4) ? 3 : 2;
        const int pct = (int)((per_sm * (smem3 + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
#endif
        resident3[cs] = max_active_clusters(kern, dim3(NT), cs, smem3);
        if (resident3[cs] <= 0) resident3[cs] = -1;
    }
    const int ncl = resident3[cs];
    if (ncl <= 0) return false;
    // runs of images per item: about six rounds of items per resident cluster
    int chunk = (int)(((long long)c.a.N1f * c.nb) / (6LL * ncl));
    if (chunk < 1) chunk = 1;
    if (chunk > c.nb) chunk = c.nb;
    const int nitems = c.a.N1f * ((c.nb + chunk - 1) / chunk);
    const int use = ncl < nitems ? ncl : nitems;
    g_col_variant = (PAIR ? 4 : (CPG == 1 && sizeof(T) == 4 ? 5 : 3)) + (DFS ? 10 : 0);
    result = launch_cluster(kern, dim3(use * cs, 1), dim3(NT), cs, smem3, c.stream, c.in, c.out, c.Df,
                            c.Sf, c.G, c.st, c.acc, stw, c.a, c.nb, chunk, c.prefetch);
    return true;
}

// k_col4 launch (single-channel dictionaries): returns false when it does not fit
constexpr int kCol4Threads = 512;
template <typename T, int N0, int E>
static bool col4_go(ColLaunch<T>& c, const C2<T>* stw, cudaError_t& result) {
    constexpr int NT = kCol4Threads, TPF = N0 / E, NG = NT / TPF;
    if constexpr (N0 > NT || NG < 1) {
        return false;
    } else {
        auto kern = k_col4<T, N0, E, NT>;
        const unsigned cs = (unsigned)((c.a.M + NG - 1) / NG);
        if (cs > 8) return false;
        const size_t smem4 = col4_smem_bytes<T, N0, E, NT>((int)cs);
        if (smem4 > 227 * 1024) return false;
        static int resident4[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // per cluster size
        if (resident4[cs] == 0) {
            resident4[cs] = max_active_clusters(kern, dim3(NT), cs, smem4);
            if (resident4[cs] <= 0) resident4[cs] = -1;
        }
        const int ncl = resident4[cs];
        if (ncl <= 0) return false;
        const long long total = (long long)c.a.N1f * c.nb;
        const int use = (long long)ncl < total ? ncl : (int)total;
        g_col_variant = 6;
#ifndef SPCSC_EMU
        if constexpr (N0 == 256 && sizeof(T) == 4) {
            // diagnosis only (SPCSC_COL4_DBG=1): cycles per phase of two probe threads per CTA, printed once
            static int dbg_left = getenv("SPCSC_COL4_DBG") ? atoi(getenv("SPCSC_COL4_DBG")) : 0;
            if (dbg_left > 0) {
                --dbg_left;
                auto kd = k_col4<T, N0, E, NT, true>;
                cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
                unsigned* dbg = nullptr;
                const size_t nw = (size_t)use * cs * 2 * 13;
                cudaMalloc(&dbg, nw * sizeof(unsigned));
                cudaMemset(dbg, 0, nw * sizeof(unsigned));
                result = launch_cluster(kd, dim3(use * cs, 1), dim3(NT), cs, smem4, c.stream, c.in, c.out, c.Df,
                                        c.Sf, c.G, c.st, c.acc, stw, c.a, c.nb, dbg);
                cudaStreamSynchronize(c.stream);
                std::vector<unsigned> hbuf(nw);
                cudaMemcpy(hbuf.data(), dbg, nw * sizeof(unsigned), cudaMemcpyDeviceToHost);
                cudaFree(dbg);
                for (int probe = 0; probe < 2; ++probe) {
                    double tot[12] = {0};
                    double ns = 0;
                    for (int bk = 0; bk < (int)(use * cs); ++bk) {
                        const unsigned* o = hbuf.data() + ((size_t)bk * 2 + probe) * 13;
                        for (int i = 0; i < 12; ++i) tot[i] += o[i];
                        ns += o[12];
                    }
                    fprintf(stderr, "k_col4 phases (probe %d, cycles per slab, %g slabs):", probe, ns);
                    double sum = 0;
                    for (int i = 0; i < 12; ++i) { fprintf(stderr, " %.0f", tot[i] / ns); sum += tot[i] / ns; }
                    fprintf(stderr, " | total %.0f\n", sum);
                }
                return true;
            }
        }
#endif
        result = launch_cluster(kern, dim3(use * cs, 1), dim3(NT), cs, smem4, c.stream, c.in, c.out, c.Df,
                                c.Sf, c.G, c.st, c.acc, stw, c.a, c.nb, (unsigned*)nullptr);
        return true;
    }
}

// k_col5 launch (single-channel dictionaries): independent thread groups sharing the staged dictionary columns
template <typename T, int N0, int E, int NGRP, bool TST = false>
static bool col5_go(ColLaunch<T>& c, const C2<T>* stw, cudaError_t& result) {
    constexpr int NT = kCol4Threads, TPF = N0 / E, NGG = (NT / NGRP) / TPF;
    if constexpr (NGG < 1) {
        return false;
    } else {
        auto kern = k_col5<T, N0, E, NT, NGRP, TST>;
        const unsigned cs = (unsigned)((c.a.M + NGG - 1) / NGG);
        if (cs > 8) return false;
        const size_t smem5 = col5_smem_bytes<T, N0, E, NT, NGRP>((int)cs);
        if (smem5 > 227 * 1024) return false;
        static int resident5[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // per cluster size
        if (resident5[cs] == 0) {
            resident5[cs] = max_active_clusters(kern, dim3(NT), cs, smem5);
            if (resident5[cs] <= 0) resident5[cs] = -1;
        }
        const int ncl = resident5[cs];
        if (ncl <= 0) return false;
        const long long total = (long long)c.a.N1f * c.nb;
        const int use = (long long)ncl < total ? ncl : (int)total;
        static int stagger = -1;
        if (stagger < 0) stagger = getenv("SPCSC_COL5_STAGGER") ? atoi(getenv("SPCSC_COL5_STAGGER")) : 0;
        g_col_variant = NGRP == 1 ? 8 + (TST ? 1 : 0) : 7;
        result = launch_cluster(kern, dim3(use * cs, 1), dim3(NT), cs, smem5, c.stream, c.in, c.out, c.Df,
                                c.Sf, c.G, c.st, c.acc, stw, c.a, c.nb, stagger);
        return true;
    }
}

template <typename T, int N0, int CD>
static cudaError_t col2_go(int mode, ColLaunch<T>& c, const C2<T>* stw) {
    constexpr int E = col2_elems<T>(), NT = kCol2Threads, CPG = col2_cpg<T>();
    constexpr int TPF = N0 / E, NG = NT / TPF;
    const int per_cta = NG * CPG;
    const unsigned cs = (unsigned)((c.a.M + per_cta - 1) / per_cta);
    c.a.N0 = N0;
    const size_t smem = ((size_t)NG * fft_region(N0) + 2 * CD * N0 + stage_tw_len(N0, E) + 2 * N0) * sizeof(C2<T>) +
                        32 * sizeof(double);
    dim3 grid(c.a.N1f * cs, c.nb);
    c.a.ntiles = c.a.N1f * c.nb;
    if (mode == COL_FWD) {
        if constexpr (CD == 1)          // forward columns only (set-up transforms, coefficient spectra)
            return launch_cluster(k_col2<T, N0, E, CPG, NT, 1, true, 0, false, false>, grid, dim3(NT), cs,
                                  smem, c.stream, c.in, c.out, c.Df, c.Sf, c.G, c.st, c.Lstep, c.acc, stw,
                                  c.a, c.sumout, c.sumin, c.ref);
        return cudaErrorInvalidValue;
    }
    if (mode == COL_GRAD_INV)       // PGM gradient step on slabs already in the frequency domain
        return launch_cluster(k_col2<T, N0, E, CPG, NT, CD, false, 2, true, false>, grid, dim3(NT), cs, smem,
                              c.stream, c.in, c.out, c.Df, c.Sf, c.G, c.st, c.Lstep, c.acc, stw, c.a,
                              c.sumout, c.sumin, c.ref);
    if (mode == COL_FWD_EVAL)       // PGM: forward columns + evaluation of the candidate
        return launch_cluster(k_col2<T, N0, E, CPG, NT, CD, true, 4, false, false>, grid, dim3(NT), cs, smem,
                              c.stream, c.in, c.out, c.Df, c.Sf, c.G, c.st, c.Lstep, c.acc, stw, c.a,
                              c.sumout, c.sumin, c.ref);
    if (mode != COL_ADMM) return cudaErrorInvalidValue;
    g_col_variant = 2;
    if constexpr (CD == 1) {
        if (c.push == 4) {          // k_col4: slab, dictionary columns and signal row staged by bulk copies
            cudaError_t e4 = cudaErrorInvalidValue;
            if (col4_go<T, N0, E>(c, stw, e4)) return e4;
        }
        if (c.push == 5) {          // k_col5: the same with two independent thread groups per CTA
            cudaError_t e5 = cudaErrorInvalidValue;
            if (col5_go<T, N0, E, 2>(c, stw, e5)) return e5;
        }
        if (c.push == 6) {          // one group: k_col4's shape with the stage twiddles in registers
            cudaError_t e5 = cudaErrorInvalidValue;
            if (col5_go<T, N0, E, 1>(c, stw, e5)) return e5;
        }
        if (c.push == 7) {          // ... and the result leaving by bulk copies (TMA stores)
            cudaError_t e5 = cudaErrorInvalidValue;
            if (col5_go<T, N0, E, 1, true>(c, stw, e5)) return e5;
        }
    }
    if (c.push && (c.push < 4 || c.push > 7) && !c.bulk) {
        // k_col3: persistent clusters over (frequency column, run of images) items; the per-frequency sums
        // travel by st.async pushes instead of cluster barriers.  push == 2 (float32): the two columns of a lane
        // group are transformed together, exchanging 16-byte elements
        cudaError_t e3 = cudaErrorInvalidValue;
        bool done = false;
        if constexpr (CD == 1) {    // + 10: the CTA's dictionary columns staged in shared memory
            if (c.push == 11) done = col3_go<T, N0, E, CPG, NT, CD, false, true>(c, stw, cs, e3);
            if constexpr (CPG == 2) {
                if (c.push == 12) done = col3_go<T, N0, E, CPG, NT, CD, true, true>(c, stw, cs, e3);
            }
        }
        if constexpr (CPG == 2) {
            if (c.push == 2) done = col3_go<T, N0, E, CPG, NT, CD, true>(c, stw, cs, e3);
            if (c.push == 3) {
                // one column per lane group: half the register payload, 3 CTAs per SM, clusters twice as large
                const unsigned cs1 = (unsigned)((c.a.M + NG - 1) / NG);
                if (cs1 <= 8) done = col3_go<T, N0, E, 1, NT, CD, false>(c, stw, cs1, e3);
            }
        }
        if (!done) done = col3_go<T, N0, E, CPG, NT, CD, false>(c, stw, cs, e3);
        if (done) return e3;
    }
    if (c.bulk) {
        // persistent clusters with the next slab prefetched by a bulk copy; in place is fine (a
        // slab is in registers before the prefetch of the next one is issued, and stored after)
        auto kern = k_col2<T, N0, E, CPG, NT, CD, true, 1, true, true>;
        const size_t smem_b = smem + 2 * sizeof(mbar_t) + (size_t)per_cta * N0 * sizeof(C2<T>);
        static int resident[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per cluster size
        if (cs < 8 && resident[cs] == 0) resident[cs] = max_active_clusters(kern, dim3(NT), cs, smem_b);
        const int ncl = cs < 8 ? resident[cs] : 0;
        if (ncl > 0) {
            const int use = ncl < c.a.ntiles ? ncl : c.a.ntiles;
            return launch_cluster(kern, dim3(use * cs, 1), dim3(NT), cs, smem_b, c.stream, c.in, c.out,
                                  c.Df, c.Sf, c.G, c.st, c.Lstep, c.acc, stw, c.a, c.sumout, c.sumin, c.ref);
        }
    }
#ifndef SPCSC_EMU
    {   // two CTAs of this kernel per SM need 2 x smem of shared memory; leave the rest of the unified
        // array to L1, which serves the second read of the dictionary slice
        static bool carved = false;
        if (!carved) {
            const int pct = (int)((2 * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
            cudaFuncSetAttribute(k_col2<T, N0, E, CPG, NT, CD, true, 1, true, false>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
            carved = true;
        }
    }
#endif
    return launch_cluster(k_col2<T, N0, E, CPG, NT, CD, true, 1, true, false>, grid, dim3(NT), cs, smem,
                          c.stream, c.in, c.out, c.Df, c.Sf, c.G, c.st, c.Lstep, c.acc, stw, c.a, c.sumout, c.sumin, c.ref);
}

template <typename T, int N0>
cudaError_t col2_launch(int mode, ColLaunch<T> c, const C2<T>* stw) {
    if constexpr (N0 >= 32 && N0 <= 512 && N0 / col2_elems<T>() <= 32) {
        if constexpr (sizeof(T) == 8) {        // float64: single-channel dictionaries only
            if (c.a.Cd != 1) return cudaErrorInvalidValue;
            return col2_go<T, N0, 1>(mode, c, stw);
        }
        switch (c.a.Cd) {
            case 1: return col2_go<T, N0, 1>(mode, c, stw);
            case 2: return col2_go<T, N0, 2>(mode, c, stw);
            case 3: return col2_go<T, N0, 3>(mode, c, stw);
            case 4: return col2_go<T, N0, 4>(mode, c, stw);
            default: return cudaErrorInvalidValue;
        }
    } else {
        return cudaErrorInvalidValue;
    }
}

}  // namespace spcsc

// launchers.cuh -- size-templated kernel launchers.  Each transform length is compiled in
// its own translation unit (size_inst.cu with -DSPCSC_SIZE=n) so the library builds in
// parallel; spcsc.cu dispatches on the runtime size.
#pragma once

#include "kernels3.cuh"
#include "kernels_gen.cuh"

namespace spcsc {

constexpr size_t kSmemLimit = 220 * 1024;     // of the 227 KB a CTA may opt into on sm_100a

template <typename T>
struct RowArgs {
    int N0, M, nb, TR;           // rows per image, filters, batch (K*Cx or Cd ...), rows per CTA
    int Cx;                      // channels handled together by the prox kernel
    int N1, gen;                 // gen: any-size direct-DFT path (N1 is then the real row length)
    const C2<T>* tw;             // exp(-2 pi i j / N1), j < N1
    cudaStream_t stream;
};

template <typename T>
struct ProxArgs {
    AdmmParams<T> prm;
    WeightView<T> wl1, wl21;
    double* acc;
    T scale;
    int nonneg, bnd0, bnd1, reg_on_y;
    int use_v2_sync;             // 1: synchronous-load row kernel (k_row_inv_prox2) even when CX == 1
    void* znext;                 // fused forward output (C2<T>*), or null
    int prox_threads;            // 128 or 256 threads per CTA for k_row_inv_prox3
};

enum ColMode {
    COL_FWD = 0,        // forward column FFT only
    COL_INV = 1,        // inverse column FFT only
    COL_ADMM = 2,       // forward, Sherman-Morrison / Woodbury solve, inverse
    COL_ADMM_NOFFT = 3, // solve only (input and output in the full 2-D frequency domain)
    COL_GRAD_INV = 4,   // gradient step (q = (Sf - s)/L) then inverse   [PGM]
    COL_FWD_SUM = 5,    // forward, write s_c = sum_m Df_c X
    COL_SUM = 6,        // write s_c = sum_m Df_c X (input already in frequency domain)
    COL_FWD_EVAL = 7    // forward, then PGM evaluation sums of the candidate (k_col SOLVE 4)
};

template <typename T>
struct ColLaunch {
    const C2<T>* in;
    C2<T>* out;
    const C2<T>* Df;
    const C2<T>* Sf;
    const C2<T>* G;
    C2<T>* sumout;
    const C2<T>* sumin;          // PGM: per-frequency sums of the current Yf
    const C2<T>* ref;            // PGM: Yf slabs the candidate is compared with
    const AdmmState<T>* st;
    T Lstep;
    double* acc;
    const C2<T>* tw;             // exp(-2 pi i j / N0)
    int nb;                      // slabs per frequency column (grid.y)
    ColArgs a;                   // MC / nchunk / parts filled by the launcher
    int gen;                     // any-size direct-DFT path (a.N0 holds the run-time length)
    int bulk;                    // k_col2: persistent clusters with bulk-copy prefetch of the next slab
    int push;                    // COL_ADMM: k_col3 (persistent clusters, sums pushed over DSMEM) instead of k_col2
    int prefetch;                // k_col3: L2 prefetch of the cluster's next slab
    cudaStream_t stream;
};

// rows: H = N1/2
template <typename T, int H>
cudaError_t row_fwd_launch(const RowArgs<T>& r, const T* A, const T* B, const AdmmState<T>* st,
                           C2<T>* Zt);
template <typename T, int H>
cudaError_t row_inv_launch(const RowArgs<T>& r, const C2<T>* Zt, T* X, T scale);
template <typename T, int H>
cudaError_t row_inv_prox_launch(const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt, T* Y,
                                T* U, const AdmmState<T>* st);
template <typename T>
struct PgmRowArgs {
    T thr_scale;                 // lmbda / L
    WeightView<T> wl1;
    double* acc;
    T scale;
    int nonneg, bnd0, bnd1;
    const C2<T>* stw;            // stage twiddles of the register plan (null: general kernel)
};
template <typename T, int H>
cudaError_t row_inv_prox_fwd_launch(const RowArgs<T>& r, const PgmRowArgs<T>& p, C2<T>* Vt, T* X);
// columns
template <typename T, int N0>
cudaError_t col_launch(int mode, ColLaunch<T> c);

// ---- any-size path (kernels_gen.cuh, k_col<T, 0, ...>)
template <typename T>
struct GenRowArgs {
    int N0, N1, M, nb, Cx, TR;
    const C2<T>* tw;
    cudaStream_t stream;
};
template <typename T>
cudaError_t row_fwd_gen_launch(const GenRowArgs<T>& r, const T* A, const T* B, const AdmmState<T>* st,
                               C2<T>* Zt);
template <typename T>
cudaError_t row_inv_gen_launch(const GenRowArgs<T>& r, const C2<T>* Zt, T* X, T scale);
template <typename T>
cudaError_t row_inv_prox_gen_launch(const GenRowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt,
                                    T* Y, T* U, const AdmmState<T>* st);
template <typename T>
cudaError_t row_inv_prox_fwd_gen_launch(const GenRowArgs<T>& r, const PgmRowArgs<T>& p, C2<T>* Vt,
                                        T* X);

// ---- kernel set v2 (kernels2.cuh): plans and eligibility ------------------------------
constexpr int kRow2Threads = 256;
constexpr int kCol2Threads = 256;
// column kernel: elements per lane and columns per lane group.  float32: 16 x 2 (32 complex values =
// 64 registers of payload per thread); float64: 8 x 1 (the same 32 registers of payload as one
// float32 column)
template <typename T> constexpr int col2_elems() { return sizeof(T) == 4 ? 16 : 8; }
template <typename T> constexpr int col2_cpg() { return sizeof(T) == 4 ? 2 : 1; }

// elements per lane of the v2 row plan (0: no v2 plan for this length); esz = sizeof(T)
constexpr int row2_elems(int H, int Cx, int esz) {
    int e = H >= 128 ? 16 : (H >= 32 ? 8 : 0);
    if ((Cx > 1 || esz == 8) && e > 8) e = 8;
    if (e != 0 && H / e > 32) e = 0;
    return e;
}
constexpr int row2_tile(int H, int Cx, int esz) {
    return row2_elems(H, Cx, esz) == 0 ? 0 : kRow2Threads / (H / row2_elems(H, Cx, esz));
}
template <typename T>
inline bool row2_ok(int H, int N0, int Cx) {
    const int tr = row2_tile(H, Cx, (int)sizeof(T));
    const int tr128 = tr / 2;                      // the 128-thread prox kernel
    return tr > 0 && N0 % tr == 0 && (tr128 == 0 || N0 % tr128 == 0) && Cx <= 4;
}
template <typename T>
inline bool col2_ok(int N0, int M, int Cd) {
    if (Cd < 1 || Cd > 4 || N0 < 32 || N0 > 512 || N0 / col2_elems<T>() > 32) return false;
    if (sizeof(T) == 8 && Cd > 1) return false;    // the Woodbury variants stay float32 for now
    const int per_cta = (kCol2Threads / (N0 / col2_elems<T>())) * col2_cpg<T>();
    return (M + per_cta - 1) / per_cta <= 8;       // portable cluster size
}

template <typename T, int H>
cudaError_t row_fwd2_launch(const RowArgs<T>& r, const T* A, const T* B, const AdmmState<T>* st,
                            C2<T>* Zt, const C2<T>* stw, int gated);
template <typename T, int H>
cudaError_t row_inv_prox2_launch(const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt, T* Y,
                                 T* U, const AdmmState<T>* st, const C2<T>* stw);
template <typename T, int N0>
cudaError_t col2_launch(int mode, ColLaunch<T> c, const C2<T>* stw);

// Rows per CTA for the row kernels (shared by launch code and memory planning).
template <typename T>
inline int row_tile(int H, int N0, int Cx) {
    int e = H >= 64 ? 8 : (H >= 16 ? 4 : 2);
    int tpf = H / e;
    int tr = 256 / tpf;
    if (tr < 1) tr = 1;
    if (tr > N0) tr = N0;
    while (tr > 1 && (size_t)Cx * tr * (H + 1) * sizeof(C2<T>) > 96 * 1024) tr >>= 1;
    return tr;
}

}  // namespace spcsc

// spcsc.cu -- host side of libspcsc.so: handle management, device memory, the per-iteration
// launch schedule and the extern "C" entry points declared in include/spcsc.h.
#include <spcsc.h>

#include <cmath>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "launchers.cuh"
#include "kernels_cdl.cuh"

#ifndef SPCSC_EMU
#include <dlfcn.h>
#endif

namespace spcsc {

// ---- dispatch on transform length -------------------------------------------------------
#define SPCSC_FOR_SIZES(X) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024)

#define SPCSC_DECL(n)                                                                           \
    extern template cudaError_t row_fwd_launch<float, n>(const RowArgs<float>&, const float*,   \
                                                         const float*, const AdmmState<float>*, \
                                                         C2<float>*);                           \
    extern template cudaError_t row_fwd_launch<double, n>(const RowArgs<double>&, const double*, \
                                                          const double*,                         \
                                                          const AdmmState<double>*, C2<double>*); \
    extern template cudaError_t row_inv_launch<float, n>(const RowArgs<float>&, const C2<float>*, \
                                                         float*, float);                        \
    extern template cudaError_t row_inv_launch<double, n>(const RowArgs<double>&,                \
                                                          const C2<double>*, double*, double);  \
    extern template cudaError_t row_inv_prox_launch<float, n>(                                  \
        const RowArgs<float>&, const ProxArgs<float>&, const C2<float>*, float*, float*,        \
        const AdmmState<float>*);                                                               \
    extern template cudaError_t row_inv_prox_launch<double, n>(                                 \
        const RowArgs<double>&, const ProxArgs<double>&, const C2<double>*, double*, double*,   \
        const AdmmState<double>*);                                                              \
    extern template cudaError_t col_launch<float, n>(int, ColLaunch<float>);                    \
    extern template cudaError_t col_launch<double, n>(int, ColLaunch<double>);
SPCSC_FOR_SIZES(SPCSC_DECL)

static bool supported_len(int n) { return n >= 2 && n <= 1024 && (n & (n - 1)) == 0; }

template <typename T>
static GenRowArgs<T> gen_args(const RowArgs<T>& r) {
    GenRowArgs<T> g;
    g.N0 = r.N0; g.N1 = r.N1; g.M = r.M; g.nb = r.nb; g.Cx = r.Cx; g.TR = r.TR; g.tw = r.tw;
    g.stream = r.stream;
    return g;
}
template <typename T>
static cudaError_t row_fwd(int H, const RowArgs<T>& r, const T* A, const T* B,
                           const AdmmState<T>* st, C2<T>* Zt) {
    if (r.gen) return row_fwd_gen_launch<T>(gen_args(r), A, B, st, Zt);
    switch (H) {
#define X(n) case n: return row_fwd_launch<T, n>(r, A, B, st, Zt);
        SPCSC_FOR_SIZES(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
template <typename T>
static cudaError_t row_inv(int H, const RowArgs<T>& r, const C2<T>* Zt, T* Xo, T scale) {
    if (r.gen) return row_inv_gen_launch<T>(gen_args(r), Zt, Xo, scale);
    switch (H) {
#define X(n) case n: return row_inv_launch<T, n>(r, Zt, Xo, scale);
        SPCSC_FOR_SIZES(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
template <typename T>
static cudaError_t row_inv_prox(int H, const RowArgs<T>& r, const ProxArgs<T>& p, const C2<T>* Zt,
                                T* Y, T* U, const AdmmState<T>* st) {
    if (r.gen) return row_inv_prox_gen_launch<T>(gen_args(r), p, Zt, Y, U, st);
    switch (H) {
#define X(n) case n: return row_inv_prox_launch<T, n>(r, p, Zt, Y, U, st);
        SPCSC_FOR_SIZES(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
template <typename T>
static cudaError_t row_inv_prox_fwd(int H, const RowArgs<T>& r, const PgmRowArgs<T>& p, C2<T>* Vt,
                                    T* Xo) {
    if (r.gen) return row_inv_prox_fwd_gen_launch<T>(gen_args(r), p, Vt, Xo);
    switch (H) {
#define X(n) case n: return row_inv_prox_fwd_launch<T, n>(r, p, Vt, Xo);
        SPCSC_FOR_SIZES(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
template <typename T>
static cudaError_t col(int N0, int mode, const ColLaunch<T>& c) {
    if (c.gen) return col_launch<T, 0>(mode, c);
    switch (N0) {
#define X(n) case n: return col_launch<T, n>(mode, c);
        SPCSC_FOR_SIZES(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

template <typename T>
static cudaError_t row_fwd2(int H, const RowArgs<T>& r, const T* A, const T* B,
                            const AdmmState<T>* st, C2<T>* Zt, const C2<T>* stw, int gated) {
    {
        switch (H) {
#define X(n) case n: return row_fwd2_launch<T, n>(r, A, B, st, Zt, stw, gated);
            SPCSC_FOR_SIZES(X)
#undef X
        }
    }
    return cudaErrorInvalidValue;
}
template <typename T>
static cudaError_t row_inv_prox2(int H, const RowArgs<T>& r, const ProxArgs<T>& p,
                                 const C2<T>* Zt, T* Y, T* U, const AdmmState<T>* st,
                                 const C2<T>* stw) {
    {
        switch (H) {
#define X(n) case n: return row_inv_prox2_launch<T, n>(r, p, Zt, Y, U, st, stw);
            SPCSC_FOR_SIZES(X)
#undef X
        }
    }
    return cudaErrorInvalidValue;
}
int g_col_variant = 0;      // set by the cluster column launchers: 2 k_col2, 3 k_col3
template <typename T>
static cudaError_t col2(int N0, int mode, const ColLaunch<T>& c, const C2<T>* stw) {
    {
        switch (N0) {
#define X(n) case n: return col2_launch<T, n>(mode, c, stw);
            SPCSC_FOR_SIZES(X)
#undef X
        }
    }
    return cudaErrorInvalidValue;
}

// complex spectra between device order [b][N1f][Mm][N0] (b = k*Cc + c) and the reference's
// (N0, N1f, Cc, Kk, Mm)
template <typename T>
SPCSC_GLOBAL void k_freq_to_ext(const C2<T>* SPCSC_RESTRICT in, C2<T>* SPCSC_RESTRICT ext, int N0,
                                int N1f, int Cc, int Kk, int Mm) {
    const size_t n = (size_t)N0 * N1f * Cc * Kk * Mm;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t t = i;
        const int m = (int)(t % Mm); t /= Mm;
        const int k = (int)(t % Kk); t /= Kk;
        const int c = (int)(t % Cc); t /= Cc;
        const int wf = (int)(t % N1f); t /= N1f;
        const int h = (int)t;
        ext[i] = in[((((size_t)k * Cc + c) * N1f + wf) * Mm + m) * N0 + h];
    }
}

}  // namespace spcsc

using namespace spcsc;

// ---- process-wide allocation pools ----------------------------------------------------
// cudaMalloc / cudaHostAlloc of the 0.5 GB arrays of a solver cost tens of milliseconds; a
// solver that is rebuilt (dictionary learning, repeated solves) gets its buffers back from here.
#include <map>
#include <mutex>
namespace {
struct MemPool {
    std::mutex mu;
    std::multimap<size_t, void*> free_;          // size -> block
    std::map<void*, std::pair<size_t, int>> live; // block -> (size, device)
    bool host;
    explicit MemPool(bool h) : host(h) {}
    static size_t round(size_t n) {
        const size_t g = n >= (1u << 20) ? (1u << 20) : 512;
        return (n + g - 1) / g * g;
    }
    cudaError_t alloc(void** p, size_t bytes, int device) {
        const size_t want = round(bytes ? bytes : 1);
        {
            std::lock_guard<std::mutex> l(mu);
            for (auto it = free_.lower_bound(want); it != free_.end() && it->first <= want + want / 4; ++it) {
                auto lv = live.find(it->second);
                if (lv != live.end() && lv->second.second == device) {
                    *p = it->second;
                    free_.erase(it);
                    return cudaSuccess;
                }
            }
        }
        void* q = nullptr;
        cudaError_t e = host ? cudaMallocHost(&q, want) : cudaMalloc(&q, want);
        if (e != cudaSuccess) {
            trim();
            e = host ? cudaMallocHost(&q, want) : cudaMalloc(&q, want);
            if (e != cudaSuccess) return e;
        }
        std::lock_guard<std::mutex> l(mu);
        live[q] = std::make_pair(want, device);
        *p = q;
        return cudaSuccess;
    }
    void release(void* p) {
        if (!p) return;
        std::lock_guard<std::mutex> l(mu);
        auto it = live.find(p);
        if (it == live.end()) return;
        free_.insert(std::make_pair(it->second.first, p));
    }
    void trim() {
        std::lock_guard<std::mutex> l(mu);
        for (auto& kv : free_) {
            if (host) cudaFreeHost(kv.second); else cudaFree(kv.second);
            live.erase(kv.second);
        }
        free_.clear();
    }
};
MemPool& dev_pool() { static MemPool* p = new MemPool(false); return *p; }
MemPool& host_pool() { static MemPool* p = new MemPool(true); return *p; }
}  // namespace

// ---- NCCL, resolved at run time (no link-time dependency) ---------------------------------
namespace {
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string err;
};
NcclApi& nccl_api(const char* libname) {
    static NcclApi api;
    if (api.ok) return api;
#ifdef SPCSC_EMU
    (void)libname;
    api.err = "NCCL is not available in the CPU emulation build";
#else
    const char* name = (libname && libname[0]) ? libname : "libnccl.so.2";
    void* hnd = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (!hnd) { api.err = std::string("dlopen failed: ") + dlerror(); return api; }
    api.GetUniqueId = (int (*)(NcclApi::UniqueId*))dlsym(hnd, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, NcclApi::UniqueId, int))dlsym(hnd, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(hnd, "ncclAllReduce");
    api.CommDestroy = (int (*)(void*))dlsym(hnd, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(hnd, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy || !api.GetErrorString) {
        api.err = "libnccl does not export the expected symbols";
        return api;
    }
    api.ok = true;
#endif
    return api;
}
}  // namespace

static thread_local std::string g_last_error;

struct spcsc_comm {
    void* comm = nullptr;
    NcclApi* api = nullptr;
    int rank = 0, nranks = 1, device = 0;
    // peer-memory all-reduce (kernels.cuh, p2p_allreduce): this rank's block -- plain cudaMalloc, IPC needs
    // it -- and the mapped blocks of all ranks.  They belong to the communicator, so every solver that
    // attaches it shares them and only the first attachment pays for the handle exchange.
    void* p2p_own = nullptr;
    void* p2p_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool p2p_ready = false;
};

struct spcsc_handle {
    std::string err;
    bool poisoned = false;
    virtual ~spcsc_handle() {}
    virtual int set_dict(const void* D) = 0;
    virtual int set_signal(const void* S) = 0;
    virtual int set_l1_weight(const void* w, const int64_t* shape) = 0;
    virtual int set_l21_weight(const void* w, const int64_t* shape) = 0;
    virtual int admm_configure(const spcsc_admm_opts* o) = 0;
    virtual int admm_reset(double rho) = 0;
    virtual int admm_set_rho(double rho) = 0;
    virtual int admm_set_iter(int k) = 0;
    virtual int admm_iterate(int n, spcsc_itstat* rows, int* n_done, int* stopped) = 0;
    virtual int admm_get_scalars(double* rho, int* k) = 0;
    virtual int admm_last_timing(float* ms, int64_t* launches) = 0;
    virtual int admm_profile(int n, float* ms4) = 0;
    virtual int admm_schedule_info(int32_t* info) = 0;
    virtual int get_array(int which, void* out) = 0;
    virtual int set_array(int which, const void* in) = 0;
    virtual int reconstruct(const void* X, void* out) = 0;
    virtual int synchronize() = 0;
    virtual int attach_comm(spcsc_comm* c, double global_nx) = 0;
    virtual int p2p_export(void* handle64) = 0;
    virtual int p2p_attach(int rank, int nranks, const void* handles64) = 0;
    virtual int set_gradreg(const void* ghg, const void* wgrd) = 0;
    virtual int pgm_set_mask(const void* W, const int64_t* shape) = 0;
    virtual int pgm_configure(const spcsc_pgm_opts* o) = 0;
    virtual int pgm_reset(const void* X0) = 0;
    virtual int pgm_trial(double L, double* out) = 0;
    virtual int pgm_accept(double coef) = 0;
    virtual int pgm_policy_stats(int store, double* out) = 0;
    virtual int pgm_combine_y(double a, double b, int save_prev) = 0;
    virtual int pgm_finish(int mode, double c0, double* out) = 0;
    virtual int ccmod_reset(const void* D0, int zero_mean) = 0;
    virtual int ccmod_setcoef_device(int source) = 0;
    virtual int ccmod_setcoef(const void* Z) = 0;
    virtual int ccmod_step(double L, double coef, int flags, double* out) = 0;
    virtual int ccmod_trial(double L, double* out) = 0;
    virtual int ccmod_accept(double coef, int flags, double* out) = 0;
    virtual int ccmod_get_dict(void* out) = 0;
    virtual int ccmod_push_dict() = 0;
    virtual int ccmod_cns_init(double rho, int y0_given, long long nb_global) = 0;
    virtual int ccmod_cns_step(double rho, double udiv, double rlx, int flags, double* out) = 0;
    virtual int ccmod_cns_get(int which, void* out) = 0;
    virtual int ccmod_set_supports(const int32_t* hw) = 0;
    virtual int ccmod_get_spectrum(int which, void* out) = 0;
};

namespace {

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            poisoned = true;                                                             \
            return SPCSC_ERR_CUDA;                                                       \
        }                                                                                \
    } while (0)

#define FAIL(code, msg)        \
    do {                       \
        err = (msg);           \
        return (code);         \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        release();
        int dev = 0;
        cudaGetDevice(&dev);
        void* q = nullptr;
        cudaError_t e = dev_pool().alloc(&q, count * sizeof(T), dev);
        if (e == cudaSuccess) {
            p = (T*)q;
            n = count;
        }
        return e;
    }
    void release() {
        if (p) cudaDeviceSynchronize();      // nothing in flight may still touch a pooled block
        if (p) dev_pool().release(p);
        p = nullptr;
        n = 0;
    }
};

template <typename T>
std::vector<C2<T>> twiddles(int n) {
    std::vector<C2<T>> w(n);
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 0; j < n; ++j) {
        const double a = -two_pi * (double)j / (double)n;
        w[j] = mk<T>((T)std::cos(a), (T)std::sin(a));
    }
    return w;
}

template <typename T>
class Engine : public spcsc_handle {
  public:
    spcsc_problem pb;
    int N0, N1, H, N1f, C, Cd, Cx, K, M;
    size_t nreal;       // K*Cx*M*N0*N1
    size_t nslab;       // K*Cx*N1f*M*N0
    cudaStream_t stream = nullptr;
    DevBuf<T> Y, U, tmp_real, wl1_buf, wl21_buf, staging;
    DevBuf<C2<T>> Zt, Zscratch, Xscratch, Df, Sf, G, tw_row, tw_col, sum_buf;
    DevBuf<C2<T>> stw_row1, stw_rowc, stw_col;     // stage twiddles of the v2 register plans
    DevBuf<C2<T>> pgA, pgB;                         // PGM: accepted Xf (= Xfprv), Yf; Zt is the candidate
    DevBuf<C2<T>> pgZ, pgYp;                        //   robust backtracking: auxiliary sequence z, Yf of the previous iteration
    DevBuf<C2<T>> pg_sx, pg_rprev, pg_sxprev;       //   step-size policies: sum_m Df Xf, remembered residual / iterate sums
    bool pg_z_init = false, pg_pol_init = false;
    spcsc_pgm_opts popts;
    DevBuf<T> cdX;                                  // CCMOD: dictionary iterate, real [Cd][M][N0][N1]
    DevBuf<C2<T>> cdXf, cdYf, cdV, cdG;             // its spectrum, the momentum point, scratch, gradient
    DevBuf<T> cd_supp;                              //   filter supports of V (sharded: what is exchanged)
    DevBuf<T> ghg_buf;                              // ConvBPDNGradReg: GHG in device order [N1f][N0]
    DevBuf<C2<T>> gw_buf;                           //   per filter (mu w_m, w_m)
    std::vector<T> gr_w;                            //   host copy of w_m (mu arrives with admm_configure)
    bool gradreg = false;
    spcsc_comm* comm_obj = nullptr;                 // attached communicator (owns the peer-memory blocks)
    P2pView p2p{};                                  // peer-memory all-reduce: view of the communicator's blocks
    bool p2p_on = false;
    DevBuf<T> mk_W, mk_r, mk_wr, mk_w2r;            // pgm.ConvBPDNMask: mask and signal-domain work planes [K][C][N0][N1]
    DevBuf<C2<T>> mk_f, mk_grad, mk_sx;             //   spectra [K][C][N1f][N0]: work, rfft(W^2 R_Y), s_X
    bool pgm_mask = false;
    DevBuf<C2<T>> cdZf;                             // coefficient spectra, slab layout [K][N1f][M][N0]
    bool cd_ready = false, cd_have_coef = false;
    int cd_zero_mean = 0;
    DevBuf<int> cd_fsupp;                           // multi-scale dictionaries: (h_m, w_m) per filter, else unallocated
    // consensus dictionary update (admm.ccmod.ConvCnstrMOD_Consensus): per-block copies of the dictionary and
    // duals [K*C][M][N0][N1], their row/column spectra, the Gram rows of the coefficient spectra, the new Y
    DevBuf<T> cnsX, cnsU, cnsYn;
    DevBuf<C2<T>> cnsZ, cnsZ2, cnsG;            // cnsZ2: the solve's output when LinSolveCheck needs its input too
    DevBuf<AdmmState<T>> cns_st;
    bool cns_ready = false, cns_gram_stale = true;
    long long cns_nb_global = 0;             // blocks over all ranks (the mean of the y step runs over them)
    bool pgm_ready = false, pgm_have_cand = false;
    bool v2_rowf = false, v2_rowp = false, v2_col = false;
    bool gen_rows = false, gen_cols = false;   // any-size direct-DFT path for this axis
    bool fuse = false;          // prox kernel also emits the next iteration's row spectra
    bool fused_batch = false, x_in_zt2 = false;
    DevBuf<C2<T>> Zt2;          // ping-pong partner of Zt when fusing
    DevBuf<double> acc;          // kAccBytes: ACC_N doubles + reproducible integer bins
    DevBuf<AdmmState<T>> st;
    DevBuf<StatRow> rows;
    WeightView<T> wl1, wl21;
    AdmmParams<T> prm;
    spcsc_admm_opts opts;
    bool have_dict = false, have_signal = false, configured = false, have_x = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* nccl_comm = nullptr;
    NcclApi* nccl = nullptr;
    int nranks = 1;
    double global_nx = 0.0;
    std::vector<cudaEvent_t> prof_ev;
    // wavefront schedule: groups of wave_g images run row-forward -> column -> row-inverse/prox back to
    // back, through a small scratch that stays in L2, round-robin over wave_s streams
    int wave_g = 0, wave_s = 1;
    bool wave_keep = true;                          // the column kernel stores the X spectra in Zt (get_array(X))
    bool wave_batch = false, wave_persist_set = false;
    bool wave_fused = false;                        // groups through column + prox only, cross-iteration fusion kept
    std::vector<cudaStream_t> wstreams;             // [0] is `stream`
    std::vector<cudaEvent_t> wjoin;
    cudaEvent_t wfork = nullptr;
    DevBuf<C2<T>> wscratch;
    float last_ms = 0.f;
    int64_t last_launches = 0;
    int last_col_kernel = 0;

    explicit Engine(const spcsc_problem& p) : pb(p) {
        N0 = p.N0; N1 = p.N1; H = N1 / 2; N1f = H + 1;
        gen_rows = (N1 & 1) || !supported_len(N1 / 2);
        gen_cols = !supported_len(N0);
        C = p.C; Cd = p.Cd; K = p.K; M = p.M;
        Cx = C - Cd + 1;
        nreal = (size_t)K * Cx * M * N0 * N1;
        nslab = (size_t)K * Cx * N1f * M * N0;
        memset(&opts, 0, sizeof(opts));
        memset(&prm, 0, sizeof(prm));
        memset(&popts, 0, sizeof(popts));
    }
    ~Engine() override {
        cudaSetDevice(pb.device);
        if (stream) cudaStreamSynchronize(stream);
        Y.release(); U.release(); tmp_real.release(); wl1_buf.release(); wl21_buf.release();
        staging.release(); Zt.release(); Zscratch.release(); Xscratch.release(); Df.release();
        Sf.release(); G.release(); tw_row.release(); tw_col.release(); sum_buf.release();
        acc.release(); st.release(); rows.release();
        stw_row1.release(); stw_rowc.release(); stw_col.release();
        pgA.release(); pgB.release(); Zt2.release();
        pgZ.release(); pgYp.release(); pg_sx.release(); pg_rprev.release(); pg_sxprev.release();
        cd_supp.release(); cdX.release(); cdXf.release(); cdYf.release(); cdV.release(); cdG.release(); cdZf.release(); ghg_buf.release(); gw_buf.release();
        cnsX.release(); cnsU.release(); cnsYn.release(); cd_fsupp.release(); cnsZ.release(); cnsZ2.release(); cnsG.release(); cns_st.release();
        mk_W.release(); mk_r.release(); mk_wr.release(); mk_w2r.release(); mk_f.release(); mk_grad.release(); mk_sx.release();
#ifndef SPCSC_EMU
#endif
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        wscratch.release();
        for (size_t i = 1; i < wstreams.size(); ++i) cudaStreamDestroy(wstreams[i]);
        for (auto e : wjoin) cudaEventDestroy(e);
        if (wfork) cudaEventDestroy(wfork);
        for (auto e : prof_ev) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }

    int init() {
        CK(cudaSetDevice(pb.device));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ev0));
        CK(cudaEventCreate(&ev1));
        CK(Y.ensure(nreal));
        CK(U.ensure(nreal));
        CK(Zt.ensure(nslab));
        CK(Df.ensure((size_t)Cd * N1f * M * N0));
        CK(Sf.ensure((size_t)K * C * N1f * N0));
        CK(G.ensure((size_t)N1f * N0 * Cd * Cd));
        CK(acc.ensure(kAccBytes / sizeof(double)));
        CK(st.ensure(1));
        CK(tw_row.ensure(N1));
        CK(tw_col.ensure(N0));
        auto wr = twiddles<T>(N1), wc = twiddles<T>(N0);
        CK(cudaMemcpyAsync(tw_row.p, wr.data(), N1 * sizeof(C2<T>), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(tw_col.p, wc.data(), N0 * sizeof(C2<T>), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        // kernel set v2 eligibility (SPCSC_KERNELS=v1 forces the general kernels)
        {
            const char* force = getenv("SPCSC_KERNELS");
            const bool allow2 = !(force && std::string(force) == "v1");
            v2_rowf = allow2 && !gen_rows && row2_ok<T>(H, N0, 1);
            v2_rowp = allow2 && !gen_rows && row2_ok<T>(H, N0, Cx);
            v2_col = allow2 && !gen_cols && col2_ok<T>(N0, M, Cd);
            const char* fz = getenv("SPCSC_FUSE");
            fuse = v2_rowf && v2_rowp && !(fz && std::string(fz) == "0");
            if (const char* wv = getenv("SPCSC_WAVE")) {
                int g = 0, ns = 1;
                if (sscanf(wv, "%d,%d", &g, &ns) >= 1 && g > 0) {
                    wave_g = g < K ? g : K;
                    wave_s = ns < 1 ? 1 : (ns > 8 ? 8 : ns);
                }
            }
            if (const char* wk = getenv("SPCSC_WAVE_KEEP")) wave_keep = atoi(wk) != 0;
            if (const char* wf = getenv("SPCSC_WAVE_FUSED")) wave_fused = atoi(wf) != 0;
            int rc;
            if (v2_rowf && (rc = upload_stage_tw(stw_row1, H, row2_elems(H, 1, (int)sizeof(T))))) return rc;
            if (v2_rowp && (rc = upload_stage_tw(stw_rowc, H, row2_elems(H, Cx, (int)sizeof(T))))) return rc;
            if (v2_col && (rc = upload_stage_tw(stw_col, N0, col2_elems<T>()))) return rc;
        }
        // default weights: scalar 1
        const T one = 1;
        CK(wl1_buf.ensure(1));
        CK(wl21_buf.ensure(1));
        CK(cudaMemcpyAsync(wl1_buf.p, &one, sizeof(T), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(wl21_buf.p, &one, sizeof(T), cudaMemcpyHostToDevice, stream));
        wl1 = WeightView<T>{wl1_buf.p, 0, 0, 0, 0, 0, 1};
        wl21 = WeightView<T>{wl21_buf.p, 0, 0, 0, 0, 0, 1};
        CK(cudaStreamSynchronize(stream));
        return admm_reset(1.0);
    }

    int upload_stage_tw(DevBuf<C2<T>>& buf, int n, int e) {
        auto tab = make_stage_twiddles<T>(n, e);
        CK(buf.ensure(tab.size()));
        CK(cudaMemcpyAsync(buf.p, tab.data(), tab.size() * sizeof(C2<T>), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }

    RowArgs<T> rowargs(int m, int nb, int cx) const {
        RowArgs<T> r;
        r.N0 = N0; r.M = m; r.nb = nb; r.Cx = cx;
        r.N1 = N1; r.gen = gen_rows ? 1 : 0;
        r.TR = gen_rows ? (N0 < 8 ? N0 : 8) : row_tile<T>(H, N0, cx);
        r.tw = tw_row.p;
        r.stream = stream;
        return r;
    }
    ColLaunch<T> colargs(int m, int nb) const {
        ColLaunch<T> c;
        memset(&c, 0, sizeof(c));
        c.Df = Df.p; c.Sf = Sf.p; c.G = G.p;
        c.tw = tw_col.p;
        c.nb = nb;
        c.a.N0 = N0; c.a.M = m; c.a.Cd = Cd; c.a.Cs = C; c.a.Cx = Cx; c.a.N1f = N1f;
        c.a.even_n1 = 1;
        c.stream = stream;
        c.Lstep = 1;
        c.gen = gen_cols ? 1 : 0;
        // persistent clusters + bulk-copy prefetch: measured no faster than one slab per cluster
        // (the kernel is issue-bound, not load-bound), so it stays opt-in
        c.bulk = (getenv("SPCSC_COLBULK") && atoi(getenv("SPCSC_COLBULK")) == 1) ? 1 : 0;
        // ADMM column kernel: 6 (default) k_col5 with one thread group -- slab, dictionary columns and signal row
        // staged by bulk copies, stage twiddles in registers; falls back to k_col2 (0) when the staged layout does
        // not fit or the dictionary has several channels.  1/2/3/11/12 k_col3 variants, 4 k_col4, 5 k_col5 with
        // two independent groups, 7 with bulk-copy stores
        c.push = getenv("SPCSC_COL3") ? atoi(getenv("SPCSC_COL3")) : 6;
        c.prefetch = getenv("SPCSC_COL3_PF") ? atoi(getenv("SPCSC_COL3_PF")) : 0;
        return c;
    }

    // rfft2 of a real array in device order [nb][m][N0][N1] into slab order [nb][N1f][m][N0]
    int forward2d(const T* real_in, C2<T>* out, int m, int nb) {
        ColLaunch<T> c = colargs(m, nb);
        c.in = out; c.out = out;
        c.a.Cd = 1;
        if (v2_rowf && v2_col && col2_ok<T>(N0, m, 1)) {        // register-plan kernels
            CK(row_fwd2<T>(H, rowargs(m, nb, 1), real_in, (const T*)nullptr,
                           (const AdmmState<T>*)nullptr, out, (const C2<T>*)stw_row1.p, 0));
            CK(col2<T>(N0, COL_FWD, c, (const C2<T>*)stw_col.p));
            return SPCSC_OK;
        }
        CK(row_fwd<T>(H, rowargs(m, nb, 1), real_in, (const T*)nullptr,
                      (const AdmmState<T>*)nullptr, out));
        CK(col<T>(N0, COL_FWD, c));
        return SPCSC_OK;
    }

    int set_dict(const void* D) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        CK(cudaSetDevice(pb.device));
        const size_t nd = (size_t)pb.hd * pb.wd * Cd * M;
        CK(staging.ensure(nd));
        CK(tmp_real.ensure((size_t)Cd * M * N0 * N1));
        CK(cudaMemcpyAsync(staging.p, D, nd * sizeof(T), cudaMemcpyHostToDevice, stream));
        CK(launch(k_pad_dict<T>, dim3(1024), dim3(256), 0, stream, (const T*)staging.p, tmp_real.p,
                  pb.hd, pb.wd, Cd, M, N0, N1));
        int rc = forward2d(tmp_real.p, Df.p, M, Cd);
        if (rc) return rc;
        CK(launch(k_gram<T>, dim3(256), dim3(128), 0, stream, (const C2<T>*)Df.p, G.p, N1f, N0, M, Cd));
        { int rg = install_ghg(); if (rg) return rg; }
        CK(cudaStreamSynchronize(stream));
        have_dict = true;
        return SPCSC_OK;
    }

    int to_internal(const void* host, T* dst, int c, int k, int m) {
        const size_t n = (size_t)N0 * N1 * c * k * m;
        CK(staging.ensure(n));
        CK(cudaMemcpyAsync(staging.p, host, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        dim3 grid((N0 * N1 + 31) / 32, (c * k * m + 31) / 32);
        CK(launch(k_to_internal<T>, grid, dim3(256), 0, stream, (const T*)staging.p, dst, N0 * N1, c, k, m));
        return SPCSC_OK;
    }
    int from_internal(const T* src, void* host, int c, int k, int m) {
        const size_t n = (size_t)N0 * N1 * c * k * m;
        CK(staging.ensure(n));
        dim3 grid((N0 * N1 + 31) / 32, (c * k * m + 31) / 32);
        CK(launch(k_from_internal<T>, grid, dim3(256), 0, stream, src, staging.p, N0 * N1, c, k, m));
        CK(cudaMemcpyAsync(host, staging.p, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }

    int set_signal(const void* S) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        CK(cudaSetDevice(pb.device));
        CK(tmp_real.ensure((size_t)K * C * N0 * N1));
        int rc = to_internal(S, tmp_real.p, C, K, 1);
        if (rc) return rc;
        rc = forward2d(tmp_real.p, Sf.p, 1, K * C);
        if (rc) return rc;
        CK(cudaStreamSynchronize(stream));
        have_signal = true;
        return SPCSC_OK;
    }

    int set_weight(DevBuf<T>& buf, WeightView<T>& view, const void* w, const int64_t d[5]) {
        const int64_t full[5] = {N0, N1, Cx, K, M};
        size_t n = 1;
        for (int i = 0; i < 5; ++i) {
            if (d[i] != 1 && d[i] != full[i]) FAIL(SPCSC_ERR_INVALID, "weight shape is not broadcastable to (N0,N1,Cx,K,M)");
            n *= (size_t)d[i];
        }
        CK(cudaSetDevice(pb.device));
        CK(cudaStreamSynchronize(stream));
        CK(buf.ensure(n));
        CK(cudaMemcpyAsync(buf.p, w, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        long long s[5];
        long long run = 1;
        for (int i = 4; i >= 0; --i) {
            s[i] = d[i] > 1 ? run : 0;
            run *= d[i];
        }
        view.p = buf.p;
        view.s0 = s[0]; view.s1 = s[1]; view.sc = s[2]; view.sk = s[3]; view.sm = s[4];
        view.spatial_uniform = (s[0] == 0 && s[1] == 0);
        return SPCSC_OK;
    }
    int set_l1_weight(const void* w, const int64_t* shape) override {
        return set_weight(wl1_buf, wl1, w, shape);
    }
    int set_l21_weight(const void* w, const int64_t* shape) override {
        const int64_t d[5] = {1, 1, 1, shape[0], shape[1]};
        return set_weight(wl21_buf, wl21, w, d);
    }

    int install_ghg() {
        if (!gradreg || Cd != 1) return SPCSC_OK;        // multi-channel dictionaries: k_gradreg_mc reads ghg_buf
        CK(launch(k_set_ghg<T>, dim3(128), dim3(256), 0, stream, G.p, (const T*)ghg_buf.p, (size_t)N1f * N0));
        return SPCSC_OK;
    }
    int set_gradreg(const void* ghg, const void* wgrd) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!ghg) { gradreg = false; return SPCSC_OK; }
        if (Cd > 4) FAIL(SPCSC_ERR_UNSUPPORTED, "gradient regularisation with more than 4 dictionary channels");
        if (!wgrd) FAIL(SPCSC_ERR_INVALID, "wgrd is NULL");
        CK(cudaSetDevice(pb.device));
        const T* g = (const T*)ghg;
        std::vector<T> t((size_t)N1f * N0);                 // (N0, N1f) -> [N1f][N0]
        for (int h = 0; h < N0; ++h)
            for (int wf = 0; wf < N1f; ++wf) t[(size_t)wf * N0 + h] = g[(size_t)h * N1f + wf];
        CK(ghg_buf.ensure(t.size()));
        CK(cudaMemcpyAsync(ghg_buf.p, t.data(), t.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        gr_w.assign((const T*)wgrd, (const T*)wgrd + M);
        gradreg = true;
        if (have_dict) return install_ghg();
        return SPCSC_OK;
    }

    int admm_configure(const spcsc_admm_opts* o) override {
        if (Cx > 4) FAIL(SPCSC_ERR_UNSUPPORTED, "more than 4 coefficient channels");
        if (Cd > 4) FAIL(SPCSC_ERR_UNSUPPORTED, "more than 4 dictionary channels");
        if (o->joint && Cd > 1) {
            // Cx == 1: the l2 norm over the channel axis is |.|, handled by the same kernel
        }
        opts = *o;
        prm.lmbda = (T)o->lmbda;
        prm.mu = (T)o->mu;
        prm.rlx = (T)o->rlx;
        prm.tau = (T)o->ar_scaling;
        prm.mur = (T)o->ar_rsdl_ratio;
        prm.xi = (T)o->ar_rsdl_target;
        prm.abs_tol = o->abs_tol;
        prm.rel_tol = o->rel_tol;
        prm.n_x = global_nx > 0.0 ? global_nx : (double)nreal;
        prm.inv_n = 1.0 / ((double)N0 * (double)N1);
        prm.autorho = o->ar_enabled;
        prm.period = o->ar_period > 0 ? o->ar_period : 1;
        prm.autoscaling = o->ar_autoscaling;
        prm.stdres = o->ar_std_residuals;
        prm.need_rsdl = (o->ar_enabled || !o->fast_solve) ? 1 : 0;
        prm.need_obj = o->fast_solve ? 0 : 1;
        prm.joint = o->joint;
        prm.enet = (o->l2_weight != 0.0) ? 1 : 0;
        prm.enet_mu = (T)o->l2_weight;
        prm.ams_m0 = M - (o->ams_maps > 0 ? o->ams_maps : 0);
        prm.gradreg = gradreg ? 1 : 0;
        prm.emit_policy = (getenv("SPCSC_EMIT") && atoi(getenv("SPCSC_EMIT")) == 0) ? 0 : 1;
        if (gradreg) {
            if (o->joint || o->l2_weight != 0.0) FAIL(SPCSC_ERR_UNSUPPORTED, "gradient regularisation with the joint or l2 penalty");
            if (Cd > 1 && (o->linsolve_check || (o->aux_var_obj && !o->fast_solve) || o->ams_maps > 0))
                FAIL(SPCSC_ERR_UNSUPPORTED, "gradient regularisation with a multi-channel dictionary: LinSolveCheck, "
                                            "AuxVarObj and AddMaskSim are not available");
            std::vector<C2<T>> gw(M);
            for (int m = 0; m < M; ++m) gw[m] = mk<T>((T)o->mu * gr_w[m], gr_w[m]);
            CK(gw_buf.ensure(M));
            CK(cudaMemcpyAsync(gw_buf.p, gw.data(), M * sizeof(C2<T>), cudaMemcpyHostToDevice, stream));
            CK(cudaStreamSynchronize(stream));
        }
        if (o->ams_maps < 0 || o->ams_maps >= M) FAIL(SPCSC_ERR_INVALID, "ams_maps out of range");
        if (o->ams_maps > 0 && o->joint) FAIL(SPCSC_ERR_UNSUPPORTED, "AddMaskSim with the joint penalty");
        prm.linsolve_check = o->linsolve_check;
        prm.dfid_direct = (o->aux_var_obj && !o->fast_solve) ? 1 : 0;
        configured = true;
        return SPCSC_OK;
    }

    int write_state(T rho, T udiv, int k, int stopped) {
        AdmmState<T> s;
        s.rho = rho; s.udiv = udiv; s.k = k; s.stopped = stopped; s.zt_stale = 1; s.emit = 0;
        CK(cudaMemcpyAsync(st.p, &s, sizeof(s), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }
    int read_state(AdmmState<T>& s) {
        CK(cudaMemcpyAsync(&s, st.p, sizeof(s), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }

    int admm_reset(double rho) override {
        CK(cudaSetDevice(pb.device));
        CK(cudaMemsetAsync(Y.p, 0, nreal * sizeof(T), stream));
        CK(cudaMemsetAsync(U.p, 0, nreal * sizeof(T), stream));
        CK(cudaMemsetAsync(acc.p, 0, kAccBytes, stream));
        have_x = false;
        return write_state((T)rho, (T)1, 0, 0);
    }
    int admm_set_rho(double rho) override {
        AdmmState<T> s;
        int rc = read_state(s);
        if (rc) return rc;
        return write_state((T)rho, s.udiv, s.k, s.stopped);
    }
    int admm_set_iter(int k) override {
        AdmmState<T> s;
        int rc = read_state(s);
        if (rc) return rc;
        if (k > 0) have_x = false;
        return write_state(s.rho, s.udiv, k, s.stopped);
    }
    int admm_get_scalars(double* rho, int* k) override {
        AdmmState<T> s;
        int rc = read_state(s);
        if (rc) return rc;
        if (rho) *rho = (double)s.rho;
        if (k) *k = s.k;
        return SPCSC_OK;
    }

    std::vector<int> prof_kind;                     // kind of the interval that ends at prof_ev[i] (-1: start)
    cudaError_t prof_mark(int kind) {
        const size_t i = prof_kind.size();
        while (prof_ev.size() <= i) {
            cudaEvent_t e;
            cudaError_t ce = cudaEventCreate(&e);
            if (ce != cudaSuccess) return ce;
            prof_ev.push_back(e);
        }
        prof_kind.push_back(kind);
        return cudaEventRecord(prof_ev[i], stream);
    }

    // Wavefront schedule (register-plan kernels): the batch is cut into groups of wave_g images; a group runs
    // row-forward -> column -> row-inverse/prox back to back through a scratch of one group's row spectra, which
    // therefore lives in L2: per iteration DRAM sees Y and U read twice and written once (6 B_r; 7 B_r when the
    // X spectra are kept for get_array(X)) instead of the 8.03 B_r of the cross-iteration-fused schedule, and a
    // change of rho costs nothing because the row spectra of Y - U are formed after the scalar kernel.  Groups
    // go round-robin over wave_s streams (each with its own scratch) so that the tail of one group's kernel
    // overlaps the next group's.
    int wave_setup() {
        const size_t gslab = (size_t)wave_g * Cx * N1f * M * N0;
        CK(wscratch.ensure(gslab * wave_s));
        if (wstreams.empty()) wstreams.push_back(stream);
        while ((int)wstreams.size() < wave_s) {
            cudaStream_t s2;
            CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            wstreams.push_back(s2);
        }
        while ((int)wjoin.size() < wave_s) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            wjoin.push_back(e);
        }
        if (!wfork) CK(cudaEventCreateWithFlags(&wfork, cudaEventDisableTiming));
#ifndef SPCSC_EMU
        if (!wave_persist_set) {
            wave_persist_set = true;
            const char* wp = getenv("SPCSC_WAVE_PERSIST");
            if (wp && atoi(wp) > 0) {
                // optional: pin each stream's scratch in the persisting part of L2
                cudaDeviceProp prop;
                CK(cudaGetDeviceProperties(&prop, pb.device));
                size_t want = gslab * sizeof(C2<T>) * wave_s;
                if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
                CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
                for (int si = 0; si < wave_s; ++si) {
                    cudaStreamAttrValue av;
                    memset(&av, 0, sizeof(av));
                    size_t bytes = gslab * sizeof(C2<T>);
                    if (bytes > (size_t)prop.accessPolicyMaxWindowSize) bytes = (size_t)prop.accessPolicyMaxWindowSize;
                    av.accessPolicyWindow.base_ptr = (void*)(wscratch.p + gslab * si);
                    av.accessPolicyWindow.num_bytes = bytes;
                    av.accessPolicyWindow.hitRatio = 1.0f;
                    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    CK(cudaStreamSetAttribute(wstreams[si], cudaStreamAttributeAccessPolicyWindow, &av));
                }
            }
        }
#endif
        return SPCSC_OK;
    }

    int launch_wave(int n, int k_base, bool prof, const ProxArgs<T>& pa0, const ColLaunch<T>& cs0) {
        int rc = wave_setup();
        if (rc) return rc;
        const int ns = prof ? 1 : wave_s;
        const size_t gslab = (size_t)wave_g * Cx * N1f * M * N0;
        const size_t img_r = (size_t)Cx * M * N0 * N1, img_z = (size_t)Cx * N1f * M * N0;
        const int ngroups = (K + wave_g - 1) / wave_g;
        C2<T>* zin = Zt.p;
        C2<T>* zoth = nullptr;
        if (wave_fused) {
            CK(Zt2.ensure(nslab));
            zoth = Zt2.p;
            fused_batch = true;
        } else {
            wave_batch = true;
        }
        if (prof) CK(prof_mark(-1));
        for (int it = 0; it < n; ++it) {
            if (wave_fused) {       // the spectra the prox kernel wrote are stale only when rho changed
                CK(row_fwd2<T>(H, rowargs(M, K * Cx, 1), (const T*)Y.p, (const T*)U.p, (const AdmmState<T>*)st.p,
                               zin, (const C2<T>*)stw_row1.p, 1));
                if (prof) CK(prof_mark(0));
                last_launches += 1;
            }
            if (ns > 1) {
                CK(cudaEventRecord(wfork, stream));
                for (int s = 1; s < ns; ++s) CK(cudaStreamWaitEvent(wstreams[s], wfork, 0));
            }
            for (int j = 0; j < ngroups; ++j) {
                const int k0 = j * wave_g, gk = (K - k0 < wave_g) ? (K - k0) : wave_g;
                const int si = j % ns;
                cudaStream_t sq = wstreams[si];
                C2<T>* zs = wscratch.p + gslab * si;
                T* yg = Y.p + img_r * k0;
                T* ug = U.p + img_r * k0;
                ColLaunch<T> cs = cs0;
                cs.stream = sq;
                cs.nb = gk * Cx;
                cs.Sf = Sf.p + (size_t)k0 * C * N1f * N0;
                ProxArgs<T> pa = pa0;
                pa.znext = nullptr;
                if (wave_fused) {
                    cs.in = zin + img_z * k0;
                    cs.out = zin + img_z * k0;
                    pa.znext = (void*)(zoth + img_z * k0);
                } else {
                    RowArgs<T> rf = rowargs(M, gk * Cx, 1);
                    rf.stream = sq;
                    CK(row_fwd2<T>(H, rf, (const T*)yg, (const T*)ug, (const AdmmState<T>*)st.p, zs,
                                   (const C2<T>*)stw_row1.p, 0));
                    if (prof) CK(prof_mark(0));
                    last_launches += 1;
                    cs.in = zs;
                    cs.out = wave_keep ? Zt.p + img_z * k0 : zs;
                }
                CK(col2<T>(N0, COL_ADMM, cs, (const C2<T>*)stw_col.p));
                last_col_kernel = g_col_variant;
                if (prof) CK(prof_mark(1));
                RowArgs<T> rp = rowargs(M, gk * Cx, Cx);
                rp.stream = sq;
                pa.wl1.p += (size_t)k0 * pa.wl1.sk;
                pa.wl21.p += (size_t)k0 * pa.wl21.sk;
                CK(row_inv_prox2<T>(H, rp, pa, (const C2<T>*)cs.out, yg, ug, (const AdmmState<T>*)st.p,
                                    (const C2<T>*)stw_rowc.p));
                if (prof) CK(prof_mark(2));
                last_launches += 2;
            }
            if (ns > 1) {
                for (int s = 1; s < ns; ++s) {
                    CK(cudaEventRecord(wjoin[s], wstreams[s]));
                    CK(cudaStreamWaitEvent(stream, wjoin[s], 0));
                }
            }
            rc = launch_scalars(k_base, n);
            if (rc) return rc;
            last_launches += 1;
            if (prof) CK(prof_mark(3));
            if (wave_fused) std::swap(zin, zoth);
        }
        return SPCSC_OK;
    }

    // sum the ACC_N accumulators over the ranks (PGM trial sums, policy scalars, dictionary-update terms)
    int reduce_acc_over_ranks() {
        if (!nccl_comm) return SPCSC_OK;
        if (p2p_on) {
            CK(launch(k_p2p_allreduce_acc<0>, dim3(1), dim3(256), 0, stream, p2p, acc.p, &st.p->stopped));
        } else {
            int nr = nccl->AllReduce(acc.p, acc.p, ACC_N, /*ncclDouble*/ 8, /*ncclSum*/ 0, nccl_comm, stream);
            if (nr != 0) {
                err = std::string("ncclAllReduce: ") + nccl->GetErrorString(nr);
                poisoned = true;
                return SPCSC_ERR_NCCL;
            }
        }
        return SPCSC_OK;
    }

    // the exchange of the residual / objective sums (multi-GPU) and the scalar kernel
    int launch_scalars(int k_base, int n) {
        P2pView pv{};
        if (p2p_on && (prm.need_rsdl || prm.need_obj)) {
            pv = p2p;
        } else if (nccl_comm && (prm.need_rsdl || prm.need_obj)) {
            // the one exchange of the path: sum the residual / objective accumulators over ranks
            CK(launch(k_fold_bins<0>, dim3(1), dim3(256), 0, stream, acc.p));
            int nr = nccl->AllReduce(acc.p, acc.p, ACC_N, /*ncclDouble*/ 8, /*ncclSum*/ 0, nccl_comm, stream);
            if (nr != 0) {
                err = std::string("ncclAllReduce: ") + nccl->GetErrorString(nr);
                poisoned = true;
                return SPCSC_ERR_NCCL;
            }
        }
        CK(launch(k_admm_scalars<T>, dim3(1), dim3(256), 0, stream, st.p, prm, acc.p, rows.p, k_base, n, pv));
        return SPCSC_OK;
    }

    // Launch n iterations on the stream.  With `prof` set, an event is recorded after every
    // kernel (prof_ev holds 4*n+1 events).
    int launch_iterations(int n, int k_base, bool prof) {
        const bool check = opts.linsolve_check != 0;
        if (check) {
            CK(Zscratch.ensure(nslab));
            CK(Xscratch.ensure(nslab));
        }
        RowArgs<T> rf = rowargs(M, K * Cx, 1);
        RowArgs<T> rp = rowargs(M, K * Cx, Cx);
        ProxArgs<T> pa;
        pa.prm = prm;
        pa.wl1 = wl1;
        pa.wl21 = wl21;
        pa.acc = acc.p;
        pa.scale = (T)(1.0 / ((double)N0 * (double)N1));
        pa.nonneg = opts.nonneg;
        pa.bnd0 = N0;
        pa.bnd1 = N1;
        if (opts.no_bndry_cross) {       // Y[1-hd:, :] = 0, Y[:, 1-wd:] = 0 (a support of 1 zeroes everything)
            pa.bnd0 = pb.hd == 1 ? 0 : N0 - (pb.hd - 1);
            pa.bnd1 = pb.wd == 1 ? 0 : N1 - (pb.wd - 1);
        }
        pa.reg_on_y = opts.aux_var_obj;
        pa.prox_threads = (getenv("SPCSC_PROX_NT") && atoi(getenv("SPCSC_PROX_NT")) == 256) ? 256 : 128;
        pa.use_v2_sync = (getenv("SPCSC_ROWPROX") && std::string(getenv("SPCSC_ROWPROX")) == "sync") ? 1 : 0;
        ColLaunch<T> cs = colargs(M, K * Cx);
        cs.st = st.p;
        cs.acc = acc.p;
        cs.Lstep = prm.enet ? prm.enet_mu : (T)0;        // SOLVE 1 reads the elastic-net weight here
        cs.a.gradreg = prm.gradreg;
        if (prm.gradreg) cs.sumin = gw_buf.p;
        cs.a.dfid_on = (prm.need_obj && !prm.dfid_direct) ? 1 : 0;
        const bool aux_eval = prm.need_obj && prm.dfid_direct;
        if (aux_eval) CK(Zscratch.ensure(nslab));
        last_launches = 0;
        C2<T>* zin = Zt.p;
        C2<T>* zoth = nullptr;
        fused_batch = false;
        wave_batch = false;
        prof_kind.clear();
        const bool wave = wave_g > 0 && v2_rowf && v2_rowp && v2_col && !check && !aux_eval && !prm.enet &&
                          !prm.gradreg && !pa.use_v2_sync && (!opts.joint || wl21.spatial_uniform || Cx > 1) &&
                          (!wave_fused || fuse);
        if (wave) return launch_wave(n, k_base, prof, pa, cs);
        if (prof) CK(prof_mark(-1));
        for (int it = 0; it < n; ++it) {
            const bool fuse_now = fuse && !check && !pa.use_v2_sync &&
                                  (!opts.joint || wl21.spatial_uniform || Cx > 1);
            if (fuse_now && !zoth) {
                CK(Zt2.ensure(nslab));
                zoth = Zt2.p;
                fused_batch = true;
            }
            if (v2_rowf)
                CK(row_fwd2<T>(H, rf, (const T*)Y.p, (const T*)U.p, (const AdmmState<T>*)st.p, zin,
                               (const C2<T>*)stw_row1.p, fuse_now ? 1 : 0));
            else
                CK(row_fwd<T>(H, rf, (const T*)Y.p, (const T*)U.p, (const AdmmState<T>*)st.p, zin));
            if (prof) CK(prof_mark(0));
            if (!check) {
                cs.in = zin; cs.out = zin;
                if (v2_col && !prm.enet && !prm.gradreg) {  // l2 / gradient terms: general kernel only
                    CK(col2<T>(N0, COL_ADMM, cs, (const C2<T>*)stw_col.p));
                    last_col_kernel = g_col_variant;
                } else if (prm.gradreg && Cd > 1) {
                    // ConvBPDNGradReg with a multi-channel dictionary: forward columns, the C x C solve with the
                    // diagonal mu w_m GHG + rho, inverse columns
                    ColLaunch<T> c1 = cs;
                    c1.a.gradreg = 0;
                    CK(col<T>(N0, COL_FWD, c1));
                    dim3 gg(N1f, (N0 + 31) / 32, K * Cx);
                    const int even = (N1 % 2 == 0) ? 1 : 0;
                    cudaError_t ge = cudaErrorInvalidValue;
                    if (Cd == 2) ge = launch(k_gradreg_mc<T, 2>, gg, dim3(256), 0, stream, zin, (const C2<T>*)Df.p, (const C2<T>*)Sf.p, (const T*)ghg_buf.p, (const C2<T>*)gw_buf.p, (const AdmmState<T>*)st.p, acc.p, N1f, M, N0, C, even, cs.a.dfid_on);
                    if (Cd == 3) ge = launch(k_gradreg_mc<T, 3>, gg, dim3(256), 0, stream, zin, (const C2<T>*)Df.p, (const C2<T>*)Sf.p, (const T*)ghg_buf.p, (const C2<T>*)gw_buf.p, (const AdmmState<T>*)st.p, acc.p, N1f, M, N0, C, even, cs.a.dfid_on);
                    if (Cd == 4) ge = launch(k_gradreg_mc<T, 4>, gg, dim3(256), 0, stream, zin, (const C2<T>*)Df.p, (const C2<T>*)Sf.p, (const T*)ghg_buf.p, (const C2<T>*)gw_buf.p, (const AdmmState<T>*)st.p, acc.p, N1f, M, N0, C, even, cs.a.dfid_on);
                    CK(ge);
                    CK(col<T>(N0, COL_INV, c1));
                    last_col_kernel = 0;
                } else {
                    CK(col<T>(N0, COL_ADMM, cs));
                    last_col_kernel = 0;
                }
                last_launches += 4;
            } else {
                ColLaunch<T> c1 = cs;
                c1.in = zin; c1.out = Zscratch.p; c1.a.Cd = Cd;
                CK(col<T>(N0, COL_FWD, c1));
                ColLaunch<T> c2 = cs;
                c2.in = Zscratch.p; c2.out = Xscratch.p;
                CK(col<T>(N0, COL_ADMM_NOFFT, c2));
                ColArgs ca = c2.a;
                CK(launch(k_linsolve_check<T>, dim3(N1f, K * Cx), dim3(128), 3 * 32 * sizeof(double), stream,
                          (const C2<T>*)Xscratch.p, (const C2<T>*)Zscratch.p, (const C2<T>*)Df.p,
                          (const C2<T>*)Sf.p, (const AdmmState<T>*)st.p, acc.p, ca, cs.Lstep,
                          (const C2<T>*)G.p, (const C2<T>*)gw_buf.p));
                ColLaunch<T> c3 = cs;
                c3.in = Xscratch.p; c3.out = zin;
                CK(col<T>(N0, COL_INV, c3));
                last_launches += 7;
            }
            if (prof) CK(prof_mark(1));
            pa.znext = fuse_now ? (void*)zoth : nullptr;
            if (v2_rowp)
                CK(row_inv_prox2<T>(H, rp, pa, (const C2<T>*)zin, Y.p, U.p,
                                    (const AdmmState<T>*)st.p, (const C2<T>*)stw_rowc.p));
            else
                CK(row_inv_prox<T>(H, rp, pa, (const C2<T>*)zin, Y.p, U.p, (const AdmmState<T>*)st.p));
            if (prof) CK(prof_mark(2));
            if (aux_eval) {
                // AuxVarObj (admm/cbpdn.py:315-344 with fEvalX False): data fidelity of the
                // auxiliary variable, 1/2 ||sum_m Df rfftn(Y) - Sf||^2, from a forward transform of Y
                CK(row_fwd<T>(H, rowargs(M, K * Cx, 1), (const T*)Y.p, (const T*)nullptr,
                              (const AdmmState<T>*)st.p, Zscratch.p));
                ColLaunch<T> ce = colargs(M, K * Cx);
                ce.in = Zscratch.p; ce.out = nullptr; ce.acc = acc.p; ce.st = st.p;
                ce.a.gradreg = prm.gradreg;
                if (prm.gradreg) ce.sumin = gw_buf.p;
                CK(col<T>(N0, COL_FWD_EVAL, ce));
                last_launches += 2;
            }
            { int rs = launch_scalars(k_base, n); if (rs) return rs; }
            if (prof) CK(prof_mark(3));
            if (fuse_now) std::swap(zin, zoth);        // the spectra just written feed the next x-step
        }
        return SPCSC_OK;
    }

    // After a fused batch of which `done` iterations actually executed (the device stops
    // launching work once the stopping test fires), point Zt at the buffer holding the next
    // x-step input and Zt2 at the one holding the last X row spectra.
    void settle_pingpong(int done) {
        if (!fused_batch) { x_in_zt2 = false; return; }
        C2<T>* p0 = Zt.p;
        C2<T>* p1 = Zt2.p;
        if (done % 2 == 1) { Zt.p = p1; Zt2.p = p0; }
        x_in_zt2 = done > 0 ? true : x_in_zt2;
        if (done == 0) { /* nothing ran: roles unchanged */ }
    }

    int admm_prepare(int n, AdmmState<T>& s0) {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!have_dict || !have_signal || !configured)
            FAIL(SPCSC_ERR_STATE, "admm_iterate before set_dict / set_signal / admm_configure");
        if (n <= 0) FAIL(SPCSC_ERR_INVALID, "n_iter must be positive");
        CK(cudaSetDevice(pb.device));
        int rc = read_state(s0);
        if (rc) return rc;
        if (s0.stopped) {
            rc = write_state(s0.rho, s0.udiv, s0.k, 0);
            if (rc) return rc;
        }
        CK(rows.ensure((size_t)n));
        return SPCSC_OK;
    }

    int admm_iterate(int n, spcsc_itstat* out_rows, int* n_done, int* stopped) override {
        AdmmState<T> s0;
        int rc = admm_prepare(n, s0);
        if (rc) return rc;
        CK(cudaEventRecord(ev0, stream));
        rc = launch_iterations(n, s0.k, false);
        if (rc) return rc;
        CK(cudaEventRecord(ev1, stream));
        AdmmState<T> s1;
        rc = read_state(s1);
        if (rc) return rc;
        CK(cudaEventElapsedTime(&last_ms, ev0, ev1));
        if (s1.stopped == 2) {
            err = "peer-memory all-reduce timed out: a rank of the group did not reach the exchange";
            poisoned = true;
            return SPCSC_ERR_NCCL;
        }
        const int done = s1.k - s0.k;
        have_x = have_x || done > 0;
        if (wave_batch && !wave_keep && done > 0) have_x = false;   // the X spectra only ever lived in the L2 scratch
        settle_pingpong(done);
        if (out_rows && done > 0) {
            std::vector<StatRow> hr(done);
            CK(cudaMemcpyAsync(hr.data(), rows.p, done * sizeof(StatRow), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            for (int i = 0; i < done; ++i) {
                spcsc_itstat& o = out_rows[i];
                o.iter = hr[i].k; o.objfun = hr[i].obj; o.dfid = hr[i].dfid; o.regl1 = hr[i].regl1;
                o.regl21 = hr[i].regl21; o.primal_rsdl = hr[i].r; o.dual_rsdl = hr[i].s;
                o.eps_primal = hr[i].epri; o.eps_dual = hr[i].edua; o.rho = hr[i].rho;
                o.xslv_relres = hr[i].xrrs; o.reserved = 0;
            }
        }
        if (n_done) *n_done = done;
        if (stopped) *stopped = s1.stopped;
        return SPCSC_OK;
    }

    int admm_last_timing(float* ms, int64_t* launches) override {
        if (ms) *ms = last_ms;
        if (launches) *launches = last_launches;
        return SPCSC_OK;
    }

    int admm_schedule_info(int32_t* info) override {
        info[0] = v2_rowf; info[1] = v2_col; info[2] = v2_rowp;
        info[3] = (fuse && !opts.linsolve_check && (!opts.joint || wl21.spatial_uniform || Cx > 1)) ? 1 : 0;
        if (wave_g > 0 && !wave_fused) info[3] = 0;
        info[4] = last_col_kernel;
        info[5] = wave_g; info[6] = wave_g > 0 ? wave_s : 0; info[7] = 0;
        return SPCSC_OK;
    }
    int admm_profile(int n, float* ms4) override {
        AdmmState<T> s0;
        int rc = admm_prepare(n, s0);
        if (rc) return rc;
        if (opts.linsolve_check) FAIL(SPCSC_ERR_INVALID, "profile with LinSolveCheck off");
        rc = launch_iterations(n, s0.k, true);
        if (rc) return rc;
        CK(cudaStreamSynchronize(stream));
        for (int j = 0; j < 4; ++j) ms4[j] = 0.f;
        for (size_t i = 1; i < prof_kind.size(); ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, prof_ev[i - 1], prof_ev[i]));
            if (prof_kind[i] >= 0 && prof_kind[i] < 4) ms4[prof_kind[i]] += ms;
        }
        have_x = !(wave_batch && !wave_keep);
        {
            AdmmState<T> s1;
            rc = read_state(s1);
            if (rc) return rc;
            settle_pingpong(s1.k - s0.k);
        }
        return SPCSC_OK;
    }

    int fold_udiv() {
        CK(launch(k_apply_udiv<T>, dim3(1184), dim3(256), 0, stream, U.p, st.p, nreal));
        CK(launch(k_reset_udiv<T>, dim3(1), dim3(1), 0, stream, st.p));
        return SPCSC_OK;
    }

    int freq_out(const C2<T>* src, void* host, int Cc, int Kk, int Mm) {
        const size_t n = (size_t)N0 * N1f * Cc * Kk * Mm;
        CK(staging.ensure(2 * n));
        CK(launch(k_freq_to_ext<T>, dim3(1184), dim3(256), 0, stream, src,
                  reinterpret_cast<C2<T>*>(staging.p), N0, N1f, Cc, Kk, Mm));
        CK(cudaMemcpyAsync(host, staging.p, n * sizeof(C2<T>), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }

    int get_array(int which, void* out) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        CK(cudaSetDevice(pb.device));
        int rc;
        switch (which) {
            case SPCSC_ARR_Y: return from_internal(Y.p, out, Cx, K, M);
            case SPCSC_ARR_U:
                rc = fold_udiv();
                if (rc) return rc;
                return from_internal(U.p, out, Cx, K, M);
            case SPCSC_ARR_X:
            case SPCSC_ARR_XF: {
                if (!have_x) FAIL(SPCSC_ERR_STATE, "X is not defined before the first iteration");
                CK(tmp_real.ensure(nreal));
                CK(row_inv<T>(H, rowargs(M, K * Cx, 1), (const C2<T>*)(x_in_zt2 ? Zt2.p : Zt.p), tmp_real.p,
                              (T)(1.0 / ((double)N0 * (double)N1))));
                if (which == SPCSC_ARR_X) return from_internal(tmp_real.p, out, Cx, K, M);
                CK(Zscratch.ensure(nslab));
                rc = forward2d(tmp_real.p, Zscratch.p, M, K * Cx);
                if (rc) return rc;
                return freq_out(Zscratch.p, out, Cx, K, M);
            }
            case SPCSC_ARR_PGM_X:
                if (!pgm_ready) FAIL(SPCSC_ERR_STATE, "PGM state not initialised");
                return from_internal(Y.p, out, Cx, K, M);
            case SPCSC_ARR_PGM_XF:
                if (!pgm_ready) FAIL(SPCSC_ERR_STATE, "PGM state not initialised");
                return freq_out(pgm_have_cand ? Zt.p : pgA.p, out, Cx, K, M);
            case SPCSC_ARR_PGM_YF:
                if (!pgm_ready) FAIL(SPCSC_ERR_STATE, "PGM state not initialised");
                return freq_out(pgB.p, out, Cx, K, M);
            case SPCSC_ARR_DF:
                if (!have_dict) FAIL(SPCSC_ERR_STATE, "no dictionary set");
                return freq_out(Df.p, out, Cd, 1, M);
            case SPCSC_ARR_SF:
                if (!have_signal) FAIL(SPCSC_ERR_STATE, "no signal set");
                return freq_out(Sf.p, out, C, K, 1);
        }
        FAIL(SPCSC_ERR_INVALID, "unknown array id");
    }

    int set_array(int which, const void* in) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        CK(cudaSetDevice(pb.device));
        int rc;
        if (which == SPCSC_ARR_Y) {
            rc = to_internal(in, Y.p, Cx, K, M);
            if (rc) return rc;
            // the row spectra of Y - U that the last prox kernel wrote no longer match Y
            CK(launch(k_mark_stale<T>, dim3(1), dim3(1), 0, stream, st.p));
        } else if (which == SPCSC_ARR_U) {
            rc = to_internal(in, U.p, Cx, K, M);
            if (rc) return rc;
            CK(launch(k_reset_udiv<T>, dim3(1), dim3(1), 0, stream, st.p));
        } else {
            FAIL(SPCSC_ERR_INVALID, "only Y and U can be set");
        }
        if (rc) return rc;
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }

    int reconstruct(const void* X, void* out) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!have_dict) FAIL(SPCSC_ERR_STATE, "no dictionary set");
        CK(cudaSetDevice(pb.device));
        size_t need = nreal > (size_t)K * C * N0 * N1 ? nreal : (size_t)K * C * N0 * N1;
        CK(tmp_real.ensure(need));
        const T* src = Y.p;
        if (X) {
            int rc = to_internal(X, tmp_real.p, Cx, K, M);
            if (rc) return rc;
            src = tmp_real.p;
        }
        CK(Zscratch.ensure(nslab));
        CK(sum_buf.ensure((size_t)K * C * N1f * N0));
        CK(row_fwd<T>(H, rowargs(M, K * Cx, 1), src, (const T*)nullptr, (const AdmmState<T>*)nullptr,
                      Zscratch.p));
        ColLaunch<T> c = colargs(M, K * Cx);
        c.in = Zscratch.p; c.out = nullptr; c.sumout = sum_buf.p;
        CK(col<T>(N0, COL_FWD_SUM, c));
        // sum_buf is [K*Cx][Cd][N1f][N0] == [K][C][N1f][1][N0]: slabs with one column each
        ColLaunch<T> ci = colargs(1, K * C);
        ci.in = sum_buf.p; ci.out = sum_buf.p; ci.a.Cd = 1;
        CK(col<T>(N0, COL_INV, ci));
        T* rec = tmp_real.p;       // stream order: the forward pass has consumed tmp_real by now
        CK(row_inv<T>(H, rowargs(1, K * C, 1), (const C2<T>*)sum_buf.p, rec,
                      (T)(1.0 / ((double)N0 * (double)N1))));
        return from_internal(rec, out, C, K, 1);
    }

    int p2p_export(void* handle64) override {
#ifdef SPCSC_EMU
        (void)handle64;
        FAIL(SPCSC_ERR_UNSUPPORTED, "peer-memory all-reduce needs real devices");
#else
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!comm_obj) FAIL(SPCSC_ERR_STATE, "p2p_export before attach_comm");
        CK(cudaSetDevice(pb.device));
        // failures here are not fatal (the NCCL all-reduce stays in use): report, do not poison the handle
        if (!comm_obj->p2p_own) {
            void* q = nullptr;
            cudaError_t e = cudaMalloc(&q, sizeof(P2pBlock));
            if (e == cudaSuccess) e = cudaMemset(q, 0, sizeof(P2pBlock));
            if (e != cudaSuccess) {
                cudaGetLastError();
                if (q) cudaFree(q);
                err = std::string("peer block allocation: ") + cudaGetErrorString(e);
                return SPCSC_ERR_UNSUPPORTED;
            }
            comm_obj->p2p_own = q;
        }
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t hnd;
        cudaError_t e = cudaIpcGetMemHandle(&hnd, comm_obj->p2p_own);
        if (e != cudaSuccess) {
            cudaGetLastError();
            err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
            return SPCSC_ERR_UNSUPPORTED;
        }
        memcpy(handle64, &hnd, 64);
        return SPCSC_OK;
#endif
    }
    // handles64 == NULL: use the mapping the communicator already holds (a solver attached earlier)
    int p2p_attach(int rank, int nr, const void* handles64) override {
#ifdef SPCSC_EMU
        (void)rank; (void)nr; (void)handles64;
        FAIL(SPCSC_ERR_UNSUPPORTED, "peer-memory all-reduce needs real devices");
#else
        if (poisoned) return SPCSC_ERR_CUDA;
        if (nr == 0) { p2p_on = false; return SPCSC_OK; }      // detach
        if (!comm_obj || !nccl_comm || nr != nranks) FAIL(SPCSC_ERR_STATE, "attach the communicator of the same group first");
        if (nr < 2 || nr > kP2pMaxRanks || rank < 0 || rank >= nr) FAIL(SPCSC_ERR_INVALID, "rank / nranks out of range");
        CK(cudaSetDevice(pb.device));
        if (!comm_obj->p2p_ready) {
            if (!handles64) {
                err = "the communicator holds no peer mapping yet";
                return SPCSC_ERR_UNSUPPORTED;
            }
            if (!comm_obj->p2p_own) FAIL(SPCSC_ERR_STATE, "p2p_attach before p2p_export");
            for (int r = 0; r < nr; ++r) {
                if (r == rank) { comm_obj->p2p_peer[r] = comm_obj->p2p_own; continue; }
                if (comm_obj->p2p_peer[r]) continue;
                cudaIpcMemHandle_t hnd;
                memcpy(&hnd, (const char*)handles64 + 64 * (size_t)r, 64);
                void* q = nullptr;
                cudaError_t e = cudaIpcOpenMemHandle(&q, hnd, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {          // not fatal: the NCCL path stays in use
                    cudaGetLastError();
                    err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
                    return SPCSC_ERR_UNSUPPORTED;
                }
                comm_obj->p2p_peer[r] = q;
            }
            comm_obj->p2p_ready = true;
        }
        P2pView v{};
        v.nranks = nr; v.rank = rank;
        // the exchange counters live in the block itself (zeroed with it at allocation) and advance only when an
        // exchange really executes -- launches skipped after the stopping test fired do not count, and flags
        // left in the peers' blocks by earlier solvers never match a new number
        for (int r = 0; r < nr; ++r) v.peer[r] = reinterpret_cast<P2pBlock*>(comm_obj->p2p_peer[r]);
        p2p = v;
        p2p_on = true;
        return SPCSC_OK;
#endif
    }
    int attach_comm(spcsc_comm* c, double gnx) override {
        if (c && c->device != pb.device) FAIL(SPCSC_ERR_INVALID, "communicator belongs to another device");
        comm_obj = c;
        if (!c) p2p_on = false;
        nccl = c ? c->api : nullptr;
        nccl_comm = c ? c->comm : nullptr;
        nranks = c ? c->nranks : 1;
        global_nx = c ? gnx : 0.0;
        prm.n_x = global_nx > 0.0 ? global_nx : (double)nreal;
        return SPCSC_OK;
    }

    // ---- dictionary update (CCMOD by PGM) --------------------------------------------------
    int ccmod_check() {
        if (poisoned) return SPCSC_ERR_CUDA;
        // Cd == 1: the channels of the signal act as further images (pgm/ccmod.py:232-237);
        // Cd == C > 1: every dictionary channel has its own gradient against the shared coefficients
        if (M > 128) FAIL(SPCSC_ERR_UNSUPPORTED, "dictionary update: more than 128 filters");
        if (!have_signal) FAIL(SPCSC_ERR_STATE, "dictionary update before set_signal");
        return SPCSC_OK;
    }
    int ccmod_reset(const void* D0, int zero_mean) override {
        int rc = ccmod_check();
        if (rc) return rc;
        CK(cudaSetDevice(pb.device));
        const size_t nd = (size_t)pb.hd * pb.wd * Cd * M, nsp = (size_t)Cd * N1f * M * N0;
        CK(staging.ensure(nd));
        CK(cdX.ensure((size_t)Cd * M * N0 * N1));
        CK(cdXf.ensure(nsp)); CK(cdYf.ensure(nsp)); CK(cdV.ensure(nsp)); CK(cdG.ensure(nsp));
        CK(cudaMemcpyAsync(staging.p, D0, nd * sizeof(T), cudaMemcpyHostToDevice, stream));
        CK(launch(k_pad_dict<T>, dim3(1024), dim3(256), 0, stream, (const T*)staging.p, cdX.p,
                  pb.hd, pb.wd, Cd, M, N0, N1));
        rc = forward2d(cdX.p, cdXf.p, M, Cd);
        if (rc) return rc;
        CK(cudaMemcpyAsync(cdYf.p, cdXf.p, nsp * sizeof(C2<T>), cudaMemcpyDeviceToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        cd_zero_mean = zero_mean;
        cd_ready = true;
        cd_grad_valid = false;
        cns_ready = false;
        return SPCSC_OK;
    }
    int ccmod_setcoef_device(int source) override {
        int rc = ccmod_check();
        if (rc) return rc;
        CK(cudaSetDevice(pb.device));
        CK(cdZf.ensure(nslab));
        if (source == SPCSC_COEF_ADMM_Y) {
            rc = forward2d(Y.p, cdZf.p, M, K * Cx);
            if (rc) return rc;
        } else if (source == SPCSC_COEF_PGM_X) {
            if (!pgA.p) FAIL(SPCSC_ERR_STATE, "no PGM iterate on this handle");
            CK(cudaMemcpyAsync(cdZf.p, pgA.p, nslab * sizeof(C2<T>), cudaMemcpyDeviceToDevice, stream));
        } else {
            FAIL(SPCSC_ERR_INVALID, "unknown coefficient source");
        }
        cd_have_coef = true;
        cd_grad_valid = false;
        cns_gram_stale = true;
        return SPCSC_OK;
    }
    int ccmod_setcoef(const void* Z) override {
        int rc = ccmod_check();
        if (rc) return rc;
        CK(cudaSetDevice(pb.device));
        CK(tmp_real.ensure(nreal));
        CK(cdZf.ensure(nslab));
        rc = to_internal(Z, tmp_real.p, Cx, K, M);
        if (rc) return rc;
        rc = forward2d(tmp_real.p, cdZf.p, M, K * Cx);
        if (rc) return rc;
        CK(cudaStreamSynchronize(stream));
        cd_have_coef = true;
        cd_grad_valid = false;
        cns_gram_stale = true;
        return SPCSC_OK;
    }
    template <bool GRAD>
    cudaError_t launch_grad(const C2<T>* yf, C2<T>* g) {
        dim3 grid(N1f, (N0 + 31) / 32);
        const int nimg = K * Cx;                                  // Cd == 1: (image, channel) pairs
        const size_t plane = (size_t)N1f * M * N0;
        for (int c = 0; c < Cd; ++c) {
            const C2<T>* yc = yf + (size_t)c * plane;
            C2<T>* gc = g ? g + (size_t)c * plane : nullptr;
            const int sfs = (Cd > 1) ? C : 1, sfc = (Cd > 1) ? c : 0;
            cudaError_t e;
            if (M <= 64)
                e = launch(k_ccmod_grad<T, 8, GRAD>, grid, dim3(256), 0, stream, (const C2<T>*)cdZf.p, yc,
                           (const C2<T>*)Sf.p, gc, acc.p, nimg, N1f, M, N0, (N1 % 2 == 0) ? 1 : 0, sfs, sfc);
            else
                e = launch(k_ccmod_grad<T, 16, GRAD>, grid, dim3(256), 0, stream, (const C2<T>*)cdZf.p, yc,
                           (const C2<T>*)Sf.p, gc, acc.p, nimg, N1f, M, N0, (N1 % 2 == 0) ? 1 : 0, sfs, sfc);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    // ---- one PGM iteration of the dictionary update = proximal trial(s) at step 1/L + acceptance.
    // Part A (sporco/pgm/pgm.py:779-811 PGMDFT.xstep with pgm/ccmod.py:295-318 grad_f): gradient at Yf (once per
    // iteration: kept in cdG across the trials of a backtracking search), Vf = Yf - g/L, V = irfftn(Vf), X = Pcn(V)
    // into cdX, Xf = rfftn(X) into cdV (the accepted iterate stays in cdXf).
    bool cd_grad_valid = false;
    double cd_fY = 0.0;                              // obfn_f(Yf) = sum |R(Yf)|^2 / 2 over the stored half spectrum (all ranks)
    int ccmod_part_a(double L) {
        const size_t nsp = (size_t)Cd * N1f * M * N0;
        const int nsupp = pb.hd * pb.wd * Cd * M;
        // Images sharded over ranks: the gradient is a sum over all of them.  Pcn only looks at the filter
        // supports and cropping is linear (cnvrep.py:953-981), so instead of all-reducing the 2 N0 N1f Cd M values of
        // the spectral gradient every rank forms V_r = irfftn(Yf / R - g_r / L), and only the hd x wd supports of
        // V = sum_r V_r are exchanged (hd wd Cd M values: 16 KB at 8x8x64) -- over the peer-memory block when the
        // ranks have mapped each other, else by one small NCCL call.
        const bool crop_reduce = nccl_comm && (p2p_on ? nsupp <= kP2pVecMax : true) &&
                                 !(getenv("SPCSC_CDL_FULLREDUCE") && atoi(getenv("SPCSC_CDL_FULLREDUCE")) == 1);
        if (!cd_grad_valid) {
            double hF = 0.0;
            CK(cudaMemsetAsync(acc.p + ACC_CDL_F, 0, 4 * sizeof(double), stream));
            CK(launch_grad<true>((const C2<T>*)cdYf.p, cdG.p));     // one pass over the coefficient spectra
            if (nccl_comm && !crop_reduce) {
                int nr = nccl->AllReduce(cdG.p, cdG.p, 2 * nsp, sizeof(T) == 4 ? 7 : 8, 0, nccl_comm, stream);
                if (nr != 0) { err = std::string("ncclAllReduce: ") + nccl->GetErrorString(nr); poisoned = true; return SPCSC_ERR_NCCL; }
            }
            if (nccl_comm) {                                         // f(Yf) of all images
                int rc = reduce_acc_over_ranks();
                if (rc) return rc;
            }
            CK(cudaMemcpyAsync(&hF, acc.p + ACC_CDL_F, sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemsetAsync(acc.p + ACC_CDL_F, 0, 4 * sizeof(double), stream));
            CK(cudaStreamSynchronize(stream));
            cd_fY = 0.5 * hF;
            cd_grad_valid = true;
        }
        CK(launch(k_ccmod_step<T>, dim3(592), dim3(256), 0, stream, (const C2<T>*)cdYf.p,
                  (const C2<T>*)cdG.p, cdV.p, (T)L, crop_reduce ? (T)(1.0 / (double)nranks) : (T)1, nsp));
        // V = irfftn(Vf): inverse columns, inverse rows
        ColLaunch<T> ci = colargs(M, Cd);
        ci.in = cdV.p; ci.out = cdV.p; ci.a.Cd = 1;
        CK(col<T>(N0, COL_INV, ci));
        CK(tmp_real.ensure((size_t)Cd * M * N0 * N1));
        CK(row_inv<T>(H, rowargs(M, Cd, 1), (const C2<T>*)cdV.p, tmp_real.p,
                      (T)(1.0 / ((double)N0 * (double)N1))));
        if (crop_reduce) {
            CK(cd_supp.ensure((size_t)nsupp + 4));
            CK(launch(k_support_copy<T>, dim3(64), dim3(256), 0, stream, tmp_real.p, cd_supp.p, Cd * M, N0, N1,
                      pb.hd, pb.wd, 0));
            if (p2p_on) {
                CK(launch(k_p2p_allreduce_vec<T>, dim3(1), dim3(1024), 0, stream, p2p, cd_supp.p, nsupp,
                          &st.p->stopped));
            } else {
                int nr = nccl->AllReduce(cd_supp.p, cd_supp.p, (size_t)nsupp, sizeof(T) == 4 ? 7 : 8, 0, nccl_comm, stream);
                if (nr != 0) { err = std::string("ncclAllReduce: ") + nccl->GetErrorString(nr); poisoned = true; return SPCSC_ERR_NCCL; }
            }
            CK(launch(k_support_copy<T>, dim3(64), dim3(256), 0, stream, tmp_real.p, cd_supp.p, Cd * M, N0, N1,
                      pb.hd, pb.wd, 1));
        }
        // X = Pcn(V) ; Xf = rfftn(X)  (into cdV, the old Xf stays in cdXf as Xfprv)
        CK(launch(k_pcn<T>, dim3(M), dim3(128), 0, stream, (const T*)tmp_real.p, cdX.p, acc.p, Cd, M,
                  N0, N1, pb.hd, pb.wd, cd_zero_mean, 0, (const int*)cd_fsupp.p));
        return forward2d(cdX.p, cdV.p, M, Cd);
    }
    // Part B (pgm/pgm.py:815-831 ystep, pgm/ccmod.py:341-376): residual against Yfprv, momentum step, the candidate
    // becomes the iterate, objective terms.
    int ccmod_part_b(double coef, int flags, double* out) {
        const size_t nsp = (size_t)Cd * N1f * M * N0;
        const int even = (N1 % 2 == 0) ? 1 : 0;
        CK(launch(k_spec_diffnorm<T>, dim3(296), dim3(256), 0, stream, (const C2<T>*)cdV.p,
                  (const C2<T>*)cdYf.p, acc.p, Cd, N1f, (size_t)M * N0, even));
        CK(launch(k_pgm_momentum<T>, dim3(592), dim3(256), 0, stream, (const C2<T>*)cdV.p,
                  (const C2<T>*)cdXf.p, cdYf.p, (T)coef, nsp));
        std::swap(cdXf.p, cdV.p);
        std::swap(cdXf.n, cdV.n);
        cd_grad_valid = false;
        // objective terms of the new iterate: data fidelity (second pass over Zf), constraint violation
        CK(cudaMemsetAsync(acc.p + ACC_CDL_F, 0, 2 * sizeof(double), stream));
        if (flags & SPCSC_CCMOD_DFID) CK(launch_grad<false>((const C2<T>*)cdXf.p, (C2<T>*)nullptr));
        if (nccl_comm && (flags & SPCSC_CCMOD_DFID)) {
            int rc = reduce_acc_over_ranks();       // the 16 accumulators over the peer block (or one small NCCL call)
            if (rc) return rc;
        }
        if (flags & SPCSC_CCMOD_CNSTR)
            CK(launch(k_pcn<T>, dim3(M), dim3(128), 0, stream, (const T*)cdX.p, (T*)nullptr, acc.p, Cd, M,
                      N0, N1, pb.hd, pb.wd, cd_zero_mean, 1, (const int*)cd_fsupp.p));
        double ha[4];
        CK(cudaMemcpyAsync(ha, acc.p + ACC_CDL_F, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        // slots ACC_CDL_* alias ADMM accumulators (ACC_AX2 ...: LinSolveCheck sums of the X step that shares
        // this handle): leave them clean
        CK(cudaMemsetAsync(acc.p + ACC_CDL_F, 0, 4 * sizeof(double), stream));
        CK(cudaStreamSynchronize(stream));
        const double inv_n = 1.0 / ((double)N0 * (double)N1);
        out[0] = 0.5 * ha[1] * inv_n;
        out[1] = std::sqrt(ha[3]);
        // the residual is computed from the (rank-identical) dictionary on every rank; the exchange above sums all
        // accumulator slots, this one included
        const bool rsdl_summed = nccl_comm && (flags & SPCSC_CCMOD_DFID);
        out[2] = ha[2] * inv_n / (rsdl_summed ? (double)nranks : 1.0);
        out[3] = cd_fY;
        return SPCSC_OK;
    }
    int ccmod_step_check(double L) {
        int rc = ccmod_check();
        if (rc) return rc;
        if (!cd_ready || !cd_have_coef) FAIL(SPCSC_ERR_STATE, "dictionary update step before ccmod_reset / setcoef");
        if (!(L > 0.0)) FAIL(SPCSC_ERR_INVALID, "L must be positive");
        CK(cudaSetDevice(pb.device));
        return SPCSC_OK;
    }
    int ccmod_step(double L, double coef, int flags, double* out) override {
        int rc = ccmod_step_check(L);
        if (rc) return rc;
        rc = ccmod_part_a(L);
        if (rc) return rc;
        return ccmod_part_b(coef, flags, out);
    }
    // One trial of a backtracking search (sporco/pgm/backtrack.py:74-107): out = { f(X) = obfn_f of the candidate,
    // f(Y), <grad f(Y), X - Y> (eval_linear_approx, pgm/pgm.py:886-894), ||X - Y||^2 } over the stored half spectra;
    // the host forms Q = f(Y) + <.,.> + (L/2) ||X - Y||^2 and decides.  spcsc_ccmod_accept then completes the iteration.
    int ccmod_trial(double L, double* out) override {
        int rc = ccmod_step_check(L);
        if (rc) return rc;
        rc = ccmod_part_a(L);
        if (rc) return rc;
        const size_t nsp = (size_t)Cd * N1f * M * N0;
        CK(cudaMemsetAsync(acc.p + ACC_CDL_F, 0, 4 * sizeof(double), stream));
        CK(launch_grad<false>((const C2<T>*)cdV.p, (C2<T>*)nullptr));             // f(X): pass over the coefficient spectra
        CK(launch(k_spec_lin<T>, dim3(296), dim3(256), 0, stream, (const C2<T>*)cdV.p, (const C2<T>*)cdYf.p,
                  (const C2<T>*)cdG.p, acc.p, nsp));
        if (nccl_comm) {
            rc = reduce_acc_over_ranks();
            if (rc) return rc;
        }
        double ha[4];
        CK(cudaMemcpyAsync(ha, acc.p + ACC_CDL_F, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemsetAsync(acc.p + ACC_CDL_F, 0, 4 * sizeof(double), stream));
        CK(cudaStreamSynchronize(stream));
        // with the gradient summed over ranks before use (no crop-before-reduce) the linear term is already global
        const bool crop_reduce = nccl_comm && (p2p_on ? pb.hd * pb.wd * Cd * M <= kP2pVecMax : true) &&
                                 !(getenv("SPCSC_CDL_FULLREDUCE") && atoi(getenv("SPCSC_CDL_FULLREDUCE")) == 1);
        const double rk = nccl_comm ? (double)nranks : 1.0;
        out[0] = 0.5 * ha[0];
        out[1] = cd_fY;
        out[2] = (nccl_comm && !crop_reduce) ? ha[2] / rk : ha[2];
        out[3] = ha[3] / rk;
        return SPCSC_OK;
    }
    int ccmod_accept(double coef, int flags, double* out) override {
        int rc = ccmod_check();
        if (rc) return rc;
        if (!cd_ready || !cd_grad_valid) FAIL(SPCSC_ERR_STATE, "ccmod_accept without a preceding ccmod_trial");
        CK(cudaSetDevice(pb.device));
        return ccmod_part_b(coef, flags, out);
    }
    // ---- consensus dictionary update (sporco/admm/ccmod.py:613-911): state and one iteration
    int ccmod_cns_init(double rho, int y0_given, long long nb_global) override {
        int rc = ccmod_check();
        if (rc) return rc;
        if (!cd_ready) FAIL(SPCSC_ERR_STATE, "ccmod_cns_init before ccmod_reset");
        if (Cd != 1 && Cd != C) FAIL(SPCSC_ERR_UNSUPPORTED, "dictionary channels must be 1 or the signal's");
        if (!(rho > 0.0)) FAIL(SPCSC_ERR_INVALID, "rho must be positive");
        CK(cudaSetDevice(pb.device));
        const int NB = K * C;
        const size_t plane = (size_t)M * N0 * N1, nb_real = (size_t)NB * plane;
        CK(cnsX.ensure(nb_real));
        CK(cnsU.ensure(nb_real));
        CK(cnsYn.ensure((size_t)Cd * plane));
        CK(cnsZ.ensure((size_t)NB * N1f * M * N0));
        CK(cnsG.ensure((size_t)K * Cx * N1f * N0));
        CK(cns_st.ensure(1));
        CK(cudaMemsetAsync(cnsYn.p, 0, (size_t)Cd * plane * sizeof(T), stream));
        if (y0_given)       // U_i = Y0 / rho for every block (ccmod.py:739-750)
            CK(launch(k_cns_uinit<T>, dim3(1184), dim3(256), 0, stream, (const T*)cdX.p, cnsU.p, NB, Cd, plane,
                      (T)(1.0 / rho)));
        else
            CK(cudaMemsetAsync(cnsU.p, 0, nb_real * sizeof(T), stream));
        CK(cudaStreamSynchronize(stream));
        cns_nb_global = nb_global > 0 ? nb_global : (long long)(NB / Cd);
        cns_ready = true;
        return SPCSC_OK;
    }
    // out: [0] DFid (on Y), [1] Cnstr (on Y), [2] ||X||^2, [3] ||X - Y||^2, [4] ||U||^2, [5] ||Y||^2, [6] ||Yprev - Y||^2
    int ccmod_cns_step(double rho, double udiv, double rlx, int flags, double* out) override {
        int rc = ccmod_check();
        if (rc) return rc;
        if (!cns_ready || !cd_have_coef) FAIL(SPCSC_ERR_STATE, "ccmod_cns_step before ccmod_cns_init / setcoef");
        if (!(rho > 0.0) || !(udiv > 0.0)) FAIL(SPCSC_ERR_INVALID, "rho and udiv must be positive");
        CK(cudaSetDevice(pb.device));
        const int NB = K * C;
        const size_t plane = (size_t)M * N0 * N1, nb_real = (size_t)NB * plane;
        const T uinv = (T)(1.0 / udiv), alpha = (T)rlx;
        if (cns_gram_stale) {       // g_i = sum_m |Zf_i,m|^2 for every block's coefficient spectra
            CK(launch(k_gram<T>, dim3(1024), dim3(128), 0, stream, (const C2<T>*)cdZf.p, cnsG.p, K * Cx * N1f, N0,
                      M, 1));
            cns_gram_stale = false;
        }
        AdmmState<T> hs;
        memset(&hs, 0, sizeof(hs));
        hs.rho = (T)rho; hs.udiv = (T)1; hs.k = 0; hs.stopped = 0; hs.zt_stale = 1; hs.emit = 0;
        CK(cudaMemcpyAsync(cns_st.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, stream));
        // x step: rfftn(Y - U_i), the column solve against block i's coefficient spectra, irfftn
        CK(launch(k_cns_yu<T>, dim3(148, NB), dim3(256), 0, stream, (const T*)cdX.p, (const T*)cnsU.p, cnsX.p, NB, Cd,
                  plane, uinv));
        const bool lscheck = (flags & SPCSC_CCMOD_LINSOLVE) != 0;
        // objective on the block variables (AuxVarObj False): the data fidelity falls out of the block solves
        // (sum_m Zf_i,m Xf_i,m - Sf_i = -rho q_i, as in ConvBPDN), the constraint term is taken on the block mean of X
        const bool obj_x = (flags & SPCSC_CCMOD_OBJ_X) != 0 && (flags & (SPCSC_CCMOD_DFID | SPCSC_CCMOD_CNSTR)) != 0;
        if (obj_x && nccl_comm) FAIL(SPCSC_ERR_UNSUPPORTED, "AuxVarObj False with the blocks sharded over ranks");
        if (obj_x) CK(cudaMemsetAsync(acc.p + ACC_DFID, 0, sizeof(double), stream));
        C2<T>* xf = cnsZ.p;
        ColLaunch<T> c = colargs(M, NB);
        c.in = cnsZ.p; c.out = cnsZ.p;
        c.Df = cdZf.p; c.G = cnsG.p; c.Sf = Sf.p;
        c.st = cns_st.p; c.acc = acc.p;
        c.a.Cd = 1; c.a.Cx = C; c.a.Cs = C;
        c.a.df_bstride = (long long)N1f * M * N0;
        c.a.g_bstride = (long long)N1f * N0;
        c.a.df_bdiv = Cd;
        c.push = 0; c.bulk = 0;
        c.Lstep = 0;
        c.a.dfid_on = obj_x ? 1 : 0;
        double hls[2] = {0.0, 0.0};
        if (lscheck) {
            // LinSolveCheck: forward columns, solve and inverse columns as separate launches of the general kernel, so
            // that the solve's input and output exist side by side in the 2-D frequency domain for the check
            const size_t nsp = (size_t)NB * N1f * M * N0;
            CK(cnsZ2.ensure(2 * nsp));
            C2<T>* zf2 = cnsZ2.p;
            C2<T>* xf2 = cnsZ2.p + nsp;
            if (v2_rowf)
                CK(row_fwd2<T>(H, rowargs(M, NB, 1), (const T*)cnsX.p, (const T*)nullptr,
                               (const AdmmState<T>*)nullptr, cnsZ.p, (const C2<T>*)stw_row1.p, 0));
            else
                CK(row_fwd<T>(H, rowargs(M, NB, 1), (const T*)cnsX.p, (const T*)nullptr,
                              (const AdmmState<T>*)nullptr, cnsZ.p));
            ColLaunch<T> c1 = c;
            c1.in = cnsZ.p; c1.out = zf2;
            CK(col<T>(N0, COL_FWD, c1));
            ColLaunch<T> c2 = c;
            c2.in = zf2; c2.out = xf2;
            CK(col<T>(N0, COL_ADMM_NOFFT, c2));
            CK(cudaMemsetAsync(acc.p + ACC_CNS_LSR, 0, 2 * sizeof(double), stream));
            dim3 grid(N1f, (N0 + 31) / 32);
            for (int cc = 0; cc < Cd; ++cc) {
                if (M <= 64)
                    CK(launch(k_cns_linsolve<T, 8>, grid, dim3(256), 0, stream, (const C2<T>*)cdZf.p, (const C2<T>*)xf2,
                              (const C2<T>*)zf2, (const C2<T>*)Sf.p, acc.p, NB, Cd, cc, N1f, M, N0, (T)rho));
                else
                    CK(launch(k_cns_linsolve<T, 16>, grid, dim3(256), 0, stream, (const C2<T>*)cdZf.p, (const C2<T>*)xf2,
                              (const C2<T>*)zf2, (const C2<T>*)Sf.p, acc.p, NB, Cd, cc, N1f, M, N0, (T)rho));
            }
            // (with the blocks sharded over ranks this is the residual of the rank's own blocks)
            CK(cudaMemcpyAsync(hls, acc.p + ACC_CNS_LSR, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemsetAsync(acc.p + ACC_CNS_LSR, 0, 2 * sizeof(double), stream));
            ColLaunch<T> c3 = c;
            c3.in = xf2; c3.out = cnsZ.p;
            CK(col<T>(N0, COL_INV, c3));
        } else if (v2_rowf && v2_col && col2_ok<T>(N0, M, 1)) {
            CK(row_fwd2<T>(H, rowargs(M, NB, 1), (const T*)cnsX.p, (const T*)nullptr,
                           (const AdmmState<T>*)nullptr, cnsZ.p, (const C2<T>*)stw_row1.p, 0));
            CK(col2<T>(N0, COL_ADMM, c, (const C2<T>*)stw_col.p));
        } else {
            CK(row_fwd<T>(H, rowargs(M, NB, 1), (const T*)cnsX.p, (const T*)nullptr,
                          (const AdmmState<T>*)nullptr, cnsZ.p));
            CK(col<T>(N0, COL_ADMM, c));
        }
        CK(row_inv<T>(H, rowargs(M, NB, 1), (const C2<T>*)xf, cnsX.p, (T)(1.0 / ((double)N0 * (double)N1))));
        // y step: supports of the mean over all blocks (of all ranks), then the constraint projection
        const double nb_glob = (double)cns_nb_global;
        CK(tmp_real.ensure((size_t)Cd * plane));
        CK(launch(k_cns_support_mean<T>, dim3(M, Cd), dim3(64), 0, stream, (const T*)cnsX.p, (const T*)cnsU.p,
                  (const T*)cdX.p, tmp_real.p, NB, Cd, M, N0, N1, pb.hd, pb.wd, alpha, uinv, (T)(1.0 / nb_glob),
                  (T)(1.0 / (double)nranks)));
        if (nccl_comm) {
            const int nsupp = pb.hd * pb.wd * Cd * M;
            CK(cd_supp.ensure((size_t)nsupp + 4));
            CK(launch(k_support_copy<T>, dim3(64), dim3(256), 0, stream, tmp_real.p, cd_supp.p, Cd * M, N0, N1,
                      pb.hd, pb.wd, 0));
            if (p2p_on && nsupp <= kP2pVecMax) {
                CK(launch(k_p2p_allreduce_vec<T>, dim3(1), dim3(1024), 0, stream, p2p, cd_supp.p, nsupp,
                          &st.p->stopped));
            } else {
                int nr = nccl->AllReduce(cd_supp.p, cd_supp.p, (size_t)nsupp, sizeof(T) == 4 ? 7 : 8, 0, nccl_comm, stream);
                if (nr != 0) { err = std::string("ncclAllReduce: ") + nccl->GetErrorString(nr); poisoned = true; return SPCSC_ERR_NCCL; }
            }
            CK(launch(k_support_copy<T>, dim3(64), dim3(256), 0, stream, tmp_real.p, cd_supp.p, Cd * M, N0, N1,
                      pb.hd, pb.wd, 1));
        }
        CK(launch(k_pcn<T>, dim3(M), dim3(128), 0, stream, (const T*)tmp_real.p, cnsYn.p, acc.p, Cd, M,
                  N0, N1, pb.hd, pb.wd, cd_zero_mean, 0, (const int*)cd_fsupp.p));
        // u step and the norms of the residuals
        CK(cudaMemsetAsync(acc.p + ACC_CNS_X2, 0, 5 * sizeof(double), stream));
        CK(launch(k_cns_update<T>, dim3(148, NB), dim3(256), 0, stream, (const T*)cnsX.p, cnsU.p, (const T*)cdX.p,
                  (const T*)cnsYn.p, acc.p, NB, Cd, plane, alpha, uinv));
        rc = reduce_acc_over_ranks();
        if (rc) return rc;
        CK(launch(k_cns_ynorms<T>, dim3(296), dim3(256), 0, stream, (const T*)cdX.p, (const T*)cnsYn.p, acc.p,
                  (size_t)Cd * plane));
        double hn[5];
        CK(cudaMemcpyAsync(hn, acc.p + ACC_CNS_X2, 5 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemsetAsync(acc.p + ACC_CNS_X2, 0, 5 * sizeof(double), stream));
        std::swap(cdX.p, cnsYn.p);
        std::swap(cdX.n, cnsYn.n);
        // the new dictionary's spectrum (hand-over to the X step, data fidelity)
        rc = forward2d(cdX.p, cdXf.p, M, Cd);
        if (rc) return rc;
        double ha[4] = {0.0, 0.0, 0.0, 0.0};
        double hq = 0.0;
        if (obj_x) {
            CK(cudaMemcpyAsync(&hq, acc.p + ACC_DFID, sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemsetAsync(acc.p + ACC_DFID, 0, sizeof(double), stream));
            CK(tmp_real.ensure((size_t)Cd * plane));
            CK(launch(k_cns_mean_x<T>, dim3(592), dim3(256), 0, stream, (const T*)cnsX.p, tmp_real.p, acc.p, NB, Cd, M, N0, N1,
                      pb.hd, pb.wd, (T)(1.0 / (double)(NB / Cd))));
            CK(launch(k_pcn<T>, dim3(M), dim3(128), 0, stream, (const T*)tmp_real.p, (T*)nullptr, acc.p, Cd, M,
                      N0, N1, pb.hd, pb.wd, cd_zero_mean, 1, (const int*)cd_fsupp.p));
            CK(cudaMemcpyAsync(ha, acc.p + ACC_CDL_F, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemsetAsync(acc.p + ACC_CDL_F, 0, 4 * sizeof(double), stream));
        } else if (flags & (SPCSC_CCMOD_DFID | SPCSC_CCMOD_CNSTR)) {
            if (flags & SPCSC_CCMOD_DFID) CK(launch_grad<false>((const C2<T>*)cdXf.p, (C2<T>*)nullptr));
            if (nccl_comm && (flags & SPCSC_CCMOD_DFID)) {
                rc = reduce_acc_over_ranks();
                if (rc) return rc;
            }
            if (flags & SPCSC_CCMOD_CNSTR)
                CK(launch(k_pcn<T>, dim3(M), dim3(128), 0, stream, (const T*)cdX.p, (T*)nullptr, acc.p, Cd, M,
                          N0, N1, pb.hd, pb.wd, cd_zero_mean, 1, (const int*)cd_fsupp.p));
            CK(cudaMemcpyAsync(ha, acc.p + ACC_CDL_F, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaMemsetAsync(acc.p + ACC_CDL_F, 0, 4 * sizeof(double), stream));
        }
        CK(cudaStreamSynchronize(stream));
        const double inv_n = 1.0 / ((double)N0 * (double)N1);
        out[0] = obj_x ? 0.5 * rho * rho * hq * inv_n : 0.5 * ha[1] * inv_n;
        out[1] = std::sqrt(ha[3]);
        out[2] = hn[0]; out[3] = hn[1]; out[4] = hn[2]; out[5] = hn[3]; out[6] = hn[4];
        out[7] = lscheck ? (hls[1] > 0.0 ? std::sqrt(hls[0] / hls[1]) : std::sqrt(hls[0])) : -1.0;
        return SPCSC_OK;
    }
    // spectra of the PGM dictionary update in device order [Cd][N1f][M][N0]: which 0 = Xf (iterate), 1 = Yf (momentum point)
    int ccmod_get_spectrum(int which, void* out) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!cd_ready) FAIL(SPCSC_ERR_STATE, "ccmod_get_spectrum before ccmod_reset");
        if (which != 0 && which != 1) FAIL(SPCSC_ERR_INVALID, "which must be 0 (Xf) or 1 (Yf)");
        CK(cudaSetDevice(pb.device));
        const size_t nsp = (size_t)Cd * N1f * M * N0;
        CK(cudaMemcpyAsync(out, which == 0 ? cdXf.p : cdYf.p, nsp * sizeof(C2<T>), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }
    // multi-scale dictionaries: the support (h_m, w_m) of every filter; NULL: all filters use the handle's hd x wd
    int ccmod_set_supports(const int32_t* hw) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        CK(cudaSetDevice(pb.device));
        if (!hw) { cd_fsupp.release(); return SPCSC_OK; }
        for (int m = 0; m < M; ++m)
            if (hw[2 * m] < 1 || hw[2 * m] > pb.hd || hw[2 * m + 1] < 1 || hw[2 * m + 1] > pb.wd)
                FAIL(SPCSC_ERR_INVALID, "filter support outside the handle's hd x wd");
        CK(cd_fsupp.ensure((size_t)2 * M));
        CK(cudaMemcpyAsync(cd_fsupp.p, hw, (size_t)2 * M * sizeof(int), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }
    // block variables in device order [K*C][M][N0][N1]: which 0 = X (after the last step), 1 = U
    int ccmod_cns_get(int which, void* out) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!cns_ready) FAIL(SPCSC_ERR_STATE, "ccmod_cns_get before ccmod_cns_init");
        if (which != 0 && which != 1) FAIL(SPCSC_ERR_INVALID, "which must be 0 (X) or 1 (U)");
        CK(cudaSetDevice(pb.device));
        const size_t n = (size_t)K * C * M * N0 * N1;
        CK(cudaMemcpyAsync(out, which == 0 ? cnsX.p : cnsU.p, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }
    int ccmod_get_dict(void* out) override {
        int rc = ccmod_check();
        if (rc) return rc;
        if (!cd_ready) FAIL(SPCSC_ERR_STATE, "ccmod_get_dict before ccmod_reset");
        CK(cudaSetDevice(pb.device));
        const size_t nd = (size_t)pb.hd * pb.wd * Cd * M;
        CK(staging.ensure(nd));
        CK(launch(k_crop_dict<T>, dim3(64), dim3(256), 0, stream, (const T*)cdX.p, staging.p, pb.hd,
                  pb.wd, Cd, M, N0, N1));
        CK(cudaMemcpyAsync(out, staging.p, nd * sizeof(T), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }
    int ccmod_push_dict() override {
        int rc = ccmod_check();
        if (rc) return rc;
        if (!cd_ready) FAIL(SPCSC_ERR_STATE, "ccmod_push_dict before ccmod_reset");
        CK(cudaSetDevice(pb.device));
        const size_t nsp = (size_t)Cd * N1f * M * N0;
        CK(cudaMemcpyAsync(Df.p, cdXf.p, nsp * sizeof(C2<T>), cudaMemcpyDeviceToDevice, stream));
        CK(launch(k_gram<T>, dim3(256), dim3(128), 0, stream, (const C2<T>*)Df.p, G.p, N1f, N0, M, Cd));
        { int rg = install_ghg(); if (rg) return rg; }
        have_dict = true;
        return SPCSC_OK;
    }

    // ---- PGM --------------------------------------------------------------------------
    int pgm_set_mask(const void* W, const int64_t* shape) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!W) { pgm_mask = false; return SPCSC_OK; }
        const int64_t full[4] = {N0, N1, C, K};
        for (int i = 0; i < 4; ++i)
            if (shape[i] != 1 && shape[i] != full[i]) FAIL(SPCSC_ERR_INVALID, "mask shape is not broadcastable to (N0,N1,C,K)");
        // expand on the host into device order [K][C][N0][N1] (set-up time, K*C planes)
        const T* w = (const T*)W;
        std::vector<T> t((size_t)K * C * N0 * N1);
        const size_t s3 = 1, s2 = (size_t)shape[3], s1 = s2 * shape[2], s0 = s1 * shape[1];
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < C; ++c)
                for (int y = 0; y < N0; ++y)
                    for (int x = 0; x < N1; ++x)
                        t[(((size_t)k * C + c) * N0 + y) * N1 + x] =
                            w[(shape[0] > 1 ? y : 0) * s0 + (shape[1] > 1 ? x : 0) * s1 +
                              (shape[2] > 1 ? c : 0) * s2 + (shape[3] > 1 ? k : 0) * s3];
        CK(cudaSetDevice(pb.device));
        const size_t nr = t.size(), nc = (size_t)K * C * N1f * N0;
        CK(mk_W.ensure(nr)); CK(mk_r.ensure(nr)); CK(mk_wr.ensure(nr)); CK(mk_w2r.ensure(nr));
        CK(mk_f.ensure(nc)); CK(mk_grad.ensure(nc)); CK(mk_sx.ensure(nc));
        CK(cudaMemcpyAsync(mk_W.p, t.data(), nr * sizeof(T), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        pgm_mask = true;
        return SPCSC_OK;
    }
    // signal-domain residual planes of a sum buffer: r = irfft2(s - Sf), [K*C][N0][N1]
    int mask_residual(const C2<T>* sums) {
        const int nb = K * C;
        const size_t nc = (size_t)nb * N1f * N0;
        CK(launch(k_spec_sub<T>, dim3(296), dim3(256), 0, stream, sums, (const C2<T>*)Sf.p, mk_f.p, nc));
        ColLaunch<T> c = colargs(1, nb);
        c.in = mk_f.p; c.out = mk_f.p; c.a.Cd = 1;
        CK(col<T>(N0, COL_INV, c));
        CK(row_inv<T>(H, rowargs(1, nb, 1), (const C2<T>*)mk_f.p, mk_r.p, (T)(1.0 / ((double)N0 * (double)N1))));
        return SPCSC_OK;
    }

    int pgm_configure(const spcsc_pgm_opts* o) override {
        popts = *o;
        return SPCSC_OK;
    }
    int pgm_reset(const void* X0) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!have_dict || !have_signal) FAIL(SPCSC_ERR_STATE, "pgm_reset before set_dict / set_signal");
        if (Cd > 4) FAIL(SPCSC_ERR_UNSUPPORTED, "more than 4 dictionary channels");
        CK(cudaSetDevice(pb.device));
        CK(pgA.ensure(nslab));
        CK(pgB.ensure(nslab));
        CK(sum_buf.ensure((size_t)K * Cx * Cd * N1f * N0));
        if (X0) {
            int rc = to_internal(X0, Y.p, Cx, K, M);
            if (rc) return rc;
            rc = forward2d(Y.p, pgA.p, M, K * Cx);
            if (rc) return rc;
            CK(cudaMemcpyAsync(pgB.p, pgA.p, nslab * sizeof(C2<T>), cudaMemcpyDeviceToDevice, stream));
        } else {
            CK(cudaMemsetAsync(Y.p, 0, nreal * sizeof(T), stream));
            CK(cudaMemsetAsync(pgA.p, 0, nslab * sizeof(C2<T>), stream));
            CK(cudaMemsetAsync(pgB.p, 0, nslab * sizeof(C2<T>), stream));
        }
        CK(cudaStreamSynchronize(stream));
        pgm_ready = true;
        pgm_have_cand = false;
        pg_z_init = false;
        pg_pol_init = false;
        return SPCSC_OK;
    }

    // ---- step-size policies, monotone and robust variants (pgm/stepsize.py, pgm/pgm.py:413-440 / 802-831,
    //      pgm/backtrack.py:110-210): a few passes over the state and scalars for the host's control flow
    template <int CD>
    cudaError_t policy_launch(int store) {
        return launch(k_pgm_policy<T, CD>, dim3(296), dim3(256), 0, stream, (const C2<T>*)sum_buf.p,
                      (const C2<T>*)pg_sx.p, (const C2<T>*)Sf.p, (const C2<T>*)G.p, pg_rprev.p, pg_sxprev.p, acc.p,
                      K * Cx, N1f, N0, 1 - (N1 & 1), store);
    }
    int pgm_policy_stats(int store, double* out) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!pgm_ready) FAIL(SPCSC_ERR_STATE, "pgm_policy_stats before pgm_reset");
        if (pgm_mask) FAIL(SPCSC_ERR_UNSUPPORTED, "step-size policies / monotone steps with a masked data fidelity");
        CK(cudaSetDevice(pb.device));
        const size_t ns = (size_t)K * Cx * Cd * N1f * N0;
        CK(sum_buf.ensure(ns));
        CK(pg_sx.ensure(ns));
        CK(pg_rprev.ensure(ns));
        CK(pg_sxprev.ensure(ns));
        if (!pg_pol_init) {
            // StepSizePolicyBB starts from xprv = gradprv = 0 (stepsize.py:108-109): grad = 0 means R = 0
            CK(cudaMemsetAsync(pg_rprev.p, 0, ns * sizeof(C2<T>), stream));
            CK(cudaMemsetAsync(pg_sxprev.p, 0, ns * sizeof(C2<T>), stream));
            pg_pol_init = true;
        }
        ColLaunch<T> cy = colargs(M, K * Cx);
        cy.in = pgB.p; cy.out = nullptr; cy.sumout = sum_buf.p;
        CK(col<T>(N0, COL_SUM, cy));
        ColLaunch<T> cx = colargs(M, K * Cx);
        cx.in = pgA.p; cx.out = nullptr; cx.sumout = pg_sx.p;
        CK(col<T>(N0, COL_SUM, cx));
        CK(cudaMemsetAsync(acc.p, 0, kAccBytes, stream));
        switch (Cd) {
            case 1: CK(policy_launch<1>(store)); break;
            case 2: CK(policy_launch<2>(store)); break;
            case 3: CK(policy_launch<3>(store)); break;
            case 4: CK(policy_launch<4>(store)); break;
            default: FAIL(SPCSC_ERR_UNSUPPORTED, "more than 4 dictionary channels");
        }
        CK(launch(k_l1_sum<T>, dim3(592), dim3(256), 0, stream, (const T*)Y.p, wl1, acc.p + 5, K, Cx, M, N0, N1));
        { int rr = reduce_acc_over_ranks(); if (rr) return rr; }
        double ha[6];
        CK(cudaMemcpyAsync(ha, acc.p, sizeof(ha), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (int i = 0; i < 4; ++i) out[i] = ha[i];
        out[4] = 0.5 * ha[4] / ((double)N0 * (double)N1);      // DFid of the accepted iterate
        out[5] = ha[5];                                         // RegL1 of the X in the real buffer
        out[6] = out[7] = 0.0;
        return SPCSC_OK;
    }
    int pgm_combine_y(double a, double b, int save_prev) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!pgm_ready) FAIL(SPCSC_ERR_STATE, "pgm_combine_y before pgm_reset");
        CK(cudaSetDevice(pb.device));
        CK(pgZ.ensure(nslab));
        CK(pgYp.ensure(nslab));
        if (!pg_z_init) {          // z_0 = x_0   (pgm/backtrack.py:163-164)
            CK(cudaMemcpyAsync(pgZ.p, pgA.p, nslab * sizeof(C2<T>), cudaMemcpyDeviceToDevice, stream));
            pg_z_init = true;
        }
        if (save_prev)             // Yfprv of the residual (pgm/pgm.py:838-841)
            CK(cudaMemcpyAsync(pgYp.p, pgB.p, nslab * sizeof(C2<T>), cudaMemcpyDeviceToDevice, stream));
        CK(launch(k_spec_axpby<T>, dim3(1184), dim3(256), 0, stream, (const C2<T>*)pgA.p, (const C2<T>*)pgZ.p,
                  pgB.p, (T)a, (T)b, nslab));
        return SPCSC_OK;
    }
    int pgm_finish(int mode, double c0, double* out) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!pgm_have_cand) FAIL(SPCSC_ERR_STATE, "pgm_finish without a candidate");
        CK(cudaSetDevice(pb.device));
        const int even = 1 - (N1 & 1);
        const double inv_n = 1.0 / ((double)N0 * (double)N1);
        double hv = 0.0;
        CK(cudaMemsetAsync(acc.p, 0, sizeof(double), stream));
        if (mode == SPCSC_PGM_FINISH_REJECT) {
            // monotone FISTA, objective went up: Xf stays, Yf = Xf + c0 (Zf - Xf); the residual is taken
            // between the kept Xf and the old Yf
            CK(launch(k_spec_wdist2<T>, dim3(592), dim3(256), 0, stream, (const C2<T>*)pgA.p, (const C2<T>*)pgB.p,
                      acc.p, K * Cx, N1f, (size_t)M * N0, even));
            CK(launch(k_spec_axpby<T>, dim3(1184), dim3(256), 0, stream, (const C2<T>*)pgA.p, (const C2<T>*)Zt.p,
                      pgB.p, (T)(1.0 - c0), (T)c0, nslab));
        } else if (mode == SPCSC_PGM_FINISH_ROBUST) {
            if (!pg_z_init) FAIL(SPCSC_ERR_STATE, "robust finish without pgm_combine_y");
            // z += c0 (x - y); x accepted; y untouched; residual against the previous iteration's y
            CK(launch(k_spec_add_diff<T>, dim3(1184), dim3(256), 0, stream, pgZ.p, (const C2<T>*)Zt.p,
                      (const C2<T>*)pgB.p, (T)c0, nslab));
            CK(launch(k_spec_wdist2<T>, dim3(592), dim3(256), 0, stream, (const C2<T>*)Zt.p, (const C2<T>*)pgYp.p,
                      acc.p, K * Cx, N1f, (size_t)M * N0, even));
            std::swap(Zt.p, pgA.p);
            std::swap(Zt.n, pgA.n);
        } else {
            FAIL(SPCSC_ERR_INVALID, "unknown finish mode");
        }
        if (nccl_comm) {
            CK(cudaMemsetAsync(acc.p + 1, 0, (ACC_N - 1) * sizeof(double), stream));
            int rr = reduce_acc_over_ranks();
            if (rr) return rr;
        }
        CK(cudaMemcpyAsync(&hv, acc.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (out) { out[0] = hv * inv_n; out[1] = 0.0; }
        pgm_have_cand = false;
        return SPCSC_OK;
    }
    int pgm_trial(double L, double* out) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!pgm_ready) FAIL(SPCSC_ERR_STATE, "pgm_trial before pgm_reset");
        if (!(L > 0.0)) FAIL(SPCSC_ERR_INVALID, "L must be positive");
        CK(cudaSetDevice(pb.device));
        CK(cudaMemsetAsync(acc.p, 0, kAccBytes, stream));
        // gradient step + inverse column transform: Zt = icol( Yf - conj(Df)(sum_m Df Yf - Sf)/L )
        ColLaunch<T> c1 = colargs(M, K * Cx);
        c1.in = pgB.p; c1.out = Zt.p; c1.sumout = sum_buf.p; c1.acc = acc.p; c1.Lstep = (T)L;
        const size_t mk_nr = (size_t)K * C * N0 * N1, mk_nc = (size_t)K * C * N1f * N0;
        if (pgm_mask) {
            // masked fidelity (pgm/cbpdn.py:461-506): the residual goes through the signal domain,
            //   grad = conj(Df) rfft(W^2 irfft(s_Y - Sf)),  F(Yf) = ||rfft(W irfft(s_Y - Sf))||^2 / 2
            ColLaunch<T> c0 = colargs(M, K * Cx);
            c0.in = pgB.p; c0.out = nullptr; c0.sumout = sum_buf.p;
            CK(col<T>(N0, COL_SUM, c0));
            int rc = mask_residual((const C2<T>*)sum_buf.p);
            if (rc) return rc;
            CK(launch(k_mask_mul<T>, dim3(296), dim3(256), 0, stream, (const T*)mk_r.p, (const T*)mk_W.p,
                      mk_wr.p, mk_w2r.p, (double*)nullptr, 0.0, mk_nr));
            rc = forward2d(mk_wr.p, mk_f.p, 1, K * C);
            if (rc) return rc;
            CK(launch(k_spec_sumsq<T>, dim3(296), dim3(256), 0, stream, (const C2<T>*)mk_f.p,
                      acc.p + ACC_PGM_FY, mk_nc));
            rc = forward2d(mk_w2r.p, mk_grad.p, 1, K * C);
            if (rc) return rc;
            c1.sumout = nullptr; c1.sumin = mk_grad.p; c1.a.pgm_mask = 1;
        }
        const bool pgm_v2 = v2_col && !pgm_mask;       // cluster column kernel (register plans)
        if (pgm_v2) CK(col2<T>(N0, COL_GRAD_INV, c1, (const C2<T>*)stw_col.p));
        else CK(col<T>(N0, COL_GRAD_INV, c1));
        // inverse rows, prox, forward rows
        PgmRowArgs<T> pr;
        pr.thr_scale = (T)popts.lmbda / (T)L;
        pr.wl1 = wl1;
        pr.acc = acc.p;
        pr.scale = (T)(1.0 / ((double)N0 * (double)N1));
        pr.nonneg = popts.nonneg;
        pr.stw = (v2_rowf && !pgm_mask) ? (const C2<T>*)stw_row1.p : nullptr;
        pr.bnd0 = N0; pr.bnd1 = N1;
        if (popts.no_bndry_cross) {
            pr.bnd0 = pb.hd == 1 ? 0 : N0 - (pb.hd - 1);
            pr.bnd1 = pb.wd == 1 ? 0 : N1 - (pb.wd - 1);
        }
        RowArgs<T> rr = rowargs(M, K * Cx, Cx);
        if (!gen_rows) rr.TR = row_tile<T>(H, N0, 1);
        CK(row_inv_prox_fwd<T>(H, rr, pr, Zt.p, Y.p));
        // forward columns + evaluation of the candidate against Yf
        ColLaunch<T> c3 = colargs(M, K * Cx);
        c3.in = Zt.p; c3.out = Zt.p; c3.sumin = sum_buf.p; c3.ref = pgB.p; c3.acc = acc.p;
        if (pgm_mask) { c3.a.pgm_mask = 1; c3.G = mk_grad.p; c3.sumout = mk_sx.p; }
        if (pgm_v2) CK(col2<T>(N0, COL_FWD_EVAL, c3, (const C2<T>*)stw_col.p));
        else CK(col<T>(N0, COL_FWD_EVAL, c3));
        if (pgm_mask) {
            // F(Xf) and DFid of the candidate, through the signal domain as well
            int rc = mask_residual((const C2<T>*)mk_sx.p);
            if (rc) return rc;
            CK(launch(k_mask_mul<T>, dim3(296), dim3(256), 0, stream, (const T*)mk_r.p, (const T*)mk_W.p,
                      mk_wr.p, (T*)nullptr, acc.p + ACC_DFID, (double)N0 * (double)N1, mk_nr));
            rc = forward2d(mk_wr.p, mk_f.p, 1, K * C);
            if (rc) return rc;
            CK(launch(k_spec_sumsq<T>, dim3(296), dim3(256), 0, stream, (const C2<T>*)mk_f.p,
                      acc.p + ACC_PGM_F, mk_nc));
        }
        { int rr = reduce_acc_over_ranks(); if (rr) return rr; }     // images sharded over ranks
        double hacc[ACC_N];
        CK(cudaMemcpyAsync(hacc, acc.p, sizeof(hacc), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        const double inv_n = 1.0 / ((double)N0 * (double)N1);
        out[SPCSC_PGM_F] = 0.5 * hacc[ACC_PGM_F];
        out[SPCSC_PGM_FY] = 0.5 * hacc[ACC_PGM_FY];
        out[SPCSC_PGM_LIN] = hacc[ACC_PGM_LIN];
        out[SPCSC_PGM_DXY2] = hacc[ACC_PGM_DXY2];
        out[SPCSC_PGM_RSDL] = hacc[ACC_PGM_RSDL] * inv_n;
        out[SPCSC_PGM_DFID] = 0.5 * hacc[ACC_DFID] * inv_n;
        out[SPCSC_PGM_REGL1] = hacc[ACC_L1];
        out[7] = 0.0;
        pgm_have_cand = true;
        return SPCSC_OK;
    }
    int pgm_accept(double coef) override {
        if (poisoned) return SPCSC_ERR_CUDA;
        if (!pgm_have_cand) FAIL(SPCSC_ERR_STATE, "pgm_accept without a candidate");
        CK(cudaSetDevice(pb.device));
        CK(launch(k_pgm_momentum<T>, dim3(1184), dim3(256), 0, stream, (const C2<T>*)Zt.p,
                  (const C2<T>*)pgA.p, pgB.p, (T)coef, nslab));
        std::swap(Zt.p, pgA.p);
        std::swap(Zt.n, pgA.n);
        pgm_have_cand = false;
        return SPCSC_OK;
    }

    int synchronize() override {
        CK(cudaSetDevice(pb.device));
        CK(cudaStreamSynchronize(stream));
        return SPCSC_OK;
    }
};

int check_problem(const spcsc_problem* p, std::string& err) {
    if (!p) { err = "null problem"; return SPCSC_ERR_INVALID; }
    if (p->dtype != SPCSC_F32 && p->dtype != SPCSC_F64) { err = "dtype must be SPCSC_F32 or SPCSC_F64"; return SPCSC_ERR_INVALID; }
    if (p->N0 < 1 || p->N1 < 1 || p->C < 1 || p->Cd < 1 || p->K < 1 || p->M < 1 || p->hd < 1 || p->wd < 1) {
        err = "non-positive dimension"; return SPCSC_ERR_INVALID;
    }
    if (p->Cd > 1 && p->Cd != p->C) { err = "multi-channel dictionary needs C == Cd"; return SPCSC_ERR_INVALID; }
    if (p->hd > p->N0 || p->wd > p->N1) { err = "filter support larger than the signal"; return SPCSC_ERR_INVALID; }
    if (p->N1 < 2 || p->N0 > 8192 || p->N1 > 8192) {
        err = "unsupported spatial size (need N1 >= 2 and both axes <= 8192)";
        return SPCSC_ERR_UNSUPPORTED;
    }
    return SPCSC_OK;
}

template <typename T>
int unit_fft2(bool inverse, int device, int batch, int N0, int N1, const void* in, void* out,
              std::string& err) {
    spcsc_problem p;
    memset(&p, 0, sizeof(p));
    p.N0 = N0; p.N1 = N1; p.C = 1; p.Cd = 1; p.K = batch; p.M = 1; p.hd = 1; p.wd = 1;
    p.dtype = sizeof(T) == 4 ? SPCSC_F32 : SPCSC_F64;
    p.device = device;
    int rc = check_problem(&p, err);
    if (rc) return rc;
    Engine<T> e(p);
    rc = e.init();
    if (rc) { err = e.err; return rc; }
    const int N1f = N1 / 2 + 1;
    const size_t nr = (size_t)batch * N0 * N1, nc = (size_t)batch * N0 * N1f;
    bool poisoned = false;
    DevBuf<C2<T>> sw;
    cudaError_t ce = sw.ensure(nc);
    if (ce != cudaSuccess) { err = cudaGetErrorString(ce); return SPCSC_ERR_NOMEM; }
#define UCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); sw.release(); (void)poisoned; return SPCSC_ERR_CUDA; } } while (0)
    if (!inverse) {
        UCK(e.tmp_real.ensure(nr));
        UCK(cudaMemcpyAsync(e.tmp_real.p, in, nr * sizeof(T), cudaMemcpyHostToDevice, e.stream));
        rc = e.forward2d(e.tmp_real.p, e.Zt.p, 1, batch);
        if (rc) { err = e.err; sw.release(); return rc; }
        // [batch][N1f][N0] -> [batch][N0][N1f]
        UCK(launch(k_swap_last2<T>, dim3(256, batch), dim3(256), 0, e.stream, (const C2<T>*)e.Zt.p, sw.p, N1f, N0));
        UCK(cudaMemcpyAsync(out, sw.p, nc * sizeof(C2<T>), cudaMemcpyDeviceToHost, e.stream));
        UCK(cudaStreamSynchronize(e.stream));
    } else {
        UCK(cudaMemcpyAsync(sw.p, in, nc * sizeof(C2<T>), cudaMemcpyHostToDevice, e.stream));
        UCK(launch(k_swap_last2<T>, dim3(256, batch), dim3(256), 0, e.stream, (const C2<T>*)sw.p, e.Zt.p, N0, N1f));
        ColLaunch<T> c = e.colargs(1, batch);
        c.in = e.Zt.p; c.out = e.Zt.p; c.a.Cd = 1;
        UCK(col<T>(N0, COL_INV, c));
        UCK(e.tmp_real.ensure(nr));
        UCK(row_inv<T>(N1 / 2, e.rowargs(1, batch, 1), (const C2<T>*)e.Zt.p, e.tmp_real.p,
                       (T)(1.0 / ((double)N0 * (double)N1))));
        UCK(cudaMemcpyAsync(out, e.tmp_real.p, nr * sizeof(T), cudaMemcpyDeviceToHost, e.stream));
        UCK(cudaStreamSynchronize(e.stream));
    }
#undef UCK
    sw.release();
    return SPCSC_OK;
}

// level-1 entry points on plain host arrays: stage, launch, copy back
struct TmpDev {
    void* p = nullptr;
    ~TmpDev() { if (p) cudaFree(p); }
    cudaError_t get(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};
#define LCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cudaGetLastError(); err = std::string(#call) + ": " + cudaGetErrorString(e_); return SPCSC_ERR_CUDA; } } while (0)
template <typename T>
int unit_solvedbi(int device, long long nf, int nk, int Cd, int M, double rho, const void* ah, const void* b,
                  void* x, std::string& err) {
    if (nf < 1 || nk < 1 || M < 1 || Cd < 1 || Cd > 4 || !(rho > 0.0)) { err = "bad argument (need nf, nk, M >= 1, 1 <= Cd <= 4, rho > 0)"; return SPCSC_ERR_INVALID; }
    LCK(cudaSetDevice(device));
    const size_t na = (size_t)nf * Cd * M * sizeof(C2<T>), nb = (size_t)nf * nk * M * sizeof(C2<T>);
    TmpDev da, db, dx;
    LCK(da.get(na)); LCK(db.get(nb)); LCK(dx.get(nb));
    LCK(cudaMemcpy(da.p, ah, na, cudaMemcpyHostToDevice));
    LCK(cudaMemcpy(db.p, b, nb, cudaMemcpyHostToDevice));
    long long warps = nf * nk;
    int blocks = (int)((warps + 7) / 8 < 1184 ? (warps + 7) / 8 : 1184);
    const C2<T>* pa = (const C2<T>*)da.p; const C2<T>* pb = (const C2<T>*)db.p; C2<T>* px = (C2<T>*)dx.p;
    switch (Cd) {
        case 1: LCK(launch(k_solvedbi<T, 1>, dim3(blocks), dim3(256), 0, (cudaStream_t)0, pa, pb, px, nf, nk, M, (T)rho)); break;
        case 2: LCK(launch(k_solvedbi<T, 2>, dim3(blocks), dim3(256), 0, (cudaStream_t)0, pa, pb, px, nf, nk, M, (T)rho)); break;
        case 3: LCK(launch(k_solvedbi<T, 3>, dim3(blocks), dim3(256), 0, (cudaStream_t)0, pa, pb, px, nf, nk, M, (T)rho)); break;
        default: LCK(launch(k_solvedbi<T, 4>, dim3(blocks), dim3(256), 0, (cudaStream_t)0, pa, pb, px, nf, nk, M, (T)rho)); break;
    }
    LCK(cudaMemcpy(x, dx.p, nb, cudaMemcpyDeviceToHost));
    return SPCSC_OK;
}
template <typename T>
int unit_prox(int device, long long n_outer, int C, long long n_inner, double alpha, double beta, int joint,
              const void* w, const void* v, void* out, std::string& err) {
    if (n_outer < 1 || C < 1 || n_inner < 1) { err = "bad argument"; return SPCSC_ERR_INVALID; }
    LCK(cudaSetDevice(device));
    const size_t nbytes = (size_t)n_outer * C * n_inner * sizeof(T);
    TmpDev dv, dw, dout;
    LCK(dv.get(nbytes)); LCK(dout.get(nbytes));
    LCK(cudaMemcpy(dv.p, v, nbytes, cudaMemcpyHostToDevice));
    if (w) { LCK(dw.get(nbytes)); LCK(cudaMemcpy(dw.p, w, nbytes, cudaMemcpyHostToDevice)); }
    LCK(launch(k_prox_l1l2<T>, dim3(592), dim3(256), 0, (cudaStream_t)0, (const T*)dv.p, (const T*)dw.p, (T*)dout.p,
               n_outer, C, n_inner, (T)alpha, (T)beta, joint));
    LCK(cudaMemcpy(out, dout.p, nbytes, cudaMemcpyDeviceToHost));
    return SPCSC_OK;
}
#undef LCK

template <typename T>
int unit_tikhonov(int device, int batch, int N0, int N1, double lmbda, int npd, const void* in,
                  void* sl, void* sh, std::string& err) {
    if (npd < 0 || npd > N0 || npd > N1) { err = "npd must lie in [0, min(N0, N1)]"; return SPCSC_ERR_INVALID; }
    const int P0 = N0 + 2 * npd, P1 = N1 + 2 * npd;
    spcsc_problem p;
    memset(&p, 0, sizeof(p));
    p.N0 = P0; p.N1 = P1; p.C = 1; p.Cd = 1; p.K = batch; p.M = 1; p.hd = 1; p.wd = 1;
    p.dtype = sizeof(T) == 4 ? SPCSC_F32 : SPCSC_F64;
    p.device = device;
    int rc = check_problem(&p, err);
    if (rc) return rc;
    Engine<T> e(p);
    rc = e.init();
    if (rc) { err = e.err; return rc; }
    const size_t nr = (size_t)batch * N0 * N1, np_ = (size_t)batch * P0 * P1;
    DevBuf<T> s_in, s_lo, s_hi;
#define UCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); s_in.release(); s_lo.release(); s_hi.release(); return SPCSC_ERR_CUDA; } } while (0)
    UCK(s_in.ensure(nr)); UCK(s_lo.ensure(nr)); UCK(s_hi.ensure(nr));
    UCK(e.tmp_real.ensure(np_));
    UCK(cudaMemcpyAsync(s_in.p, in, nr * sizeof(T), cudaMemcpyHostToDevice, e.stream));
    UCK(launch(k_pad_symmetric<T>, dim3(592), dim3(256), 0, e.stream, (const T*)s_in.p, e.tmp_real.p, batch,
               N0, N1, npd));
    rc = e.forward2d(e.tmp_real.p, e.Zt.p, 1, batch);
    if (rc) { err = e.err; s_in.release(); s_lo.release(); s_hi.release(); return rc; }
    UCK(launch(k_tikhonov_divide<T>, dim3(592), dim3(256), 0, e.stream, e.Zt.p, batch, P1 / 2 + 1, P0, P1,
               (T)lmbda));
    ColLaunch<T> c = e.colargs(1, batch);
    c.in = e.Zt.p; c.out = e.Zt.p; c.a.Cd = 1;
    UCK(col<T>(P0, COL_INV, c));
    UCK(row_inv<T>(P1 / 2, e.rowargs(1, batch, 1), (const C2<T>*)e.Zt.p, e.tmp_real.p,
                   (T)(1.0 / ((double)P0 * (double)P1))));
    UCK(launch(k_crop_split<T>, dim3(592), dim3(256), 0, e.stream, (const T*)e.tmp_real.p, (const T*)s_in.p,
               s_lo.p, s_hi.p, batch, N0, N1, npd));
    UCK(cudaMemcpyAsync(sl, s_lo.p, nr * sizeof(T), cudaMemcpyDeviceToHost, e.stream));
    UCK(cudaMemcpyAsync(sh, s_hi.p, nr * sizeof(T), cudaMemcpyDeviceToHost, e.stream));
    UCK(cudaStreamSynchronize(e.stream));
#undef UCK
    s_in.release(); s_lo.release(); s_hi.release();
    return SPCSC_OK;
}

}  // namespace

// =========================================================================================
extern "C" {

int spcsc_version(void) { return 100; }

int spcsc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int spcsc_device_name(int device, char* buf, int buflen) {
    if (!buf || buflen <= 0) return SPCSC_ERR_INVALID;
#ifdef SPCSC_EMU
    snprintf(buf, buflen, "spcsc CPU emulation (test only)");
    (void)device;
    return SPCSC_OK;
#else
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return SPCSC_ERR_CUDA; }
    snprintf(buf, buflen, "%s", prop.name);
    return SPCSC_OK;
#endif
}

int spcsc_memory_info(int device, uint64_t* free_bytes, uint64_t* total_bytes) {
    size_t f = 0, t = 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) {
        g_last_error = "cudaMemGetInfo failed";
        return SPCSC_ERR_CUDA;
    }
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return SPCSC_OK;
}

const char* spcsc_last_error(const spcsc_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int spcsc_create(const spcsc_problem* prob, spcsc_handle** out) {
    if (!out) { g_last_error = "null output pointer"; return SPCSC_ERR_INVALID; }
    *out = nullptr;
    int rc = check_problem(prob, g_last_error);
    if (rc) return rc;
    int ndev = spcsc_device_count();
    if (ndev <= 0) { g_last_error = "no CUDA device available"; return SPCSC_ERR_CUDA; }
    if (prob->device < 0 || prob->device >= ndev) { g_last_error = "device ordinal out of range"; return SPCSC_ERR_INVALID; }
    spcsc_handle* h = nullptr;
    int irc;
    if (prob->dtype == SPCSC_F32) {
        Engine<float>* e = new (std::nothrow) Engine<float>(*prob);
        if (!e) { g_last_error = "out of host memory"; return SPCSC_ERR_NOMEM; }
        irc = e->init();
        h = e;
    } else {
        Engine<double>* e = new (std::nothrow) Engine<double>(*prob);
        if (!e) { g_last_error = "out of host memory"; return SPCSC_ERR_NOMEM; }
        irc = e->init();
        h = e;
    }
    if (irc) {
        g_last_error = h->err;
        delete h;
        return irc;
    }
    *out = h;
    return SPCSC_OK;
}

int spcsc_destroy(spcsc_handle* h) {
    delete h;
    return SPCSC_OK;
}

#define H_CALL(expr)                                                   \
    if (!h) { g_last_error = "null handle"; return SPCSC_ERR_INVALID; } \
    return (expr)

int spcsc_synchronize(spcsc_handle* h) { H_CALL(h->synchronize()); }
int spcsc_pgm_configure(spcsc_handle* h, const spcsc_pgm_opts* o) { H_CALL(o ? h->pgm_configure(o) : SPCSC_ERR_INVALID); }
int spcsc_pgm_reset(spcsc_handle* h, const void* X0) { H_CALL(h->pgm_reset(X0)); }
int spcsc_pgm_trial(spcsc_handle* h, double L, double out[8]) { H_CALL(out ? h->pgm_trial(L, out) : SPCSC_ERR_INVALID); }
int spcsc_pgm_accept(spcsc_handle* h, double coef) { H_CALL(h->pgm_accept(coef)); }
int spcsc_pgm_policy_stats(spcsc_handle* h, int32_t store, double out[8]) { H_CALL(out ? h->pgm_policy_stats(store, out) : SPCSC_ERR_INVALID); }
int spcsc_pgm_combine_y(spcsc_handle* h, double a, double b, int32_t save_prev) { H_CALL(h->pgm_combine_y(a, b, save_prev)); }
int spcsc_pgm_finish(spcsc_handle* h, int32_t mode, double c0, double out[2]) { H_CALL(h->pgm_finish(mode, c0, out)); }
int spcsc_p2p_export(spcsc_handle* h, void* handle64) { H_CALL(handle64 ? h->p2p_export(handle64) : SPCSC_ERR_INVALID); }
int spcsc_p2p_attach(spcsc_handle* h, int32_t rank, int32_t nranks, const void* handles64) { H_CALL(h->p2p_attach(rank, nranks, handles64)); }
int spcsc_set_gradreg(spcsc_handle* h, const void* ghg, const void* wgrd) { H_CALL(h->set_gradreg(ghg, wgrd)); }
int spcsc_pgm_set_mask(spcsc_handle* h, const void* W, const int64_t shape[4]) { H_CALL((W && !shape) ? SPCSC_ERR_INVALID : h->pgm_set_mask(W, shape)); }
int spcsc_ccmod_reset(spcsc_handle* h, const void* D0, int32_t zm) { H_CALL(D0 ? h->ccmod_reset(D0, zm) : SPCSC_ERR_INVALID); }
int spcsc_ccmod_setcoef_device(spcsc_handle* h, int32_t source) { H_CALL(h->ccmod_setcoef_device(source)); }
int spcsc_ccmod_setcoef(spcsc_handle* h, const void* Z) { H_CALL(Z ? h->ccmod_setcoef(Z) : SPCSC_ERR_INVALID); }
int spcsc_ccmod_step(spcsc_handle* h, double L, double coef, int32_t flags, double out[4]) { H_CALL(out ? h->ccmod_step(L, coef, flags, out) : SPCSC_ERR_INVALID); }
int spcsc_ccmod_trial(spcsc_handle* h, double L, double out[4]) { H_CALL(out ? h->ccmod_trial(L, out) : SPCSC_ERR_INVALID); }
int spcsc_ccmod_accept(spcsc_handle* h, double coef, int32_t flags, double out[4]) { H_CALL(out ? h->ccmod_accept(coef, flags, out) : SPCSC_ERR_INVALID); }
int spcsc_ccmod_get_dict(spcsc_handle* h, void* D_out) { H_CALL(D_out ? h->ccmod_get_dict(D_out) : SPCSC_ERR_INVALID); }
int spcsc_ccmod_push_dict(spcsc_handle* h) { H_CALL(h->ccmod_push_dict()); }
int spcsc_ccmod_cns_init(spcsc_handle* h, double rho, int32_t y0_given, int64_t nb_global) { H_CALL(h->ccmod_cns_init(rho, y0_given, (long long)nb_global)); }
int spcsc_ccmod_get_spectrum(spcsc_handle* h, int32_t which, void* out) { H_CALL(out ? h->ccmod_get_spectrum(which, out) : SPCSC_ERR_INVALID); }
int spcsc_ccmod_set_supports(spcsc_handle* h, const int32_t* hw) { H_CALL(h->ccmod_set_supports(hw)); }
int spcsc_ccmod_cns_get(spcsc_handle* h, int32_t which, void* out) { H_CALL(out ? h->ccmod_cns_get(which, out) : SPCSC_ERR_INVALID); }
int spcsc_ccmod_cns_step(spcsc_handle* h, double rho, double udiv, double rlx, int32_t flags, double out[8]) { H_CALL(out ? h->ccmod_cns_step(rho, udiv, rlx, flags, out) : SPCSC_ERR_INVALID); }
int spcsc_comm_unique_id(const char* nccl_lib, void* id128) {
    if (!id128) { g_last_error = "null id buffer"; return SPCSC_ERR_INVALID; }
    NcclApi& api = nccl_api(nccl_lib);
    if (!api.ok) { g_last_error = api.err; return SPCSC_ERR_NCCL; }
    NcclApi::UniqueId uid;
    int r = api.GetUniqueId(&uid);
    if (r != 0) { g_last_error = std::string("ncclGetUniqueId: ") + api.GetErrorString(r); return SPCSC_ERR_NCCL; }
    memcpy(id128, &uid, sizeof(uid));
    return SPCSC_OK;
}
int spcsc_comm_create(const char* nccl_lib, const void* id128, int32_t rank, int32_t nranks,
                      int32_t device, spcsc_comm** out) {
    if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) {
        g_last_error = "bad argument";
        return SPCSC_ERR_INVALID;
    }
    *out = nullptr;
    NcclApi& api = nccl_api(nccl_lib);
    if (!api.ok) { g_last_error = api.err; return SPCSC_ERR_NCCL; }
    if (cudaSetDevice(device) != cudaSuccess) { g_last_error = "cudaSetDevice failed"; return SPCSC_ERR_CUDA; }
    NcclApi::UniqueId uid;
    memcpy(&uid, id128, sizeof(uid));
    void* comm = nullptr;
    int r = api.CommInitRank(&comm, nranks, uid, rank);
    if (r != 0) { g_last_error = std::string("ncclCommInitRank: ") + api.GetErrorString(r); return SPCSC_ERR_NCCL; }
    spcsc_comm* c = new spcsc_comm();
    c->comm = comm; c->api = &api; c->rank = rank; c->nranks = nranks; c->device = device;
    *out = c;
    return SPCSC_OK;
}
int spcsc_comm_destroy(spcsc_comm* c) {
    if (c) {
#ifndef SPCSC_EMU
        cudaSetDevice(c->device);
        for (int r = 0; r < 8; ++r)
            if (c->p2p_peer[r] && c->p2p_peer[r] != c->p2p_own) cudaIpcCloseMemHandle(c->p2p_peer[r]);
        if (c->p2p_own) cudaFree(c->p2p_own);
        cudaGetLastError();
#endif
        if (c->comm && c->api) c->api->CommDestroy(c->comm);
        delete c;
    }
    return SPCSC_OK;
}
int spcsc_attach_comm(spcsc_handle* h, spcsc_comm* c, double global_nx) { H_CALL(h->attach_comm(c, global_nx)); }
int spcsc_host_alloc(uint64_t bytes, void** out) {
    if (!out) return SPCSC_ERR_INVALID;
    cudaError_t e = host_pool().alloc(out, (size_t)bytes, -1);
    if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return SPCSC_ERR_NOMEM; }
    return SPCSC_OK;
}
int spcsc_host_free(void* p) { host_pool().release(p); return SPCSC_OK; }
int spcsc_trim_pools(void) { dev_pool().trim(); host_pool().trim(); return SPCSC_OK; }
int spcsc_set_dict(spcsc_handle* h, const void* D) { H_CALL(D ? h->set_dict(D) : SPCSC_ERR_INVALID); }
int spcsc_set_signal(spcsc_handle* h, const void* S) { H_CALL(S ? h->set_signal(S) : SPCSC_ERR_INVALID); }
int spcsc_set_l1_weight(spcsc_handle* h, const void* w, const int64_t shape[5]) {
    H_CALL((w && shape) ? h->set_l1_weight(w, shape) : SPCSC_ERR_INVALID);
}
int spcsc_set_l21_weight(spcsc_handle* h, const void* w, const int64_t shape[2]) {
    H_CALL((w && shape) ? h->set_l21_weight(w, shape) : SPCSC_ERR_INVALID);
}
int spcsc_admm_configure(spcsc_handle* h, const spcsc_admm_opts* o) {
    H_CALL(o ? h->admm_configure(o) : SPCSC_ERR_INVALID);
}
int spcsc_admm_reset(spcsc_handle* h, double rho) { H_CALL(h->admm_reset(rho)); }
int spcsc_admm_set_rho(spcsc_handle* h, double rho) { H_CALL(h->admm_set_rho(rho)); }
int spcsc_admm_set_iter(spcsc_handle* h, int32_t k) { H_CALL(k >= 0 ? h->admm_set_iter(k) : SPCSC_ERR_INVALID); }
int spcsc_admm_iterate(spcsc_handle* h, int32_t n_iter, spcsc_itstat* rows, int32_t* n_done,
                       int32_t* stopped) {
    H_CALL(h->admm_iterate(n_iter, rows, n_done, stopped));
}
int spcsc_admm_get_scalars(spcsc_handle* h, double* rho, int32_t* k) {
    H_CALL(h->admm_get_scalars(rho, k));
}
int spcsc_admm_last_timing(spcsc_handle* h, float* ms, int64_t* launches) {
    H_CALL(h->admm_last_timing(ms, launches));
}
int spcsc_admm_schedule_info(spcsc_handle* h, int32_t info[4]) { H_CALL(info ? h->admm_schedule_info(info) : SPCSC_ERR_INVALID); }
int spcsc_admm_profile(spcsc_handle* h, int32_t n_iter, float kernel_ms[4]) {
    H_CALL(kernel_ms ? h->admm_profile(n_iter, kernel_ms) : SPCSC_ERR_INVALID);
}
int spcsc_get_array(spcsc_handle* h, int32_t which, void* out) {
    H_CALL(out ? h->get_array(which, out) : SPCSC_ERR_INVALID);
}
int spcsc_set_array(spcsc_handle* h, int32_t which, const void* in) {
    H_CALL(in ? h->set_array(which, in) : SPCSC_ERR_INVALID);
}
int spcsc_reconstruct(spcsc_handle* h, const void* X, void* out) {
    H_CALL(out ? h->reconstruct(X, out) : SPCSC_ERR_INVALID);
}

int spcsc_rfft2(int32_t dtype, int32_t device, int32_t batch, int32_t N0, int32_t N1, const void* x,
                void* xf) {
    if (!x || !xf || batch < 1) { g_last_error = "bad argument"; return SPCSC_ERR_INVALID; }
    return dtype == SPCSC_F32 ? unit_fft2<float>(false, device, batch, N0, N1, x, xf, g_last_error)
                              : unit_fft2<double>(false, device, batch, N0, N1, x, xf, g_last_error);
}
int spcsc_tikhonov_filter(int32_t dtype, int32_t device, int32_t batch, int32_t N0, int32_t N1, double lmbda,
                          int32_t npd, const void* s, void* sl, void* sh) {
    if (!s || !sl || !sh || batch < 1) { g_last_error = "bad argument"; return SPCSC_ERR_INVALID; }
    return dtype == SPCSC_F32
               ? unit_tikhonov<float>(device, batch, N0, N1, lmbda, npd, s, sl, sh, g_last_error)
               : unit_tikhonov<double>(device, batch, N0, N1, lmbda, npd, s, sl, sh, g_last_error);
}
int spcsc_solvedbi_sm(int32_t dtype, int32_t device, int64_t nf, int32_t nk, int32_t Cd, int32_t M, double rho,
                      const void* ah, const void* b, void* x) {
    if (!ah || !b || !x) { g_last_error = "bad argument"; return SPCSC_ERR_INVALID; }
    return dtype == SPCSC_F32 ? unit_solvedbi<float>(device, nf, nk, Cd, M, rho, ah, b, x, g_last_error)
                              : unit_solvedbi<double>(device, nf, nk, Cd, M, rho, ah, b, x, g_last_error);
}
int spcsc_prox_l1(int32_t dtype, int32_t device, int64_t n, double alpha, const void* w, const void* v, void* out) {
    if (!v || !out) { g_last_error = "bad argument"; return SPCSC_ERR_INVALID; }
    return dtype == SPCSC_F32 ? unit_prox<float>(device, 1, 1, n, alpha, 0.0, 0, w, v, out, g_last_error)
                              : unit_prox<double>(device, 1, 1, n, alpha, 0.0, 0, w, v, out, g_last_error);
}
int spcsc_prox_sl1l2(int32_t dtype, int32_t device, int64_t n_outer, int32_t C, int64_t n_inner, double alpha,
                     double beta, const void* v, void* out) {
    if (!v || !out) { g_last_error = "bad argument"; return SPCSC_ERR_INVALID; }
    return dtype == SPCSC_F32 ? unit_prox<float>(device, n_outer, C, n_inner, alpha, beta, 1, nullptr, v, out, g_last_error)
                              : unit_prox<double>(device, n_outer, C, n_inner, alpha, beta, 1, nullptr, v, out, g_last_error);
}
int spcsc_irfft2(int32_t dtype, int32_t device, int32_t batch, int32_t N0, int32_t N1, const void* xf,
                 void* x) {
    if (!x || !xf || batch < 1) { g_last_error = "bad argument"; return SPCSC_ERR_INVALID; }
    return dtype == SPCSC_F32 ? unit_fft2<float>(true, device, batch, N0, N1, xf, x, g_last_error)
                              : unit_fft2<double>(true, device, batch, N0, N1, xf, x, g_last_error);
}

}  // extern "C"

// platform.h -- the one place where the kernels meet the toolchain.
//
// Product build: nvcc, sm_100a, real CUDA runtime.
// SPCSC_EMU build: g++ only, used by tests/emu to execute the *same kernel source* on the
// CPU (cooperative fibres stand in for CUDA threads) so index arithmetic and control
// flow can be checked in a container without a GPU.  The emulation layer itself lives
// under tests/emu/ and is never built into libspcsc.so.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <utility>

#ifdef SPCSC_EMU
#include "cuda_emu.h"            // tests/emu/cuda_emu.h (test infrastructure)
#define SPCSC_HD inline
#define SPCSC_DEV inline
#define SPCSC_GLOBAL
#define SPCSC_LAUNCH_BOUNDS(t)
#define SPCSC_RESTRICT
#define SPCSC_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#define SPCSC_UNROLL
#else
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cuda_pipeline.h>
#define SPCSC_HD __host__ __device__ __forceinline__
#define SPCSC_DEV __device__ __forceinline__
#define SPCSC_GLOBAL __global__
#define SPCSC_LAUNCH_BOUNDS(t) __launch_bounds__(t)
#define SPCSC_RESTRICT __restrict__
#define SPCSC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SPCSC_UNROLL _Pragma("unroll")
#endif

namespace spcsc {

// ---- kernel launch -----------------------------------------------------------------
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          cudaStream_t stream, Args... args) {
#ifdef SPCSC_EMU
    (void)stream;
    emu::launch(grid, block, smem, [=]() { kern(args...); });
    return cudaSuccess;
#else
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, block, smem, stream>>>(args...);
    return cudaGetLastError();
#endif
}

// ---- thread-block clusters (distributed shared memory) ----------------------------------
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, unsigned cs,
                                  size_t smem, cudaStream_t stream, Args... args) {
#ifdef SPCSC_EMU
    (void)stream;
    emu::launch_cluster(grid, block, cs, smem, [=]() { kern(args...); });
    return cudaSuccess;
#else
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
#endif
}

#ifdef SPCSC_EMU
inline unsigned cluster_rank() { return emu::cluster_rank(); }
inline unsigned cluster_size() { return emu::cluster_size(); }
inline void cluster_arrive() { emu::cluster_arrive(); }
inline void cluster_wait() { emu::cluster_wait(); }
template <typename P> inline P* cluster_peer(P* p, unsigned rank) {
    return reinterpret_cast<P*>(emu::map_shared_rank((void*)p, rank));
}
#else
SPCSC_DEV unsigned cluster_rank() { return cooperative_groups::this_cluster().block_rank(); }
SPCSC_DEV unsigned cluster_size() { return cooperative_groups::this_cluster().num_blocks(); }
SPCSC_DEV void cluster_arrive() { cooperative_groups::this_cluster().barrier_arrive(); }
SPCSC_DEV void cluster_wait() { cooperative_groups::this_cluster().barrier_wait(); }
template <typename P> SPCSC_DEV P* cluster_peer(P* p, unsigned rank) {
    return cooperative_groups::this_cluster().map_shared_rank(p, rank);
}
#endif

// ---- asynchronous global -> shared copies (LDGSTS) -------------------------------------
#ifdef SPCSC_EMU
template <int BYTES> inline void cp_async(void* dst, const void* src) { memcpy(dst, src, BYTES); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
#else
template <int BYTES> SPCSC_DEV void cp_async(void* dst, const void* src) {
    __pipeline_memcpy_async(dst, src, BYTES);
}
SPCSC_DEV void cp_async_commit() { __pipeline_commit(); }
template <int N> SPCSC_DEV void cp_async_wait() { __pipeline_wait_prior(N); }
#endif

// ---- complex value type with natural vector alignment (8 B for float, 16 B for double)
template <typename T>
struct alignas(2 * sizeof(T)) C2 {
    T re, im;
};

template <typename T> SPCSC_HD C2<T> mk(T a, T b) { C2<T> r; r.re = a; r.im = b; return r; }
#if defined(SPCSC_EMU) || defined(SPCSC_NO_F32X2)
template <typename T> SPCSC_HD C2<T> operator+(C2<T> a, C2<T> b) { return mk<T>(a.re + b.re, a.im + b.im); }
template <typename T> SPCSC_HD C2<T> operator-(C2<T> a, C2<T> b) { return mk<T>(a.re - b.re, a.im - b.im); }
#else
// Blackwell packed FP32 (FADD2 / FFMA2): one instruction per complex add / subtract.  The
// results are bit-identical to the scalar forms (a - b as fma(b, -1, a) is exact).
SPCSC_DEV float2 as_f2(C2<float> a) { return make_float2(a.re, a.im); }
SPCSC_DEV C2<float> as_c2(float2 a) { C2<float> r; r.re = a.x; r.im = a.y; return r; }
template <typename T> SPCSC_HD C2<T> operator+(C2<T> a, C2<T> b) { return mk<T>(a.re + b.re, a.im + b.im); }
template <typename T> SPCSC_HD C2<T> operator-(C2<T> a, C2<T> b) { return mk<T>(a.re - b.re, a.im - b.im); }
template <> SPCSC_HD C2<float> operator+<float>(C2<float> a, C2<float> b) {
#ifdef __CUDA_ARCH__
    return as_c2(__fadd2_rn(as_f2(a), as_f2(b)));
#else
    return mk<float>(a.re + b.re, a.im + b.im);
#endif
}
template <> SPCSC_HD C2<float> operator-<float>(C2<float> a, C2<float> b) {
#ifdef __CUDA_ARCH__
    return as_c2(__ffma2_rn(as_f2(b), make_float2(-1.0f, -1.0f), as_f2(a)));
#else
    return mk<float>(a.re - b.re, a.im - b.im);
#endif
}
#endif
template <typename T> SPCSC_HD C2<T> operator*(C2<T> a, C2<T> b) {
    return mk<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename T> SPCSC_HD C2<T> operator*(T s, C2<T> a) { return mk<T>(s * a.re, s * a.im); }
template <typename T> SPCSC_HD C2<T> conj(C2<T> a) { return mk<T>(a.re, -a.im); }
// a * conj(b)
template <typename T> SPCSC_HD C2<T> mulc(C2<T> a, C2<T> b) {
    return mk<T>(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im);
}
// multiply by +i / -i
template <typename T> SPCSC_HD C2<T> mul_i(C2<T> a) { return mk<T>(-a.im, a.re); }
template <typename T> SPCSC_HD C2<T> mul_mi(C2<T> a) { return mk<T>(a.im, -a.re); }
template <typename T> SPCSC_HD T abs2(C2<T> a) { return a.re * a.re + a.im * a.im; }

// ---- warp / block reductions (double accumulators) -------------------------------------
SPCSC_DEV double warp_sum(double v) {
    SPCSC_UNROLL
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum NV doubles per thread over the whole block, then one atomicAdd per value from
// thread 0 into acc[0..NV).  `red` is block-shared scratch of at least NV*32 doubles.
// All threads of the block must call this (blockDim.x a multiple of 32).
template <int NV>
SPCSC_DEV void block_accumulate(const double (&v)[NV], double* red, double* acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    double w[NV];
    SPCSC_UNROLL
    for (int i = 0; i < NV; ++i) w[i] = warp_sum(v[i]);
    __syncthreads();                       // scratch may alias buffers used earlier
    if (lane == 0) {
        SPCSC_UNROLL
        for (int i = 0; i < NV; ++i) red[i * 32 + warp] = w[i];
    }
    __syncthreads();
    if (warp == 0) {
        SPCSC_UNROLL
        for (int i = 0; i < NV; ++i) {
            double x = (lane < nwarp) ? red[i * 32 + lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0 && x != 0.0) atomicAdd(acc + i, x);
        }
    }
}


// ---- order-independent (bit-reproducible) accumulation of non-negative doubles ------------
// A value is split over three 64-bit integer bins selected by its exponent (bin width 2^32,
// 8 guard bits: up to 2^24 contributions per bin cannot overflow); integer atomics commute, so
// the total does not depend on the order in which thread blocks finish, and the three-way split
// is exact to double precision.  det_bins_value() turns a bin row back into a double (highest
// bin first, a fixed order).
constexpr int kDetBins = 64;
constexpr int kDetBias = 1087 + 8;

// 2^e as a double, built from the exponent bits (|e| within the normal range).
SPCSC_DEV double det_pow2(int e) {
    long long bits = (long long)(e + 1023) << 52;
    double d;
    memcpy(&d, &bits, sizeof(d));
    return d;
}
SPCSC_DEV void det_accumulate(double v, unsigned long long* bins) {
    if (!(v > 0.0)) return;
    long long bits;
    memcpy(&bits, &v, sizeof(bits));
    const int e = (int)((bits >> 52) & 0x7ff) - 1022;      // v = f * 2^e, f in [0.5, 1) (normal v)
    int b = (e + 1087) >> 5;                               // bin of the leading bits
    if (b > kDetBins - 1) b = kDetBins - 1;
    if (b < 5) b = 5;                                      // keeps 2^(+-unit exponent) a normal double; terms < 2^-927 count as 0
    double r = v;
    SPCSC_UNROLL
    for (int lvl = 0; lvl < 3; ++lvl) {
        const int ue = ((b - lvl) << 5) - kDetBias;        // unit = 2^ue
        const double q = floor(r * det_pow2(-ue));         // level 0: < 2^40, then < 2^32 (exact scaling)
        if (q > 0.0) atomicAdd(bins + (b - lvl), (unsigned long long)q);
        r -= q * det_pow2(ue);                             // exact
    }
}
SPCSC_DEV double det_bin_term(const unsigned long long* bins, int b) {
    const unsigned long long q = bins[b];
    const int ue = (b << 5) - kDetBias;
    if (q == 0ull || ue < -1022) return 0.0;
    return (double)q * det_pow2(ue);
}
SPCSC_DEV double det_bins_value(const unsigned long long* bins) {
    double s = 0.0;
    for (int b = kDetBins - 1; b >= 0; --b) s += det_bin_term(bins, b);
    return s;
}

// As block_accumulate, but the block totals go into the reproducible bins (row i of `bins`
// has kDetBins entries) instead of floating-point atomics; lane i of warp 0 handles value i.
template <int NV>
SPCSC_DEV void block_accumulate_det(const double (&v)[NV], double* red, unsigned long long* bins) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    double w[NV];
    SPCSC_UNROLL
    for (int i = 0; i < NV; ++i) w[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
        SPCSC_UNROLL
        for (int i = 0; i < NV; ++i) red[i * 32 + warp] = w[i];
    }
    __syncthreads();
    if (warp == 0) {
        double mine = 0.0;
        SPCSC_UNROLL
        for (int i = 0; i < NV; ++i) {
            double x = (lane < nwarp) ? red[i * 32 + lane] : 0.0;
            x = warp_sum(x);
            if (lane == i) mine = x;
        }
        if (lane < NV) det_accumulate(mine, bins + lane * kDetBins);
    }
}

}  // namespace spcsc

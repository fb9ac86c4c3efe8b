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

// number of clusters of `cs` CTAs that can be resident at once (persistent grids)
template <typename... KArgs>
inline int max_active_clusters(void (*kern)(KArgs...), dim3 block, unsigned cs, size_t smem) {
#ifdef SPCSC_EMU
    (void)kern; (void)block; (void)cs; (void)smem;
    return 3;
#else
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(cs, 1, 1);
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
#endif
}

#ifdef SPCSC_EMU
inline unsigned cluster_rank() { return emu::cluster_rank(); }
inline unsigned cluster_size() { return emu::cluster_size(); }
inline void cluster_arrive() { emu::cluster_arrive(); }
inline void cluster_arrive_relaxed() { emu::cluster_arrive(); }
inline void cluster_wait() { emu::cluster_wait(); }
template <typename P> inline P* cluster_peer(P* p, unsigned rank) {
    return reinterpret_cast<P*>(emu::map_shared_rank((void*)p, rank));
}
#else
// arrival that publishes nothing (the caller has only finished READING its peers' memory): no
// release fence, which otherwise costs a GPU-scope MEMBAR + ERRBAR per arrival
SPCSC_DEV void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
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

// ---- bulk asynchronous copy global -> shared (TMA, 1-D) completing on an mbarrier ------------
typedef unsigned long long mbar_t;
#ifdef SPCSC_EMU
// emulated mbarrier: transaction count, pending arrivals, arrival count per phase, phase parity.  A
// cluster runs as cooperative fibres on one OS thread, so plain fields suffice.
struct EmuBar { int tx; unsigned short pending; unsigned char init, phase; };
static_assert(sizeof(EmuBar) == sizeof(mbar_t), "emulated barrier must fit the real one");
inline void emu_bar_check(EmuBar* b) {
    if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->init; }
}
inline void mbar_init(mbar_t* bar, unsigned count) {
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    b->tx = 0; b->pending = (unsigned short)count; b->init = (unsigned char)count; b->phase = 0;
}
inline void mbar_expect_tx(mbar_t* bar, unsigned bytes) {
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    b->tx += (int)bytes; b->pending -= 1; emu_bar_check(b);
}
inline void mbar_complete_tx(mbar_t* bar, unsigned bytes) {
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    b->tx -= (int)bytes; emu_bar_check(b);
}
inline void bulk_load(void* dst, const void* src, unsigned bytes, mbar_t* bar) {
    mbar_expect_tx(bar, bytes);
    memcpy(dst, src, bytes);
    mbar_complete_tx(bar, bytes);
}
// copy only: the bytes must have been announced with mbar_expect_tx (several copies, one arrival)
inline void bulk_copy(void* dst, const void* src, unsigned bytes, mbar_t* bar) {
    memcpy(dst, src, bytes);
    mbar_complete_tx(bar, bytes);
}
inline void mbar_wait(mbar_t* bar, unsigned parity) {
    while (reinterpret_cast<EmuBar*>(bar)->phase == (unsigned char)parity) emu::yield();
}
// bulk copy shared -> global (TMA store), tracked by the issuing thread's bulk group
inline void fence_async_smem() {}
inline void bulk_store(void* gdst, const void* ssrc, unsigned bytes) { memcpy(gdst, ssrc, bytes); }
inline void bulk_store_commit() {}
inline void bulk_store_wait_read() {}
inline void bulk_store_wait_all() {}
#else
SPCSC_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SPCSC_DEV void mbar_init(mbar_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: arrive on the barrier and announce `bytes` of asynchronous writes for the current phase
SPCSC_DEV void mbar_expect_tx(mbar_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
// one thread: announce `bytes` on the barrier and start the copy (16-byte aligned, multiple of 16)
SPCSC_DEV void bulk_load(void* dst, const void* src, unsigned bytes, mbar_t* bar) {
    mbar_expect_tx(bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// copy only: the bytes must have been announced with mbar_expect_tx (several copies, one arrival)
SPCSC_DEV void bulk_copy(void* dst, const void* src, unsigned bytes, mbar_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// ---- bulk copy shared -> global (TMA store), tracked by the issuing thread's bulk group -----------------
// generic-proxy writes to shared memory become visible to the asynchronous proxy (call before the barrier that
// precedes bulk_store)
SPCSC_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SPCSC_DEV void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)),
                 "r"(bytes)
                 : "memory");
}
SPCSC_DEV void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's stores have finished READING shared memory (the source may be overwritten)
SPCSC_DEV void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SPCSC_DEV void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
SPCSC_DEV void mbar_wait(mbar_t* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
}
#endif

// ---- pushing values into a peer CTA's shared memory (st.async over distributed shared memory) ---------
// The store completes `sizeof(value)` transaction bytes on an mbarrier of the RECEIVING CTA, which waits on
// its own barrier: no cluster-wide barrier, hence none of the cluster-scope fences (and the L1 invalidation)
// that barrier.cluster.arrive.release / wait.acquire bring along.
// ---- named barrier over a subset of the CTA's warps (independent thread groups inside one CTA) -------------
#ifdef SPCSC_EMU
inline void group_barrier(int id, int nthreads) { emu::named_barrier(id, nthreads); }
inline void nap(unsigned) {}
#else
SPCSC_DEV void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
SPCSC_DEV void nap(unsigned ns) { __nanosleep(ns); }
#endif

#ifdef SPCSC_EMU
typedef unsigned char* rptr_t;                       // address in a peer's shared memory
template <typename P> inline rptr_t cluster_remote(P* p, unsigned rank) {
    return reinterpret_cast<rptr_t>(emu::map_shared_rank((void*)p, rank));
}
inline rptr_t rptr_add(rptr_t r, size_t bytes) { return r + bytes; }
template <typename V> inline void push_remote(rptr_t dst, V v, rptr_t bar) {
    memcpy(dst, &v, sizeof(V));
    mbar_complete_tx(reinterpret_cast<mbar_t*>(bar), (unsigned)sizeof(V));
}
#else
typedef unsigned rptr_t;                             // shared::cluster address
template <typename P> SPCSC_DEV rptr_t cluster_remote(P* p, unsigned rank) {
    rptr_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr(p)), "r"(rank));
    return r;
}
SPCSC_DEV rptr_t rptr_add(rptr_t r, size_t bytes) { return r + (unsigned)bytes; }
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
// elementwise helpers on (re, im) used as a pair of reals: a*s, a*s + c, a*b + c (lane by lane)
template <typename T> SPCSC_HD C2<T> pmul(C2<T> a, T s) { return mk<T>(a.re * s, a.im * s); }
template <typename T> SPCSC_HD C2<T> pfma(C2<T> a, T s, C2<T> c) { return mk<T>(a.re * s + c.re, a.im * s + c.im); }
template <typename T> SPCSC_HD C2<T> pfma(C2<T> a, C2<T> b, C2<T> c) {
    return mk<T>(a.re * b.re + c.re, a.im * b.im + c.im);
}
#if !defined(SPCSC_EMU) && !defined(SPCSC_NO_F32X2)
// Packed forms: FMUL2 takes a scalar broadcast operand and FFMA2 a lane-swapped, per-lane negated
// one, so a complex product is two instructions (re = fma(-a.im, b.im, a.re*b.re), im =
// fma(a.re, b.im, a.im*b.re)) instead of four.
template <> SPCSC_HD C2<float> operator*<float>(C2<float> a, C2<float> b) {
#ifdef __CUDA_ARCH__
    const float2 t = __fmul2_rn(as_f2(a), make_float2(b.re, b.re));
    return as_c2(__ffma2_rn(make_float2(a.im, a.re), make_float2(-b.im, b.im), t));
#else
    return mk<float>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
#endif
}
template <> SPCSC_HD C2<float> operator*<float>(float s, C2<float> a) {
#ifdef __CUDA_ARCH__
    return as_c2(__fmul2_rn(as_f2(a), make_float2(s, s)));
#else
    return mk<float>(s * a.re, s * a.im);
#endif
}
template <> SPCSC_HD C2<float> mulc<float>(C2<float> a, C2<float> b) {
#ifdef __CUDA_ARCH__
    const float2 t = __fmul2_rn(as_f2(a), make_float2(b.re, b.re));
    return as_c2(__ffma2_rn(make_float2(a.im, a.re), make_float2(b.im, -b.im), t));
#else
    return mk<float>(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im);
#endif
}
template <> SPCSC_HD C2<float> pmul<float>(C2<float> a, float s) {
#ifdef __CUDA_ARCH__
    return as_c2(__fmul2_rn(as_f2(a), make_float2(s, s)));
#else
    return mk<float>(a.re * s, a.im * s);
#endif
}
template <> SPCSC_HD C2<float> pfma<float>(C2<float> a, float s, C2<float> c) {
#ifdef __CUDA_ARCH__
    return as_c2(__ffma2_rn(as_f2(a), make_float2(s, s), as_f2(c)));
#else
    return mk<float>(a.re * s + c.re, a.im * s + c.im);
#endif
}
template <> SPCSC_HD C2<float> pfma<float>(C2<float> a, C2<float> b, C2<float> c) {
#ifdef __CUDA_ARCH__
    return as_c2(__ffma2_rn(as_f2(a), as_f2(b), as_f2(c)));
#else
    return mk<float>(a.re * b.re + c.re, a.im * b.im + c.im);
#endif
}
#endif
// multiply by +i / -i
template <typename T> SPCSC_HD C2<T> mul_i(C2<T> a) { return mk<T>(-a.im, a.re); }
template <typename T> SPCSC_HD C2<T> mul_mi(C2<T> a) { return mk<T>(a.im, -a.re); }
template <typename T> SPCSC_HD T abs2(C2<T> a) { return a.re * a.re + a.im * a.im; }

// Streaming (read-once) global load: evict-first in L1 so that re-used operands (dictionary
// spectra) keep their lines.
#ifdef SPCSC_EMU
template <typename T> inline C2<T> ld_stream(const C2<T>* p) { return *p; }
#else
SPCSC_DEV C2<float> ld_stream(const C2<float>* p) {
    const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
    C2<float> r; r.re = v.x; r.im = v.y; return r;
}
SPCSC_DEV C2<double> ld_stream(const C2<double>* p) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    C2<double> r; r.re = v.x; r.im = v.y; return r;
}
#endif

// Bring a line into L2 ahead of its use (the next slab of a persistent CTA).
#ifdef SPCSC_EMU
inline void prefetch_l2(const void*) {}
#else
SPCSC_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// Re-used read-only operand (dictionary spectra, read twice per slab by the same thread): ask L1 to
// keep the line (evict-last), the counterpart of ld_stream.
#ifdef SPCSC_EMU
template <typename T> inline C2<T> ld_keep(const C2<T>* p) { return *p; }
#else
SPCSC_DEV C2<float> ld_keep(const C2<float>* p) {
    C2<float> r;
    asm volatile("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(r.re), "=f"(r.im) : "l"(p));
    return r;
}
SPCSC_DEV C2<double> ld_keep(const C2<double>* p) { return *p; }
#endif

#ifndef SPCSC_EMU
SPCSC_DEV void push_remote(rptr_t dst, C2<float> v, rptr_t bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(dst),
                 "f"(v.re), "f"(v.im), "r"(bar)
                 : "memory");
}
SPCSC_DEV void push_remote(rptr_t dst, C2<double> v, rptr_t bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(dst),
                 "d"(v.re), "d"(v.im), "r"(bar)
                 : "memory");
}
#endif

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

// Eight float values per lane summed over the warp with 9 shuffles (halving exchange: after the
// steps with offsets 16, 8, 4 every lane carries one value, then two plain butterfly steps):
// lane 4q ends up with the warp total of value q.  Fixed order, hence reproducible.
SPCSC_DEV float warp_sum8(const float (&s)[8]) {
    const int lane = threadIdx.x & 31;
    float a[4], b[2];
    const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0, u4 = (lane & 4) != 0;
    SPCSC_UNROLL
    for (int i = 0; i < 4; ++i) {
        const float send = u16 ? s[i] : s[i + 4], keep = u16 ? s[i + 4] : s[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    SPCSC_UNROLL
    for (int i = 0; i < 2; ++i) {
        const float send = u8 ? a[i] : a[i + 2], keep = u8 ? a[i + 2] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float c = (u4 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, u4 ? b[0] : b[1], 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}

// Block totals of up to eight float values per thread into the reproducible bins: warp totals by
// warp_sum8, then lane i of warp 0 adds the per-warp partials of value i in warp order.
// `red`: block-shared scratch of 8*32 floats.
template <int NV>
SPCSC_DEV void block_accumulate_det_f(const float (&v)[8], float* red, unsigned long long* bins) {
    static_assert(NV <= 8, "at most eight values");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    const float c = warp_sum8(v);
    __syncthreads();                       // scratch may alias buffers used earlier
    if ((lane & 3) == 0) red[(lane >> 2) * 32 + warp] = c;
    __syncthreads();
    if (warp == 0 && lane < NV) {
        double mine = 0.0;
        for (int w = 0; w < nwarp; ++w) mine += (double)red[lane * 32 + w];
        det_accumulate(mine, bins + lane * kDetBins);
    }
}

}  // namespace spcsc

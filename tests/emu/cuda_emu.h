// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny stand-in for the CUDA runtime and SIMT execution model so that the kernel
// sources under sporco_b200/csrc can be compiled with plain g++ (-DSPCSC_EMU) and executed
// on the CPU by tests/ in a container without a GPU.  CUDA threads are cooperative
// fibres (ucontext); blocks are spread over a pool of OS threads.  Nothing here is ever
// linked into the product library libspcsc.so, and nothing measured comes from it.
#pragma once

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3 { unsigned x, y, z; };

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct emuEvent* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost,
                      cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };

namespace emu {
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
unsigned char* dyn_smem();
void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> fn);
void launch_cluster(dim3 grid, dim3 block, unsigned cs, size_t smem, std::function<void()> fn);
unsigned cluster_rank();
unsigned cluster_size();
void cluster_arrive();
void cluster_wait();
void* map_shared_rank(void* p, unsigned rank);
void yield();                        // let the other fibres of the cluster run (used by emulated waits)
void sync_threads();
void named_barrier(int id, int nthreads);   // bar.sync id, nthreads
void sync_warp();
// exchange 16-byte payloads between lanes of the calling warp
void warp_exchange(const void* mine, void* out, int src_lane, size_t nbytes);
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
#define __shared__ static thread_local

inline void __syncthreads() { emu::sync_threads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
    T out;
    emu::warp_exchange(&v, &out, src & 31, sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int mask) {
    return __shfl_sync(m, v, (int)((emu::t_threadIdx.x & 31) ^ mask));
}
template <typename T>
inline T __shfl_down_sync(unsigned m, T v, int d) {
    int lane = (int)(emu::t_threadIdx.x & 31);
    int src = lane + d;
    T out;
    emu::warp_exchange(&v, &out, src < 32 ? src : lane, sizeof(T));
    return out;
}

template <typename T>
inline T __ldg(const T* p) { return *p; }

inline double atomicAdd(double* p, double v) {
    std::atomic<double>* a = reinterpret_cast<std::atomic<double>*>(p);
    double old = a->load();
    while (!a->compare_exchange_weak(old, old + v)) {}
    return old;
}
inline float atomicAdd(float* p, float v) {
    std::atomic<float>* a = reinterpret_cast<std::atomic<float>*>(p);
    float old = a->load();
    while (!a->compare_exchange_weak(old, old + v)) {}
    return old;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- runtime API subset -------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t n);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t n);
cudaError_t cudaFreeHost(void* p);
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0);
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)1 << 34; return cudaSuccess; }

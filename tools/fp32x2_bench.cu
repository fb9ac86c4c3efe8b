// Microbenchmark: issue throughput of scalar vs packed FP32 on sm_100a (FADD/FFMA vs FADD2/FFMA2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_bench tools/fp32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* out, int iters, float2 seed) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, -0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
            if (MODE == 1) { a[i] = __fadd2_rn(a[i], c); }
            if (MODE == 2) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            if (MODE == 3) { a[i] = __ffma2_rn(a[i], m, c); }
        }
    }
    float2 s = make_float2(0, 0);
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
    float2* d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(float2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    k<MODE><<<148 * 8, 1024>>>(d, 16, make_float2(1, 2));
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 1024>>>(d, iters, make_float2(1, 2));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 148.0 * 8 * 1024 * iters * 8 * 2 * ((MODE >= 2) ? 2 : 1);
    printf("%-8s %.3f ms  %.1f TFLOP/s  (%.2f G lane-ops/s)\n", name, ms, flops / ms / 1e9,
           148.0 * 8 * 1024 * iters * 16 / ms / 1e6);
    cudaFree(d);
}
int main() { run<0>("FADD"); run<1>("FADD2"); run<2>("FFMA"); run<3>("FFMA2"); return 0; }

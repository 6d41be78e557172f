// Microbenchmark: FP64 vector FMA vs mma.sync.m8n8k4.f64 (DMMA) throughput on B200, alone and mixed.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>  // 0: dfma only, 1: dmma only, 2: both interleaved (8 dfma : 1 dmma per acc set)
__global__ void k(double *out, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001;
    double f[8], c[16];
    for (int i = 0; i < 8; ++i) f[i] = a + i;
    for (int i = 0; i < 16; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __fma_rn(f[i], b, 1e-9);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += f[i];
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, double flop_per_thread_iter) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 1 << 14;
    double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, best, flop_per_thread_iter * iters * (double)blocks * threads / (best * 1e-3) * 1e-12);
    cudaFree(d);
}
int main() {
    run<0>("dfma only (8/iter)", 2.0 * 8);
    // one warp-level m8n8k4 = 8*8*4 FMAs = 512 flop per warp = 16 flop per thread
    run<1>("dmma only (8/iter)", 16.0 * 8);
    run<2>("dfma+dmma mixed", 2.0 * 8 + 16.0 * 8);
    return 0;
}

// Constant-bank (LDCU -> uniform register) operand delivery for the lanes-are-elements contraction (B200).
// Each warp repeatedly sweeps 27 "Gauss points"; per point it pulls 24 warp-uniform doubles from __constant__ memory
// and issues 60 DFMA with them.  Patterns: the strided row layout [g][k][MEP], a per-tile contiguous stream, a table
// small enough for the first-level constant cache, and 128-bit loads.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double ctab[7680];   // 60 kB
template <int MODE>   // 0 strided rows, 1 contiguous per tile, 2 small table, 3 contiguous double2, 4 no constant loads (DFMA only)
__global__ void __launch_bounds__(512, 1) k(double *out, int iters, long long *cyc) {
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    double acc[16], b[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = 1.0 + 1e-9 * (threadIdx.x + j);
    const int ti = warp % 9, tj = (warp * 5) % 9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 3
        for (int g = 0; g < 27; ++g) {
            double y[12], x[12];
            if (MODE == 0) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) { y[r * 4 + i] = ctab[(g * 4 + r) * 36 + 4 * ti + i]; x[r * 4 + i] = ctab[(g * 4 + r + 1) * 36 + 4 * tj + i]; }
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 12; ++i) { y[i] = ctab[(ti * 27 + g) * 12 + i]; x[i] = ctab[3000 + (tj * 27 + g) * 12 + i]; }
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 12; ++i) { y[i] = ctab[((g & 3) * 4 + (ti & 1)) * 12 + i]; x[i] = ctab[((g & 3) * 4 + 2 + (tj & 1)) * 12 + i]; }
            } else if (MODE == 3) {
                const double2 *py = reinterpret_cast<const double2 *>(ctab + (ti * 27 + g) * 12), *px = reinterpret_cast<const double2 *>(ctab + 3000 + (tj * 27 + g) * 12);
#pragma unroll
                for (int i = 0; i < 6; ++i) { double2 v = py[i]; y[2 * i] = v.x; y[2 * i + 1] = v.y; double2 w = px[i]; x[2 * i] = w.x; x[2 * i + 1] = w.y; }
            } else {
#pragma unroll
                for (int i = 0; i < 12; ++i) { y[i] = 1.0 + i * 1e-9; x[i] = 1.0 - i * 1e-9; }
            }
            double b1[4], b2[4], bw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { b1[j] = __fma_rn(b[0], x[j], -(b[1] * x[4 + j])); b2[j] = __fma_rn(b[2], x[j], -(b[3] * x[4 + j])); bw[j] = x[8 + j] * b[0]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i * 4 + j] = __fma_rn(y[i], b1[j], __fma_rn(-y[4 + i], b2[j], acc[i * 4 + j]));
                    acc[(i * 4 + j + 5) & 15] = __fma_rn(y[8 + i], bw[j], acc[(i * 4 + j + 5) & 15]);
                }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char *name) {
    const int blocks = 148, threads = 512, iters = 200;
    double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    long long *dc, hc; cudaMalloc(&dc, 8);
    k<MODE><<<blocks, threads>>>(d, iters, dc);
    k<MODE><<<blocks, threads>>>(d, iters, dc);
    cudaDeviceSynchronize(); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    // per SM: 16 warps x iters x 27 points; ideal FP64: 68 warp-instr x 2 cycles / 4 SMSP = 34 cycles per warp-point SM-wide
    printf("%-40s %7.1f cycles per warp-point SM-wide (FP64-pipe bound: 34)  %s\n", name, (double)hc / (16.0 * iters * 27), cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(dc);
}
int main() {
    static double h[7680];
    for (int i = 0; i < 7680; ++i) h[i] = 1.0 + 1e-7 * i;
    cudaMemcpyToSymbol(ctab, h, sizeof(h));
    run<4>("no constant loads");
    run<0>("strided rows [g][k][36]");
    run<1>("contiguous per tile, LDCU.64");
    run<3>("contiguous per tile, 128-bit");
    run<2>("small table (768 B)");
    return 0;
}

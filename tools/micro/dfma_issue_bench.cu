// dfma_issue_bench.cu -- why do the accumulate chains of contract_kernel issue every 2.44 cycles instead of 2?  (B200, sm_100a)
//
// tile_bench.cu (hoist mode) measured the contraction's inner loop WITHOUT any loads at 2.44 cycles per DFMA per SMSP (82 % of the
// nominal FP64 rate), where a pure FMA loop runs at 2.05.  In that loop every DFMA reads a reused operand (y, .reuse flag), a
// distinct b[j] and a distinct accumulator.  This microbenchmark varies how many operands are new per instruction and how
// many accumulators are live, 16 warps per SM, no memory traffic in the loop:
//     P0  acc[k] = fma(y0, b0, acc[k])              both multiplicands reused, only the accumulator is new          (best case)
//     P1  acc[i][j] = fma(y[i], b[j], acc[i][j])    the contraction's pattern: y reused over j, b[j] and acc new
//     P2  the same with j outer, i inner            b reused over i, y[i] and acc new
//     P3  P1 followed by the M chain acc2[i][j] = fma(y3[i], bw[j], acc2[i][j])   (32 accumulators, as in the kernel)
//     P4  P3 with the K part as two dependent FMAs per pair (exactly the kernel's K update)
// For every kernel the host prints cycles per DFMA per SMSP.  `cuobjdump -sass dfma_issue_bench | grep DFMA` shows the register
// numbers ptxas chose; tools/micro/README.md explains how to relate them to the timing (operand-bank hypothesis: a DFMA whose
// new multiplicand and accumulator sit in the same 64-bit bank, (reg/2) % 2, costs an extra cycle).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o dfma_issue_bench dfma_issue_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }

template <int P, int NACC>
__global__ void __launch_bounds__(512, 1) k(const double *in, double *out, int iters, long long *cyc) {
    double y[4], b[4], y3[4], bw[4], acc[NACC], acc2[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) { y[i] = in[i]; b[i] = in[4 + i]; y3[i] = in[8 + i]; bw[i] = in[12 + i]; }
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = in[16 + (i & 15)];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc2[i] = in[16 + i];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (P == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int kk = 0; kk < NACC; ++kk) acc[kk] = dfma(y[0], b[0], acc[kk]);
        } else if (P == 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i * 4 + j] = dfma(y[i], b[j], acc[i * 4 + j]);
        } else if (P == 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i * 4 + j] = dfma(y[i], b[j], acc[i * 4 + j]);
        } else if (P == 3) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[i * 4 + j] = dfma(y[i], b[j], acc[i * 4 + j]);
                        acc2[i * 4 + j] = dfma(y3[i], bw[j], acc2[i * 4 + j]);
                    }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[i * 4 + j] = dfma(y[i], b[j], dfma(-y3[i], bw[j], acc[i * 4 + j]));
                        acc2[i * 4 + j] = dfma(y3[i], b[j], acc2[i * 4 + j]);
                    }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc2[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int P, int NACC>
void run(const char *what, int dfma_per_iter) {
    const int blocks = 148, threads = 512, iters = 4000;
    double h[32];
    for (int i = 0; i < 32; ++i) h[i] = 1.0 + 1e-9 * i;
    double *din, *d; long long *dc, hc = 0;
    cudaMalloc(&din, sizeof(h)); cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&d, sizeof(double) * blocks * threads); cudaMalloc(&dc, 8);
    k<P, NACC><<<blocks, threads>>>(din, d, iters, dc);
    k<P, NACC><<<blocks, threads>>>(din, d, iters, dc);
    cudaDeviceSynchronize(); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<P, NACC>);
    // 16 warps per SM = 4 per SMSP; each SMSP issues its warps' DFMAs one after the other
    printf("P%d acc=%2d regs=%3d  %5.3f cycles per DFMA per SMSP   %s\n", P, NACC, fa.numRegs, (double)hc / (4.0 * iters * dfma_per_iter), what);
    cudaFree(din); cudaFree(d); cudaFree(dc);
}

int main() {
    run<0, 8>("both multiplicands reused, 8 accumulators", 32);
    run<0, 16>("both multiplicands reused, 16 accumulators", 64);
    run<1, 16>("y reused over j, b[j] and acc new (the contraction's pattern)", 64);
    run<2, 16>("b reused over i, y[i] and acc new", 64);
    run<3, 16>("K-like and M-like chains interleaved, 32 accumulators", 64);
    run<4, 16>("the kernel's update: K = two dependent FMAs per pair, M one", 96);
    return 0;
}

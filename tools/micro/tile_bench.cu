// tile_bench.cu -- would taller tiles pay in contract_kernel?  (B200, sm_100a; question for round 2)
//
// contract_kernel (movfem_b200/csrc/contract.cuh) gives a warp one 4x4 tile of slot pairs for 32 elements (lanes).  Per
// tile and Gauss point it issues 68 FP64 instructions (20 to form b1, b2, bw from the column operands and the per-lane
// 2x2 block of Q and T, 48 for the 16 pairs) against 12 broadcast LDS.128 + 5 lane-distinct LDS.64, and ncu shows the
// FP64 pipe (63 %) and the shared-memory return path (68 %) balanced against each other.  A tile of R x 4 pairs
// (R = 8, 12: two or three row groups of the SAME direction, which share the column operands and b1/b2/bw) needs
//     FP64 instructions   20 + 12 R      (R=4: 68, 8: 116, 12: 164)   -> 4.25 / 3.63 / 3.42 per pair
//     LDS.128 broadcast   6 + 1.5 R      (12 / 18 / 24)               -> 0.75 / 0.56 / 0.50 per pair
//     LDS.64 per lane     5
// at the price of 2*4*R accumulator doubles per lane (R=8: 128 registers, R=12: 192), i.e. fewer resident warps.
// This microbenchmark runs exactly that inner loop (same operand layout, same dfma chains, static operands: no TMA ring)
// for R = 4, 8, 12 and several warp counts and prints pairs*Gauss points per cycle per SM and the FP64 issue utilisation
// (an SM issues at most 2 FP64 warp instructions per cycle).  me = 36 geometry: MEP = 36 slots, 27 Gauss points.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -Xptxas -v -o tile_bench tile_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int MEP = 36, NGP = 27, CB = NGP * 32;

__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ void ld4(double (&v)[4], const double *p) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <int R, int W, int H>   // H: 0 the real loop; 1 the 12 broadcast LDS.128 hoisted out of the Gauss-point loop; 2 no loads in the loop at all
__global__ void __launch_bounds__(W * 32, 1) k(double *out, int iters, long long *cyc) {
    extern __shared__ __align__(16) double sm[];
    double *s_tab = sm;                       // [g][k][MEP]
    double *s_stage = sm + NGP * 4 * MEP;     // [5][g][32]
    for (int i = threadIdx.x; i < NGP * 4 * MEP; i += W * 32) s_tab[i] = 1.0 + 1e-3 * (i % 97);
    for (int i = threadIdx.x; i < 5 * CB; i += W * 32) s_stage[i] = 0.5 + 1e-3 * (i % 89);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double *S = s_stage + lane;
    constexpr int RG = R / 4;                 // row groups per tile
    double total = 0.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        // walk (ti, tj) like the kernel's tile stream: column group tj, first row group ti (same direction block)
        const int tj = (it + warp) % 3, ti = 3 + ((it + warp) / 3) % (4 - RG) ;
        const int k1I = 2, k2I = 0, k1J = 2, k2J = 1;
        double accK[R * 4], accM[R * 4];
#pragma unroll
        for (int i = 0; i < R * 4; ++i) { accK[i] = 0.0; accM[i] = 0.0; }
        const double *Y1 = s_tab + k1I * MEP + 4 * ti, *Y2 = s_tab + k2I * MEP + 4 * ti, *Y3 = s_tab + 3 * MEP + 4 * ti;
        const double *X1 = s_tab + k1J * MEP + 4 * tj, *X2 = s_tab + k2J * MEP + 4 * tj, *X3 = s_tab + 3 * MEP + 4 * tj;
#pragma unroll 3
        for (int g = 0; g < NGP; ++g) {
            const int o = H >= 1 ? 0 : g * 4 * MEP;
            const int gq = H >= 2 ? 0 : g;
            const double q00 = S[(0 * NGP + gq) * 32], q01 = S[(1 * NGP + gq) * 32], q10 = S[(2 * NGP + gq) * 32],
                         q11 = S[(3 * NGP + gq) * 32], tt = S[(4 * NGP + gq) * 32];
            double b1[4], b2[4], bw[4], xa[4], xb[4], xc[4];
            ld4(xa, X1 + o); ld4(xb, X2 + o); ld4(xc, X3 + o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b1[j] = dfma(q00, xa[j], -(q01 * xb[j]));
                b2[j] = dfma(q10, xa[j], -(q11 * xb[j]));
                bw[j] = xc[j] * tt;
            }
#pragma unroll
            for (int rg = 0; rg < RG; ++rg) {
                double ya[4], yb[4], yc[4];
                ld4(ya, Y1 + o + 4 * rg); ld4(yb, Y2 + o + 4 * rg); ld4(yc, Y3 + o + 4 * rg);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double y1 = ya[i], y2 = yb[i], y3 = yc[i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int a = (rg * 4 + i) * 4 + j;
                        accK[a] = dfma(y1, b1[j], dfma(-y2, b2[j], accK[a]));
                        accM[a] = dfma(y3, bw[j], accM[a]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < R * 4; ++i) total += accK[i] + accM[i];
    }
    const long long t1 = clock64();
    out[(size_t)blockIdx.x * W * 32 + threadIdx.x] = total;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int R, int W, int H = 0>
void run() {
    const int blocks = 148, iters = 400;
    const size_t smem = sizeof(double) * (NGP * 4 * MEP + 5 * CB);
    double *d; cudaMalloc(&d, sizeof(double) * blocks * W * 32);
    long long *dc, hc = 0; cudaMalloc(&dc, 8);
    cudaFuncSetAttribute(k<R, W, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<R, W, H>);
    k<R, W, H><<<blocks, W * 32, smem>>>(d, iters, dc);
    k<R, W, H><<<blocks, W * 32, smem>>>(d, iters, dc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    const double tile_g = (double)W * iters * NGP;              // (tile, Gauss point) steps per SM
    const double instr = tile_g * (20 + 12 * R);                // FP64 warp instructions per SM (the final adds excluded)
    printf("H=%d R=%2d W=%2d regs=%3d spill=%zu B  %7.2f cycles/(tile,g) SM-wide  %6.3f pairs*g/cycle/SM  FP64 issue %5.1f %%  %s\n", H, R, W, fa.numRegs,
           (size_t)fa.localSizeBytes, hc / tile_g, tile_g * R * 4 / hc, 100.0 * instr / (2.0 * hc), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(dc);
}

int main(int argc, char **) {
    if (argc > 1) {   // what caps the FP64 issue rate at ~74 %?  take the loads out of the loop step by step
        run<4, 16, 0>(); run<4, 16, 1>(); run<4, 16, 2>();
        run<8, 8, 0>(); run<8, 8, 1>(); run<8, 8, 2>();
        return 0;
    }
    run<4, 16>(); run<4, 15>(); run<4, 12>(); run<4, 8>();
    run<8, 12>(); run<8, 10>(); run<8, 8>(); run<8, 6>();
    run<12, 8>(); run<12, 6>(); run<12, 4>();
    return 0;
}

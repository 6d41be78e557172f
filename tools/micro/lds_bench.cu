// Shared-memory load throughput for the access patterns of the contraction phase (B200).
#include <cstdio>
#include <cuda_runtime.h>
template <int W>  // W = bytes per lane (8 or 16)
__global__ void k(double *out, const int *idx, int iters, long long *cyc) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int base = idx[threadIdx.x & 31];   // in doubles
    double acc0 = 0, acc1 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int o = (base + u * 64 + (it & 7) * 512) & 4095 & ~1;
            if (W == 16) { double2 v = *reinterpret_cast<const double2 *>(sm + o); acc0 += v.x; acc1 += v.y; }
            else { acc0 += sm[o]; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int W>
void run(const char *name, const int *hidx) {
    int *didx; cudaMalloc(&didx, 32 * 4); cudaMemcpy(didx, hidx, 128, cudaMemcpyHostToDevice);
    const int blocks = 148, threads = 512, iters = 2000;
    double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    long long *dc, hc; cudaMalloc(&dc, 8);
    k<W><<<blocks, threads, 32768>>>(d, didx, iters, dc);
    k<W><<<blocks, threads, 32768>>>(d, didx, iters, dc);
    cudaDeviceSynchronize(); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    // per SM: 16 warps x iters x 8 loads
    printf("%-44s %6.2f cycles per warp-load (SM-wide)\n", name, (double)hc / (16.0 * iters * 8));
    cudaFree(d); cudaFree(didx); cudaFree(dc);
}
int main() {
    int a[32];
    for (int l = 0; l < 32; ++l) a[l] = (l / 4) * 4 + (((l / 4) >> 2) & 1) * 2; run<16>("LDS.128 8 distinct chunks, swizzled (4 lanes each)", a);
    for (int l = 0; l < 32; ++l) a[l] = (l / 4) * 4; run<16>("LDS.128 8 distinct chunks stride 32B", a);
    for (int l = 0; l < 32; ++l) a[l] = 0; run<16>("LDS.128 full broadcast", a);
    for (int l = 0; l < 32; ++l) a[l] = l * 2; run<16>("LDS.128 32 distinct consecutive", a);
    for (int l = 0; l < 32; ++l) a[l] = (l / 4) * 2; run<8>("LDS.64 8 distinct words (4 lanes each)", a);
    for (int l = 0; l < 32; ++l) a[l] = (l / 2) * 2; run<8>("LDS.64 16 distinct words (2 lanes each)", a);
    for (int l = 0; l < 32; ++l) a[l] = 0; run<8>("LDS.64 full broadcast", a);
    for (int l = 0; l < 32; ++l) a[l] = l * 2 ; run<8>("LDS.64 32 distinct stride 16B", a);
    for (int l = 0; l < 32; ++l) a[l] = l; run<8>("LDS.64 32 distinct consecutive (odd->even masked)", a);
    for (int l = 0; l < 32; ++l) a[l] = (l & 1) * 432 * 0 + (l / 16) * 1296 ; run<8>("LDS.64 2 distinct (2 elements)", a);
    return 0;
}

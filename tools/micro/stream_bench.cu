// stream_bench.cu -- what the memory system gives the access patterns of gather_finalize_kernel (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu && ./stream_bench
// N entries of 16 bytes; CTAs of 128 threads, one entry per thread unless stated.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_write(double2 *a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = make_double2((double)i, 1.0);
}
__global__ void k_read(const double2 *a, int64_t n, double *sink) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double2 v = i < n ? a[i] : make_double2(0, 0);
    if (v.x == -1.2345) *sink = v.y;
}
__global__ void k_copy(const double2 *b, double2 *a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = b[i];
}
template <int E>
__global__ void k_copy_multi(const double2 *b, double2 *a, int64_t n) {   // E entries per thread, all loads first
    const int64_t i0 = ((int64_t)blockIdx.x * E) * blockDim.x + threadIdx.x;
    double2 v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { const int64_t i = i0 + (int64_t)e * blockDim.x; v[e] = i < n ? b[i] : make_double2(0, 0); }
#pragma unroll
    for (int e = 0; e < E; ++e) { const int64_t i = i0 + (int64_t)e * blockDim.x; if (i < n) a[i] = v[e]; }
}
// 2 reads : 1 write (K/M value + index-sized stream -> a)
__global__ void k_triad(const double2 *b, const double2 *c, double2 *a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double2 x = b[i], y = c[i]; a[i] = make_double2(x.x + y.x, x.y + y.y); }
}
// dependent chain: per-CTA base (8 B) -> per-thread index (4 B) -> value (16 B, address from the index; idx[i] == i) -> a
__global__ void k_chain(const int64_t *base, const uint32_t *idx, const double2 *b, double2 *a, int64_t n, int levels) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c0 = (int64_t)blockIdx.x * blockDim.x;
    if (levels >= 3) c0 = base[blockIdx.x];
    uint32_t j = (uint32_t)i;
    if (levels >= 2) j = idx[c0 + threadIdx.x];
    a[i] = b[j];
}

int main() {
    const int64_t n = 200 << 20;   // 3.36 GB per array
    double2 *a, *b, *c; double *sink; int64_t *base; uint32_t *idx;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&c, n * 16); cudaMalloc(&sink, 8);
    cudaMalloc(&base, (n / 128 + 1) * 8); cudaMalloc(&idx, n * 4);
    cudaMemset(a, 0, n * 16); cudaMemset(b, 0, n * 16); cudaMemset(c, 0, n * 16);
    {   // idx[i] = i, base[k] = 128 k
        uint32_t *hi = (uint32_t *)malloc(n * 4); int64_t *hb = (int64_t *)malloc((n / 128 + 1) * 8);
        for (int64_t i = 0; i < n; ++i) hi[i] = (uint32_t)i;
        for (int64_t k = 0; k <= n / 128; ++k) hb[k] = k * 128;
        cudaMemcpy(idx, hi, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(base, hb, (n / 128 + 1) * 8, cudaMemcpyHostToDevice);
        free(hi); free(hb);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, double bytes, auto launch) {
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float t; cudaEventElapsedTime(&t, e0, e1); if (r && t < best) best = t;
        }
        printf("%-44s %8.3f ms  %7.1f GB/s\n", name, best, bytes / best / 1e6);
    };
    const unsigned g128 = (unsigned)((n + 127) / 128);
    run("write 16 B/thread, 128-thread CTAs", n * 16.0, [&] { k_write<<<g128, 128>>>(a, n); });
    run("write 16 B/thread, 512-thread CTAs", n * 16.0, [&] { k_write<<<(unsigned)((n + 511) / 512), 512>>>(a, n); });
    run("read 16 B/thread, 128-thread CTAs", n * 16.0, [&] { k_read<<<g128, 128>>>(b, n, sink); });
    run("copy 16 B/thread, 128-thread CTAs", n * 32.0, [&] { k_copy<<<g128, 128>>>(b, a, n); });
    run("copy 2 entries/thread", n * 32.0, [&] { k_copy_multi<2><<<(g128 + 1) / 2, 128>>>(b, a, n); });
    run("copy 4 entries/thread", n * 32.0, [&] { k_copy_multi<4><<<(g128 + 3) / 4, 128>>>(b, a, n); });
    run("copy 8 entries/thread", n * 32.0, [&] { k_copy_multi<8><<<(g128 + 7) / 8, 128>>>(b, a, n); });
    run("triad (2 reads : 1 write)", n * 48.0, [&] { k_triad<<<g128, 128>>>(b, c, a, n); });
    run("chain: value -> a (1 level)", n * 32.0, [&] { k_chain<<<g128, 128>>>(base, idx, b, a, n, 1); });
    run("chain: index -> value -> a (2 levels)", n * 36.0, [&] { k_chain<<<g128, 128>>>(base, idx, b, a, n, 2); });
    run("chain: base -> index -> value -> a (3 levels)", n * 36.0, [&] { k_chain<<<g128, 128>>>(base, idx, b, a, n, 3); });
    return 0;
}

// movfem_b200/csrc/finalize.cuh -- deterministic gather-reduce, per-frequency A = K + i*w32*M,
// the float32 round trip of ga_sort_sparse and the zero strip.
//
// Replaces:
//   MoVFEM_3DMT.f90:221-263  local_vfem  (a(idd) += alocal ; b(...) += blocal)
//   global_assembly.f90:43-79 assign_aij / assign_bi
//   global_assembly.f90:152-181 ga_sort_sparse  (Q10: single-precision scratch, transposed values)
//   global_assembly.f90:123-150 find_zeros / rem_zeros (Q11)
//
// Every matrix entry is summed by ONE thread over its <= 4 element contributions in ascending
// element id -- the order in which the reference's serial loop executes a(idd)=a(idd)+aij -- so the
// result does not depend on scheduling (no atomics).  Entries are already in delivery order
// (upper triangle, row-major), so no sort is needed: ga_sort_sparse's net effect is this order
// plus the float32 rounding, and "the lower-triangle value appears at the upper position".
#pragma once
#include "common.cuh"

namespace movfem {

constexpr int kGatherSub = 1;         // blocks of kFinThreads entries per gather CTA (2: twice the reads in flight per thread, -2 %; 4: +27 %)
#ifndef FIN_THREADS
#define FIN_THREADS 128
#endif
constexpr int kFinThreads = FIN_THREADS;      // entries per gather block.  Measured on config 2: 64: 0.400 ms, 128: 0.334, 256: 0.342, 512: 0.361

// Contribution index, compressed once per mesh: per block of kFinThreads entries the 64-bit position of its first
// contribution (cblk) and per entry a 16-bit offset from it (an entry has <= 4 contributions, so a block has <= 4*kFinThreads).
__global__ void __launch_bounds__(kFinThreads)
compress_cptr_kernel(int64_t nzu, const int64_t *__restrict__ cptr, int64_t *__restrict__ cblk, uint16_t *__restrict__ off16) {
    const int64_t i = (int64_t)blockIdx.x * kFinThreads + threadIdx.x;
    const int64_t b0 = cptr[(int64_t)blockIdx.x * kFinThreads];
    if (threadIdx.x == 0) {
        cblk[blockIdx.x] = b0;
        if (blockIdx.x == gridDim.x - 1) cblk[gridDim.x] = cptr[nzu];
    }
    if (i < nzu) off16[i] = (uint16_t)(cptr[i] - b0);
}

// mode 0 (T2): values rounded through float32, per-block count of surviving (non-zero) entries
// mode 1 (T1): double values, nothing stripped
// What does NOT bound this kernel (B200, 205 M entries of config 5 at half scale, 2.50 ms = 4.0 TB/s of algorithmic traffic;
// profiles/r02_summary.md): the scattered 16-byte reads (reading K/M sequentially instead: -3 %), the instruction count
// (-8.5 % instructions, issue 73 -> 66 %: +-0), the reads in flight (two blocks per CTA: -2 %, four: +27 % from the lost
// occupancy), CTA turnover (a persistent grid-stride CTA per slot: +10 to +30 %), a three-stage software pipeline over the
// blocks of a persistent CTA (+4 to +6 %), 256 entries per block (+-0; 512: +9 %), L2 prefetch of the index lines of a block 2-8 k blocks ahead (+6 %).  DRAM 48 %, L2 48 %, issue 66 %: the index -> value -> sum chain of 16 resident
// 128-entry blocks per SM is where it stands: without the sum (first contribution only) 2.17 ms, without the K/M reads 1.17 ms,
// of which 0.85 ms is the launch rate of 1.6 M 128-thread CTAs (tools/micro/stream_bench.cu: a dependent base -> index -> value
// -> store chain of this shape streams 5.1 TB/s, a plain copy 6.9 TB/s).
// Phase 1: the block's <= 4*kFinThreads contributions are fetched by all threads (independent random 16-byte reads, up to four
// in flight per thread) into shared memory; phase 2: one thread per entry sums its contributions in ascending order.
// cache: 0 none; 1 fill kmg[i] = gathered (K, M) of every entry; 2: stream_finalize_kernel (below) has delivered the entries from
// that cache -- nothing to do -- unless the node kernel saw Re(sigma) change (flags[1]), in which case this call refills it.
constexpr int kGatherMinBlocks = kGatherSub == 1 ? 2048 / kFinThreads : (kGatherSub == 2 ? 12 : 6);
// LOOP: grid-stride over the blocks (the refill launch behind stream_finalize_kernel: a small grid whose CTAs return at once
// when there is nothing to refill -- 10-30 % slower than one CTA per block when it does work, see above)
template <bool LOOP>
__global__ void __launch_bounds__(kFinThreads, kGatherMinBlocks)
gather_finalize_kernel(int64_t nzu, double w32, const int64_t *__restrict__ cblk, const uint16_t *__restrict__ off16,
                       const uint32_t *__restrict__ src, const double2 *__restrict__ KM, double2 *__restrict__ a,
                       int *__restrict__ blk_nonzero, int mode, int cache,
                       double2 *__restrict__ kmg, const int *__restrict__ flags, int nblk,
                       unsigned long long *__restrict__ total_nonzero, int NP, int W, uint32_t *__restrict__ pairflags,
                       uint32_t *__restrict__ batchany, uint32_t *__restrict__ forcek, unsigned long long *__restrict__ n_doubt,
                       double doubt_abs_k, double doubt_abs_m /* test hook: doubt every entry below these sizes; < 0 = off */) {
    // kGatherSub blocks of kFinThreads entries per CTA: twice the scattered reads in flight per thread at the same occupancy
    __shared__ double2 vals[kGatherSub][4 * kFinThreads + 4];
    __shared__ uint16_t offs[kGatherSub][kFinThreads + 1];
    if (cache == 2) {                       // a later frequency of a sweep: stream_finalize_kernel does the work unless Re(sigma) changed
        if (flags[1] == 0) return;
        cache = 1;
    }
    const int ngroups = (nblk + kGatherSub - 1) / kGatherSub;
    for (int bid = blockIdx.x; bid < ngroups; bid += LOOP ? (int)gridDim.x : ngroups) {
    int64_t c0s[kGatherSub];
    int ns[kGatherSub];
    {
#pragma unroll
        for (int sb = 0; sb < kGatherSub; ++sb) {
            const int blk = bid * kGatherSub + sb;
            c0s[sb] = 0; ns[sb] = 0;
            if (blk >= nblk) continue;
            const int64_t i = (int64_t)blk * kFinThreads + threadIdx.x;
            c0s[sb] = cblk[blk];
            ns[sb] = (int)(cblk[blk + 1] - c0s[sb]);
            if (i < nzu) offs[sb][threadIdx.x] = off16[i];
            uint32_t sidx[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = threadIdx.x + k * kFinThreads;
                sidx[k] = c < ns[sb] ? src[c0s[sb] + c] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = threadIdx.x + k * kFinThreads;
                if (c < ns[sb]) {   // LDGSTS: global -> shared without the register round trip, L1 bypassed (.cg); measured -10 % against
                                    // plain loads, __ldcs +15 %, ld.global.nc.L1::no_allocate +1 % (profiles/r02_ab_results.md)
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(&vals[sb][c]);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(KM + sidx[k]) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
#pragma unroll
    for (int sb = 0; sb < kGatherSub; ++sb) {
    const int blk = bid * kGatherSub + sb;
    if (blk >= nblk) break;
    const int64_t i = (int64_t)blk * kFinThreads + threadIdx.x;
    const int64_t c0 = c0s[sb];
    const int n = ns[sb];
    int nzflag = 0;
    {
        if (i < nzu) {
            const int lo = offs[sb][threadIdx.x];
            const int hi = (threadIdx.x + 1 < kFinThreads && i + 1 < nzu) ? offs[sb][threadIdx.x + 1] : n;
            // Contributions re-evaluated in the reference's operation order (exact.cuh) are recognised by magnitude (both parts
            // scaled by 2^-600); they carry K_e and the imaginary part with w32 already inside.  An entry made of such
            // contributions only is summed exactly as the reference's a(idd)=a(idd)+aij does, so its (0,0) test agrees.
            double k = 0.0, mm = 0.0, kx = 0.0, mx = 0.0, ak = 0.0, am = 0.0;
            int ninexact = 0, nlazy = 0;
            // Fast path (the kernel is issue bound: 326 instructions per warp before this, profiles/r02_summary.md): the <= 4
            // contributions are read without a loop or a branch (absent ones are +0.0: x + 0.0 leaves the running sum of
            // a(idd)=a(idd)+aij bit-identical) and the magnitude test is done on the exponent words; the loop below runs only in
            // a warp that holds a re-evaluated (or exactly zero) contribution
            const int nc = hi - lo;
            // unconditional loads (the staging array has four spare slots; a slot past the entry's last contribution holds
            // another entry's value or stale data and is replaced by +0.0 before any arithmetic): selects, no branches
            const double2 *vp = &vals[sb][lo];
            double2 v0 = vp[0], v1 = vp[1], v2 = vp[2], v3 = vp[3];
            auto tagged = [](const double2 &v) {   // |v.y| < 2^-500 and |v.x| < 2^-200, on the high words
                return (((unsigned)__double2hiint(v.y) & 0x7fffffffu) < 0x20b00000u) & (((unsigned)__double2hiint(v.x) & 0x7fffffffu) < 0x33700000u);
            };
            const bool slow = ((nc > 0) & tagged(v0)) | ((nc > 1) & tagged(v1)) | ((nc > 2) & tagged(v2)) | ((nc > 3) & tagged(v3));
            auto keep = [](double2 &v, bool on) {             // bitwise select (+0.0 when absent)
                const long long mk = on ? -1ll : 0ll;
                v.x = __longlong_as_double(__double_as_longlong(v.x) & mk); v.y = __longlong_as_double(__double_as_longlong(v.y) & mk);
            };
            keep(v0, nc > 0); keep(v1, nc > 1); keep(v2, nc > 2); keep(v3, nc > 3);
            if (!__any_sync(__activemask(), slow)) {
                k = (((k + v0.x) + v1.x) + v2.x) + v3.x;
                mm = (((mm + v0.y) + v1.y) + v2.y) + v3.y;
                ak = ((fabs(v0.x) + fabs(v1.x)) + fabs(v2.x)) + fabs(v3.x);
                am = ((fabs(v0.y) + fabs(v1.y)) + fabs(v2.y)) + fabs(v3.y);
                ninexact = nc;
            } else
            for (int c = lo; c < hi; ++c) {
                const double2 v = vals[sb][c];
                const double ax = fabs(v.x);
                if (fabs(v.y) < 0x1p-500 && ax < 0x1p-200) {          // re-evaluated slot: exact imaginary part, scaled by 2^-600
                    mx = mx + v.y * 0x1p+600;
                    if (ax < 0x1p-500) kx = kx + v.x * 0x1p+600;      // ... and exact K_e
                    else { kx = kx + v.x * 0x1p+300; ++nlazy; }       // ... beside the fast path's K_e (scaled by 2^-300): inexact
                } else {
                    k = k + v.x; mm = mm + v.y;
                    ak += fabs(v.x); am += fabs(v.y);
                    ++ninexact;
                }
            }
            // The entry's zero test is in doubt when it cancels ACROSS elements down to round-off (fast-path contributions), or
            // when its exact imaginary parts sum to zero while a K_e beside them is the fast path's: flag the contributions for
            // (full) re-evaluation; the host re-runs exact_kernel and this gather (none on the BASELINE meshes)
            const bool doubt_x = ninexact && ((fabs(k) <= 1e-9 * ak && fabs(mm) <= 1e-9 * am) || (fabs(k) <= doubt_abs_k && fabs(mm) <= doubt_abs_m));
            const bool doubt_l = nlazy && !ninexact && mx == 0.0;
            if ((doubt_x || doubt_l) && pairflags) {
                for (int c = lo; c < hi; ++c) {
                    const double2 v = vals[sb][c];
                    const bool reev = fabs(v.y) < 0x1p-500 && fabs(v.x) < 0x1p-200;
                    if (reev && fabs(v.x) < 0x1p-500) continue;        // fully re-evaluated already
                    const uint32_t sx = src[c0 + c];
                    const int64_t row = (int64_t)((sx >> 5) / (uint32_t)NP) * 32 + (sx & 31);
                    const int pr = (int)((sx >> 5) % (uint32_t)NP);
                    atomicOr(pairflags + row * W + (pr >> 5), 1u << (pr & 31));
                    atomicOr(forcek + row * W + (pr >> 5), 1u << (pr & 31));
                    atomicOr(batchany + (row >> 5), 1u << (row & 31));
                }
                atomicAdd(n_doubt, 1ull);
            }
            if (cache == 1) kmg[i] = make_double2(k, mm);
            double re = k + kx, im = w32 * mm + mx;
            if (mode == 0) { re = f32r(re); im = f32r(im); }
            a[i] = make_double2(re, im);
            nzflag = !(re == 0.0 && im == 0.0);
        }
    }
    const int cnt = __syncthreads_count(nzflag);
    if (threadIdx.x == 0) {
        blk_nonzero[blk] = cnt;
        // find_zeros: the number of STRIPPED entries is accumulated (integer: order-independent), so a block without exact zeros --
        // nearly every block -- touches no shared counter.  (Counting the survivors instead cost one same-address atomic per
        // 128 entries: 12.8 M serialised atomics on config 5.)
        const int valid = (int)min((int64_t)kFinThreads, nzu - (int64_t)blk * kFinThreads);
        if (total_nonzero && mode == 0 && cnt != valid) atomicAdd(total_nonzero, (unsigned long long)(valid - cnt));
    }
    // signature of the stripped set (which entries rem_zeros removes): lets MOVFEM_MODE_KEEP_PATTERN tell whether the caller's
    // irn/jcn still match the delivered pattern.  Wrapping integer sums: order independent.
    if (total_nonzero && mode == 0 && i < nzu && !nzflag) {
        atomicAdd(total_nonzero + 1, (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull);
        atomicAdd(total_nonzero + 2, ((unsigned long long)(i + 1) * 0xC2B2AE3D27D4EB4Full) ^ (unsigned long long)(i >> 7));
    }
    }
    }
}

// A later frequency of a sweep (cache == 2): every K_e, M_e is frequency independent (Q18), so every entry streams from the
// gathered cache kmg.  Four entries per thread, loads first: a CTA per 128 entries with one 16-byte load per thread is bound by
// the CTA launch rate (tools/micro/stream_bench.cu: 3.9 TB/s; 6.9 TB/s from 512-thread CTAs).  Returns at once when the node
// kernel saw Re(sigma) change: gather_finalize_kernel, launched behind it, then refills the cache.
constexpr int kStreamThreads = 256, kStreamPer = 4;
__global__ void __launch_bounds__(kStreamThreads)
stream_finalize_kernel(int64_t nzu, double w32, const double2 *__restrict__ kmg, double2 *__restrict__ a, int *__restrict__ blk_nonzero,
                       int mode, const int *__restrict__ flags, int nblk, unsigned long long *__restrict__ total_nonzero) {
    static_assert(kFinThreads == 128 && kStreamThreads == 2 * kFinThreads, "block bookkeeping below");
    if (flags[1] != 0) return;
    __shared__ int cnt[2 * kStreamPer];
    if (threadIdx.x < 2 * kStreamPer) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (kStreamThreads * kStreamPer);
    double2 v[kStreamPer];
#pragma unroll
    for (int e = 0; e < kStreamPer; ++e) {
        const int64_t i = base + e * kStreamThreads + threadIdx.x;
        v[e] = i < nzu ? kmg[i] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int e = 0; e < kStreamPer; ++e) {
        const int64_t i = base + e * kStreamThreads + threadIdx.x;
        double re = v[e].x, im = w32 * v[e].y;
        if (mode == 0) { re = f32r(re); im = f32r(im); }
        const bool live = i < nzu, nz = live && !(re == 0.0 && im == 0.0);
        if (live) a[i] = make_double2(re, im);
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if ((threadIdx.x & 31) == 0) atomicAdd(&cnt[2 * e + (threadIdx.x >> 7)], __popc(bal));
        if (total_nonzero && mode == 0 && live && !nz) {      // signature of the stripped set, as in gather_finalize_kernel
            atomicAdd(total_nonzero + 1, (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull);
            atomicAdd(total_nonzero + 2, ((unsigned long long)(i + 1) * 0xC2B2AE3D27D4EB4Full) ^ (unsigned long long)(i >> 7));
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kStreamPer) {
        const int blk = blockIdx.x * (2 * kStreamPer) + threadIdx.x;
        if (blk < nblk) {
            const int c = cnt[threadIdx.x];
            blk_nonzero[blk] = c;
            const int valid = (int)min((int64_t)kFinThreads, nzu - (int64_t)blk * kFinThreads);
            if (total_nonzero && mode == 0 && c != valid) atomicAdd(total_nonzero, (unsigned long long)(valid - c));
        }
    }
}

// Clears the pair flags of a cold pass: only the rows batchany marks carry bits (every writer of pairflags / forcek also sets
// the row's bit there), so the bitmaps -- 772 MB on config 5 -- are not memset but walked through their 4 MB summary
__global__ void clear_flags_kernel(int64_t nrows /* km_rows */, int W, const uint32_t *__restrict__ batchany, uint32_t *__restrict__ pairflags,
                                   uint32_t *__restrict__ forcek) {
    // one thread per row: a warp tests one summary word; a marked row clears its W words of both bitmaps
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows || !((batchany[row >> 5] >> (row & 31)) & 1u)) return;
    for (int w = 0; w < W; ++w) { pairflags[row * W + w] = 0u; forcek[row * W + w] = 0u; }
}

// order-preserving compaction of the entries whose value is not exactly (0,0)
__global__ void __launch_bounds__(kFinThreads)
compact_kernel(int64_t nzu, const int64_t *__restrict__ blk_off, const int *__restrict__ irn, const int *__restrict__ jcn,
               const double2 *__restrict__ a, int *__restrict__ irn_c, int *__restrict__ jcn_c, double2 *__restrict__ a_c,
               int copy_pattern /* 0: irn_c / jcn_c already hold this stripped set (same signature as the last compaction) */) {
    __shared__ int wsum[kFinThreads / 32];
    const int64_t i = (int64_t)blockIdx.x * kFinThreads + threadIdx.x;
    double2 v = make_double2(0.0, 0.0);
    if (i < nzu) v = a[i];
    const int keep = (i < nzu) && !(v.x == 0.0 && v.y == 0.0);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = __popc(bal & ((1u << lane) - 1));
    for (int k = 0; k < w; ++k) off += wsum[k];
    if (keep) {
        const int64_t p = blk_off[blockIdx.x] + off;
        if (copy_pattern) { irn_c[p] = irn[i]; jcn_c[p] = jcn[i]; }
        a_c[p] = v;
    }
}

// float32-exact values (tap T2, global_assembly.f90:157) as complex64 for the host link: half the bytes, widened back to
// complex128 on the host without loss (api.cu: pipe_d2h)
__global__ void narrow_kernel(int64_t n, const double2 *__restrict__ a, float2 *__restrict__ a32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double2 v = a[i]; a32[i] = make_float2((float)v.x, (float)v.y); }
}

// CSR row pointers of delivered (row-sorted, 1-based) triplets: rowptr[r] = first entry with irn >= row_lo + r + 1
__global__ void csr_rowptr_kernel(int nrows, int row_lo, int64_t nz, const int *__restrict__ irn, int64_t *__restrict__ rowptr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nrows) return;
    const int key = row_lo + r + 1;
    int64_t lo = 0, hi = nz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (irn[mid] < key) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = lo;
}

// A device consumer of the CSR hand-off: y = A x for the complex SYMMETRIC matrix whose upper triangle is the
// delivered result (what a GPU sparse solver's residual check or an iterative refinement step needs).  Pass 1: one warp per row
// over its stored entries (diagonal and right of it); pass 2: the mirrored entries, y_c += a_rc x_r, by atomics.
__device__ __forceinline__ double2 zmul2(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__global__ void spmv_upper_rows_kernel(int nrows, const int64_t *__restrict__ rowptr, const int *__restrict__ jcn, const double2 *__restrict__ a,
                                       const double2 *__restrict__ x, double2 *__restrict__ y) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= nrows) return;
    double sx = 0.0, sy = 0.0;
    for (int64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) {
        const double2 t = zmul2(a[k], x[jcn[k] - 1]);
        sx += t.x; sy += t.y;
    }
    for (int o = 16; o; o >>= 1) { sx += __shfl_down_sync(0xffffffffu, sx, o); sy += __shfl_down_sync(0xffffffffu, sy, o); }
    if (lane == 0) y[r] = make_double2(sx, sy);
}
__global__ void spmv_upper_mirror_kernel(int64_t nz, const int *__restrict__ irn, const int *__restrict__ jcn, const double2 *__restrict__ a,
                                         const double2 *__restrict__ x, double2 *__restrict__ y) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nz) return;
    const int r = irn[k] - 1, c = jcn[k] - 1;
    if (r == c) return;
    const double2 t = zmul2(a[k], x[r]);
    atomicAdd(&y[c].x, t.x); atomicAdd(&y[c].y, t.y);
}

// b(gne + (d-1)*nne) += blocal(d): sum of the <= 4 sharing elements in ascending element order
__global__ void rhs_kernel(int nrows, const int *__restrict__ rown, const double4 *__restrict__ be, double2 *__restrict__ rhs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;   // row local to the handle's slab; rhs is [2][nrows]
    if (r >= nrows) return;
    double b0r = 0, b0i = 0, b1r = 0, b1i = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int o = rown[(int64_t)r * 4 + k];
        if (o < 0) break;
        const double4 v = be[o];
        b0r = b0r + v.x; b0i = b0i + v.y; b1r = b1r + v.z; b1i = b1i + v.w;
    }
    rhs[r] = make_double2(b0r, b0i);
    rhs[(int64_t)nrows + r] = make_double2(b1r, b1i);
}

}  // namespace movfem

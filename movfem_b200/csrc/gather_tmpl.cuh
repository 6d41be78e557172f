// movfem_b200/csrc/gather_tmpl.cuh -- structured-mesh path of the deterministic gather (cold assemblies, OPT-IN).
//
// Status: bit-identical to the indexed gather (tests/test_gpu_parity.py::test_structured_gather_path_is_bitwise_neutral)
// but measured slower on B200 (config 2: 284 + 68 us against 320 us; config 1: 141 + 28 against 93): it removes the
// LSU-wavefront cost of the scattered reads and pays more in instructions per entry.  Enabled with
// MOVFEM_GATHER_TEMPLATE=1; kept as the starting point for a cheaper formulation.
//
// gather_finalize_kernel (finalize.cuh) follows a contribution index: every lane's 16-byte read of the K/M store is
// its own LSU wavefront, which bounds that kernel well below the HBM roof.  On the structured mesh of MoVFEM_3DMT the
// rows owned by an interior element have the same structure up to translation: entry q of element e sums the packed
// pairs p_1..p_k of the elements e + off_1..e + off_k (off in {0,1}^3), the same (off, p) list for every such e.  This
// kernel therefore runs LANES = 32 CONSECUTIVE ELEMENTS (one column segment): the K/M store is interleaved over 32
// elements, so each warp load is one contiguous 512-byte segment, no index array is read, and the results are
// transposed through shared memory so that every element's entries leave as contiguous 512-byte stores.
//
// Which elements conform is VERIFIED, not assumed: at create time template_verify_kernel compares, for every candidate
// element, its row lengths, contribution counts and contribution indices with the template extracted from one interior
// element; only exact matches use this path, so the sums (and their order: ascending element id, the reference's
// a(idd)=a(idd)+aij order) are bit-identical to the indexed gather's.  Blocks of 256 entries that contain any entry of
// a non-conforming element stay with gather_finalize_kernel (blk_generic), which also owns their non-zero counts.
#pragma once
#include "common.cuh"
#include "finalize.cuh"

namespace movfem {

constexpr int kTmplNone = 0xffff;
struct __align__(8) TmplEntry { uint16_t c[4]; };   // contribution k: (off << 11) | p, off = (dx*2+dy)*2+dz; kTmplNone = absent

__host__ __device__ __forceinline__ uint32_t km_index(int kr, int p, int NP) {
    return (uint32_t)((((int64_t)(kr >> 5) * NP + p) << 5) + (kr & 31));
}

// estart[e_local] = first entry of the first row element e owns (e = e_base + e_local, owned elements only, +1 sentinel)
__global__ void elem_entry_start_kernel(int n_own, int e_base, const int64_t *__restrict__ ebase /* global, ne+1 */, int row_lo,
                                        const int64_t *__restrict__ rowptr, int64_t *__restrict__ estart) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_own) estart[i] = rowptr[ebase[e_base + i] - row_lo];
}

// conform[e_local] = 1 iff the rows element e owns reproduce the template exactly
__global__ void template_verify_kernel(MeshDims m, int n_own, int e_base, int e_end, const int64_t *__restrict__ ebase, int row_lo,
                                       const int64_t *__restrict__ rowptr, const int64_t *__restrict__ cptr, const uint32_t *__restrict__ src,
                                       const int *__restrict__ kmrow, int NP, int nrows_t, const int *__restrict__ rowlen_t, int NQ,
                                       const TmplEntry *__restrict__ tmpl, uint8_t *__restrict__ conform) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const int e = e_base + i;
    int ie, je, ke;
    elem_ijk(m, e, ie, je, ke);
    bool ok = ie >= 2 && ie < m.nx && je >= 2 && je < m.ny && ke >= 2 && ke < m.nz && e + m.ny * m.nz + m.nz + 1 < e_end;
    if (ok) {
        const int64_t r0 = ebase[e] - row_lo, r1 = ebase[e + 1] - row_lo;
        ok = (r1 - r0) == nrows_t;
        int64_t q = 0;
        for (int r = 0; ok && r < nrows_t; ++r) {
            ok = (rowptr[r0 + r + 1] - rowptr[r0 + r]) == rowlen_t[r];
        }
        if (ok) {
            const int64_t i0 = rowptr[r0];
            const int eo[8] = {0, 1, m.nz, m.nz + 1, m.ny * m.nz, m.ny * m.nz + 1, m.ny * m.nz + m.nz, m.ny * m.nz + m.nz + 1};
            for (q = 0; ok && q < NQ; ++q) {
                const int64_t c0 = cptr[i0 + q], c1 = cptr[i0 + q + 1];
                const TmplEntry t = tmpl[q];
                int k = 0;
                for (; k < 4 && t.c[k] != kTmplNone; ++k) {
                    if (c0 + k >= c1) { ok = false; break; }
                    const int off = t.c[k] >> 11, p = t.c[k] & 0x7ff;
                    if (src[c0 + k] != km_index(kmrow[e + eo[off] - e_base], p, NP)) { ok = false; break; }
                }
                if (ok && c1 - c0 != k) ok = false;
            }
        }
    }
    conform[i] = ok ? 1 : 0;
}

// blk_generic[b] = 1 iff block b (kFinThreads entries) holds an entry of a non-conforming element
__global__ void template_block_kernel(int64_t nzu, int nblk, int e_base, int n_own, const int *__restrict__ irn, int row_lo_global_unused,
                                      const int *__restrict__ ownE, const int64_t *__restrict__ estart, const uint8_t *__restrict__ conform,
                                      uint8_t *__restrict__ blk_generic) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int64_t i0 = (int64_t)b * kFinThreads, i1 = min(nzu, i0 + (int64_t)kFinThreads);
    int el = ownE[irn[i0] - 1] - e_base;   // owner element of the block's first entry (irn is 1-based, global row)
    bool gen = false;
    while (el < n_own && estart[el] < i1) {
        if (!conform[el]) { gen = true; break; }
        ++el;
    }
    blk_generic[b] = gen ? 1 : 0;
}

constexpr int kTmplWarps = 8;

// one warp per (group of <= 32 consecutive conforming elements of one column, chunk of 32 template entries).
// blk_nonzero arrives holding the entry count of every block; exact zeros (rare) are subtracted.
__global__ void __launch_bounds__(kTmplWarps * 32)
gather_template_kernel(int ngroups, const int2 *__restrict__ groups, MeshDims m, int e_base, int NQ, const TmplEntry *__restrict__ tmpl,
                       const int64_t *__restrict__ estart, const int *__restrict__ kmrow, int NP, const double2 *__restrict__ KM,
                       double2 *__restrict__ a, const uint8_t *__restrict__ blk_generic, int *__restrict__ blk_nonzero, double w32, int mode) {
    __shared__ double2 tile[kTmplWarps][32][5];    // 4 entries per element and step (+1: conflict-free 16-byte columns)
    __shared__ int s_kr[kTmplWarps][8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (NQ + 31) / 32;
    const int64_t item = (int64_t)blockIdx.x * kTmplWarps + warp;
    const int g = (int)(item / nchunks), q0 = (int)(item % nchunks) * 32;
    if (g >= ngroups) return;
    const int2 grp = groups[g];                     // x: first element (global id), y: count
    const bool live = lane < grp.y;
    const int el = grp.x - e_base + (live ? lane : 0);
    const int eo[8] = {0, 1, m.nz, m.nz + 1, m.ny * m.nz, m.ny * m.nz + 1, m.ny * m.nz + m.nz, m.ny * m.nz + m.nz + 1};
#pragma unroll
    for (int o = 0; o < 8; ++o) s_kr[warp][o][lane] = kmrow[el + eo[o]];
    const int64_t s = estart[el];
    const int nq = min(32, NQ - q0);
    TmplEntry mine;
    mine.c[0] = mine.c[1] = mine.c[2] = mine.c[3] = kTmplNone;
    if (lane < nq) mine = tmpl[q0 + lane];
    const unsigned lo32 = mine.c[0] | ((unsigned)mine.c[1] << 16), hi32 = mine.c[2] | ((unsigned)mine.c[3] << 16);
    __syncwarp();
    // four template entries per step: their <= 16 independent loads are issued before any is consumed
    for (int qq = 0; qq < nq; qq += 4) {
        double2 v[4][4];
        unsigned code[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned tlo = __shfl_sync(0xffffffffu, lo32, (qq + u) & 31), thi = __shfl_sync(0xffffffffu, hi32, (qq + u) & 31);
            code[u][0] = tlo & 0xffffu; code[u][1] = tlo >> 16; code[u][2] = thi & 0xffffu; code[u][3] = thi >> 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                v[u][c] = make_double2(0.0, 0.0);
                if (qq + u < nq && code[u][c] != (unsigned)kTmplNone)
                    v[u][c] = KM[km_index(s_kr[warp][code[u][c] >> 11][lane], (int)(code[u][c] & 0x7ffu), NP)];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            // the indexed gather's sum, operation for operation: 0.0 + first + second ... in ascending element order
            double k = 0.0, mm = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (code[u][c] != (unsigned)kTmplNone) { k = k + v[u][c].x; mm = mm + v[u][c].y; }
            double re = k, im = w32 * mm;
            if (mode == 0) { re = f32r(re); im = f32r(im); }
            tile[warp][lane][u] = make_double2(re, im);
        }
        __syncwarp();
        // transposed write-out: 8 elements x 4 consecutive entries (64-byte runs) per store instruction
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e2 = r * 8 + (lane >> 2), u = lane & 3;
            const int64_t s2 = __shfl_sync(0xffffffffu, s, e2);
            if (e2 < grp.y && qq + u < nq) {
                const int64_t i = s2 + q0 + qq + u;
                if (!blk_generic[i / kFinThreads]) {
                    const double2 val = tile[warp][e2][u];
                    a[i] = val;
                    if (mode == 0 && val.x == 0.0 && val.y == 0.0) atomicSub(&blk_nonzero[i / kFinThreads], 1);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace movfem

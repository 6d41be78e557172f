// movfem_b200/csrc/pattern.cuh -- DOF numbering and sparsity pattern, built on the device once per mesh.
//
// Replaces:
//   global_assembly.f90:183-195,231-475  ga_cgne / c_gne12 / c_gne36 / c_gne54   -> gne(ne,me), nne
//   global_assembly.f90:197-229,488-1906 ga_nzindx / shr_nzindx12/36/54         -> the pattern
//   global_assembly.f90:81-121           ga_assemble_nze                          -> irn/jcn
//
// The reference numbers DOFs by a serial first-encounter sweep and enumerates shared matrix slots
// with 1400 lines of hand-written neighbour cases.  Here both are closed forms of the structured
// mesh: a DOF is OWNED by the lexicographically first element that contains it, owned DOFs are
// numbered by an exclusive scan over elements, and a matrix row is assembled by its owner from the
// <= 4 elements that share the DOF (owner + its +x/+y/+z neighbours).  The delivered pattern is the
// upper triangle in row-major order, which is what ZMUMPS receives after ga_sort_sparse (Q9-Q11).
#pragma once
#include "common.cuh"

namespace movfem {

// ------------------------------------------------------------------------------------------
// exclusive scan of int32 counts into int64 offsets (out has n+1 entries)
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

__global__ void scan_block_sums(const int *__restrict__ in, int64_t n, int64_t *__restrict__ bsum) {
    __shared__ int64_t red[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t s = 0;
    for (int i = threadIdx.x; i < kScanTile; i += kScanThreads)
        if (base + i < n) s += in[base + i];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += red[w];
        bsum[blockIdx.x] = t;
    }
}

__global__ void scan_block_offsets(int64_t *bsum, int nb) {   // single block, serial carry
    __shared__ int64_t buf[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        int64_t v = i < nb ? bsum[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int64_t t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        const int64_t incl = buf[threadIdx.x];
        if (i < nb) bsum[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}

__global__ void scan_finish(const int *__restrict__ in, int64_t n, const int64_t *__restrict__ bsum,
                            int64_t *__restrict__ out) {
    __shared__ int64_t wsum[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    int64_t incl = s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int64_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int64_t off = bsum[blockIdx.x];
    for (int k = 0; k < w; ++k) off += wsum[k];
    off += incl - s;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { if (base + k < n) out[base + k] = off; off += v[k]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = bsum[gridDim.x];
}

// ------------------------------------------------------------------------------------------
// DOF ownership
// ------------------------------------------------------------------------------------------
// boundary_conds.f90:262-390 edge_bdary: first matching face in the reference's if-chain order
__device__ __forceinline__ int dirichlet_face(const MeshDims &m, const ShareTables &st, int ie, int je, int ke, int im) {
    const unsigned fm = st.face[im];
    if (ie == 1 && (fm & 1u)) return 1;
    if (je == 1 && (fm & 2u)) return 2;
    if (ke == 1 && (fm & 4u)) return 3;
    if (ie == m.nx && (fm & 8u)) return 4;
    if (je == m.ny && (fm & 16u)) return 5;
    if (ke == m.nz && (fm & 32u)) return 6;
    return 0;
}

__device__ __forceinline__ bool dof_owned(const MeshDims &m, const ShareTables &st, int ie, int je, int ke, int im) {
    if (m.dirichlet && dirichlet_face(m, st, ie, je, ke, im)) return false;
    if (ie > 1 && st.back[0][im]) return false;
    if (je > 1 && st.back[1][im]) return false;
    if (ke > 1 && st.back[2][im]) return false;
    return true;
}

__global__ void gne_count_kernel(MeshDims m, const ShareTables *__restrict__ stp, int *__restrict__ cnt) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m.ne) return;
    int ie, je, ke;
    elem_ijk(m, e, ie, je, ke);
    int c = 0;
    for (int im = 1; im <= m.me; ++im) c += dof_owned(m, *stp, ie, je, ke, im);
    cnt[e] = c;
}

// one thread per (element, local DOF): resolve the owner through the -x > -y > -z copy priority of
// c_gne* (the last assignment in the reference wins) and number it from the owner's scan offset.
__global__ void gne_assign_kernel(MeshDims m, const ShareTables *__restrict__ stp, const int64_t *__restrict__ base,
                                  int *__restrict__ gne, int *__restrict__ ownE, uint8_t *__restrict__ ownL) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)m.ne * m.me) return;
    const int im = (int)(t / m.ne) + 1, e = (int)(t % m.ne);
    const ShareTables &st = *stp;
    int ie, je, ke;
    elem_ijk(m, e, ie, je, ke);
    if (m.dirichlet) {
        const int f = dirichlet_face(m, st, ie, je, ke, im);
        if (f) { gne[t] = -f; return; }
    }
    int ce = e, cim = im;
    for (;;) {
        if (ie > 1 && st.back[0][cim]) { cim = st.back[0][cim]; --ie; ce -= m.ny * m.nz; }
        else if (je > 1 && st.back[1][cim]) { cim = st.back[1][cim]; --je; ce -= m.nz; }
        else if (ke > 1 && st.back[2][cim]) { cim = st.back[2][cim]; --ke; ce -= 1; }
        else break;
    }
    int rank = 0;
    for (int k = 1; k < cim; ++k) rank += dof_owned(m, st, ie, je, ke, k);
    const int id = (int)base[ce] + rank + 1;
    gne[t] = id;
    if (ce == e && cim == im) { ownE[id - 1] = e; ownL[id - 1] = (uint8_t)im; }
}

// ------------------------------------------------------------------------------------------
// rows: one warp per global DOF r
// ------------------------------------------------------------------------------------------
constexpr int kRowWarps = 4;
constexpr uint64_t kKeyNone = ~0ull;

// packed lower-by-local-index pair id of local DOFs a,b (0-based)
__host__ __device__ __forceinline__ int pair_index(int a, int b) {
    const int hi = a > b ? a : b, lo = a > b ? b : a;
    return hi * (hi + 1) / 2 + lo;
}

template <int CAP>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t *k, int lane) {
    for (int size = 2; size <= CAP; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < CAP / 2; t += 32) {
                const int lo = 2 * t - (t & (stride - 1));   // index with bit `stride` cleared
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint64_t a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
            __syncwarp();
        }
}

/*
 * FILL=false: rowcnt[r] = number of structural upper entries (cols >= r) of row r,
 *             rowcand[r] = number of element contributions to them.
 * FILL=true : cols / irn / jcn (1-based) per entry, cptr (first contribution of each entry), src (the
 *             contributions of each entry in ASCENDING ELEMENT ORDER = the reference's summation order,
 *             encoded as the index into the K/M store), rown (the <=4 (element, local DOF) owners of row r).
 */
template <int CAP, bool FILL>
__global__ void __launch_bounds__(kRowWarps * 32)
row_kernel(MeshDims m, const ShareTables *__restrict__ stp, const int *__restrict__ gne, const int *__restrict__ ownE,
           const uint8_t *__restrict__ ownL, int row_lo, int nrows, int e_base, int NP, int *__restrict__ rowcnt, int *__restrict__ rowcand,
           const int64_t *__restrict__ row_ptr, const int64_t *__restrict__ cbase, int *__restrict__ irn,
           int *__restrict__ jcn, int64_t *__restrict__ cptr, uint32_t *__restrict__ src, int *__restrict__ rown,
           const int *__restrict__ kmrow /* element (slab-local) -> row of the K/M store, see contract.cuh */) {
    __shared__ uint64_t s_keys[kRowWarps][CAP];
    __shared__ int s_el[kRowWarps][4], s_ll[kRowWarps][4], s_n[kRowWarps];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rl = blockIdx.x * kRowWarps + w;      // row index local to this handle's slab
    if (rl >= nrows) return;
    const int r = row_lo + rl;                      // global row (0-based DOF id)
    const ShareTables &st = *stp;
    uint64_t *keys = s_keys[w];
    if (lane == 0) {
        // closure of the sharing relation starting from the owner
        int el[4], ll[4], n = 1;
        el[0] = ownE[r]; ll[0] = ownL[r];
        const int stride[3] = {m.ny * m.nz, m.nz, 1};
        for (int i = 0; i < n; ++i) {
            int ie, je, ke;
            elem_ijk(m, el[i], ie, je, ke);
            const bool has[3] = {ie < m.nx, je < m.ny, ke < m.nz};
            for (int a = 0; a < 3; ++a) {
                const int t = st.fwd[a][ll[i]];
                if (!t || !has[a]) continue;
                const int en = el[i] + stride[a];
                bool seen = false;
                for (int q = 0; q < n; ++q) seen |= (el[q] == en);
                if (!seen && n < 4) { el[n] = en; ll[n] = t; ++n; }
            }
        }
        for (int i = 1; i < n; ++i)   // ascending element id
            for (int q = i; q > 0 && el[q - 1] > el[q]; --q) {
                int t = el[q]; el[q] = el[q - 1]; el[q - 1] = t;
                t = ll[q]; ll[q] = ll[q - 1]; ll[q - 1] = t;
            }
        for (int i = 0; i < 4; ++i) { s_el[w][i] = i < n ? el[i] : -1; s_ll[w][i] = i < n ? ll[i] : 0; }
        s_n[w] = n;
    }
    __syncwarp();
    const int n = s_n[w], me = m.me, total = n * me;
    for (int idx = lane; idx < CAP; idx += 32) {
        uint64_t key = kKeyNone;
        if (idx < total) {
            const int k = idx / me, jm = idx - k * me;
            const int c = gne[(int64_t)jm * m.ne + s_el[w][k]];
            if (c >= r + 1) key = ((uint64_t)c << 16) | ((uint64_t)k << 8) | (uint64_t)jm;
        }
        keys[idx] = key;
    }
    __syncwarp();
    warp_bitonic_sort<CAP>(keys, lane);
    int nuniq = 0, ncand = 0;
    for (int base = 0; base < CAP; base += 32) {
        const int idx = base + lane;
        const uint64_t key = keys[idx];
        const bool valid = key != kKeyNone;
        const bool first = valid && (idx == 0 || (keys[idx - 1] >> 16) != (key >> 16));
        const unsigned fb = __ballot_sync(0xffffffffu, first), vb = __ballot_sync(0xffffffffu, valid);
        if (FILL && valid) {
            const int k = (int)((key >> 8) & 0xff), jm = (int)(key & 0xff), c = (int)(key >> 16);
            const int64_t cpos = cbase[rl] + idx;   // valid keys sort to the front, so idx is the rank
            // K/M store layout [row / 32][packed pair][row % 32] (32 elements interleaved: coalesced by the contraction)
            const int kr = kmrow[s_el[w][k] - e_base];
            src[cpos] = (uint32_t)((((int64_t)(kr >> 5) * NP + pair_index(s_ll[w][k] - 1, jm)) << 5) + (kr & 31));
            if (first) {
                const int64_t pos = row_ptr[rl] + nuniq + __popc(fb & ((1u << lane) - 1));
                irn[pos] = r + 1; jcn[pos] = c; cptr[pos] = cpos;
            }
        }
        nuniq += __popc(fb); ncand += __popc(vb);
    }
    if (lane == 0) {
        if (!FILL) { rowcnt[rl] = nuniq; rowcand[rl] = ncand; }
        else
            for (int i = 0; i < 4; ++i) rown[(int64_t)rl * 4 + i] = s_el[w][i] < 0 ? -1 : (int)be_index(kmrow[s_el[w][i] - e_base], me, s_ll[w][i] - 1);
    }
}

}  // namespace movfem

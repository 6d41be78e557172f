// movfem_b200/csrc/contract.cuh -- the element-matrix contraction (the FP64-bound kernel of the path).
//
// Replaces alocal / f1 / f2 (integration.f90:76-86,154-238) for all pairs of one element:
//     K_e[i][j] = sum_g  c_i(g)^T Q_e(g) c_j(g)            (GPML layers: dphi_i(g)^T P_e(g)[(.,d_i)][(.,d_j)] dphi_j(g))
//     M_e[i][j] = sum_g  phi_i(g) phi_j(g) T_e(g)[d_i][d_j]
// with the per-Gauss-point tensors Q|P and T produced by geometry_kernel (element.cuh) and the reference-element
// factors c = dphi x e_d, dphi, phi constants of the element type.
//
// Mapping.  LANES ARE ELEMENTS, WARPS ARE TILES: a warp owns one 4x4 tile (row slots 4ti.., column slots 4tj..) of
// the lower triangle in slot space and its 32 lanes carry the same tile of 32 different elements.  Both tile
// operands are then warp-uniform and are read from a shared-memory copy of the constant table with BROADCAST
// 128-bit loads (2.1 cycles per LDS.128 against 4.2 for lane-distinct addresses, tools/micro/lds_bench.cu); the only
// per-lane loads are the 2x2 block of Q (3x3 of P) and the one entry of T the tile's direction pair needs -- 5 (10)
// conflict-free LDS.64 per 60 (104) DFMA.  (The previous thread-per-tile kernel pulled 17-29 lane-distinct doubles
// per tile and Gauss point through the shared-memory return path and was bound by it at 34-38 % FP64-pipe
// utilisation; feeding the operands from the constant bank instead (LDCU -> uniform registers) is slower still:
// 2-5 cycles per constant load SM-wide, tools/micro/ldcu_bench.cu.)
//
// Data flow.  Tiles are sorted by direction-pair class (dI,dJ) (6 classes).  The Q|P,T records of a batch of 32
// elements sit in an L2-resident scratch, component-major: qt[batch][component][g][lane].  One producer lane
// streams, per (batch, class), the 5 (10) component blocks that class needs into a shared-memory ring with
// cp.async.bulk (TMA 1-D bulk copies, 6.9 kB each) signalled through mbarriers; W consumer warps walk the classes
// in order, each taking the tiles t = w (mod W) of the class, and hand the ring slot back through a second mbarrier.
// Results go out as (K,M) double2, packed lower triangle by LOCAL DOF index, the 32 elements of a batch interleaved
// ([batch][pair][lane]) so that every store instruction of a warp writes 512 contiguous bytes.
#pragma once
#include "common.cuh"

namespace movfem {

constexpr int kConUnroll = 3;   // unroll of the Gauss-point loop (1: -10 %, 9: +1.5 % and twice the code; profiles/r02_ab_results.md)

constexpr int kMaxTiles = 120;   // me=54: 15 groups of 4 slots

// Constants of the element type (one resident table per device, see api.cu: const_table_acquire)
struct ContractTables {
    double at[kMaxGp * 4 * kMaxSlots];   // [g][k][MEP]: k = 0..2 dphi/dxi_k, 3 phi, in slot order
    short slot_dof[kMaxSlots];           // slot -> local DOF (0-based) or -1
    unsigned char tile_ti[kMaxTiles], tile_tj[kMaxTiles];   // tiles sorted by class
    short cls_begin[8];                  // first tile of class c (c = 0..5), cls_begin[6] = number of tiles
    unsigned char comp[2][6][10];        // [pml][class][k]: scratch component streamed to stage block k
};
__constant__ ContractTables c_ct;
// The CTAs fill their shared-memory operand table from this global copy of c_ct.at: the fill from the constant bank is a
// lane-distinct LDC (serialised), 4 % of contract_kernel's warp samples in round 1; measured -3 % on the kernel.
__device__ double g_ct_at[kMaxGp * 4 * kMaxSlots];

// class c <-> (dI, dJ), dI >= dJ
__host__ __device__ __forceinline__ constexpr int cls_dI(int c) { return c == 0 ? 0 : (c <= 2 ? 1 : 2); }
__host__ __device__ __forceinline__ constexpr int cls_dJ(int c) { return c == 0 ? 0 : (c == 1 ? 0 : (c == 2 ? 1 : c - 3)); }

struct ContractArgs {
    const double *qt;     // [nbatch][NCMP][NGP][32]
    int nlist;            // elements in this launch (scratch order = K/M row order)
    double2 *KM;          // K/M store at this launch's first row: [batch][NP][32 lanes]
    const int *flags;     // flags[1]: Re sigma changed (cache refresh launches)
    int skip_unless_changed;
    int skip_if_simple;   // linear elements: exit when fused12_kernel handles the list (flags[0] == 0 and flags[2] == 0)
    // tiny-pair flags (exact.cuh): a pair whose K_e and M_e are both below kTinyRel of the element's scale is a round-off
    // residue of a mathematically zero entry; the reference's residue decides what rem_zeros strips, so it is re-evaluated
    const double2 *escale;          // [list position] (K scale, M scale) from geometry_kernel
    uint32_t *pairflags;            // [row][W] at this launch's first row
    uint32_t *batchany;             // [row0/32 + batch]
    unsigned long long *nflag;      // number of flagged (element, pair)s
    int W;
};
constexpr double kTinyRelC = 1e-9, kFlagAbsC = 0x1p-400;

// four consecutive doubles (16-byte aligned) as two 128-bit loads
__device__ __forceinline__ void ld4(double (&v)[4], const double *p) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <int ME_, int MEP_, int NGP_, bool PML_, int W_, int STAGES_, int MINB_ = 1>
struct ContractCfg {
    static constexpr int ME = ME_, MEP = MEP_, NGP = NGP_, W = W_, STAGES = STAGES_, MINB = MINB_;
    static constexpr bool PML = PML_;
    static constexpr int NC = PML ? 10 : 5;          // component blocks per stage
    static constexpr int NCMP = PML ? 51 : 12;       // scratch components per (element, Gauss point): P(45)|Q(6), T(6)
    static constexpr int CB = NGP * 32;              // doubles per component block
    static constexpr int STAGE_D = NC * CB;
    static constexpr int THREADS = (W + 1) * 32;     // W consumer warps + the producer warp
    static constexpr int TAB_D = NGP * 4 * MEP;      // constant operand table, broadcast-read from shared memory
    static constexpr size_t SMEM = sizeof(double) * ((size_t)STAGES * STAGE_D + TAB_D) + sizeof(uint64_t) * 2 * STAGES;
    static constexpr int NT = MEP / 4, NTILES = NT * (NT + 1) / 2, NP = ME * (ME + 1) / 2;
};

template <class CFG>
__global__ void __launch_bounds__(CFG::THREADS, CFG::MINB) contract_kernel(ContractArgs A) {
    constexpr int MEP = CFG::MEP, NGP = CFG::NGP, W = CFG::W, STAGES = CFG::STAGES, NC = CFG::NC, NCMP = CFG::NCMP;
    constexpr int CB = CFG::CB, STAGE_D = CFG::STAGE_D, NP = CFG::NP;
    constexpr bool PML = CFG::PML;
    if (A.skip_unless_changed && A.flags[1] == 0) return;
    if (A.skip_if_simple && A.flags[0] == 0 && A.flags[2] == 0) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_stage = reinterpret_cast<double *>(smem_raw);
    double *s_tab = s_stage + (size_t)STAGES * STAGE_D;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_tab + CFG::TAB_D), *empty = full + STAGES;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int nbatch = (A.nlist + 31) / 32;
    for (int i = threadIdx.x; i < CFG::TAB_D; i += CFG::THREADS) s_tab[i] = g_ct_at[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], W); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Work items are (batch, class) pairs; every CTA takes a contiguous range of the item sequence, so the grid is
    // balanced to within one class block instead of one 32-element batch.  Inside the CTA the tiles of its items form
    // one stream and warp w takes the stream positions = w (mod W).
    const int64_t nitems = (int64_t)nbatch * 6;
    const int n0 = (int)(nitems * blockIdx.x / gridDim.x), n1 = (int)(nitems * (blockIdx.x + 1) / gridDim.x);
    if (warp == W) {
        // ---- producer: one lane streams the class blocks of this CTA's items through the ring ----
        if (lane == 0) {
            for (int n = n0; n < n1; ++n) {
                const int b = n / 6, c = n - b * 6, k = n - n0;
                const double *src = A.qt + (size_t)b * NCMP * CB;
                const int slot = k % STAGES, round = k / STAGES;
                if (round > 0) mbar_wait(&empty[slot], (unsigned)((round - 1) & 1));
                mbar_expect_tx(&full[slot], (unsigned)(STAGE_D * sizeof(double)));
                double *dst = s_stage + (size_t)slot * STAGE_D;
#pragma unroll 1
                for (int q = 0; q < NC; ++q)
                    bulk_g2s(dst + q * CB, src + (size_t)c_ct.comp[PML ? 1 : 0][c][q] * CB, (unsigned)(CB * sizeof(double)), &full[slot]);
            }
        }
        return;
    }

    // ---- consumers ----
    const int pos0 = (n0 / 6) * CFG::NTILES + c_ct.cls_begin[n0 % 6];   // stream position of the CTA's first tile
    {
#pragma unroll 1
        for (int n = n0; n < n1; ++n) {
            const int b = n / 6, c = n - b * 6, k = n - n0;
            const bool live = b * 32 + lane < A.nlist;
            double2 *KMo = A.KM + (size_t)b * NP * 32 + lane;
            const int slot = k % STAGES, round = k / STAGES;
            mbar_wait(&full[slot], (unsigned)(round & 1));
            const double *S = s_stage + (size_t)slot * STAGE_D + lane;
            const int dI = cls_dI(c), dJ = cls_dJ(c);
            const int t_lo = c_ct.cls_begin[c], t_hi = c_ct.cls_begin[c + 1];
            // rows of the constant table this class reads: c_d = tau_d (dphi_k1, -dphi_k2) on the axes perpendicular to d
            //   d=0: (dphi_2, -dphi_1)   d=1: -(dphi_2, -dphi_0)   d=2: (dphi_1, -dphi_0)
            const int k1I = dI == 2 ? 1 : 2, k2I = dI == 0 ? 1 : 0, k1J = dJ == 2 ? 1 : 2, k2J = dJ == 0 ? 1 : 0;
            const double tau = ((dI == 1) != (dJ == 1)) ? -1.0 : 1.0;
            const int first = (b * CFG::NTILES + t_lo - pos0) % W;   // warp that owns the class's first tile
#pragma unroll 1
            for (int t = t_lo + ((warp - first + W) % W); t < t_hi; t += W) {
                const int ti = c_ct.tile_ti[t], tj = c_ct.tile_tj[t];
                double accK[16], accM[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { accK[i] = 0.0; accM[i] = 0.0; }
                if (!PML) {
                    const double *Y1 = s_tab + k1I * MEP + 4 * ti, *Y2 = s_tab + k2I * MEP + 4 * ti, *Y3 = s_tab + 3 * MEP + 4 * ti;
                    const double *X1 = s_tab + k1J * MEP + 4 * tj, *X2 = s_tab + k2J * MEP + 4 * tj, *X3 = s_tab + 3 * MEP + 4 * tj;
#pragma unroll kConUnroll
                    for (int g = 0; g < NGP; ++g) {
                        const int o = g * 4 * MEP;
                        const double q00 = S[(0 * NGP + g) * 32], q01 = S[(1 * NGP + g) * 32], q10 = S[(2 * NGP + g) * 32],
                                     q11 = S[(3 * NGP + g) * 32], tt = S[(4 * NGP + g) * 32];
                        double b1[4], b2[4], bw[4], xa[4], xb[4], xc[4], ya[4], yb[4], yc[4];
                        ld4(xa, X1 + o); ld4(xb, X2 + o); ld4(xc, X3 + o);
                        ld4(ya, Y1 + o); ld4(yb, Y2 + o); ld4(yc, Y3 + o);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            b1[j] = dfma(q00, xa[j], -(q01 * xb[j]));
                            b2[j] = dfma(q10, xa[j], -(q11 * xb[j]));
                            bw[j] = xc[j] * tt;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double y1 = ya[i], y2 = yb[i], y3 = yc[i];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                accK[i * 4 + j] = dfma(y1, b1[j], dfma(-y2, b2[j], accK[i * 4 + j]));
                                accM[i * 4 + j] = dfma(y3, bw[j], accM[i * 4 + j]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) accK[i] *= tau;
                } else {
                    const double *Y = s_tab + 4 * ti, *X = s_tab + 4 * tj;
#pragma unroll kConUnroll
                    for (int g = 0; g < NGP; ++g) {
                        const int o = g * 4 * MEP;
                        double P[9];
#pragma unroll
                        for (int k = 0; k < 9; ++k) P[k] = S[(k * NGP + g) * 32];
                        const double tt = S[(9 * NGP + g) * 32];
                        double bb[3][4], bw[4], xv[4][4], yv[4][4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) { ld4(xv[k], X + o + k * MEP); ld4(yv[k], Y + o + k * MEP); }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double x0 = xv[0][j], x1 = xv[1][j], x2 = xv[2][j];
#pragma unroll
                            for (int u = 0; u < 3; ++u) bb[u][j] = dfma(P[u * 3], x0, dfma(P[u * 3 + 1], x1, P[u * 3 + 2] * x2));
                            bw[j] = xv[3][j] * tt;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double y0 = yv[0][i], y1 = yv[1][i], y2 = yv[2][i], y3 = yv[3][i];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                accK[i * 4 + j] = dfma(y0, bb[0][j], dfma(y1, bb[1][j], dfma(y2, bb[2][j], accK[i * 4 + j])));
                                accM[i * 4 + j] = dfma(y3, bw[j], accM[i * 4 + j]);
                            }
                        }
                    }
                }
                // write-out: packed lower triangle by LOCAL DOF index, 32 elements interleaved -> 512-byte coalesced stores
                if (live) {
                    int nfl = 0;
                    // tiny-pair test against the element's scale (re-read per tile: nothing extra lives across the Gauss-point loop)
                    const double2 sc = A.escale ? __ldg(A.escale + b * 32 + lane) : make_double2(-1.0, -1.0);
                    const double thrK = kTinyRelC * sc.x, thrM = kTinyRelC * sc.y;
                    uint32_t *pfl = A.pairflags + (size_t)(b * 32 + lane) * A.W;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int si = 4 * ti + i, im = c_ct.slot_dof[si];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int sj = 4 * tj + j, jm = c_ct.slot_dof[sj];
                            if (im >= 0 && jm >= 0 && sj <= si) {
                                const int hi = im > jm ? im : jm, lo = im > jm ? jm : im;
                                const int p = hi * (hi + 1) / 2 + lo;
                                const double kv = accK[i * 4 + j], mv = accM[i * 4 + j];
                                KMo[p * 32] = make_double2(kv, mv);
                                const double ak = fabs(kv), am = fabs(mv);
                                if ((ak <= thrK && am <= thrM) || (ak < kFlagAbsC && am < kFlagAbsC)) {
                                    atomicOr(pfl + (p >> 5), 1u << (p & 31));
                                    ++nfl;
                                }
                            }
                        }
                    }
                    if (nfl) { atomicOr(A.batchany + b, 1u << lane); atomicAdd(A.nflag, (unsigned long long)nfl); }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    }
}

}  // namespace movfem

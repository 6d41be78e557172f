// movfem_b200/csrc/exact.cuh -- reference-order re-evaluation of the element-matrix entries that decide the
// delivered sparsity pattern (SURVEY Q11, section 7 "exact-zero set").
//
// find_zeros / rem_zeros (global_assembly.f90:123-150, caller MoVFEM_3DMT.f90:85-97) strip the entries whose
// float32-rounded value is exactly (0,0).  For the 20-node element 36 of the 666 pairs of a brick-like element are
// mathematically zero (symmetry of the quadrature), and the reference delivers their ROUND-OFF RESIDUE: on BASELINE
// config 2 1.71 M of 24.7 M entries are such residues (<= 5e-16 of the matrix scale) and 254 of them happen to cancel
// to exactly zero and are stripped.  Which ones is a property of the reference's operation order -- alocal
// (integration.f90:76-86) summing wgt*(f1 + i*w32*f2) over the Gauss points, f1 the 36-term expansion
// (integration.f90:154-209), f2 the 9-term one (integration.f90:211-238), on top of nf_jacobian / nf_inv_jac
// (n_fem.f90:355-388), mix_grad_ln / vf_elem_curl / vf_elem_ve (v_fem.f90:38-60,470-484) and p_intmodels
// (problem.f90:139-142) -- so the tensor formulation of contract.cuh cannot reproduce it.
//
// This kernel therefore re-evaluates exactly those pairs the way the reference does: IEEE double, no FMA (the library is
// built -fmad=false), the reference's association and term order, alocal(im,jm) with im the local DOF of the LARGER global
// id (MoVFEM_3DMT.f90:241-250 evaluates gne(im) >= gne(jm) only).  Which pairs: contract_kernel flags every pair whose
// K_e and M_e are both below 1e-9 of the element's scale (the residues sit 7 orders below that, the smallest genuine
// entries 4 orders above), and gather_finalize_kernel flags the contributions of any entry that cancels ACROSS elements
// (none on the BASELINE meshes).  The re-evaluated values replace the pair's slot of the K/M store, scaled by 2^-600
// (exact) so that the gather recognises them by magnitude: K_e(ref) = sum_g wgt*f1 and the imaginary part
// sum_g wgt*(w32*f2) with the float32 omega INSIDE the Gauss-point sum, as the reference rounds it -- hence the kernel
// runs every frequency on meshes that have flagged pairs.  Linear (8-node) and Lagrangian (27-node) bricks have no such
// pairs and pay one flag test per pair.
#pragma once
#include "common.cuh"
#include "element.cuh"

namespace movfem {

constexpr double kExactScale = 0x1p-600, kExactUnscale = 0x1p+600, kExactBelow = 0x1p-500, kFlagAbs = 0x1p-400;
constexpr double kLazyScale = 0x1p-300, kLazyUnscale = 0x1p+300, kLazyBelow = 0x1p-200;   // K_e of the fast path kept beside a re-evaluated imaginary part
constexpr double kTinyRel = 1e-9;

struct ExactArgs {
    MeshDims m;
    PmlParams pml;
    double omega;
    const ElemTables *T;
    const NodeRec *nodes;
    const double *xp, *yp;
    const int *list;               // element ids of this list, in K/M row order
    int nlist;
    int64_t row0;                  // K/M row of list position 0
    const uint32_t *batchany;      // [km_rows/32]: bit l = row 32*b+l has flagged pairs
    const uint32_t *pairflags;     // [km_rows][W]: bit p = packed pair p of the row is to be re-evaluated
    const uint32_t *forcek;        // [km_rows][W]: ... including its K_e even if its imaginary part is non-zero (set by the gather)
    int W, NP;
    const int *gne;                // gne(ne,me): [im][e]
    double2 *KM;                   // whole K/M store: [row/32][NP][32]
    int stretched;                 // this list goes through the GPML form of f1/f2 with h != 1 (scheme 0 layers)
};

// per (element, Gauss point) cache, the module variables of integration.f90:16-17 restated: nf_ji, wgt, mf1, Re mf2 ...
struct ExactGp {
    double ji[3][3], wgt, m1[6], m2[6];
};
// ... and gpml (stretched lists only): h1, h2, h3 and h1*h3/h2, h1*h2/h3, h2*h3/h1 (integration.f90:171-188)
struct ExactH {
    double h[3], hf[3];
};

#ifndef EXACT_EB
#define EXACT_EB 7       // measured on config 2 (1.85 M flagged pairs), ms of both launches: EB 14 / 512 items / 2 CTAs per SM 1.078,
#define EXACT_ITEMS 256  // 10 / 512 / 3: 1.081, 9 / 512 / 3: 1.027, 7 / 256 / 3: 0.937, 7 / 256 / 4 (64 registers, 152 B of spills): 0.894 --
#define EXACT_MINB 4     // the non-FMA reference-order chains want warps in flight more than they want registers
#endif
template <int MN, int ME, int NGP>
struct ExactCfg {
    // A group of up to EB flagged elements is worked on at a time, chosen so that the group's flagged pairs fill whole
    // rounds of the CTA's threads (20-node bricks have 36 such pairs each: 7 elements = 252 of 256 thread slots).
    // (the 27-node element keeps the larger groups: its flagged pairs sit in few elements -- config 3: 0.085 ms against 0.099)
    static constexpr int EB = ME == 54 ? 14 : EXACT_EB, THREADS = 256, ITEMS = ME == 54 ? 512 : EXACT_ITEMS, MINB = ME == 54 ? 2 : EXACT_MINB;
    static constexpr int NDD = 13;                       // staged per node: z, mu^-1 (6), Re sigma (6)
    static constexpr size_t SMEM = sizeof(ExactGp) * EB * NGP + sizeof(double) * (EB * MN * NDD + EB * 6) + sizeof(int) * (EB * 8 + 4 + 32 + THREADS);
    static constexpr size_t SMEM_H = sizeof(ExactH) * EB * NGP;   // added for stretched lists
};

// alocal(im, jm) of one element (integration.f90:76-86): sum over the Gauss points of wgt*(f1 + i*w32*f2) in the reference's
// operation order.  DIAG: only the 11, 22, 33 components of mu^-1 and Re sigma are non-zero anywhere in the element, so the
// terms of f1 / f2 that carry another component are exact +-0 and are left out (the sums are unchanged).
template <bool DIAG, bool WANT_K, bool WANT_M>
__device__ __forceinline__ void exact_pair(const ElemTables &T, const ExactGp *__restrict__ Pg, const ExactH *__restrict__ Hg, int ngp, int im, int jm,
                                           int di, int dj, bool gpml_form, double w32, double &are, double &aim) {
    constexpr int zm = DIAG ? 0x2929 : 0x3f3f;
    for (int g = 0; g < ngp; ++g) {
        const ExactGp &P = Pg[g];
        // mix_grad_ln v_fem.f90:478-483, grad_xi :515-518, vf_elem_curl :55-59, vf_elem_ve :41-43
        double a1[3] = {0, 0, 0}, a2[3] = {0, 0, 0}, b1[3] = {0, 0, 0}, b2[3] = {0, 0, 0}, va[3] = {0, 0, 0}, vb[3] = {0, 0, 0};
        {
            double dn[3] = {0, 0, 0}, v[3];
#pragma unroll
            for (int mm = 0; mm < 3; ++mm) {
                if (WANT_K) {
                    double sacc = 0.0;
#pragma unroll
                    for (int nn = 0; nn < 3; ++nn) sacc = sacc + P.ji[mm][nn] * T.dphi[g][im][nn];
                    dn[mm] = sacc;
                }
                v[mm] = P.ji[mm][di];
            }
            if (WANT_K) {
                a1[0] = dn[1] * v[2]; a2[0] = dn[2] * v[1];
                a1[1] = dn[2] * v[0]; a2[1] = dn[0] * v[2];
                a1[2] = dn[0] * v[1]; a2[2] = dn[1] * v[0];
            }
            if (WANT_M) {
                const double ph = T.phi[g][im];
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) va[mm] = ph * v[mm];
            }
        }
        {
            double dn[3] = {0, 0, 0}, v[3];
#pragma unroll
            for (int mm = 0; mm < 3; ++mm) {
                if (WANT_K) {
                    double sacc = 0.0;
#pragma unroll
                    for (int nn = 0; nn < 3; ++nn) sacc = sacc + P.ji[mm][nn] * T.dphi[g][jm][nn];
                    dn[mm] = sacc;
                }
                v[mm] = P.ji[mm][dj];
            }
            if (WANT_K) {
                b1[0] = dn[1] * v[2]; b2[0] = dn[2] * v[1];
                b1[1] = dn[2] * v[0]; b2[1] = dn[0] * v[2];
                b1[2] = dn[0] * v[1]; b2[2] = dn[1] * v[0];
            }
            if (WANT_M) {
                const double ph = T.phi[g][jm];
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) vb[mm] = ph * v[mm];
            }
        }
        const double *mu = P.m1, *sg = P.m2;
        double v1 = 0.0, v2 = 0.0;
        // f1, integration.f90:171-207: A(p,s) = cv1 of im, B(q,t) = cv2 of jm
#define MOVFEM_T(sign, hfac, mk, aa, bb)                                            \
    if (zm & (1 << (mk))) { const double t_ = (((hfac) * mu[mk]) * (aa)) * (bb); r = (sign) > 0 ? r + t_ : r - t_; }
#define MOVFEM_U(sign, mk, aa, bb)                                                  \
    if (zm & (1 << (mk))) { const double t_ = (mu[mk] * (aa)) * (bb); r = (sign) > 0 ? r + t_ : r - t_; }
        if (gpml_form) {
            double h1 = 1.0, h2 = 1.0, h3 = 1.0, f13_2 = 1.0, f12_3 = 1.0, f23_1 = 1.0;   // unstretched: (1*1)/1 = 1 exactly
            if (Hg) {
                const ExactH &H = Hg[g];
                h1 = H.h[0]; h2 = H.h[1]; h3 = H.h[2]; f13_2 = H.hf[0]; f12_3 = H.hf[1]; f23_1 = H.hf[2];
            }
            double r = 0.0;
            if (WANT_K) {
            MOVFEM_T(+1, f13_2, 0, a1[0], b1[0]) MOVFEM_T(-1, h1, 0, a2[0], b1[0]) MOVFEM_T(-1, h1, 0, a1[0], b2[0]) MOVFEM_T(+1, f12_3, 0, a2[0], b2[0])
            MOVFEM_T(+1, h1, 1, a1[0], b1[1]) MOVFEM_T(-1, f12_3, 1, a2[0], b1[1]) MOVFEM_T(-1, h3, 1, a1[0], b2[1]) MOVFEM_T(+1, h2, 1, a2[0], b2[1])
            MOVFEM_T(+1, h3, 2, a1[0], b1[2]) MOVFEM_T(-1, h2, 2, a2[0], b1[2]) MOVFEM_T(-1, f13_2, 2, a1[0], b2[2]) MOVFEM_T(+1, h1, 2, a2[0], b2[2])
            MOVFEM_T(+1, h1, 1, a1[1], b1[0]) MOVFEM_T(-1, h3, 1, a2[1], b1[0]) MOVFEM_T(-1, f12_3, 1, a1[1], b2[0]) MOVFEM_T(+1, h2, 1, a2[1], b2[0])
            MOVFEM_T(+1, f12_3, 3, a1[1], b1[1]) MOVFEM_T(-1, h2, 3, a2[1], b1[1]) MOVFEM_T(-1, h2, 3, a1[1], b2[1]) MOVFEM_T(+1, f23_1, 3, a2[1], b2[1])
            MOVFEM_T(+1, h2, 4, a1[1], b1[2]) MOVFEM_T(-1, f23_1, 4, a2[1], b1[2]) MOVFEM_T(-1, h1, 4, a1[1], b2[2]) MOVFEM_T(+1, h3, 4, a2[1], b2[2])
            MOVFEM_T(+1, h3, 2, a1[2], b1[0]) MOVFEM_T(-1, f13_2, 2, a2[2], b1[0]) MOVFEM_T(-1, h2, 2, a1[2], b2[0]) MOVFEM_T(+1, h1, 2, a2[2], b2[0])
            MOVFEM_T(+1, h2, 4, a1[2], b1[1]) MOVFEM_T(-1, h1, 4, a2[2], b1[1]) MOVFEM_T(-1, f23_1, 4, a1[2], b2[1]) MOVFEM_T(+1, h3, 4, a2[2], b2[1])
            MOVFEM_T(+1, f23_1, 5, a1[2], b1[2]) MOVFEM_T(-1, h3, 5, a2[2], b1[2]) MOVFEM_T(-1, h3, 5, a1[2], b2[2]) MOVFEM_T(+1, f13_2, 5, a2[2], b2[2])
            }
            v1 = r;
            // f2, integration.f90:228-232: h1*h2*h3*cv2(q)*m(k)*cv1(p), nine terms summed left to right
            const double hhh = (h1 * h2) * h3;
            double q = 0.0;
            if (WANT_M) {
#define MOVFEM_M(mk, bq, ap) if (zm & (256 << (mk))) q = q + ((hhh * vb[bq]) * sg[mk]) * va[ap];
            MOVFEM_M(0, 0, 0) MOVFEM_M(1, 1, 0) MOVFEM_M(2, 2, 0) MOVFEM_M(1, 0, 1) MOVFEM_M(3, 1, 1) MOVFEM_M(4, 2, 1)
            MOVFEM_M(2, 0, 2) MOVFEM_M(4, 1, 2) MOVFEM_M(5, 2, 2)
            }
#undef MOVFEM_M
            v2 = q;
        } else {
            double r = 0.0;
            if (WANT_K) {
            MOVFEM_U(+1, 0, a1[0], b1[0]) MOVFEM_U(-1, 0, a2[0], b1[0]) MOVFEM_U(-1, 0, a1[0], b2[0]) MOVFEM_U(+1, 0, a2[0], b2[0])
            MOVFEM_U(+1, 1, a1[0], b1[1]) MOVFEM_U(-1, 1, a2[0], b1[1]) MOVFEM_U(-1, 1, a1[0], b2[1]) MOVFEM_U(+1, 1, a2[0], b2[1])
            MOVFEM_U(+1, 2, a1[0], b1[2]) MOVFEM_U(-1, 2, a2[0], b1[2]) MOVFEM_U(-1, 2, a1[0], b2[2]) MOVFEM_U(+1, 2, a2[0], b2[2])
            MOVFEM_U(+1, 1, a1[1], b1[0]) MOVFEM_U(-1, 1, a2[1], b1[0]) MOVFEM_U(-1, 1, a1[1], b2[0]) MOVFEM_U(+1, 1, a2[1], b2[0])
            MOVFEM_U(+1, 3, a1[1], b1[1]) MOVFEM_U(-1, 3, a2[1], b1[1]) MOVFEM_U(-1, 3, a1[1], b2[1]) MOVFEM_U(+1, 3, a2[1], b2[1])
            MOVFEM_U(+1, 4, a1[1], b1[2]) MOVFEM_U(-1, 4, a2[1], b1[2]) MOVFEM_U(-1, 4, a1[1], b2[2]) MOVFEM_U(+1, 4, a2[1], b2[2])
            MOVFEM_U(+1, 2, a1[2], b1[0]) MOVFEM_U(-1, 2, a2[2], b1[0]) MOVFEM_U(-1, 2, a1[2], b2[0]) MOVFEM_U(+1, 2, a2[2], b2[0])
            MOVFEM_U(+1, 4, a1[2], b1[1]) MOVFEM_U(-1, 4, a2[2], b1[1]) MOVFEM_U(-1, 4, a1[2], b2[1]) MOVFEM_U(+1, 4, a2[2], b2[1])
            MOVFEM_U(+1, 5, a1[2], b1[2]) MOVFEM_U(-1, 5, a2[2], b1[2]) MOVFEM_U(-1, 5, a1[2], b2[2]) MOVFEM_U(+1, 5, a2[2], b2[2])
            }
            v1 = r;
            // f2 Dirichlet form, integration.f90:234-236: three parenthesised rows (every product is formed: +-0 where a component
            // is zero, the association of the reference is kept)
            if (WANT_M) {
                double rows[3];
#pragma unroll
                for (int pp = 0; pp < 3; ++pp) {
                    const int k0 = sym3(0, pp), k1 = sym3(1, pp), k2 = sym3(2, pp);
                    rows[pp] = ((vb[0] * sg[k0]) * va[pp] + (vb[1] * sg[k1]) * va[pp]) + (vb[2] * sg[k2]) * va[pp];
                }
                v2 = (rows[0] + rows[1]) + rows[2];
            }
        }
#undef MOVFEM_T
#undef MOVFEM_U
        // alocal, integration.f90:84: a = a + wgt*(f1 + cmplx(0,omega)*f2), cmplx() single precision (Q2)
        if (WANT_K) are = are + P.wgt * v1;
        if (WANT_M) aim = aim + P.wgt * (w32 * v2);
    }
}

template <int MN, int ME, int NGP>
__global__ void __launch_bounds__(256, (ExactCfg<MN, ME, NGP>::MINB)) exact_kernel(ExactArgs A) {
    using CFG = ExactCfg<MN, ME, NGP>;
    constexpr int EB = CFG::EB, NDD = CFG::NDD, NORD = MN == 8 ? 2 : 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ExactGp *s_gp = reinterpret_cast<ExactGp *>(smem_raw);                    // [EB][NGP]
    double *s_nd = reinterpret_cast<double *>(s_gp + EB * NGP);               // [EB][MN][NDD]
    double *s_xy = s_nd + EB * MN * NDD;                                      // [EB][6]: xs[3], ys[3]
    int *s_el = reinterpret_cast<int *>(s_xy + EB * 6);                       // [EB][8]: element, lane, flags[3], npairs, prefix, zero masks
    int *s_n = s_el + EB * 8;                                                 // [4]
    int *s_cnt = s_n + 4;                                                     // [32] flagged pairs of every lane of the batch
    uint32_t *s_any = reinterpret_cast<uint32_t *>(s_cnt + 32);               // [THREADS] batch words of the next 256 steps of the walk
    ExactH *s_h = reinterpret_cast<ExactH *>(s_any + CFG::THREADS);           // [EB][NGP], stretched lists only

    const ElemTables &T = *A.T;
    const MeshDims &m = A.m;
    const int tid = threadIdx.x;
    const bool gpml_form = !m.dirichlet;   // integration.f90:169,224: the GPML form of f1/f2 is taken whenever dirichlet is false
    const double w32 = f32r(A.omega);
    const int nbatch = (A.nlist + 31) / 32;

    // The CTA walks the batches b = blockIdx.x + k*gridDim.x.  Meshes without residue pairs (all linear-element BASELINE
    // configs) only pay a scan: 256 steps of the walk are looked at at once, one batch word per thread.
    for (int k0 = 0; (int64_t)blockIdx.x + (int64_t)k0 * gridDim.x < nbatch; k0 += CFG::THREADS) {
        __syncthreads();   // s_any of the previous 256 steps no longer read
        const int64_t bmine = (int64_t)blockIdx.x + (int64_t)(k0 + tid) * gridDim.x;
        const uint32_t mine = bmine < nbatch ? A.batchany[(A.row0 >> 5) + bmine] : 0u;
        s_any[tid] = mine;
        if (!__syncthreads_or(mine != 0u)) continue;
    for (int kk = 0; kk < CFG::THREADS; ++kk) {
        const int64_t b64 = (int64_t)blockIdx.x + (int64_t)(k0 + kk) * gridDim.x;
        if (b64 >= nbatch) break;
        const int b = (int)b64;
        const int64_t brow = A.row0 + (int64_t)b * 32;
        const uint32_t any = s_any[kk];
        if (any == 0) continue;
        __syncthreads();   // previous batch consumed
        if (tid < 32) {
            int c = 0;
            if (((any >> tid) & 1u) && b * 32 + tid < A.nlist) {
                const uint32_t *fl = A.pairflags + (brow + tid) * A.W;
                for (int w = 0; w < A.W; ++w) c += __popc(fl[w]);
            }
            s_cnt[tid] = c;
        }
        if (tid == 0) s_n[2] = 0;   // next lane to look at
        for (;;) {
            __syncthreads();   // previous group consumed; s_cnt / s_n[2] visible
            if (tid == 0) {
                int n = 0, total = 0, l = s_n[2];
                for (; l < 32 && n < EB; ++l) {
                    const int c = s_cnt[l];
                    if (c == 0) continue;
                    if (n > 0 && total + c > CFG::ITEMS) break;
                    s_el[n * 8] = A.list[b * 32 + l];
                    s_el[n * 8 + 1] = l;
                    s_el[n * 8 + 5] = c; s_el[n * 8 + 6] = total; s_el[n * 8 + 7] = 0;
                    total += c;
                    ++n;
                }
                s_n[2] = l; s_n[0] = n; s_n[1] = total;
            }
            __syncthreads();
            const int n = s_n[0];
            if (n == 0) break;
            // ---- phase 0: stage the node data of the group's elements; flags; pair counts ----
            if (tid < n) {
                const int e = s_el[tid * 8];
                int f[3] = {0, 0, 0};
                if (A.stretched) effective_pml(m, A.pml, e, f);
                s_el[tid * 8 + 2] = f[0]; s_el[tid * 8 + 3] = f[1]; s_el[tid * 8 + 4] = f[2];
            }
            for (int i = tid; i < n * MN; i += CFG::THREADS) {
                const int s = i / MN, l = i % MN;
                int ie, je, ke;
                elem_ijk(m, s_el[s * 8], ie, je, ke);
                const int g1 = m.nord - 1;
                const int64_t id0 = (int64_t)(ie - 1) * g1 * m.nyz + (int64_t)(je - 1) * g1 * m.nnz + (ke - 1) * g1;
                const NodeRec &nr = A.nodes[id0 + T.node_off[l]];
                double *d = s_nd + (s * MN + l) * NDD;
                d[0] = nr.z;
#pragma unroll
                for (int k = 0; k < 6; ++k) { d[1 + k] = nr.inmu[k]; d[7 + k] = nr.sre[k]; }
                if (l < 3) {
                    s_xy[s * 6 + l] = l < NORD ? A.xp[(ie - 1) * g1 + l] : 0.0;
                    s_xy[s * 6 + 3 + l] = l < NORD ? A.yp[(je - 1) * g1 + l] : 0.0;
                }
            }
            __syncthreads();
            // ---- phase 1: one thread per (Gauss point, element): int_elem_params, integration.f90:60-74,108-137 ----
            for (int i = tid; i < n * NGP; i += CFG::THREADS) {
                const int g = i / n, s = i % n;
                const double *nd = s_nd + s * MN * NDD;
                ExactGp &P = s_gp[s * NGP + g];
                // nf_jacobian, n_fem.f90:359-366
                double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, xg[3] = {0, 0, 0};
                for (int l = 0; l < MN; ++l) {
                    const double x = s_xy[s * 6 + T.node_i[l]], y = s_xy[s * 6 + 3 + T.node_j[l]], z = nd[l * NDD];
#pragma unroll
                    for (int mm = 0; mm < 3; ++mm) {
                        const double dn = T.dN[g][l][mm];
                        J[mm][0] = J[mm][0] + dn * x; J[mm][1] = J[mm][1] + dn * y; J[mm][2] = J[mm][2] + dn * z;
                    }
                    const double ln = T.N[g][l];   // g_rw, integration.f90:120-124
                    xg[0] = xg[0] + ln * x; xg[1] = xg[1] + ln * y; xg[2] = xg[2] + ln * z;
                }
                // nf_det n_fem.f90:393-394, wgt integration.f90:71, nf_inv_jac n_fem.f90:378-386 (Q6: / dabs(det))
                const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                                   J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
                P.wgt = det * T.rw[g][3];
                const double ad = fabs(det);
                P.ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / ad;
                P.ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / ad;
                P.ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / ad;
                P.ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / ad;
                P.ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / ad;
                P.ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / ad;
                P.ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / ad;
                P.ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / ad;
                P.ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / ad;
                // p_intmodels, problem.f90:139-142 (accumulating into zeroed mf1 / mf2; only Re mf2 reaches f2, Q3)
                double m1[6] = {0, 0, 0, 0, 0, 0}, m2[6] = {0, 0, 0, 0, 0, 0};
                for (int l = 0; l < MN; ++l) {
                    const double ln = T.N[g][l];
#pragma unroll
                    for (int k = 0; k < 6; ++k) { m1[k] = m1[k] + ln * nd[l * NDD + 1 + k]; m2[k] = m2[k] + ln * nd[l * NDD + 7 + k]; }
                }
                int zm = 0;   // which tensor components are non-zero at some Gauss point: a zero factor makes a term of f1 / f2 an exact +-0
#pragma unroll
                for (int k = 0; k < 6; ++k) { P.m1[k] = m1[k]; P.m2[k] = m2[k]; zm |= (m1[k] != 0.0 ? 1 << k : 0) | (m2[k] != 0.0 ? 256 << k : 0); }
                atomicOr(&s_el[s * 8 + 7], zm);
                // gpml(i,:) = Re gpml_h (integration.f90:16,125, Q18) with the lagging flags (Q17)
                double h1 = 1.0, h2 = 1.0, h3 = 1.0;
                if (A.stretched) {
                    h1 = gpml_axis(A.pml, s_el[s * 8 + 2], 0, xg[0], A.omega).x;
                    h2 = gpml_axis(A.pml, s_el[s * 8 + 3], 1, xg[1], A.omega).x;
                    h3 = gpml_axis(A.pml, s_el[s * 8 + 4], 2, xg[2], A.omega).x;
                }
                if (A.stretched) {
                    ExactH &H = s_h[s * NGP + g];
                    H.h[0] = h1; H.h[1] = h2; H.h[2] = h3;
                    H.hf[0] = (h1 * h3) / h2; H.hf[1] = (h1 * h2) / h3; H.hf[2] = (h2 * h3) / h1;
                }
            }
            __syncthreads();
            // ---- phase 2: one thread per flagged (element, pair): alocal, integration.f90:76-86 ----
            const int total = s_n[1];
            int gmask = 0;
            for (int q = 0; q < n; ++q) gmask |= s_el[q * 8 + 7];
            const bool diag = (gmask & ~0x2929) == 0;   // bits 0-5: mu^-1 components, 8-13: Re sigma; 0x29 = {11, 22, 33}
            for (int item = tid; item < total; item += CFG::THREADS) {
                int s = 0;
                while (s + 1 < n && s_el[(s + 1) * 8 + 6] <= item) ++s;
                int rank = item - s_el[s * 8 + 6];
                const int e = s_el[s * 8], lane = s_el[s * 8 + 1];
                const uint32_t *fl = A.pairflags + (brow + lane) * A.W;
                int p = -1;
                for (int w = 0; w < A.W; ++w) {
                    const uint32_t word = fl[w];
                    const int c = __popc(word);
                    if (rank < c) { p = w * 32 + (int)__fns(word, 0, rank + 1); break; }
                    rank -= c;
                }
                if (p < 0 || p >= A.NP) continue;
                int hi = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
                while (hi * (hi + 1) / 2 > p) --hi;
                while ((hi + 1) * (hi + 2) / 2 <= p) ++hi;
                const int lo = p - hi * (hi + 1) / 2;
                const int gh = A.gne[(size_t)hi * m.ne + e], gl = A.gne[(size_t)lo * m.ne + e];
                if (gh <= 0 || gl <= 0) continue;          // Dirichlet DOF: the pair never reaches the matrix
                const int im = gh >= gl ? hi : lo, jm = gh >= gl ? lo : hi;
                const int di = T.edir[im], dj = T.edir[jm];
                double are = 0.0, aim = 0.0;
                const ExactGp *Pg = s_gp + s * NGP;
                const ExactH *Hg = A.stretched ? s_h + s * NGP : nullptr;
                // isotropic-diagonal mu^-1 and Re sigma (components 11, 22, 33 only) in the whole group: 12 of the 36 terms of f1
                // and 3 of the 9 of f2 can be non-zero; otherwise every term is evaluated (an exact zero factor adds +-0)
                // The imaginary part first (f2: a quarter of the work).  If it is non-zero the entry survives whatever K_e is, so
                // the fast path's K_e stays -- marked as NOT re-evaluated by its scaling (2^-300 instead of 2^-600) -- unless the
                // gather found an entry whose exact imaginary parts cancel and asked for this pair's K_e (forcek).
                if (diag) exact_pair<true, false, true>(T, Pg, Hg, NGP, im, jm, di, dj, gpml_form, w32, are, aim);
                else exact_pair<false, false, true>(T, Pg, Hg, NGP, im, jm, di, dj, gpml_form, w32, are, aim);
                const int64_t row = brow + lane;
                double2 *slot = A.KM + ((((row >> 5) * A.NP + p) << 5) + (row & 31));
                const double kold = slot->x, akold = fabs(kold);
                const bool forced = (A.forcek[(brow + lane) * A.W + (p >> 5)] >> (p & 31)) & 1u;
                // the slot's K_e: re-evaluated at an earlier frequency (0 < |.| < 2^-500; K_e(ref) does not depend on omega: keep),
                // the fast path's value marked lazy (2^-500 <= |.| < 2^-200) or unmarked, or an exact zero (always evaluated)
                double kout;
                bool evaluate = forced || aim == 0.0 || akold == 0.0;
                if (akold != 0.0 && akold < kExactBelow) { kout = kold; evaluate = false; }
                else if (!evaluate) {
                    kout = akold < kLazyBelow ? kold : kold * kLazyScale;
                    if (!(fabs(kout) >= kExactBelow && fabs(kout) < kLazyBelow)) evaluate = true;   // cannot carry the mark
                }
                if (evaluate) {
                    if (diag) exact_pair<true, true, false>(T, Pg, Hg, NGP, im, jm, di, dj, gpml_form, w32, are, aim);
                    else exact_pair<false, true, false>(T, Pg, Hg, NGP, im, jm, di, dj, gpml_form, w32, are, aim);
                    kout = are * kExactScale;
                }
                *slot = make_double2(kout, aim * kExactScale);
            }
        }
    }
    }
}

}  // namespace movfem

// movfem_b200/csrc/fused12.cuh -- the 8-node / 12-DOF (linear) element path in ONE warp-specialised kernel.
//
// BASELINE configs 1, 4 and 5 are linear-element meshes (5: 32 M elements).  There the two-kernel element path of
// element.cuh + contract.cuh is bound by its own HBM round trips and block barriers, not by arithmetic: per element
// 768 B of Q|T scratch are written by geometry_kernel and re-read by contract_kernel next to 1 248 B of K_e/M_e --
// against 10 944 algorithmic flops -- and ncu shows 30 % of the warp samples at barriers or waiting for the
// staged node records.  This kernel keeps the per-Gauss-point tensors on chip and has NO block-wide barrier in its loop:
//
//   batch = 32 consecutive elements of the unstretched list (LANES ARE ELEMENTS everywhere), CTA = 8 warps
//   producers (warps 6, 7; four Gauss points each): the node fields of a batch arrive by cp.async (16 bytes per lane and
//       chunk), issued one batch ahead; interpolation to the Gauss points, then J,
//       det, G = J^-1, Q = (w/det^2) J mu^-1 J^T, T = G^T S G and the source R in CLOSED FORM: x and y are tensor-product
//       lines, so J = [[a,0,p],[0,b,q],[0,0,r]] with a = dx/2, b = dy/2 and (p,q,r) the xi-gradient of z (the n_fem.f90:193
//       typo is in the table).  Q|T go to one of two shared-memory buffers [component][g][lane], signalled by an mbarrier;
//       R stays in a third block from which the same two warps form b_e = sum_g phi_j R[d_j] (integration.f90:96-104), one
//       polarisation each
//   consumers (warps 0-5): one 4x4 tile of the lower triangle each (the 12 DOFs are three direction-uniform groups of four,
//       so the six tiles are the six direction classes), operands warp-uniform broadcast loads -- the inner loop of
//       contract.cuh -- from the buffer the producers filled while the previous batch was contracted; K_e, M_e leave as
//       512-byte coalesced stores, with the tiny-pair test of exact.cuh in the epilogue
//
// The closed forms hold when mu = mu0 I and sigma is diagonal at every node (node_kernel reports both; every
// linear-element BASELINE mesh): otherwise the kernel returns at once and the generic two-kernel path does the work
// (api.cu launches both; exactly one of them runs).  They agree with the reference's arithmetic to rounding (<= 1e-15);
// the entries whose last bits matter are re-evaluated in the reference's own operation order by exact_kernel.
// The RHS-only pass of a cached frequency (DO_KM = false) is the same code with the two producer warps only.
//
// Measured on the way (B200, 4 M elements of config 5 at half scale, ms of this kernel; profiles/r02_summary.md): block-phased
// version (all warps geometry, then all warps contraction, TMA-staged records) 4.30; the same with a tensor-core (DMMA)
// interpolation phase and a reused buffer 5.30 (two more block barriers); warp-specialised with direct global loads 3.77;
// + cp.async staging 3.67; + no global-load chain at the top of a batch, one reciprocal per Gauss point, no sums of
// identically-zero source components, lower half only of diagonal tiles 3.58; + element ids two batches ahead and a partly
// rolled node loop 3.36 (this file); four producer warps of two Gauss points (10 warps, 96 registers, spills in the tile
// loop) 4.05; b_e formed by four of the consumer warps from a double-buffered R block 3.38 (no gain: the producers wait for
// their scattered 16-byte node-field reads, 27 sectors per request, not for their own instructions -- the next step is a
// field-major node layout so that a warp's request is 256 contiguous bytes).  geometry_kernel + contract_kernel: 5.59.
//
// Replaces for linear elements: MoVFEM_3DMT.f90:193-211 (element loop body), integration.f90:60-106 (int_elem_params,
// alocal, blocal), n_fem.f90:66-102,355-395, v_fem.f90:38-60, problem.f90:70-149.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "contract.cuh"
#include "element.cuh"

namespace movfem {

struct Fused12Args {
    MeshDims m;
    double omega;
    const ElemTables *T;
    const NodeRec *nodes;
    const double *xp, *yp;
    const int *list;               // element ids (0-based) of this launch, K/M row order
    int nlist;
    int e_base;
    double2 *KM;                   // K/M store at this launch's first row: [batch][78][32]
    double *be;                    // [element - e_base][12][4]
    int *status;
    const int *flags;              // flags[0]: any dmu != 0, [1]: Re sigma changed, [2]: any off-diagonal sigma (node_kernel)
    uint32_t *pairflags;           // [row][W] at this launch's first row
    uint32_t *batchany;            // at this launch's first batch
    unsigned long long *nflag;
    int W;
    int no_l1;                     // test hook: no element-level tiny-pair flags
    int skip_unless_changed;       // launch is a cache refresh: exit unless flags[1] (Re sigma changed)
};

template <bool DO_KM>
struct Fused12Cfg {
    static constexpr int MN = 8, ME = 12, NGP = 8, NP = 78;
    static constexpr int NW = DO_KM ? 8 : 2, THREADS = NW * 32, MINB = DO_KM ? 2 : 8;
    static constexpr size_t BLK_D = (size_t)12 * NGP * 32;                  // one [component][g][lane] block of doubles
    static constexpr size_t TAB_D = (size_t)NGP * 4 * ME, DN_D = (size_t)MN * 4 * NGP, PHI_D = (size_t)NGP * ME;
    static constexpr int NCH = MN * 6;                                      // staged 16-byte chunks per element: 6 per node record
    static constexpr size_t STG_D = (size_t)NCH * 32 * 2;                   // node-field staging [chunk][lane] of double2
    static constexpr size_t SMEM = sizeof(double) * ((DO_KM ? 3 : 1) * BLK_D + STG_D + TAB_D + DN_D + PHI_D) + sizeof(unsigned long long) * (2 * 32 * 2 + 4) +
                                   sizeof(int) * (2 * ME + MN);
};

__device__ __forceinline__ void bar_sync_producers() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

template <bool DO_KM>
__global__ void __launch_bounds__(Fused12Cfg<DO_KM>::THREADS, Fused12Cfg<DO_KM>::MINB) fused12_kernel(Fused12Args A) {
    using C = Fused12Cfg<DO_KM>;
    constexpr int MN = C::MN, ME = C::ME, NGP = C::NGP, NP = C::NP, NW = C::NW;
    constexpr int QS = NGP * 32;                                      // component stride inside a block
    if (A.skip_unless_changed && A.flags[1] == 0) return;
    if (A.flags[0] != 0 || A.flags[2] != 0) return;                   // mu != mu0 or off-diagonal sigma: the generic path runs instead

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_R = reinterpret_cast<double *>(smem_raw);               // [12 = d*4 + (pol, re|im)][NGP][32]
    double *s_qt = s_R + C::BLK_D;                                    // [2][12][NGP][32]: Q (0-5, sym3 order), T (6-11)   (DO_KM)
    double2 *s_stage = reinterpret_cast<double2 *>(s_R + (DO_KM ? 3 : 1) * C::BLK_D);   // [48 = node*6 + field][32 lanes]: (z,e), sigma re 01|23|45, im 01|23
    double *s_tab = s_R + (DO_KM ? 3 : 1) * C::BLK_D + C::STG_D;      // [NGP][4][ME]: dphi (0-2), phi (3), slot order
    double *s_dN = s_tab + C::TAB_D;                                  // [MN][4][NGP]: dN/dxi (0-2), N (3)
    double *s_phi = s_dN + C::DN_D;                                   // [NGP][ME] phi in slot order
    unsigned long long *s_scale = reinterpret_cast<unsigned long long *>(s_phi + C::PHI_D);   // [2][32][2]
    uint64_t *full = reinterpret_cast<uint64_t *>(s_scale + 2 * 32 * 2), *empty = full + 2;
    int *s_slot = reinterpret_cast<int *>(empty + 2);                 // [ME] slot -> local DOF
    int *s_sdir = s_slot + ME;                                        // [ME] slot -> direction
    int *s_noff = s_sdir + ME;                                        // [MN] node offset from the element's base node

    const ElemTables &T = *A.T;
    const MeshDims &m = A.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    for (int i = tid; i < NGP * ME; i += C::THREADS) s_phi[i] = T.phi[i / ME][T.slot_dof[i % ME]];
    if (DO_KM)
        for (int i = tid; i < (int)C::TAB_D; i += C::THREADS) s_tab[i] = g_ct_at[i];
    for (int i = tid; i < ME; i += C::THREADS) { s_slot[i] = T.slot_dof[i]; s_sdir[i] = T.slot_dir[i]; }
    for (int i = tid; i < MN; i += C::THREADS) s_noff[i] = T.node_off[i];
    for (int i = tid; i < MN * 4 * NGP; i += C::THREADS) s_dN[i] = T.dNt[(i / NGP) * 32 + (i % NGP)];
    if (tid == 0 && DO_KM) {
        for (int b = 0; b < 2; ++b) { mbar_init(&full[b], 2); mbar_init(&empty[b], 6); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nbatch = (A.nlist + 31) / 32;
    if (warp >= NW - 2) {
        // ================= producers: Gauss points g0 .. g0+3, polarisation pw of the RHS =================
        const int pw = warp - (NW - 2), g0 = 4 * pw;
        const double psig = f32r(A.omega * kEps0);   // pset_pmodel, problem.f90:250
        const double w32 = f32r(A.omega);            // cmplx(0.d0,-omega), problem.f90:112
        // The node fields of a batch arrive by cp.async (16 bytes per lane and chunk, L1 bypassed), issued one batch ahead by
        // the two producer warps (four nodes each): the global-memory latency of the scattered records (27 sectors per
        // request) is hidden behind the closed forms, the RHS and the consumers' contraction of the previous batch.
        // Three-deep software pipeline of the producers' own inputs: the element id of batch k+2 is loaded while the node fields
        // (cp.async) and the half-widths a = dx/2, b = dy/2 of batch k+1 are fetched with the id loaded one batch earlier, so no
        // dependent global-memory round trip (list -> x/y lines and node addresses) sits inside a batch
        int e_nx = 0, e_nx2 = 0;
        double a_nx = 0.0, b_nx = 0.0;
        auto load_id = [&](int batch_) {
            const int pos_ = batch_ * 32 + lane;
            return batch_ < nbatch ? (pos_ < A.nlist ? A.list[pos_] : A.list[A.nlist - 1]) : 0;
        };
        auto prefetch = [&](int batch_) {      // batch_ is the batch whose id is in e_nx2
            e_nx = e_nx2;
            e_nx2 = load_id(batch_ + (int)gridDim.x);
            int ie_, je_, ke_;
            elem_ijk(m, e_nx, ie_, je_, ke_);
            a_nx = 0.5 * (__ldg(A.xp + ie_) - __ldg(A.xp + ie_ - 1)); b_nx = 0.5 * (__ldg(A.yp + je_) - __ldg(A.yp + je_ - 1));
            const NodeRec *base_ = A.nodes + ((int64_t)(ie_ - 1) * m.nyz + (int64_t)(je_ - 1) * m.nnz + (ke_ - 1));
#pragma unroll
            for (int l4 = 0; l4 < 4; ++l4) {
                const int l = 4 * pw + l4;
                const double2 *r2 = reinterpret_cast<const double2 *>(base_ + s_noff[l]);
                constexpr int foff[6] = {0, 4, 5, 6, 7, 8};
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(s_stage + (l * 6 + f) * 32 + lane);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(r2 + foff[f]) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        e_nx2 = load_id(blockIdx.x);
        if ((int)blockIdx.x < nbatch) prefetch(blockIdx.x);
        // mu = mu0 I at every node (flags[0] == 0): one scalar for the whole mesh
        const double mu0inv = DO_KM ? __ldg(reinterpret_cast<const double *>(A.nodes + ((int64_t)(A.e_base / (m.ny * m.nz)) * m.nyz)) + 2) : 0.0;
        double wt[4];
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) wt[gi] = T.rw[g0 + gi][3];
        int k = 0;
        for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x, ++k) {
            const int buf = k & 1;
            const int pos = batch * 32 + lane;
            const bool live = pos < A.nlist;
            const int e = e_nx;
            const double a = a_nx, b = b_nx;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            bar_sync_producers();                          // both warps' copies have landed
            // interpolation to this warp's four Gauss points (p_intmodels problem.f90:139-142; N_l-weighted part of p_source
            // :424-457) and the xi-gradient of z (nf_jacobian, n_fem.f90:359-366)
            double s0[4], s3[4], s5[4], e12[4], e15[4], e18[4], e21[4], zp[4], zq[4], zr[4];
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) { s0[gi] = s3[gi] = s5[gi] = e12[gi] = e15[gi] = e18[gi] = e21[gi] = zp[gi] = zq[gi] = zr[gi] = 0.0; }
#pragma unroll 2
            for (int l = 0; l < MN; ++l) {   // (not fully unrolled: the producers' code competes with the consumers' for the instruction cache)
                const double2 *st = s_stage + (l * 6) * 32 + lane;
                const double2 ze = st[0], s01 = st[32], s23 = st[64], i01 = st[128], i23 = st[160];
                const double s33 = DO_KM ? st[96].y : 0.0;
                const double di1 = i01.x - psig, di2 = i23.y - psig;   // Im(dsigma) on the diagonal (pdelta_model, problem.f90:329-331)
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    const int g = g0 + gi;
                    const double ln = s_dN[(l * 4 + 3) * NGP + g];
                    const double le = ln * ze.y;                       // N_l e_l, e_l = f32(omega b0 z_l)
                    if (DO_KM) { s0[gi] = dfma(ln, s01.x, s0[gi]); s3[gi] = dfma(ln, s23.y, s3[gi]); s5[gi] = dfma(ln, s33, s5[gi]); }
                    e12[gi] = dfma(le, di1, e12[gi]); e15[gi] = dfma(le, s01.x, e15[gi]);
                    e18[gi] = dfma(le, di2, e18[gi]); e21[gi] = dfma(le, s23.y, e21[gi]);
                    zp[gi] = dfma(s_dN[(l * 4 + 0) * NGP + g], ze.x, zp[gi]);
                    zq[gi] = dfma(s_dN[(l * 4 + 1) * NGP + g], ze.x, zq[gi]);
                    zr[gi] = dfma(s_dN[(l * 4 + 2) * NGP + g], ze.x, zr[gi]);
                }
            }
            bar_sync_producers();                          // the staged fields are consumed: refill for the next batch
            if (batch + (int)gridDim.x < nbatch) prefetch(batch + gridDim.x);
            if (DO_KM && k >= 2) mbar_wait(&empty[buf], (unsigned)(((k >> 1) - 1) & 1));   // the consumers are done with this buffer
            if (DO_KM && pw == 0) { s_scale[(buf * 32 + lane) * 2] = 0ull; s_scale[(buf * 32 + lane) * 2 + 1] = 0ull; }
            if (DO_KM) bar_sync_producers();                                               // scales zeroed before either warp's max
            double trq = 0.0, trt = 0.0;
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
                const int g = g0 + gi;
                const double p = zp[gi], q = zq[gi], r = zr[gi];
                const double det = (a * b) * r;
                if (det == 0.0 && live) atomicCAS(A.status, 0, -3);
                const double w = det * wt[gi];
                const double rad = 1.0 / fabs(det);                    // Q6: nf_ji = adj(J)/abs(det)
                const double G00 = (b * r) * rad, G11 = (a * r) * rad, G22 = (a * b) * rad, G02 = -(p * b) * rad, G12 = -(a * q) * rad;
                if (DO_KM) {
                    double *qo = s_qt + (size_t)buf * C::BLK_D + g * 32 + lane;
                    const double fm = (det < 0.0 ? -wt[gi] : wt[gi]) * rad * mu0inv;   // (w/det^2) mu^-1 = (wt/det) mu^-1: Q = fm J J^T
                    const double q00 = fm * dfma(a, a, p * p), q11 = fm * dfma(b, b, q * q), q22 = fm * (r * r);
                    qo[0 * QS] = q00; qo[1 * QS] = fm * (p * q); qo[2 * QS] = fm * (p * r);
                    qo[3 * QS] = q11; qo[4 * QS] = fm * (q * r); qo[5 * QS] = q22;
                    // T = G^T S G, S = w Re sigma_g (integration.f90:234-236, Q3), sigma diagonal
                    const double S0 = w * s0[gi], S1 = w * s3[gi], S2 = w * s5[gi];
                    const double t00 = S0 * (G00 * G00), t11 = S1 * (G11 * G11);
                    const double t22 = dfma(S0 * G02, G02, dfma(S1 * G12, G12, (S2 * G22) * G22));
                    qo[6 * QS] = t00; qo[7 * QS] = 0.0; qo[8 * QS] = (S0 * G00) * G02;
                    qo[9 * QS] = t11; qo[10 * QS] = (S1 * G11) * G12; qo[11 * QS] = t22;
                    trq = fmax(trq, fabs(q00) + fabs(q11) + fabs(q22));
                    trt = fmax(trt, fabs(t00) + fabs(t11) + fabs(t22));
                }
                // R[d][pol] = G[:,d] . (w src_pol);  src = dmpf * cmplx32(0,-omega) (problem.f90:112); for a diagonal sigma
                // src_1 = (A1, 0, 0), src_2 = (0, A2, 0):  pol 1 dmpf = (+Im ds*e, -Re ds*e), pol 2 = (-Im ds*e, +Re ds*e)
                const double a1x = w * (-e15[gi] * w32), a1y = w * (-(e12[gi] * w32)), a2x = w * (e21[gi] * w32), a2y = w * (e18[gi] * w32);
                double *Ro = s_R + g * 32 + lane;
                Ro[0 * QS] = G00 * a1x; Ro[1 * QS] = G00 * a1y; Ro[2 * QS] = 0.0;       Ro[3 * QS] = 0.0;
                Ro[4 * QS] = 0.0;       Ro[5 * QS] = 0.0;       Ro[6 * QS] = G11 * a2x; Ro[7 * QS] = G11 * a2y;
                Ro[8 * QS] = G02 * a1x; Ro[9 * QS] = G02 * a1y; Ro[10 * QS] = G12 * a2x; Ro[11 * QS] = G12 * a2y;
            }
            if (DO_KM) {
                // the element's K / M scale for the tiny-pair test of the epilogue: max over the Gauss points, order independent
                atomicMax(&s_scale[(buf * 32 + lane) * 2], (unsigned long long)__double_as_longlong(trq));
                atomicMax(&s_scale[(buf * 32 + lane) * 2 + 1], (unsigned long long)__double_as_longlong(trt));
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[buf]);    // release: this warp's Q|T and scales are visible to the consumers
            }
            bar_sync_producers();                          // R of all eight Gauss points is in place
            // ---- RHS, polarisation pw: b_e(j) = sum_g phi_j(g) R[d_j][pol](g)  (blocal / f3, integration.f90:96-104,258-263) ----
            if (live) {
                double2 *bo = reinterpret_cast<double2 *>(A.be) + ((size_t)(e - A.e_base) * ME) * 2 + pw;
#pragma unroll
                for (int q4 = 0; q4 < ME; q4 += 4) {
                    const int cd = s_sdir[q4];
                    double br[4] = {0, 0, 0, 0}, bi[4] = {0, 0, 0, 0};
                    // diagonal sigma: the source of polarisation 1 is along x, of polarisation 2 along y, so R[d][pol] is
                    // identically zero for (d, pol) = (1, 1) and (0, 2) -- written as zeros above, not summed here
                    if (cd == 2 || cd == pw) {
                        const double *Rr = s_R + (size_t)(cd * 4 + 2 * pw) * QS + lane, *Ri = Rr + QS;
#pragma unroll
                        for (int g = 0; g < NGP; ++g) {
                            const double rr = Rr[g * 32], ri = Ri[g * 32];
                            const double2 p01 = *reinterpret_cast<const double2 *>(s_phi + g * ME + q4), p23 = *reinterpret_cast<const double2 *>(s_phi + g * ME + q4 + 2);
                            br[0] = dfma(p01.x, rr, br[0]); bi[0] = dfma(p01.x, ri, bi[0]);
                            br[1] = dfma(p01.y, rr, br[1]); bi[1] = dfma(p01.y, ri, bi[1]);
                            br[2] = dfma(p23.x, rr, br[2]); bi[2] = dfma(p23.x, ri, bi[2]);
                            br[3] = dfma(p23.y, rr, br[3]); bi[3] = dfma(p23.y, ri, bi[3]);
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) bo[(size_t)s_slot[q4 + kk] * 2] = make_double2(br[kk], bi[kk]);
                }
            }
            bar_sync_producers();                          // R may be overwritten
        }
        return;
    }

    // ================= consumers: tile `warp` = direction class `warp` (inner loop of contract.cuh) =================
    if (DO_KM) {
        const int c = warp;
        const int ti = c_ct.tile_ti[c], tj = c_ct.tile_tj[c];
        const int dI = cls_dI(c), dJ = cls_dJ(c);
        const int k1I = dI == 2 ? 1 : 2, k2I = dI == 0 ? 1 : 0, k1J = dJ == 2 ? 1 : 2, k2J = dJ == 0 ? 1 : 0;
        const double tau = ((dI == 1) != (dJ == 1)) ? -1.0 : 1.0;
        const int cq0 = c_ct.comp[0][c][0], cq1 = c_ct.comp[0][c][1], cq2 = c_ct.comp[0][c][2], cq3 = c_ct.comp[0][c][3], cq4 = c_ct.comp[0][c][4];
        const double *Y1 = s_tab + k1I * ME + 4 * ti, *Y2 = s_tab + k2I * ME + 4 * ti, *Y3 = s_tab + 3 * ME + 4 * ti;
        const double *X1 = s_tab + k1J * ME + 4 * tj, *X2 = s_tab + k2J * ME + 4 * tj, *X3 = s_tab + 3 * ME + 4 * tj;
        int k = 0;
        for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x, ++k) {
            const int buf = k & 1;
            mbar_wait(&full[buf], (unsigned)((k >> 1) & 1));
            const double *S = s_qt + (size_t)buf * C::BLK_D + lane;
            double accK[16], accM[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { accK[i] = 0.0; accM[i] = 0.0; }
            // a diagonal tile (classes xx, yy, zz) stores only its lower half: the six pairs above the diagonal are not formed
            auto contract = [&](auto diag_tag) {
                constexpr bool DIAG = decltype(diag_tag)::value;
#pragma unroll 4
                for (int g = 0; g < NGP; ++g) {
                    const int o = g * 4 * ME;
                    const double q00 = S[(cq0 * NGP + g) * 32], q01 = S[(cq1 * NGP + g) * 32], q10 = S[(cq2 * NGP + g) * 32],
                                 q11 = S[(cq3 * NGP + g) * 32], tt = S[(cq4 * NGP + g) * 32];
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
                            if (DIAG && j > i) continue;
                            accK[i * 4 + j] = dfma(y1, b1[j], dfma(-y2, b2[j], accK[i * 4 + j]));
                            accM[i * 4 + j] = dfma(y3, bw[j], accM[i * 4 + j]);
                        }
                    }
                }
            };
            if (ti == tj) contract(std::true_type{}); else contract(std::false_type{});
            const double thrK = A.no_l1 ? -1.0 : kTinyRelC * NGP * __longlong_as_double((long long)s_scale[(buf * 32 + lane) * 2]);
            const double thrM = A.no_l1 ? -1.0 : kTinyRelC * NGP * __longlong_as_double((long long)s_scale[(buf * 32 + lane) * 2 + 1]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[buf]);       // the buffer (and its scales) may be refilled
            // write-out: packed lower triangle by LOCAL DOF index, 32 elements interleaved -> 512-byte coalesced stores
            if (batch * 32 + lane < A.nlist) {
                double2 *KMo = A.KM + (size_t)batch * NP * 32 + lane;
                uint32_t *pfl = A.pairflags + (size_t)(batch * 32 + lane) * A.W;
                int nfl = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int si = 4 * ti + i, im = c_ct.slot_dof[si];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int sj = 4 * tj + j, jm = c_ct.slot_dof[sj];
                        if (sj <= si) {
                            const int hi = im > jm ? im : jm, lo = im > jm ? jm : im;
                            const int p = hi * (hi + 1) / 2 + lo;
                            const double kv = accK[i * 4 + j] * tau, mv = accM[i * 4 + j];
                            KMo[p * 32] = make_double2(kv, mv);
                            const double ak = fabs(kv), am = fabs(mv);
                            if ((ak <= thrK && am <= thrM) || (ak < kFlagAbsC && am < kFlagAbsC)) {
                                atomicOr(pfl + (p >> 5), 1u << (p & 31));
                                ++nfl;
                            }
                        }
                    }
                }
                if (nfl) { atomicOr(A.batchany + batch, 1u << lane); atomicAdd(A.nflag, (unsigned long long)nfl); }
            }
        }
    }
}

}  // namespace movfem

// movfem_b200/csrc/fused12.cuh -- the 8-node / 12-DOF (linear) element path in ONE kernel.
//
// BASELINE configs 1, 4 and 5 are linear-element meshes (5: 32 M elements).  There the two-kernel element path of
// element.cuh + contract.cuh is bound by its own HBM round trips, not by arithmetic: per element 768 B of Q|T scratch
// are written by geometry_kernel and re-read by contract_kernel next to 1 248 B of K_e/M_e -- against 10 944
// algorithmic flops.  This kernel keeps the per-Gauss-point tensors on chip:
//
//   batch = 32 consecutive elements of the unstretched list, one CTA of 8 warps
//   G   thread (Gauss point g = warp, element s = lane): material / source interpolation, J, det, G = J^-1,
//       Q = (w/det^2) J mu^-1 J^T, T = G^T S G, R (source) -- the formulation of element.cuh -- into a shared-memory
//       block [component][g][lane]; the element's K / M scale for the tiny-pair test (exact.cuh) by an atomic max
//   C   warps 0-5: one 4x4 tile of the lower triangle each (the 12 DOFs are three direction-uniform groups of four, so
//       the six tiles are the six direction classes), LANES ARE ELEMENTS, operands warp-uniform broadcast loads -- the
//       inner loop of contract.cuh -- reading Q|T from the block; K_e, M_e leave as 512-byte coalesced stores
//   R   warps 6-7 at the same time: b_e = sum_g phi_j R[d_j]  (integration.f90:96-104)
//
// Two CTAs per SM, so the FP64-bound phase C of one overlaps the latency-bound phase G of the other.  Node records
// arrive by TMA bulk copies (cp.async.bulk + mbarrier), requested one phase ahead.  Stretched (GPML scheme 0) elements
// and the RHS-only pass of a cached frequency stay on the generic kernels.
//
// Replaces for linear elements: MoVFEM_3DMT.f90:193-211 (element loop body), integration.f90:60-86 (int_elem_params,
// alocal), n_fem.f90:66-102,355-395, v_fem.f90:38-60, problem.f90:70-149.
#pragma once
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
    const int *flags;              // flags[0]: any dmu != 0, [2]: any off-diagonal sigma component (node_kernel)
    uint32_t *pairflags;           // [row][W] at this launch's first row
    uint32_t *batchany;            // at this launch's first batch
    unsigned long long *nflag;
    int W;
    int no_l1;                     // test hook: no element-level tiny-pair flags
    int skip_unless_changed;       // launch is a cache refresh: exit unless flags[1] (Re sigma changed)
};

struct Fused12Cfg {
    static constexpr int MN = 8, ME = 12, NGP = 8, EB = 32, THREADS = 256, NP = 78;
    static constexpr int NREC = 20, NDW = NREC + 2, NSTR = MN * NDW + 2;   // staged node record (+ x, y), per-element stride
    static constexpr int RST = 14;                                         // R record: 12 doubles padded to an odd number of 16-byte chunks
    static constexpr size_t NODES_D = (size_t)EB * NSTR, R_D = (size_t)NGP * EB * RST, QT_D = (size_t)12 * NGP * 32;
    static constexpr size_t TAB_D = (size_t)NGP * 4 * ME, DN_D = (size_t)MN * 4 * NGP, PHI_D = (size_t)NGP * ME;
    static constexpr size_t SMEM = sizeof(double) * (NODES_D + R_D + QT_D + TAB_D + DN_D + PHI_D) + sizeof(unsigned long long) * (2 * EB + 1) +
                                   sizeof(int64_t) * EB + sizeof(int) * (EB * 3 + 2 * ME + 3 * MN);
};

template <bool DO_KM>
__global__ void __launch_bounds__(256, 2) fused12_kernel(Fused12Args A) {
    using C = Fused12Cfg;
    constexpr int MN = C::MN, ME = C::ME, NGP = C::NGP, EB = C::EB, NDW = C::NDW, NREC = C::NREC, RST = C::RST, NP = C::NP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_qt = reinterpret_cast<double *>(smem_raw);              // [12][NGP][32]: Q (0-5, sym3 order), T (6-11)
    double *s_R = s_qt + C::QT_D;                                     // [NGP][EB][RST]
    double *s_nodes = s_R + C::R_D;                                   // [EB][NSTR]
    double *s_tab = s_nodes + C::NODES_D;                             // [NGP][4][ME]: dphi (0-2), phi (3), slot order
    double *s_dN = s_tab + C::TAB_D;                                  // [MN][4][NGP]: dN/dxi (0-2), N (3)
    double *s_phi = s_dN + C::DN_D;                                   // [NGP][ME] phi in slot order
    unsigned long long *s_scale = reinterpret_cast<unsigned long long *>(s_phi + C::PHI_D);   // [EB][2]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_scale + 2 * EB);
    int64_t *s_rbase = reinterpret_cast<int64_t *>(s_bar + 1);        // [EB]
    int *s_el = reinterpret_cast<int *>(s_rbase + EB);                // [EB]
    int *s_rxy = s_el + EB;                                           // [EB][2]
    int *s_slot = s_rxy + 2 * EB;                                     // [ME] slot -> local DOF
    int *s_sdir = s_slot + ME;                                        // [ME] slot -> direction
    int *s_noff = s_sdir + ME;                                        // [MN][3]

    const ElemTables &T = *A.T;
    const MeshDims &m = A.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (A.skip_unless_changed && A.flags[1] == 0) return;

    for (int i = tid; i < NGP * ME; i += C::THREADS) {
        const int g = i / ME, sl = i % ME;
        s_phi[i] = T.phi[g][T.slot_dof[sl]];
    }
    for (int i = tid; i < (int)C::TAB_D; i += C::THREADS) s_tab[i] = g_ct_at[i];
    for (int i = tid; i < ME; i += C::THREADS) { s_slot[i] = T.slot_dof[i]; s_sdir[i] = T.slot_dir[i]; }
    for (int i = tid; i < MN; i += C::THREADS) { s_noff[i * 3] = T.node_off[i]; s_noff[i * 3 + 1] = T.node_i[i]; s_noff[i * 3 + 2] = T.node_j[i]; }
    for (int i = tid; i < MN * 4 * NGP; i += C::THREADS) {
        const int g = i % NGP, lm = i / NGP;
        s_dN[i] = T.dNt[lm * 32 + g];
    }

    const int nbatch = (A.nlist + EB - 1) / EB;
    auto prepare_request = [&](int b) {
        const int first = b * EB, nb = min(EB, A.nlist - first);
        if (tid < nb) {
            int ie, je, ke;
            elem_ijk(m, A.list[first + tid], ie, je, ke);
            s_rbase[tid] = (int64_t)(ie - 1) * m.nyz + (int64_t)(je - 1) * m.nnz + (ke - 1);
            s_rxy[tid * 2] = ie - 1; s_rxy[tid * 2 + 1] = je - 1;
        }
    };
    auto request_nodes = [&](int b) {
        const int nb = min(EB, A.nlist - b * EB);
        if (tid == 0) mbar_expect_tx(s_bar, (unsigned)(nb * MN * NREC * sizeof(double)));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int i = tid; i < nb * MN; i += C::THREADS) {
            const int l = i % MN, s = i / MN;
            double *dst = s_nodes + s * C::NSTR + l * NDW;
            bulk_g2s(dst, A.nodes + (s_rbase[s] + s_noff[l * 3]), (unsigned)(NREC * sizeof(double)), s_bar);
            *reinterpret_cast<double2 *>(dst + NREC) = make_double2(A.xp[s_rxy[s * 2] + s_noff[l * 3 + 1]], A.yp[s_rxy[s * 2 + 1] + s_noff[l * 3 + 2]]);
        }
    };
    if (tid == 0) { mbar_init(s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if ((int)blockIdx.x < nbatch) prepare_request(blockIdx.x);
    __syncthreads();
    if ((int)blockIdx.x < nbatch) request_nodes(blockIdx.x);
    unsigned node_phase = 0;
    const double psig = f32r(A.omega * kEps0);   // pset_pmodel, problem.f90:250
    const double w32 = f32r(A.omega);            // cmplx(0.d0,-omega), problem.f90:112
    const int has_dmu = A.flags[0];
    // mu = mu0 I at every node and sigma diagonal at every node (every linear-element BASELINE mesh): the interpolation needs 7
    // of its 24 columns and J, G, Q, T, R have closed forms in the five non-zero entries of J (x and y are tensor-product
    // lines, so J = [[a,0,p],[0,b,q],[0,0,r]] with a = dx/2, b = dy/2 and (p,q,r) the xi-gradient of z, n_fem.f90:193 included)
    const bool simple = !has_dmu && A.flags[2] == 0;

    for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int first = batch * EB;
        const int nb = min(EB, A.nlist - first);
        __syncthreads();   // previous batch fully consumed (s_qt, s_R, s_el, s_scale)
        if (tid < EB) {
            s_el[tid] = tid < nb ? A.list[first + tid] : -1;
            s_scale[tid * 2] = 0ull; s_scale[tid * 2 + 1] = 0ull;
        }
        mbar_wait(s_bar, node_phase);
        node_phase ^= 1;
        __syncthreads();
        if (batch + (int)gridDim.x < nbatch) prepare_request(batch + gridDim.x);   // s_rbase / s_rxy: no reader until the request below

        // ---- phase G: thread = (Gauss point g = warp, element s = lane) ----
        if (simple && lane < nb) {
            const int g = warp, s = lane;
            const double *nd = s_nodes + s * C::NSTR;
            double s0 = 0.0, s3 = 0.0, s5 = 0.0, e12 = 0.0, e15 = 0.0, e18 = 0.0, e21 = 0.0, p = 0.0, q = 0.0, r = 0.0;
#pragma unroll
            for (int l = 0; l < MN; ++l) {
                const double2 *r2 = reinterpret_cast<const double2 *>(nd + l * NDW);
                const double2 ze = r2[0], s01 = r2[4], s23 = r2[5], s45 = r2[6], i01 = r2[7], i23 = r2[8];
                const double ln = s_dN[(l * 4 + 3) * NGP + g];
                const double le = ln * ze.y;
                if (DO_KM) { s0 = dfma(ln, s01.x, s0); s3 = dfma(ln, s23.y, s3); s5 = dfma(ln, s45.y, s5); }
                e12 = dfma(le, i01.x - psig, e12); e15 = dfma(le, s01.x, e15);
                e18 = dfma(le, i23.y - psig, e18); e21 = dfma(le, s23.y, e21);
                p = dfma(s_dN[(l * 4 + 0) * NGP + g], ze.x, p);
                q = dfma(s_dN[(l * 4 + 1) * NGP + g], ze.x, q);
                r = dfma(s_dN[(l * 4 + 2) * NGP + g], ze.x, r);
            }
            const double a = 0.5 * (nd[NREC] - nd[2 * NDW + NREC]), b = 0.5 * (nd[NDW + NREC + 1] - nd[NREC + 1]);
            const double det = (a * b) * r;
            if (det == 0.0) atomicCAS(A.status, 0, -3);
            const double w = det * T.rw[g][3];
            const double rad = 1.0 / fabs(det);
            const double G00 = (b * r) * rad, G11 = (a * r) * rad, G22 = (a * b) * rad, G02 = -(p * b) * rad, G12 = -(a * q) * rad;
            if (DO_KM) {
                double *qo = s_qt + g * 32 + s;
                constexpr int QS = NGP * 32;
                const double fm = (w / (det * det)) * nd[2];      // (w/det^2) * mu^-1 (the same scalar at every node)
                const double q00 = fm * dfma(a, a, p * p), q11 = fm * dfma(b, b, q * q), q22 = fm * (r * r);
                qo[0 * QS] = q00; qo[1 * QS] = fm * (p * q); qo[2 * QS] = fm * (p * r);
                qo[3 * QS] = q11; qo[4 * QS] = fm * (q * r); qo[5 * QS] = q22;
                const double S0 = w * s0, S1 = w * s3, S2 = w * s5;
                const double t00 = S0 * (G00 * G00), t11 = S1 * (G11 * G11);
                const double t22 = dfma(S0 * G02, G02, dfma(S1 * G12, G12, (S2 * G22) * G22));
                qo[6 * QS] = t00; qo[7 * QS] = 0.0; qo[8 * QS] = (S0 * G00) * G02;
                qo[9 * QS] = t11; qo[10 * QS] = (S1 * G11) * G12; qo[11 * QS] = t22;
                atomicMax(&s_scale[s * 2], (unsigned long long)__double_as_longlong(fabs(q00) + fabs(q11) + fabs(q22)));
                atomicMax(&s_scale[s * 2 + 1], (unsigned long long)__double_as_longlong(fabs(t00) + fabs(t11) + fabs(t22)));
            }
            // R[d][pol] = G[:,d] . (w src_pol): src_1 = (A1, 0, 0), src_2 = (0, A2, 0) for a diagonal sigma
            const double a1x = w * (-e15 * w32), a1y = w * (-(e12 * w32)), a2x = w * (e21 * w32), a2y = w * (e18 * w32);
            double *Ro = s_R + (size_t)(g * EB + s) * RST;
            *reinterpret_cast<double2 *>(Ro + 0) = make_double2(G00 * a1x, G00 * a1y);
            *reinterpret_cast<double2 *>(Ro + 2) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2 *>(Ro + 4) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2 *>(Ro + 6) = make_double2(G11 * a2x, G11 * a2y);
            *reinterpret_cast<double2 *>(Ro + 8) = make_double2(G02 * a1x, G02 * a1y);
            *reinterpret_cast<double2 *>(Ro + 10) = make_double2(G12 * a2x, G12 * a2y);
        } else if (lane < nb) {
            const int g = warp, s = lane;
            const double *nd = s_nodes + s * C::NSTR;
            // interpolation to the Gauss point (p_intmodels problem.f90:139-142; N_l-weighted part of p_source :424-457).
            // Measured against a tensor-core (mma.sync.m8n8k4.f64) interpolation phase with a reused buffer: the extra two
            // block barriers cost more than the eightfold re-read of the node records (4.3 vs 5.3 ms on 4 M elements).
            double mu[6] = {0, 0, 0, 0, 0, 0}, sr[6] = {0, 0, 0, 0, 0, 0};
            double c12[3] = {0, 0, 0}, c15[3] = {0, 0, 0}, c18[3] = {0, 0, 0}, c21[3] = {0, 0, 0};
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            const double xs0 = nd[2 * NDW + NREC], xs1 = nd[NREC], ys0 = nd[NREC + 1], ys1 = nd[NDW + NREC + 1];
#pragma unroll
            for (int l = 0; l < MN; ++l) {
                // record: z, e | mu^-1 (6) | Re sigma (6) | Im sigma (6), 16-byte aligned: 128-bit loads (conflict-free over the
                // 32 lanes, whose records are 89 16-byte chunks apart)
                const double2 *r2 = reinterpret_cast<const double2 *>(nd + l * NDW);
                const double2 ze = r2[0], s01 = r2[4], s23 = r2[5], s45 = r2[6], i01 = r2[7], i23 = r2[8], i45 = r2[9];
                const double ln = s_dN[(l * 4 + 3) * NGP + g];
                const double le = ln * ze.y;             // N_l * e_l,  e_l = f32(omega b0 z_l)
                if (DO_KM) {
                    const double2 m01 = r2[1], m23 = r2[2], m45 = r2[3];
                    mu[0] = dfma(ln, m01.x, mu[0]); mu[1] = dfma(ln, m01.y, mu[1]); mu[2] = dfma(ln, m23.x, mu[2]);
                    mu[3] = dfma(ln, m23.y, mu[3]); mu[4] = dfma(ln, m45.x, mu[4]); mu[5] = dfma(ln, m45.y, mu[5]);
                    sr[0] = dfma(ln, s01.x, sr[0]); sr[1] = dfma(ln, s01.y, sr[1]); sr[2] = dfma(ln, s23.x, sr[2]);
                    sr[3] = dfma(ln, s23.y, sr[3]); sr[4] = dfma(ln, s45.x, sr[4]); sr[5] = dfma(ln, s45.y, sr[5]);
                }
                // Im(dsigma) = Im(sigma) - psig on the diagonal (pdelta_model, problem.f90:329-331)
                c12[0] = dfma(le, i01.x - psig, c12[0]); c12[1] = dfma(le, i01.y, c12[1]); c12[2] = dfma(le, i23.x, c12[2]);
                c15[0] = dfma(le, s01.x, c15[0]);        c15[1] = dfma(le, s01.y, c15[1]); c15[2] = dfma(le, s23.x, c15[2]);
                c18[0] = dfma(le, i01.y, c18[0]);        c18[1] = dfma(le, i23.y - psig, c18[1]); c18[2] = dfma(le, i45.x, c18[2]);
                c21[0] = dfma(le, s01.y, c21[0]);        c21[1] = dfma(le, s23.y, c21[1]); c21[2] = dfma(le, s45.x, c21[2]);
                // nf_jacobian, n_fem.f90:359-366: l ascending, no FMA (reference bits for J, det, w)
                const double x = (kNodeI27[l] ? xs1 : xs0), y = (kNodeJ27[l] ? ys1 : ys0), z = ze.x;
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) {
                    const double dn = s_dN[(l * 4 + mm) * NGP + g];
                    J[mm][0] = J[mm][0] + dn * x; J[mm][1] = J[mm][1] + dn * y; J[mm][2] = J[mm][2] + dn * z;
                }
            }
            const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
            if (det == 0.0) atomicCAS(A.status, 0, -3);
            const double w = det * T.rw[g][3];
            const double rad = 1.0 / fabs(det);
            double G[3][3];
            G[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * rad;
            G[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * rad;
            G[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * rad;
            G[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * rad;
            G[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * rad;
            G[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * rad;
            G[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * rad;
            G[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * rad;
            G[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * rad;
            double pc1[3] = {0, 0, 0}, pc2[3] = {0, 0, 0};
            if (has_dmu) {   // p_pcurl, problem.f90:362-374 (mu != mu0 only)
                int ie, je, ke;
                elem_ijk(m, s_el[s], ie, je, ke);
                const int64_t base = (int64_t)(ie - 1) * m.nyz + (int64_t)(je - 1) * m.nnz + (ke - 1);
                for (int l = 0; l < MN; ++l) {
                    const double *vc = A.nodes[base + s_noff[l * 3]].vc;
                    const double t0 = s_dN[(l * 4 + 0) * NGP + g], t1 = s_dN[(l * 4 + 1) * NGP + g], t2 = s_dN[(l * 4 + 2) * NGP + g];
                    double dn[3];
#pragma unroll
                    for (int mm = 0; mm < 3; ++mm) dn[mm] = G[mm][0] * t0 + G[mm][1] * t1 + G[mm][2] * t2;
                    pc1[0] += vc[2] * dn[1] - vc[1] * dn[2]; pc1[1] += vc[0] * dn[2] - vc[2] * dn[0]; pc1[2] += vc[1] * dn[0] - vc[0] * dn[1];
                    pc2[0] += vc[5] * dn[1] - vc[4] * dn[2]; pc2[1] += vc[3] * dn[2] - vc[5] * dn[0]; pc2[2] += vc[4] * dn[0] - vc[3] * dn[1];
                }
            }
            double *qo = s_qt + g * 32 + s;
            constexpr int QS = NGP * 32;
            double trq = 0.0, trt = 0.0;
            if (DO_KM) {   // Q = (w/det^2) J mu^-1 J^T
                const double f = w / (det * det);
                double Jm[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        Jm[a][q] = dfma(J[a][0], mu[sym3(0, q)], dfma(J[a][1], mu[sym3(1, q)], J[a][2] * mu[sym3(2, q)]));
                int q6 = 0;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = a; b < 3; ++b) {
                        const double qv = f * dfma(Jm[a][0], J[b][0], dfma(Jm[a][1], J[b][1], Jm[a][2] * J[b][2]));
                        qo[(q6++) * QS] = qv;
                        if (a == b) trq += fabs(qv);
                    }
            }
            if (DO_KM) {   // T = G^T S G, S = w Re sigma_g (integration.f90:234-236, Q3)
                double S[6], SG[3][3];
#pragma unroll
                for (int k = 0; k < 6; ++k) S[k] = w * sr[k];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int d2 = 0; d2 < 3; ++d2)
                        SG[a][d2] = dfma(S[sym3(a, 0)], G[0][d2], dfma(S[sym3(a, 1)], G[1][d2], S[sym3(a, 2)] * G[2][d2]));
                int q6 = 0;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = a; b < 3; ++b) {
                        const double tv = dfma(G[0][a], SG[0][b], dfma(G[1][a], SG[1][b], G[2][a] * SG[2][b]));
                        qo[(6 + q6++) * QS] = tv;
                        if (a == b) trt += fabs(tv);
                    }
            }
            if (DO_KM) {
                atomicMax(&s_scale[s * 2], (unsigned long long)__double_as_longlong(trq));
                atomicMax(&s_scale[s * 2 + 1], (unsigned long long)__double_as_longlong(trt));
            }
            {   // R[d][pol] = G[:,d] . (w src_pol);  src = (dmpf + pcrl) * cmplx32(0,-omega)  (problem.f90:112)
                // pol 1 dmpf = (+Im ds*e, -Re ds*e), pol 2 = (-Im ds*e, +Re ds*e)
                double a1x[3], a1y[3], a2x[3], a2y[3];
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) {
                    const double d1r = c12[mm], d1i = -c15[mm], d2r = -c18[mm], d2i = c21[mm];
                    a1x[mm] = w * (d1i * w32); a1y[mm] = w * (-((d1r + pc1[mm]) * w32));
                    a2x[mm] = w * (d2i * w32); a2y[mm] = w * (-((d2r + pc2[mm]) * w32));
                }
                double *Ro = s_R + (size_t)(g * EB + s) * RST;
#pragma unroll
                for (int d2 = 0; d2 < 3; ++d2) {
                    const double r0 = dfma(G[0][d2], a1x[0], dfma(G[1][d2], a1x[1], G[2][d2] * a1x[2]));
                    const double r1 = dfma(G[0][d2], a1y[0], dfma(G[1][d2], a1y[1], G[2][d2] * a1y[2]));
                    const double r2 = dfma(G[0][d2], a2x[0], dfma(G[1][d2], a2x[1], G[2][d2] * a2x[2]));
                    const double r3 = dfma(G[0][d2], a2y[0], dfma(G[1][d2], a2y[1], G[2][d2] * a2y[2]));
                    *reinterpret_cast<double2 *>(Ro + d2 * 4) = make_double2(r0, r1);
                    *reinterpret_cast<double2 *>(Ro + d2 * 4 + 2) = make_double2(r2, r3);
                }
            }
        }
        __syncthreads();
        if (batch + (int)gridDim.x < nbatch) request_nodes(batch + gridDim.x);   // s_nodes is dead: lands during phase C / R

        if (DO_KM && warp < 6) {
            // ---- phase C: tile `warp` = direction class `warp`; lanes are elements (inner loop of contract.cuh) ----
            const int c = warp;
            const int ti = c_ct.tile_ti[c], tj = c_ct.tile_tj[c];
            const int dI = cls_dI(c), dJ = cls_dJ(c);
            const int k1I = dI == 2 ? 1 : 2, k2I = dI == 0 ? 1 : 0, k1J = dJ == 2 ? 1 : 2, k2J = dJ == 0 ? 1 : 0;
            const double tau = ((dI == 1) != (dJ == 1)) ? -1.0 : 1.0;
            const double *S = s_qt + lane;
            const int cq0 = c_ct.comp[0][c][0], cq1 = c_ct.comp[0][c][1], cq2 = c_ct.comp[0][c][2], cq3 = c_ct.comp[0][c][3], cq4 = c_ct.comp[0][c][4];
            double accK[16], accM[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { accK[i] = 0.0; accM[i] = 0.0; }
            const double *Y1 = s_tab + k1I * ME + 4 * ti, *Y2 = s_tab + k2I * ME + 4 * ti, *Y3 = s_tab + 3 * ME + 4 * ti;
            const double *X1 = s_tab + k1J * ME + 4 * tj, *X2 = s_tab + k2J * ME + 4 * tj, *X3 = s_tab + 3 * ME + 4 * tj;
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
                        accK[i * 4 + j] = dfma(y1, b1[j], dfma(-y2, b2[j], accK[i * 4 + j]));
                        accM[i * 4 + j] = dfma(y3, bw[j], accM[i * 4 + j]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) accK[i] *= tau;
            if (lane < nb) {
                double2 *KMo = A.KM + (size_t)batch * NP * 32 + lane;
                const double thrK = A.no_l1 ? -1.0 : kTinyRelC * NGP * __longlong_as_double((long long)s_scale[lane * 2]);
                const double thrM = A.no_l1 ? -1.0 : kTinyRelC * NGP * __longlong_as_double((long long)s_scale[lane * 2 + 1]);
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
                if (nfl) { atomicOr(A.batchany + batch, 1u << lane); atomicAdd(A.nflag, (unsigned long long)nfl); }
            }
        } else {
            // ---- phase R: one thread per (element, group of four slots of one direction): blocal / f3,
            //      integration.f90:96-104,258-263 (warps 6-7 beside phase C; every warp in the RHS-only pass) ----
            for (int i = DO_KM ? tid - 192 : tid; i < nb * 3; i += DO_KM ? 64 : 256) {
                const int cs = i / 3, q4 = (i % 3) * 4;
                const int cd = s_sdir[q4];
                double bacc[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) bacc[k][cc] = 0.0;
                const double *R0 = s_R + (size_t)cs * RST + cd * 4, *ph = s_phi + q4;
#pragma unroll
                for (int g = 0; g < NGP; ++g) {
                    const double2 p01 = *reinterpret_cast<const double2 *>(ph + g * ME), p23 = *reinterpret_cast<const double2 *>(ph + g * ME + 2);
                    const double2 r01 = *reinterpret_cast<const double2 *>(R0 + (size_t)g * EB * RST), r23 = *reinterpret_cast<const double2 *>(R0 + (size_t)g * EB * RST + 2);
                    const double phi[4] = {p01.x, p01.y, p23.x, p23.y}, R[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) bacc[k][cc] = dfma(phi[k], R[cc], bacc[k][cc]);
                }
                const int64_t e = s_el[cs];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    reinterpret_cast<double4 *>(A.be)[(e - A.e_base) * ME + s_slot[q4 + k]] = make_double4(bacc[k][0], bacc[k][1], bacc[k][2], bacc[k][3]);
            }
        }
    }
}

}  // namespace movfem

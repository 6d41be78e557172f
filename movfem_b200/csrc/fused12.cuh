// movfem_b200/csrc/fused12.cuh -- the 8-node / 12-DOF (linear) element path in ONE kernel.
//
// BASELINE configs 1, 4 and 5 are linear-element meshes (5: 32 M elements).  There the two-kernel element path of
// element.cuh + contract.cuh is bound by its own HBM round trips and block barriers, not by arithmetic: per element
// 768 B of Q|T scratch are written by geometry_kernel and re-read by contract_kernel next to 1 248 B of K_e/M_e --
// against 10 944 algorithmic flops.  This kernel keeps the per-Gauss-point tensors on chip and has no
// block-wide barrier in its loop:
//
//   batch = 32 consecutive elements of the unstretched list (LANES ARE ELEMENTS everywhere), CTA = 8 warps, 2 CTAs per SM.
//   Per batch every warp does two things, the second one batch behind the first:
//   (1) geometry of ONE Gauss point (warp w = Gauss point w) of batch k+1: interpolation of the staged node fields, then J,
//       det, G = J^-1, Q = (w/det^2) J mu^-1 J^T, T = G^T S G and the source R in CLOSED FORM -- x and y are tensor-product
//       lines, so J = [[a,0,p],[0,b,q],[0,0,r]] with a = dx/2, b = dy/2 and (p,q,r) the xi-gradient of z (the n_fem.f90:193
//       typo is in the table).  Q|T and R go to one of two shared-memory buffers [component][g][lane]; every thread
//       arrives on the buffer's `full` mbarrier
//   (2) of batch k -- warps 0-5: one 4x4 tile of the lower triangle each (the 12 DOFs are three direction-uniform groups of
//       four, so the six tiles are the six direction classes; operands are warp-uniform broadcast loads: the inner loop of
//       contract.cuh), K_e, M_e leave as 512-byte coalesced stores with the tiny-pair test of exact.cuh in the epilogue,
//       then the threads arrive on `empty`; warps 6, 7: b_e = sum_g phi_j R[d_j] (integration.f90:96-104), the x and y slot
//       groups on one warp, the z group on the other, stored by K/M row (be_index: 1 kB runs of 32-byte stores), then the
//       asynchronous copies (cp.async, completion on the `staged` mbarrier) of the node fields of batch k+2 from the
//       FIELD-MAJOR node arrays node_kernel writes next to its records: the 32 lanes of a copy are 32 consecutive nodes of a
//       k-column, 256 contiguous bytes
//
// The closed forms hold when mu = mu0 I and sigma is diagonal at every node (node_kernel reports both; every
// linear-element BASELINE mesh): otherwise the kernel returns at once and the generic two-kernel path does the work
// (api.cu launches both; exactly one of them runs).  They agree with the reference's arithmetic to rounding (<= 1e-15);
// the entries whose last bits matter are re-evaluated in the reference's own operation order by exact_kernel.
// The RHS-only pass of a cached frequency (DO_KM = false) is the same geometry with the six RHS units on warps 0-5.
//
// How it got here (B200, 4.2 M elements of config 5 at half scale, ms of this kernel; profiles/r02_summary.md).  Warp
// specialised, two producer warps (four Gauss points each) feeding six tile warps: 3.37 -- ncu: instruction-cache hit rate
// 82 %, the GPC-level instruction cache at 95 % of its request rate (62 kB of straight-line SASS executed once per batch),
// tile warps waiting for the producers half of the time; producer alone 2.44, tile warps alone 1.75.  One Gauss point per
// warp, one code path for the tiles, compact epilogue (this structure): 3.18 -- instruction cache 99.7 %, now the
// shared-memory / LSU data pipe at 81-86 % of its wavefront rate is the limit, so what followed removes wavefronts:
// field-major node arrays instead of 208-byte records (32 sectors -> 2-3 per copy) and b_e by K/M row (32 -> 8 wavefronts
// per store): 2.38; diagonal tiles reuse their row operands, Gauss-point weights as two 16-byte loads, float scales: 2.25;
// isotropic-sigma variant (four staged fields instead of seven): 2.05; b_e of both polarisations in one 32-byte store per slot
// (st.global.v4.f64: 8 wavefronts per kB instead of 16): 2.03 = 22.6 TFLOP/s algorithmic, 0.65 of the measured FP64 peak.
// geometry_kernel + contract_kernel on the same elements: 5.59.  Tried and dropped: DMMA interpolation phase in a block-phased
// version (block barriers: 5.30), four producer warps (register spills: 4.05), asynchronous id / line prefetch in the producers
// (3.47), producers on one SM sub-partition (3.50), rolled Gauss-point / node loops (2 x unrolled: 3.53), two Gauss points x 16
// elements per warp in the geometry (half the distinct addresses per field load, same wavefronts: 2.07 against 2.05), the
// interpolation as mma.m8n8k4.f64 (A = N or dN/dxi from a fragment table, B = node fields of 8 elements, all rows 33 doubles apart so
// that fragment loads and stores are conflict free; partner warps keep one accumulator column each): a third of the geometry's
// shared-memory loads, LSU 79 %, parity green -- and 2.21 against 2.05 (long-scoreboard stalls double behind the DMMA chains).
//
// Replaces for linear elements: MoVFEM_3DMT.f90:193-211 (element loop body), integration.f90:60-106 (int_elem_params,
// alocal, blocal), n_fem.f90:66-102,355-395, v_fem.f90:38-60, problem.f90:70-149.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "contract.cuh"
#include "element.cuh"

#ifndef F12_UNROLL_G
#define F12_UNROLL_G 4
#endif
#ifndef F12_UNROLL_L
#define F12_UNROLL_L 8
#endif

namespace movfem {

constexpr int kF12UnrollG = F12_UNROLL_G, kF12UnrollL = F12_UNROLL_L;

struct Fused12Args {
    MeshDims m;
    double omega;
    const ElemTables *T;
    const NodeRec *nodes;
    const double *xp, *yp, *zp;
    const double *soa;             // field-major node fields [6][soa_stride]: e, Re sigma 00 11 22, Im sigma 00 11 (node_kernel)
    size_t soa_stride;
    const int *list;               // element ids (0-based) of this launch, K/M row order
    int nlist;
    int e_base;
    double2 *KM;                   // K/M store at this launch's first row: [batch][78][32]
    double *be;                    // by K/M row (be_index): [row / 32][12][row % 32][4]
    int *status;
    const int *flags;              // flags[0]: any dmu != 0, [1]: Re sigma changed, [2]: any off-diagonal sigma, [3]: unequal diagonal (node_kernel)
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
    static constexpr int NW = 8, THREADS = NW * 32, MINB = DO_KM ? 2 : 3;
    static constexpr size_t QT_D = (size_t)12 * NGP * 32;                   // one Q|T block [component][g][lane] of doubles
    static constexpr size_t R_D = (size_t)8 * NGP * 32;                     // one source block: the 8 not identically zero components
    static constexpr size_t SF_D = (size_t)7 * MN * 32;                     // staged node fields [field][node][lane]: z, e, Re sigma 00 11 22, Im sigma 00 11
    static constexpr size_t TR_D = (size_t)2 * NGP * 32;                    // [2][g][lane] float2 (tr Q, tr T): a scale, rounded up
    static constexpr size_t TAB_D = (size_t)NGP * 4 * ME, DN_D = (size_t)MN * 4 * NGP, PHI_D = (size_t)NGP * ME, XY_D = 4 * 32;
    static constexpr size_t SMEM = sizeof(double) * ((DO_KM ? 2 * QT_D + TR_D : 0) + 2 * R_D + SF_D + TAB_D + DN_D + PHI_D + XY_D) +
                                   sizeof(unsigned long long) * 8 + sizeof(int) * (2 * ME + MN + 4 * 32 + 6 * 16);
};

// ISO: sigma = s I at every node (flags[3] == 0; every linear-element BASELINE mesh on its first frequency): three of the seven node fields are copies
// of others and are neither staged nor interpolated -- the results are bit-identical to the general variant's, which runs
// when some node has unequal diagonal entries (api.cu launches both; the device flag picks one -- both bodies in ONE kernel behind
// a branch on the flag: 2.48 against 2.03 ms, spills and twice the code in one image)
template <bool DO_KM, bool ISO>
__global__ void __launch_bounds__(Fused12Cfg<DO_KM>::THREADS, Fused12Cfg<DO_KM>::MINB) fused12_kernel(Fused12Args A) {
    using C = Fused12Cfg<DO_KM>;
    constexpr int MN = C::MN, ME = C::ME, NGP = C::NGP, NP = C::NP;
    constexpr int QS = NGP * 32;                                      // component stride inside a block
    if (A.skip_unless_changed && A.flags[1] == 0) return;
    if (A.flags[0] != 0 || A.flags[2] != 0) return;                   // mu != mu0 or off-diagonal sigma: the generic path runs instead
    if ((A.flags[3] == 0) != ISO) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_R = reinterpret_cast<double *>(smem_raw);               // [2][8][NGP][32]: (d0,pol1) (d1,pol2) (d2,pol1) (d2,pol2), re|im each
    double *s_qt = s_R + 2 * C::R_D;                                  // [2][12][NGP][32]: Q (0-5, sym3 order), T (6-11)   (DO_KM)
    double *s_tr = s_qt + (DO_KM ? 2 * C::QT_D : 0);                  // [2][NGP][32] float2                                  (DO_KM)
    double *s_sf = s_tr + (DO_KM ? C::TR_D : 0);                      // [7][MN][32]: z, e, Re sigma 00, 11, 22, Im sigma 00, 11
    double *s_tab = s_sf + C::SF_D;                                   // [NGP][4][ME]: dphi (0-2), phi (3), slot order
    double *s_dN = s_tab + C::TAB_D;                                  // [NGP][MN][4]: dN/dxi (0-2), N (3) -- the four weights of (g, node) in two 16-byte loads
    double *s_phi = s_dN + C::DN_D;                                   // [NGP][ME] phi in slot order
    double *s_xy = s_phi + C::PHI_D;                                  // [4][32]: x0, x1, y0, y1 of the staged batch
    uint64_t *full = reinterpret_cast<uint64_t *>(s_xy + C::XY_D), *empty = full + 2, *staged = full + 4;
    int *s_slot = reinterpret_cast<int *>(full + 8);                  // [ME] slot -> local DOF
    int *s_sdir = s_slot + ME;                                        // [ME] slot -> direction
    int *s_noff = s_sdir + ME;                                        // [MN] node offset from the element's base node
    int *s_ids = s_noff + MN;                                         // [4][32] element ids of batches j & 3
    int *s_pidx = s_ids + 4 * 32;                                     // [6 tiles][16]: packed-triangle index of pair (i, j) of the tile, -1: not stored

    const ElemTables &T = *A.T;
    const MeshDims &m = A.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int nbatch = (A.nlist + 31) / 32;
    const int G = (int)gridDim.x;
    const int nb = (int)blockIdx.x < nbatch ? (nbatch - (int)blockIdx.x + G - 1) / G : 0;   // batches of this CTA: blockIdx.x + j G

    for (int i = tid; i < NGP * ME; i += C::THREADS) s_phi[i] = T.phi[i / ME][T.slot_dof[i % ME]];
    if (DO_KM)
        for (int i = tid; i < (int)C::TAB_D; i += C::THREADS) s_tab[i] = g_ct_at[i];
    for (int i = tid; i < ME; i += C::THREADS) { s_slot[i] = T.slot_dof[i]; s_sdir[i] = T.slot_dir[i]; }
    for (int i = tid; i < MN; i += C::THREADS) s_noff[i] = T.node_off[i];
    for (int i = tid; i < MN * 4 * NGP; i += C::THREADS) s_dN[i] = T.dNt[(((i >> 2) % MN) * 4 + (i & 3)) * 32 + i / (4 * MN)];
    auto id_src = [&](int j_) {                                       // list entry of lane `lane` in this CTA's j-th batch (clamped)
        const int pos_ = ((int)blockIdx.x + j_ * G) * 32 + lane;
        return A.list + (pos_ < A.nlist ? pos_ : A.nlist - 1);
    };
    if (warp == 0 && nb > 0) s_ids[lane] = *id_src(0);
    if (DO_KM && tid < 6 * 16) {
        const int c_ = tid >> 4, i_ = (tid >> 2) & 3, j_ = tid & 3;
        const int si = 4 * c_ct.tile_ti[c_] + i_, sj = 4 * c_ct.tile_tj[c_] + j_, im = c_ct.slot_dof[si], jm = c_ct.slot_dof[sj];
        const int hi = im > jm ? im : jm, lo = im > jm ? jm : im;
        s_pidx[tid] = sj <= si ? hi * (hi + 1) / 2 + lo : -1;
    }
    if (tid == 0) {
        // every thread arrives for itself (release of its own stores / loads): 8 x 32 writers of a batch, 6 x 32 readers
        for (int b = 0; b < 2; ++b) { mbar_init(&full[b], 8 * 32); mbar_init(&empty[b], 6 * 32); }
        mbar_init(staged, 64);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (nb == 0) return;

    const bool fetcher = warp >= 6;                                   // warps 6, 7: node-field copies (four nodes each) and, with DO_KM, the RHS
    const int pw = warp - 6;
    const double psig = f32r(A.omega * kEps0);   // pset_pmodel, problem.f90:250
    const double w32 = f32r(A.omega);            // cmplx(0.d0,-omega), problem.f90:112
    // mu = mu0 I at every node (flags[0] == 0): one scalar for the whole mesh
    const double mu0inv = DO_KM ? __ldg(reinterpret_cast<const double *>(A.nodes + ((int64_t)(A.e_base / (m.ny * m.nz)) * m.nyz)) + 2) : 0.0;
    const double wt = T.rw[warp][3];

    // ---- node fields, x / y lines of batch j and the element ids of batch j+1: asynchronous copies, completion on `staged` ----
    auto cp_async = [&](void *dst_, const void *src_, auto bytes_tag) {
        constexpr int BYTES = decltype(bytes_tag)::value;
        const unsigned d_ = (unsigned)__cvta_generic_to_shared(dst_);
        if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d_), "l"(src_) : "memory");
        else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d_), "l"(src_) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d_), "l"(src_) : "memory");
    };
    auto prefetch = [&](int j_) {
        const int e_ = s_ids[(j_ & 3) * 32 + lane];
        int ie_, je_, ke_;
        elem_ijk(m, e_, ie_, je_, ke_);
        // Field-major node arrays: the 32 lanes of a request are 32 consecutive elements of a k-column (nearly always), i.e. 32
        // consecutive nodes -- 256 contiguous bytes per copy instead of 32 sectors of an array of records
        const int64_t n0_ = (int64_t)(ie_ - 1) * m.nyz + (int64_t)(je_ - 1) * m.nnz + (ke_ - 1);
#pragma unroll
        for (int l4 = 0; l4 < 4; ++l4) {
            const int l = 4 * pw + l4;
            const int64_t n_ = n0_ + s_noff[l];
            cp_async(s_sf + (0 * MN + l) * 32 + lane, A.zp + n_, std::integral_constant<int, 8>{});
#pragma unroll
            for (int f = 0; f < 6; ++f)
                if (!ISO || f == 0 || f == 1 || f == 4) cp_async(s_sf + ((1 + f) * MN + l) * 32 + lane, A.soa + f * A.soa_stride + n_, std::integral_constant<int, 8>{});
        }
        if (pw == 0) {
            cp_async(s_xy + lane, A.xp + ie_ - 1, std::integral_constant<int, 8>{}); cp_async(s_xy + 32 + lane, A.xp + ie_, std::integral_constant<int, 8>{});
            cp_async(s_xy + 64 + lane, A.yp + je_ - 1, std::integral_constant<int, 8>{}); cp_async(s_xy + 96 + lane, A.yp + je_, std::integral_constant<int, 8>{});
            if (j_ + 1 < nb) cp_async(s_ids + ((j_ + 1) & 3) * 32 + lane, id_src(j_ + 1), std::integral_constant<int, 4>{});
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(staged)) : "memory");
    };

    // ---- geometry of Gauss point `warp` of batch j: interpolation, J, det, G = J^-1, Q, T and the source R in closed form ----
    auto geometry = [&](int j_) {
        const int buf = j_ & 1, g = warp;
        mbar_wait(staged, (unsigned)(j_ & 1));
        if (j_ >= 2) mbar_wait(&empty[buf], (unsigned)(((j_ >> 1) - 1) & 1));      // the readers of this buffer (batch j-2) are done
        const bool live = ((int)blockIdx.x + j_ * G) * 32 + lane < A.nlist;
        const double a = 0.5 * (s_xy[32 + lane] - s_xy[lane]), b = 0.5 * (s_xy[96 + lane] - s_xy[64 + lane]);
        // interpolation (p_intmodels problem.f90:139-142; N_l-weighted part of p_source :424-457) and the xi-gradient of z
        // (nf_jacobian, n_fem.f90:359-366)
        double s0 = 0.0, s3 = 0.0, s5 = 0.0, e12 = 0.0, e15 = 0.0, e18 = 0.0, e21 = 0.0, zp = 0.0, zq = 0.0, zr = 0.0;
#pragma unroll kF12UnrollL
        for (int l = 0; l < MN; ++l) {
            const double2 ze = make_double2(s_sf[(0 * MN + l) * 32 + lane], s_sf[(1 * MN + l) * 32 + lane]);
            const double r00 = s_sf[(2 * MN + l) * 32 + lane], r11 = ISO ? r00 : s_sf[(3 * MN + l) * 32 + lane];
            const double di1 = s_sf[(5 * MN + l) * 32 + lane] - psig, di2 = ISO ? di1 : s_sf[(6 * MN + l) * 32 + lane] - psig;   // Im(dsigma), pdelta_model problem.f90:329-331
            const double2 dn01 = *reinterpret_cast<const double2 *>(s_dN + (g * MN + l) * 4), dn2n = *reinterpret_cast<const double2 *>(s_dN + (g * MN + l) * 4 + 2);
            const double ln = dn2n.y;
            const double le = ln * ze.y;                               // N_l e_l, e_l = f32(omega b0 z_l)
            if (DO_KM) { s0 = dfma(ln, r00, s0); if (!ISO) { s3 = dfma(ln, r11, s3); s5 = dfma(ln, s_sf[(4 * MN + l) * 32 + lane], s5); } }
            e12 = dfma(le, di1, e12); e15 = dfma(le, r00, e15);
            if (!ISO) { e18 = dfma(le, di2, e18); e21 = dfma(le, r11, e21); }
            zp = dfma(dn01.x, ze.x, zp);
            zq = dfma(dn01.y, ze.x, zq);
            zr = dfma(dn2n.x, ze.x, zr);
        }
        if (ISO) { s3 = s0; s5 = s0; e18 = e12; e21 = e15; }       // the same sums of the same values
        const double p = zp, q = zq, r = zr;
        const double det = (a * b) * r;
        if (det == 0.0 && live) atomicCAS(A.status, 0, -3);
        const double w = det * wt;
        const double rad = 1.0 / fabs(det);                            // Q6: nf_ji = adj(J)/abs(det)
        const double G00 = (b * r) * rad, G11 = (a * r) * rad, G22 = (a * b) * rad, G02 = -(p * b) * rad, G12 = -(a * q) * rad;
        if (DO_KM) {
            double *qo = s_qt + (size_t)buf * C::QT_D + g * 32 + lane;
            const double fm = (det < 0.0 ? -wt : wt) * rad * mu0inv;   // (w/det^2) mu^-1 = (wt/det) mu^-1: Q = fm J J^T
            const double q00 = fm * dfma(a, a, p * p), q11 = fm * dfma(b, b, q * q), q22 = fm * (r * r);
            qo[0 * QS] = q00; qo[1 * QS] = fm * (p * q); qo[2 * QS] = fm * (p * r);
            qo[3 * QS] = q11; qo[4 * QS] = fm * (q * r); qo[5 * QS] = q22;
            // T = G^T S G, S = w Re sigma_g (integration.f90:234-236, Q3), sigma diagonal
            const double S0 = w * s0, S1 = w * s3, S2 = w * s5;
            const double t00 = S0 * (G00 * G00), t11 = S1 * (G11 * G11);
            const double t22 = dfma(S0 * G02, G02, dfma(S1 * G12, G12, (S2 * G22) * G22));
            qo[6 * QS] = t00; qo[7 * QS] = 0.0; qo[8 * QS] = (S0 * G00) * G02;
            qo[9 * QS] = t11; qo[10 * QS] = (S1 * G11) * G12; qo[11 * QS] = t22;
            // the element's K / M scale for the tiny-pair test of the epilogue (maximum over the Gauss points, taken by the readers)
            reinterpret_cast<float2 *>(s_tr)[(buf * NGP + g) * 32 + lane] =
                make_float2(__double2float_ru(fabs(q00) + fabs(q11) + fabs(q22)), __double2float_ru(fabs(t00) + fabs(t11) + fabs(t22)));
        }
        // R[d][pol] = G[:,d] . (w src_pol);  src = dmpf * cmplx32(0,-omega) (problem.f90:112); for a diagonal sigma
        // src_1 = (A1, 0, 0), src_2 = (0, A2, 0):  pol 1 dmpf = (+Im ds*e, -Re ds*e), pol 2 = (-Im ds*e, +Re ds*e);
        // R[1][pol 1] and R[0][pol 2] are identically zero and not stored
        const double a1x = w * (-e15 * w32), a1y = w * (-(e12 * w32)), a2x = w * (e21 * w32), a2y = w * (e18 * w32);
        double *Ro = s_R + (size_t)buf * C::R_D + g * 32 + lane;
        Ro[0 * QS] = G00 * a1x; Ro[1 * QS] = G00 * a1y; Ro[2 * QS] = G11 * a2x; Ro[3 * QS] = G11 * a2y;
        Ro[4 * QS] = G02 * a1x; Ro[5 * QS] = G02 * a1y; Ro[6 * QS] = G12 * a2x; Ro[7 * QS] = G12 * a2y;
        mbar_arrive(&full[buf]);                                       // release: this thread's share of the batch is in place
    };

    // ---- RHS unit (four slots q4..q4+3, polarisation pol) of batch j: b_e(j) = sum_g phi_j(g) R[d_j][pol](g)
    //      (blocal / f3, integration.f90:96-104,258-263) ----
    auto rhs_sums = [&](int buf, int q4, int cd, int pol, double (&br)[4], double (&bi)[4]) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) { br[kk] = 0.0; bi[kk] = 0.0; }
        // diagonal sigma: the source of polarisation 1 is along x, of polarisation 2 along y, so R[d][pol] is identically zero
        // for (d, pol) = (1, 1) and (0, 2): zeros are delivered, nothing is summed
        if (cd == 2 || cd == pol) {
            const double *Rr = s_R + (size_t)buf * C::R_D + (size_t)(cd == 2 ? 4 + 2 * pol : 2 * pol) * QS + lane, *Ri = Rr + QS;
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
    };
    // one polarisation of four slots (the RHS-only pass: six such units on six warps), 16-byte stores
    auto rhs_unit = [&](int j_, int q4, int pol) {
        if (!(((int)blockIdx.x + j_ * G) * 32 + lane < A.nlist)) return;
        // b_e by K/M row: lanes are 32 consecutive rows, so a store covers 1 kB (the other polarisation fills the gaps)
        double2 *bo = reinterpret_cast<double2 *>(A.be) + ((size_t)((int)blockIdx.x + j_ * G) * ME * 32 + lane) * 2 + pol;
        double br[4], bi[4];
        rhs_sums(j_ & 1, q4, s_sdir[q4], pol, br, bi);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) bo[(size_t)s_slot[q4 + kk] * 64] = make_double2(br[kk], bi[kk]);
    };
    // both polarisations of four slots, one 32-byte store per slot (st.global.v4.f64): a warp's store is 1 kB without gaps --
    // 8 wavefronts for 32 bytes per lane instead of 2 x 8 for 2 x 16
    auto rhs_group = [&](int j_, int q4) {
        if (!(((int)blockIdx.x + j_ * G) * 32 + lane < A.nlist)) return;
        double *bo = A.be + ((size_t)((int)blockIdx.x + j_ * G) * ME * 32 + lane) * 4;
        const int cd = s_sdir[q4];
        double br0[4], bi0[4], br1[4], bi1[4];
        rhs_sums(j_ & 1, q4, cd, 0, br0, bi0);
        rhs_sums(j_ & 1, q4, cd, 1, br1, bi1);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(bo + (size_t)s_slot[q4 + kk] * 128), "d"(br0[kk]), "d"(bi0[kk]), "d"(br1[kk]),
                         "d"(bi1[kk]) : "memory");
    };

    // ---- tile `c` (direction class c) of batch j: the inner loop of contract.cuh on the Q|T the eight warps left in shared memory ----
    const int c = warp < 6 ? warp : 0;
    const int ti = c_ct.tile_ti[c], tj = c_ct.tile_tj[c];
    const int dI = cls_dI(c), dJ = cls_dJ(c);
    const int k1I = dI == 2 ? 1 : 2, k2I = dI == 0 ? 1 : 0, k1J = dJ == 2 ? 1 : 2, k2J = dJ == 0 ? 1 : 0;
    const double tau = ((dI == 1) != (dJ == 1)) ? -1.0 : 1.0;
    const int cq0 = c_ct.comp[0][c][0], cq1 = c_ct.comp[0][c][1], cq2 = c_ct.comp[0][c][2], cq3 = c_ct.comp[0][c][3], cq4 = c_ct.comp[0][c][4];
    const double *Y1 = s_tab + k1I * ME + 4 * ti, *Y2 = s_tab + k2I * ME + 4 * ti, *Y3 = s_tab + 3 * ME + 4 * ti;
    const double *X1 = s_tab + k1J * ME + 4 * tj, *X2 = s_tab + k2J * ME + 4 * tj, *X3 = s_tab + 3 * ME + 4 * tj;
    auto tile = [&](int j_) {
        const int buf = j_ & 1, batch = (int)blockIdx.x + j_ * G;
        const double *S = s_qt + (size_t)buf * C::QT_D + lane;
        double accK[16], accM[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { accK[i] = 0.0; accM[i] = 0.0; }
        // a diagonal tile (classes xx, yy, zz): the column operands are the row operands (six broadcast loads less per Gauss
        // point: the kernel is bound by the shared-memory pipe) and the six pairs above the diagonal are not formed
        auto contract = [&](auto diag_tag) {
            constexpr bool DIAG = decltype(diag_tag)::value;
#pragma unroll kF12UnrollG
            for (int g = 0; g < NGP; ++g) {
                const int o = g * 4 * ME;
                const double q00 = S[(cq0 * NGP + g) * 32], q01 = S[(cq1 * NGP + g) * 32], q10 = S[(cq2 * NGP + g) * 32],
                             q11 = S[(cq3 * NGP + g) * 32], tt = S[(cq4 * NGP + g) * 32];
                double b1[4], b2[4], bw[4], xa[4], xb[4], xc[4], ya[4], yb[4], yc[4];
                ld4(ya, Y1 + o); ld4(yb, Y2 + o); ld4(yc, Y3 + o);
                if (DIAG) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { xa[j] = ya[j]; xb[j] = yb[j]; xc[j] = yc[j]; }
                } else { ld4(xa, X1 + o); ld4(xb, X2 + o); ld4(xc, X3 + o); }
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
        float trq = 0.f, trt = 0.f;
#pragma unroll
        for (int g = 0; g < NGP; ++g) {
            const float2 t2 = reinterpret_cast<const float2 *>(s_tr)[(buf * NGP + g) * 32 + lane];
            trq = fmaxf(trq, t2.x); trt = fmaxf(trt, t2.y);
        }
        const double thrK = A.no_l1 ? -1.0 : kTinyRelC * NGP * (double)trq, thrM = A.no_l1 ? -1.0 : kTinyRelC * NGP * (double)trt;
        mbar_arrive(&empty[buf]);                           // the buffer (and its scales) may be refilled
        // write-out: packed lower triangle by LOCAL DOF index, 32 elements interleaved -> 512-byte coalesced stores
        if (batch * 32 + lane < A.nlist) {
            double2 *KMo = A.KM + (size_t)batch * NP * 32 + lane;
            const int *pi = s_pidx + c * 16;
            unsigned fm = 0u;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int p = pi[t];
                if (p >= 0) {
                    const double kv = accK[t] * tau, mv = accM[t];
                    KMo[p * 32] = make_double2(kv, mv);
                    const double ak = fabs(kv), am = fabs(mv);
                    if ((ak <= thrK && am <= thrM) || (ak < kFlagAbsC && am < kFlagAbsC)) fm |= 1u << t;
                }
            }
            if (fm) {                                       // tiny pairs (rare): flagged for exact_kernel
                uint32_t *pfl = A.pairflags + (size_t)(batch * 32 + lane) * A.W;
                atomicAdd(A.nflag, (unsigned long long)__popc(fm));
                atomicOr(A.batchany + batch, 1u << lane);
                while (fm) {
                    const int t = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const int p = pi[t];
                    atomicOr(pfl + (p >> 5), 1u << (p & 31));
                }
            }
        }
    };

    // ================= the pipeline: geometry of batch k+1 (all warps), then the tile / RHS / copies of batch k =================
    if (fetcher) prefetch(0);
    geometry(0);
    if (fetcher && nb > 1) { mbar_wait(&full[0], 0u); prefetch(1); }   // (one staging buffer: batch j+1 is copied once every warp has read batch j)
    for (int k = 0; k < nb; ++k) {
        if (k + 1 < nb) geometry(k + 1);
        const unsigned par = (unsigned)((k >> 1) & 1);
        if (!fetcher) {
            mbar_wait(&full[k & 1], par);
            if (DO_KM) tile(k);
            else {
                rhs_unit(k, 4 * (warp >> 1), warp & 1);
                mbar_arrive(&empty[k & 1]);                            // R of this buffer may be overwritten
            }
        } else {
            if (DO_KM) {
                mbar_wait(&full[k & 1], par);
                // warp 6: the x and y slot groups (one non-zero polarisation each), warp 7: the z group (both)
#pragma unroll 1
                for (int q4 = 0; q4 < ME; q4 += 4)
                    if ((s_sdir[q4] == 2) == (pw == 1)) rhs_group(k, q4);
            }
            if (k + 2 < nb) {
                mbar_wait(&full[(k + 1) & 1], (unsigned)(((k + 1) >> 1) & 1));   // every warp has read the staged fields of batch k+1
                prefetch(k + 2);
            }
        }
    }
}

}  // namespace movfem

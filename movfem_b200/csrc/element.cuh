// movfem_b200/csrc/element.cuh -- the hot kernels: per-node fields and element matrices.
//
// Replaces, per element:
//   n_fem.f90:66-102,355-395      nf_get_r, nf_jacobian, nf_inv_jac, nf_det
//   v_fem.f90:38-60,470-519       vf_elem_ve, vf_elem_curl, mix_grad_ln, grad_xi
//   problem.f90:70-149,247-457    p_elem_fields, p_intmodels, p_source (+ helpers)
//   boundary_conds.f90:72-186     get_pml (consumed one element late, Q17), gpml_h
//   integration.f90:60-265        int_elem_params, alocal/f1/f2, blocal/f3
//
// Formulation.  The reference evaluates one alocal(im,jm) at a time (36-term f1, 9-term f2, per
// Gauss point, 3*me+mn+1 Jacobian rebuilds per point).  Here every Gauss point gets ONE Jacobian
// J, G = J^-1 (both with the reference's exact operation order), and the edge basis
// N_e = phi_e * grad(xi_d) = phi_e * G[:,d] is never formed per DOF.  Instead the tensor-product
// structure is pushed into three small per-Gauss-point tensors:
//
//   mass    N_i . S N_j            = phi_i phi_j T[d_i][d_j],         T = G^T S G        (S = w Re[h1h2h3 sigma])
//   curl    curl N_j = (G dphi_j) x (G e_d) = (1/det J) J^T (dphi_j x e_d) = (1/det J) J^T c_j
//           curl N_i . D curl N_j  = c_i^T Q c_j,                      Q = (w/det^2) J mu^-1 J^T
//           (c_j = dphi_j x e_d is a CONSTANT of the reference element with two non-zero components)
//   source  N_j . (w h1h2h3 src_p) = phi_j R[d_j][p],                  R[d][p] = G[:,d] . (w h src_p)
//
// so that per (DOF pair, Gauss point) the element matrices cost 2 (K) + 1 (M) FMAs instead of the
// 3 + 3 of the B^T D B form (and 143 flops of the reference's expanded f1/f2).  Inside the GPML
// layers the stretched half-curls do not collapse to a curl; there K uses the 9x9 real tensor
// P = H^T D6 H with H[a][(u,d)] = s_a G[x_a][u] G[y_a][d] and costs 3 FMAs per pair and point.
//
// geometry_kernel (this file) produces Q|P, T per (element, Gauss point) into an L2-resident scratch and the
// element RHS b_e; contract_kernel (contract.cuh) turns the scratch into K_e, M_e.
// A_e = K_e + i*f32(omega)*M_e is formed later, per frequency (finalize.cuh), so K_e, M_e of the
// unstretched elements are frequency independent and cached in HBM across a sweep.
#pragma once
#include "common.cuh"

namespace movfem {

// ------------------------------------------------------------------------------------------
// node kernel: problem.f90:257-358 per grid node instead of per (element, node)
// (Measured and dropped: writing only the frequency-dependent half of a record -- f32(omega b0 z), sigma -- after the first
// assembly of a handle, the rest depending on g_zp and g_mu only: the partial-sector stores into the 208-byte records make the
// kernel 1.8 x SLOWER, 0.134 against 0.073 ms on config 4, although it moves 35 % fewer bytes.)
// ------------------------------------------------------------------------------------------
__global__ void node_kernel(int n0, int npt, double omega, const double *__restrict__ zp, const double *__restrict__ mu,
                            const double2 *__restrict__ sigma, NodeRec *__restrict__ out, int *__restrict__ status,
                            int *__restrict__ flags /* [0]: any dmu != 0, [1]: Re sigma changed, [2]: any off-diagonal sigma, [3]: sigma_11 != sigma_22 or _33 */, int check_re,
                            double *__restrict__ soa /* linear elements: field-major copy [6][npt_all] of e, Re sigma 00 11 22, Im sigma 00 11 (fused12.cuh) */,
                            size_t soa_stride) {
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;   // [n0, npt): the node planes this handle's slab touches
    if (i >= npt) return;
    double a[6];
    double2 s[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { a[k] = mu[(size_t)6 * i + k]; s[k] = sigma[(size_t)6 * i + k]; }
    // cdet / det singularity stops of problem.f90:260-271 (sigma^-1 itself is unused for pe_sch=1)
    auto cm = [](double2 x, double2 y) { return make_double2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x); };
    auto cs = [](double2 x, double2 y) { return make_double2(x.x - y.x, x.y - y.y); };
    auto ca = [](double2 x, double2 y) { return make_double2(x.x + y.x, x.y + y.y); };
    const double2 cd = ca(ca(cm(s[0], cs(cm(s[3], s[5]), cm(s[4], s[4]))), cm(s[1], cs(cm(s[2], s[4]), cm(s[1], s[5])))),
                          cm(s[2], cs(cm(s[1], s[4]), cm(s[3], s[2]))));
    const double det = a[0] * (a[3] * a[5] - a[4] * a[4]) + a[1] * (a[2] * a[4] - a[1] * a[5]) + a[2] * (a[1] * a[4] - a[3] * a[2]);
    if ((cd.x == 0.0 && cd.y == 0.0) || det == 0.0) { atomicCAS(status, 0, -4); return; }
    NodeRec r;
    r.z = zp[i];
    r.e = f32r(omega * kB0 * r.z);
    r.inmu[0] = (a[3] * a[5] - a[4] * a[4]) / det;
    r.inmu[1] = (a[2] * a[4] - a[1] * a[5]) / det;
    r.inmu[2] = (a[1] * a[4] - a[2] * a[3]) / det;
    r.inmu[3] = (a[0] * a[5] - a[2] * a[2]) / det;
    r.inmu[4] = (a[2] * a[1] - a[0] * a[4]) / det;
    r.inmu[5] = (a[3] * a[0] - a[1] * a[1]) / det;
    bool changed = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (check_re && out[i].sre[k] != s[k].x) changed = true;
        r.sre[k] = s[k].x; r.sim[k] = s[k].y;
    }
    // pdelta_model + pe_modelcurl, problem.f90:329-334,391-403
    const double pmu = 4 * kPi * 1.e-7;
    const double d[6] = {a[0] - pmu, a[1], a[2], a[3] - pmu, a[4], a[5] - pmu};
    const double *im = r.inmu;
    const double hp = f32r(kB0) / (4 * kPi * 1.e-7);
    // column 2 of m = mu^-1 dmu (pol 1: Hp along y), column 1 (pol 2: Hp along x)
    r.vc[0] = (im[0] * d[1] + im[1] * d[3] + im[2] * d[4]) * hp;
    r.vc[1] = (im[1] * d[1] + im[3] * d[3] + im[4] * d[4]) * hp;
    r.vc[2] = (im[2] * d[1] + im[4] * d[3] + im[5] * d[4]) * hp;
    r.vc[3] = (im[0] * d[0] + im[1] * d[1] + im[2] * d[2]) * hp;
    r.vc[4] = (im[1] * d[0] + im[3] * d[1] + im[4] * d[2]) * hp;
    r.vc[5] = (im[2] * d[0] + im[4] * d[1] + im[5] * d[2]) * hp;
    bool anyd = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) anyd |= (d[k] != 0.0);
    if (anyd) flags[0] = 1;
    if (changed) flags[1] = 1;
    if (s[1].x != 0.0 || s[1].y != 0.0 || s[2].x != 0.0 || s[2].y != 0.0 || s[4].x != 0.0 || s[4].y != 0.0) flags[2] = 1;
    if (s[0].x != s[3].x || s[0].x != s[5].x || s[0].y != s[3].y) flags[3] = 1;   // the diagonal entries fused12_kernel reads differ
    out[i] = r;
    if (soa) {
        soa[i] = r.e; soa[soa_stride + i] = r.sre[0]; soa[2 * soa_stride + i] = r.sre[3]; soa[3 * soa_stride + i] = r.sre[5];
        soa[4 * soa_stride + i] = r.sim[0]; soa[5 * soa_stride + i] = r.sim[3];
    }
}

// ------------------------------------------------------------------------------------------
// element kernel
// ------------------------------------------------------------------------------------------
struct ElemArgs {
    MeshDims m;
    PmlParams pml;
    double omega;
    const ElemTables *T;
    const NodeRec *nodes;
    const double *xp, *yp;
    const int *list;      // element ids (0-based) this launch handles
    int nlist;
    int e_base;           // first element stored in KM / be (slab handles keep only their slab + halo)
    int64_t be_row0;      // K/M row of list[0] (b_e is stored by row, be_index)
    double *qt;           // scratch [nbatch32][NCMP][NGP][32]: Q|P and T per (element, Gauss point), see contract.cuh
    double *be;           // [ne][ME][4]  (re,im) x 2 polarisations
    double2 *escale;      // [list position]: (ngp * max_g tr Q|P, ngp * max_g tr T): the element's K / M magnitude, the scale the
                          // contraction's tiny-pair test (exact.cuh) is relative to
    int *status;
    const int *flags;     // flags[0] any dmu, flags[1] Re sigma changed
    int skip_unless_changed;   // launch is a cache refresh: exit unless flags[1]
    int skip_if_simple;        // linear elements: exit when fused12_kernel handles the list (flags[0] == 0 and flags[2] == 0)
    int phase_mask;            // profiling aid (MOVFEM_PHASE_MASK): bit0 B, bit1 RHS; default 3
};

// columns interpolated to the Gauss points by phase B1 (record offsets into NodeRec, see common.cuh):
//   0-5 mu^-1, 6-11 Re sigma, 12-14 e*Im(dsigma col 1), 15-17 e*Re(sigma col 1), 18-20 e*Im(dsigma col 2),
//   21-23 e*Re(sigma col 2), [GPML only] 24-29 Im sigma
__constant__ int kColOff[30] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 8, 9, 10, 15, 17, 18, 9, 11, 12, 14, 15, 16, 17, 18, 19};
__constant__ int kColFlag[30] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 1, 1, 1, 1, 1, 1, 3, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0};   // bit0: times e, bit1: minus psig first

// x / y line index of local node l in the 27-node numbering (n_fem.f90:36-59, i1-1 / j1-1 for nord = 3; the 8- and
// 20-node tables are prefixes, with 2 -> nord-1)
__device__ constexpr int kNodeI27[27] = {2, 2, 0, 0, 2, 2, 0, 0, 2, 1, 0, 1, 2, 1, 0, 1, 2, 2, 0, 0, 2, 1, 0, 1, 1, 1, 1};
__device__ constexpr int kNodeJ27[27] = {0, 2, 2, 0, 0, 2, 2, 0, 1, 2, 1, 0, 1, 2, 1, 0, 0, 2, 2, 0, 1, 2, 1, 0, 1, 1, 1};

template <int MN_, int ME_, int MEP_, int NGP_, int EB_, int THREADS_, int MINB_, bool PML_>
struct ElemCfg {
    static constexpr int MN = MN_, ME = ME_, MEP = MEP_, NGP = NGP_, EB = EB_, THREADS = THREADS_, MINB = MINB_;
    static constexpr bool PML = PML_;
    static constexpr int NREC = 20;                        // doubles of the node record staged in smem (z,e,mu^-1,Re/Im sigma; not vc)
    static constexpr int NDW = NREC + 2;                   // staged record + x + y
    static constexpr int NCOL = PML ? 30 : 24;             // columns interpolated by phase B1
    // per (Gauss point, element) record: the columns, then R (12) in place; stored [g][element] with an odd number of
    // 16-byte chunks per record, so the 8 lanes of a quarter warp (8 elements) hit distinct bank groups
    static constexpr int GEO = (NCOL / 2) % 2 ? NCOL : NCOL + 2;
    static constexpr int NCMP = PML ? 51 : 12;             // scratch components: P(45)|Q(6), T(6)
    static constexpr int MNP = (MN + 1) & ~1;              // row stride of the N table (16-byte aligned rows)
    static constexpr int NGPP = (NGP + 1) & ~1;            // row stride of the dN table
    // shared memory: phi [NGP][MEP] + dN|N [MN][4][NGPP] | records [EB][NGP][GEO] | node records [EB][MN][NDW]
    static constexpr size_t ATAB_D = (size_t)NGP * MEP + (size_t)MN * 4 * NGPP, GEO_D = (size_t)EB * NGP * GEO;
    static constexpr int NSTR = MN * NDW + 2;              // per-element stride of the node records (+16 B: bank shift)
    static constexpr size_t NODES_D = (size_t)EB * NSTR;
    static constexpr size_t ZXY_D = 0;
    static constexpr size_t SMEM = sizeof(double) * (ATAB_D + GEO_D + NODES_D + ZXY_D + EB + 1 + 2 * EB) + sizeof(int) * (EB * 4 + 2 * MEP + 2 * EB + 3 * MN);
    static_assert((ATAB_D % 2) == 0 && (GEO_D % 2) == 0, "16-byte alignment of the smem regions");
    static_assert(32 % EB == 0, "a batch of the contraction (32 lanes) is a whole number of geometry batches");
};

// gpml_h for one axis, boundary_conds.f90:94-123, AS STORED by the reference: integration.f90:16 declares the
// per-Gauss-point stretch array `gpml` REAL(kind=double), so `gpml(i,:)=gpml_h(...)` (integration.f90:125) keeps
// Re(h) only and f1/f2/f3 read it back with a zero imaginary part (found by executing the reference
// source).  Scheme 0 (Fang): Re[hx0*cmplx(1,-bx/(omega*eps))] = hx0 = 1 + a0*rho^n; scheme 1 (Zhou):
// Re[1 + cmplx(0,bx)] = 1.  The imaginary parts (the sin^2 profile, the Smith division) never reach the matrix.
__device__ __forceinline__ double2 gpml_axis(const PmlParams &p, int flag, int axis, double r, double omega) {
    if (flag == 0 || p.sch != 0) return make_double2(1.0, 0.0);
    const int s = flag < 0 ? 0 : 1;
    const double dl = p.b[axis][s] - p.a[axis][s], rr_pml = sqrt(dl * dl);
    const double dr = r - p.a[axis][s], rr = sqrt(dr * dr);
    const double rho = rr / rr_pml;
    const double pw = (p.nn == 2.0) ? rho * rho : (p.nn == 1.0 ? rho : pow(rho, p.nn));
    const double hx0 = 1.0 + p.a0 * pw;
    return make_double2(hx0 * 1.0, 0.0);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cdivf(double2 a, double2 b) {   // Fortran rules (Smith)
    if (fabs(b.x) < fabs(b.y)) {
        const double ratio = b.x / b.y, div = (b.x * ratio) + b.y;
        return make_double2(((a.x * ratio) + a.y) / div, ((a.y * ratio) - a.x) / div);
    }
    const double ratio = b.y / b.x, div = (b.y * ratio) + b.x;
    return make_double2(((a.y * ratio) + a.x) / div, (a.y - (a.x * ratio)) / div);
}

// index of (r,c) in a packed symmetric 3x3 (11,12,13,22,23,33)
__host__ __device__ __forceinline__ constexpr int sym3(int r, int c) {
    return r == c ? (r == 0 ? 0 : (r == 1 ? 3 : 5)) : ((r + c == 1) ? 1 : ((r + c == 2) ? 2 : 4));
}
// index of (r,c), r <= c, in a packed upper-triangular 9x9 stored row-major
__host__ __device__ __forceinline__ constexpr int up9(int r, int c) { return r * 9 - r * (r - 1) / 2 + (c - r); }

// DO_QT = false: RHS-only variant for a later frequency of a sweep (K_e, M_e are cached -- of the stretched elements
// too, because the stored stretch Re(h) does not depend on omega, Q18): only the source columns are interpolated and
// Q|P, T are neither formed nor written; the RHS of a stretched element still carries h1*h2*h3.
template <class CFG, bool DO_QT>
__global__ void __launch_bounds__(CFG::THREADS, CFG::MINB) geometry_kernel(ElemArgs A) {
    constexpr int MN = CFG::MN, ME = CFG::ME, MEP = CFG::MEP, EB = CFG::EB, NGP = CFG::NGP;
    constexpr int GEO = CFG::GEO, NDW = CFG::NDW, NCMP = CFG::NCMP;
    constexpr bool PML = CFG::PML;
    constexpr int GR = 0;                                   // R overwrites the record's first 12 columns
    constexpr int CLO = DO_QT ? 0 : 12, CHI = DO_QT ? CFG::NCOL : 24, NCOLA = CHI - CLO;   // active columns
    if (A.skip_unless_changed && A.flags[1] == 0) return;
    if (A.skip_if_simple && A.flags[0] == 0 && A.flags[2] == 0) return;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_phi = reinterpret_cast<double *>(smem_raw);             // [NGP][MEP] phi in slot order
    double *s_dN = s_phi + NGP * MEP;                              // [MN][4][NGPP]: dN/dxi (0..2), N (3); Gauss point fastest
    double *s_geo = s_phi + CFG::ATAB_D;                              // [NGP][EB][GEO]
    double *s_nodes = s_geo + CFG::GEO_D;                             // [EB][MN][NDW]
    int64_t *s_rbase = reinterpret_cast<int64_t *>(s_nodes + CFG::NODES_D + CFG::ZXY_D);   // [EB] base node id of the batch being prefetched
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_rbase + EB);     // mbarrier of the node-record bulk copies
    unsigned long long *s_scale = reinterpret_cast<unsigned long long *>(s_bar + 1);   // [EB][2]: max_g tr Q|P, max_g tr T (bit patterns of >= 0 doubles)
    int *s_el = reinterpret_cast<int *>(s_scale + 2 * EB);            // [EB][4]: element id, GPML flags
    int *s_slot = s_el + EB * 4;                                      // [MEP] slot -> local DOF (0-based) or -1
    int *s_sdir = s_slot + MEP;                                       // [MEP] slot -> direction (0-based)
    int *s_rxy = s_sdir + MEP;                                        // [EB][2] x / y line index of the prefetched elements' base node
    int *s_noff = s_rxy + 2 * EB;                                     // [MN][3] node offset, i1-1, j1-1
    constexpr int NGPP = CFG::NGPP, NREC = CFG::NREC;

    const ElemTables &T = *A.T;
    const MeshDims &m = A.m;
    const int tid = threadIdx.x;

    for (int i = tid; i < NGP * MEP; i += CFG::THREADS) {
        const int g = i / MEP, sl = i % MEP;
        const int dof = T.slot_dof[sl];
        s_phi[i] = dof >= 0 ? T.phi[g][dof] : 0.0;
    }
    for (int i = tid; i < MEP; i += CFG::THREADS) { s_slot[i] = T.slot_dof[i]; s_sdir[i] = T.slot_dir[i]; }
    for (int i = tid; i < MN; i += CFG::THREADS) { s_noff[i * 3] = T.node_off[i]; s_noff[i * 3 + 1] = T.node_i[i]; s_noff[i * 3 + 2] = T.node_j[i]; }
    for (int i = tid; i < MN * 4 * NGPP; i += CFG::THREADS) {
        const int g = i % NGPP, lm = i / NGPP;
        s_dN[i] = g < NGP ? T.dNt[lm * 32 + g] : 0.0;
    }

    // phase-B1 GEMM fragments (mma.m8n8k4: A row = lane/4, k = lane%4; B k = lane%4, n = lane/4; C row = lane/4, cols 2(lane%4)+{0,1})
    constexpr int NW = CFG::THREADS / 32, MB = (NGP + 7) / 8, MH = 1, MBH = MB / MH, KB = (MN + 3) / 4, NB = (NCOLA + 7) / 8;
    const int lane = tid & 31;
    constexpr int b1_mh = 0;
    const double psig = f32r(A.omega * kEps0);   // pset_pmodel, problem.f90:250
    int b1_off[NB], b1_flag[NB];
#pragma unroll
    for (int nbk = 0; nbk < NB; ++nbk) {
        const int c = CLO + 8 * nbk + (lane >> 2);
        b1_off[nbk] = c < CHI ? kColOff[c] : -1; b1_flag[nbk] = c < CHI ? kColFlag[c] : 0;
    }

    const int nbatch = (A.nlist + EB - 1) / EB;
    // asynchronous gather of one batch's node records into s_nodes (16-byte cp.async pieces, coalesced per record)
    // (threads < EB first resolve the batch's elements to base node ids: prepare_request, one barrier earlier)
    auto prepare_request = [&](int b) {
        const int first = b * EB, nb = min(EB, A.nlist - first);
        if (tid < nb) {
            int ie, je, ke;
            elem_ijk(m, A.list[first + tid], ie, je, ke);
            const int g1 = m.nord - 1;
            s_rbase[tid] = (int64_t)(ie - 1) * g1 * m.nyz + (int64_t)(je - 1) * g1 * m.nnz + (ke - 1) * g1;
            s_rxy[tid * 2] = (ie - 1) * g1; s_rxy[tid * 2 + 1] = (je - 1) * g1;
        }
    };
    // one 160-byte bulk copy (TMA 1-D) per node record, completion counted on s_bar; x,y lines by plain stores
    auto request_nodes = [&](int b) {
        const int nb = min(EB, A.nlist - b * EB);
        if (tid == 0) mbar_expect_tx(s_bar, (unsigned)(nb * MN * NREC * sizeof(double)));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of s_nodes vs the async writes
        for (int i = tid; i < nb * MN; i += CFG::THREADS) {
            const int l = i % MN, s = i / MN;
            double *dst = s_nodes + s * CFG::NSTR + l * NDW;
            bulk_g2s(dst, A.nodes + (s_rbase[s] + s_noff[l * 3]), (unsigned)(NREC * sizeof(double)), s_bar);
            *reinterpret_cast<double2 *>(dst + NREC) = make_double2(A.xp[s_rxy[s * 2] + s_noff[l * 3 + 1]], A.yp[s_rxy[s * 2 + 1] + s_noff[l * 3 + 2]]);
        }
    };
    if (tid == 0) { mbar_init(s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if ((int)blockIdx.x < nbatch) prepare_request(blockIdx.x);
    __syncthreads();
    if ((int)blockIdx.x < nbatch) request_nodes(blockIdx.x);
    unsigned node_phase = 0;
    for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int first = batch * EB;
        const int nb = min(EB, A.nlist - first);
        __syncthreads();   // previous batch fully consumed (s_geo / s_el reuse); also publishes the tables
        if (tid < EB) {
            const int e = tid < nb ? A.list[first + tid] : -1;
            s_el[tid * 4] = e;
            int f[3] = {0, 0, 0};
            if (PML && e >= 0) effective_pml(m, A.pml, e, f);
            s_el[tid * 4 + 1] = f[0]; s_el[tid * 4 + 2] = f[1]; s_el[tid * 4 + 3] = f[2];
            s_scale[tid * 2] = 0ull; s_scale[tid * 2 + 1] = 0ull;
        }

        // ---- phase A: the node records of this batch were requested with bulk copies while the previous batch was in
        //      its RHS phase (s_nodes is dead after phase B2); wait for them here ----
        mbar_wait(s_bar, node_phase);
        node_phase ^= 1;
        __syncthreads();

        // ---- phase B1: interpolate node data to the Gauss points (p_intmodels problem.f90:139-142 and the N_l-weighted
        //      part of p_source problem.f90:424-457): per element the small GEMM  out[g][c] = sum_l N[g][l] V[l][c]
        //      (NGP x MN) x (MN x NCOL) on the FP64 tensor-core path, mma.sync.m8n8k4.f64.  A = N is a constant of the
        //      element type and stays in registers for the whole kernel (the warp always owns the same Gauss-point
        //      blocks); B = the node-record columns, read straight from the staged records with their per-column
        //      transform; one warp per (element, half of the Gauss-point blocks). ----
        if (A.phase_mask & 1) {
            const int wid = tid >> 5;
            // A fragments (N, a constant of the element type) are re-read from L1 at the start of the phase instead of
            // occupying registers through B2 and the RHS phase
            double b1_a[MBH][KB];
            if (wid < nb) {
#pragma unroll
                for (int mi = 0; mi < MBH; ++mi)
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
                        const int g = 8 * mi + (lane >> 2), l = 4 * kb + (lane & 3);
                        b1_a[mi][kb] = (g < NGP && l < MN) ? T.N[g][l] : 0.0;
                    }
            }
            for (int task = wid; task < nb * MH; task += NW) {
                const int s = task / MH;
                const double *nd = s_nodes + s * CFG::NSTR;
                double acc[MBH][NB][2];
#pragma unroll
                for (int mi = 0; mi < MBH; ++mi)
#pragma unroll
                    for (int nbk = 0; nbk < NB; ++nbk) acc[mi][nbk][0] = acc[mi][nbk][1] = 0.0;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    const int l = 4 * kb + (lane & 3);
                    const bool lv = l < MN;
                    const double el = lv ? nd[l * NDW + 1] : 0.0;
#pragma unroll
                    for (int nbk = 0; nbk < NB; ++nbk) {
                        double bv = 0.0;
                        if (lv && b1_off[nbk] >= 0) {
                            bv = nd[l * NDW + b1_off[nbk]];
                            if (b1_flag[nbk] & 2) bv = bv - psig;       // Im(dsigma) on the diagonal (pdelta_model)
                            if (b1_flag[nbk] & 1) bv = bv * el;         // times e_l = f32(omega b0 z_l): |Ep| at the node
                        }
#pragma unroll
                        for (int mi = 0; mi < MBH; ++mi)
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(acc[mi][nbk][0]), "+d"(acc[mi][nbk][1]) : "d"(b1_a[mi][kb]), "d"(bv));
                    }
                }
#pragma unroll
                for (int mi = 0; mi < MBH; ++mi) {
                    const int g = 8 * (b1_mh * MBH + mi) + (lane >> 2);
#pragma unroll
                    for (int nbk = 0; nbk < NB; ++nbk) {
                        const int c = CLO + 8 * nbk + 2 * (lane & 3);
                        if (g < NGP && c < CHI)   // NCOL and CLO are even: the pair (c, c+1) is in range together
                            *reinterpret_cast<double2 *>(s_geo + (size_t)(g * EB + s) * GEO + c) = make_double2(acc[mi][nbk][0], acc[mi][nbk][1]);
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase B2: one thread per (Gauss point, element): J, G, GPML, source -> Q|P, T (scratch), R (in place) ----
        if (batch + (int)gridDim.x < nbatch) prepare_request(batch + gridDim.x);
        if (A.phase_mask & 1) {
            const int has_dmu = A.flags[0];
            const double w32 = f32r(A.omega);            // cmplx(0.d0,-omega), problem.f90:112
            for (int i = tid; i < nb * NGP; i += CFG::THREADS) {
                const int g = i / nb, s = i % nb;       // element fastest: runs of EB lanes in the scratch
                const double *nd = s_nodes + s * CFG::NSTR;
                double *geo = s_geo + (g * EB + s) * GEO;
                // scratch position of this (element, Gauss point): qt[batch32][component][g][lane]
                const int pos = first + s;
                double *qo = A.qt + ((size_t)(pos >> 5) * NCMP * NGP + g) * 32 + (pos & 31);
                constexpr size_t QS = (size_t)NGP * 32;   // component stride
                // nf_jacobian, n_fem.f90:359-366: J(m,n) = sum_l dN_l/dxi_m * r_l(n), l ascending, no FMA.  x and y are
                // tensor-product lines (2 or 3 distinct values per element): loaded once, selected per node at compile time
                constexpr int NORD = MN == 8 ? 2 : 3;
                double xs[3], ys[3];
                xs[0] = nd[2 * NDW + NREC]; xs[NORD - 1] = nd[NREC]; ys[0] = nd[NREC + 1]; ys[NORD - 1] = nd[NDW + NREC + 1];
                if (NORD == 3) { xs[1] = nd[9 * NDW + NREC]; ys[1] = nd[8 * NDW + NREC + 1]; }
                double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, xg[3] = {0, 0, 0};
#pragma unroll
                for (int l = 0; l < MN; ++l) {
                    const double x = xs[kNodeI27[l] * (NORD - 1) / 2], y = ys[kNodeJ27[l] * (NORD - 1) / 2], z = nd[l * NDW];
#pragma unroll
                    for (int mm = 0; mm < 3; ++mm) {
                        const double dn = s_dN[(l * 4 + mm) * NGPP + g];
                        J[mm][0] = J[mm][0] + dn * x; J[mm][1] = J[mm][1] + dn * y; J[mm][2] = J[mm][2] + dn * z;
                    }
                    if (PML) {   // g_rw, integration.f90:120-125 (reference order, no FMA: feeds the float32-rounded h)
                        const double ln = s_dN[(l * 4 + 3) * NGPP + g];
                        xg[0] = xg[0] + ln * x; xg[1] = xg[1] + ln * y; xg[2] = xg[2] + ln * z;
                    }
                }
                // nf_det, n_fem.f90:393-394 ; wgt, integration.f90:71
                const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                                   J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
                if (det == 0.0) atomicCAS(A.status, 0, -3);
                const double w = det * T.rw[g][3];
                const double rad = 1.0 / fabs(det);   // Q6: nf_ji = adj(J)/abs(det); one reciprocal (G only feeds fused products)
                double G[3][3];                // nf_ji: G[m][n] = d xi_n / d x_m
                G[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * rad;
                G[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * rad;
                G[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * rad;
                G[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * rad;
                G[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * rad;
                G[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * rad;
                G[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * rad;
                G[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * rad;
                G[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * rad;
                // interpolated columns of phase B1 (this thread's record is overwritten below)
                double mu[6], sr[6], si[6] = {0, 0, 0, 0, 0, 0};
                double dm1r[3], dm1i[3], dm2r[3], dm2i[3];
                if (DO_QT) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) { mu[k] = geo[k]; sr[k] = geo[6 + k]; if (PML) si[k] = geo[24 + k]; }
                }
                // signs: pol 1 dmpf = (+Im ds*e, -Re ds*e), pol 2 = (-Im ds*e, +Re ds*e)
#pragma unroll
                for (int k = 0; k < 3; ++k) { dm1r[k] = geo[12 + k]; dm1i[k] = -geo[15 + k]; dm2r[k] = -geo[18 + k]; dm2i[k] = geo[21 + k]; }
                double pc1[3] = {0, 0, 0}, pc2[3] = {0, 0, 0};
                if (has_dmu) {   // p_pcurl, problem.f90:362-374: grad N_l x (mu^-1 dmu Hp)_l  (mu != mu0 only: vc is read from L2)
                    int ie, je, ke;
                    elem_ijk(m, s_el[s * 4], ie, je, ke);
                    const int g1 = m.nord - 1;
                    const int64_t id0 = (int64_t)(ie - 1) * g1 * m.nyz + (int64_t)(je - 1) * g1 * m.nnz + (ke - 1) * g1;
                    for (int l = 0; l < MN; ++l) {
                        const double *vc = A.nodes[id0 + T.node_off[l]].vc;
                        const double t0 = s_dN[(l * 4 + 0) * NGPP + g], t1 = s_dN[(l * 4 + 1) * NGPP + g], t2 = s_dN[(l * 4 + 2) * NGPP + g];
                        double dn[3];
#pragma unroll
                        for (int mm = 0; mm < 3; ++mm) dn[mm] = G[mm][0] * t0 + G[mm][1] * t1 + G[mm][2] * t2;
                        pc1[0] += vc[2] * dn[1] - vc[1] * dn[2]; pc1[1] += vc[0] * dn[2] - vc[2] * dn[0]; pc1[2] += vc[1] * dn[0] - vc[0] * dn[1];
                        pc2[0] += vc[5] * dn[1] - vc[4] * dn[2]; pc2[1] += vc[3] * dn[2] - vc[5] * dn[0]; pc2[2] += vc[4] * dn[0] - vc[3] * dn[1];
                    }
                }
                double trq = 0.0, trt = 0.0;   // traces of Q|P and T at this Gauss point (magnitude of the element's K / M)
                // GPML stretch (boundary_conds.f90:84-186) with the LAGGING flags (Q17)
                double2 hhh = make_double2(1.0, 0.0);
                double2 h1 = hhh, h2 = hhh, h3 = hhh;
                if (PML) {
                    h1 = gpml_axis(A.pml, s_el[s * 4 + 1], 0, xg[0], A.omega);
                    h2 = gpml_axis(A.pml, s_el[s * 4 + 2], 1, xg[1], A.omega);
                    h3 = gpml_axis(A.pml, s_el[s * 4 + 3], 2, xg[2], A.omega);
                    hhh = cmul(cmul(h1, h2), h3);
                }
                if (PML && DO_QT) {
                    // Re G_ij, G_ij = h1h2h3/(h_i h_j): the factor of the half-curl pair with derivative axes i,j
                    double Gr[6];
                    Gr[0] = cdivf(cmul(h2, h3), h1).x; Gr[3] = cdivf(cmul(h1, h3), h2).x; Gr[5] = cdivf(cmul(h1, h2), h3).x;
                    Gr[1] = h3.x; Gr[2] = h2.x; Gr[4] = h1.x;
                    // half-curl a = (p,s): derivative axis x_a, grad-xi component y_a, sign s_a (v_fem.f90:57-59,
                    // integration.f90:171-188):  (1,1):(y,z,+) (1,2):(z,y,-) (2,1):(z,x,+) (2,2):(x,z,-) (3,1):(x,y,+) (3,2):(y,x,-)
                    constexpr int xa[6] = {1, 2, 2, 0, 0, 1}, ya[6] = {2, 1, 0, 2, 1, 0}, pa[6] = {0, 0, 1, 1, 2, 2};
                    constexpr double sa[6] = {1.0, -1.0, 1.0, -1.0, 1.0, -1.0};
                    // D6[a][b] = w mu^-1[p_a][p_b] Re G[x_a][x_b]   (symmetric 6x6; formed on the fly)
                    // P[(u,d)][(v,e)] = sum_ab H[a][u,d] D6[a][b] H[b][v,e],  H[a][u,d] = s_a G[x_a][u] G[y_a][d]
                    // evaluated per column (v,e) as DH = D6 H[:,(v,e)], then P[(u,d)][(v,e)] = (G^T W G)[u][d] with the
                    // antisymmetric-pattern W[x_a][y_a] = s_a DH[a]; everything statically indexed (registers)
                    double wmu[6], D6[21];
#pragma unroll
                    for (int k = 0; k < 6; ++k) wmu[k] = w * mu[k];
#pragma unroll
                    for (int a = 0; a < 6; ++a)
#pragma unroll
                        for (int b = a; b < 6; ++b) D6[a * 6 - a * (a - 1) / 2 + (b - a)] = wmu[sym3(pa[a], pa[b])] * Gr[sym3(xa[a], xa[b])];
#pragma unroll
                    for (int col = 0; col < 9; ++col) {
                        const int v = col / 3, e2 = col % 3;
                        double Hc[6], DH[6];
#pragma unroll
                        for (int b = 0; b < 6; ++b) Hc[b] = sa[b] * G[xa[b]][v] * G[ya[b]][e2];
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
                            double acc = 0.0;
#pragma unroll
                            for (int b = 0; b < 6; ++b) {
                                const int lo = a < b ? a : b, hi = a < b ? b : a;
                                acc = dfma(D6[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)], Hc[b], acc);
                            }
                            DH[a] = acc;
                        }
                        // E[x][d] = sum_y W[x][y] G[y][d];  W01=+DH4 W02=-DH3 W10=-DH5 W12=+DH0 W20=+DH2 W21=-DH1
                        double E[3][3];
#pragma unroll
                        for (int d2 = 0; d2 < 3; ++d2) {
                            E[0][d2] = dfma(DH[4], G[1][d2], -(DH[3] * G[2][d2]));
                            E[1][d2] = dfma(DH[0], G[2][d2], -(DH[5] * G[0][d2]));
                            E[2][d2] = dfma(DH[2], G[0][d2], -(DH[1] * G[1][d2]));
                        }
#pragma unroll
                        for (int row = 0; row <= col; ++row) {
                            const int u = row / 3, d2 = row % 3;
                            const double pv = dfma(G[0][u], E[0][d2], dfma(G[1][u], E[1][d2], G[2][u] * E[2][d2]));
                            qo[up9(row, col) * QS] = pv;
                            if (row == col) trq += fabs(pv);
                        }
                    }
                } else if (DO_QT) {
                    // Q = (w/det^2) J mu^-1 J^T  (curl N = (1/det J) J^T (dphi x e_d))
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
                // T = G^T S G with the mass tensor S = w Re[h1h2h3 sigma_g] (integration.f90:228-236, Q3)
                if (DO_QT) {
                    double S[6], SG[3][3];
#pragma unroll
                    for (int k = 0; k < 6; ++k) S[k] = PML ? w * (hhh.x * sr[k] - hhh.y * si[k]) : w * sr[k];
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
                            qo[((PML ? 45 : 6) + q6++) * QS] = tv;
                            if (a == b) trt += fabs(tv);
                        }
                    // element scales for the tiny-pair test of the contraction: max over the Gauss points, order independent
                    atomicMax(&s_scale[s * 2], (unsigned long long)__double_as_longlong(trq));
                    atomicMax(&s_scale[s * 2 + 1], (unsigned long long)__double_as_longlong(trt));
                }
                // R[d][pol] = G[:,d] . (w h1h2h3 src_pol);  src = (dmpf + pcrl) * cmplx32(0,-omega)  (problem.f90:112)
                {
                    const double2 whh = make_double2(w * hhh.x, w * hhh.y);
                    double2 a1[3], a2[3];
#pragma unroll
                    for (int mm = 0; mm < 3; ++mm) {
                        a1[mm] = cmul(whh, make_double2(dm1i[mm] * w32, -((dm1r[mm] + pc1[mm]) * w32)));
                        a2[mm] = cmul(whh, make_double2(dm2i[mm] * w32, -((dm2r[mm] + pc2[mm]) * w32)));
                    }
#pragma unroll
                    for (int d2 = 0; d2 < 3; ++d2) {
                        geo[GR + d2 * 4 + 0] = dfma(G[0][d2], a1[0].x, dfma(G[1][d2], a1[1].x, G[2][d2] * a1[2].x));
                        geo[GR + d2 * 4 + 1] = dfma(G[0][d2], a1[0].y, dfma(G[1][d2], a1[1].y, G[2][d2] * a1[2].y));
                        geo[GR + d2 * 4 + 2] = dfma(G[0][d2], a2[0].x, dfma(G[1][d2], a2[1].x, G[2][d2] * a2[2].x));
                        geo[GR + d2 * 4 + 3] = dfma(G[0][d2], a2[0].y, dfma(G[1][d2], a2[1].y, G[2][d2] * a2[2].y));
                    }
                }
            }
        }
        __syncthreads();
        if (DO_QT && A.escale && tid < nb)
            A.escale[first + tid] = make_double2(NGP * __longlong_as_double((long long)s_scale[tid * 2]), NGP * __longlong_as_double((long long)s_scale[tid * 2 + 1]));

        if (batch + (int)gridDim.x < nbatch) request_nodes(batch + gridDim.x);   // lands during the RHS phase / next wait

        // ---- RHS: blocal / f3, integration.f90:96-104,258-263: one thread per (element, slot) -- EB*MEP tasks keep every warp
        //      busy (one thread per group of four slots left 72 of 256 threads working on the 20-node element and 29 % of the
        //      kernel's warp samples at barriers; measured -4 % on the kernel, profiles/r02_ab_results.md) ----
        if (A.phase_mask & 2) {
            for (int i = tid; i < nb * MEP; i += CFG::THREADS) {
                const int cs = i / MEP, sl = i % MEP;
                const int cdof = s_slot[sl];
                if (cdof < 0) continue;
                const int cd = s_sdir[sl];
                double bacc[4] = {0.0, 0.0, 0.0, 0.0};
                const double *R0 = s_geo + cs * GEO + GR + cd * 4, *ph = s_phi + sl;
#pragma unroll
                for (int g = 0; g < NGP; ++g) {
                    const double phi = ph[g * MEP];
                    const double2 r01 = *reinterpret_cast<const double2 *>(R0 + g * EB * GEO), r23 = *reinterpret_cast<const double2 *>(R0 + g * EB * GEO + 2);
                    bacc[0] = dfma(phi, r01.x, bacc[0]); bacc[1] = dfma(phi, r01.y, bacc[1]);
                    bacc[2] = dfma(phi, r23.x, bacc[2]); bacc[3] = dfma(phi, r23.y, bacc[3]);
                }
                reinterpret_cast<double4 *>(A.be)[be_index(A.be_row0 + first + cs, ME, cdof)] = make_double4(bacc[0], bacc[1], bacc[2], bacc[3]);
            }
        }
    }
}

}  // namespace movfem

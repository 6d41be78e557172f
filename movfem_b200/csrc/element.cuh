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
// Gauss point, 3*me+mn+1 Jacobian rebuilds per point).  Here every Gauss point gets ONE Jacobian,
// the me basis vectors / curls are formed once (bit-identical to the reference's cve1-cve2, ve),
// and the element matrices are the two symmetric contractions
//        K_e = sum_g C_g^T (w mu^-1) C_g         M_e = sum_g V_g^T (w Re[h1h2h3 sigma]) V_g
// (6-row half-curl form with the 6x6 real GPML tensor inside the stretched layers), computed as
// register-tiled 4x4 FP64 FMA blocks over the lower triangle from shared-memory B-matrices.
// A_e = K_e + i*f32(omega)*M_e is formed later, per frequency (finalize.cuh), so K_e, M_e of the
// unstretched elements are frequency independent and cached in HBM across a sweep.
#pragma once
#include "common.cuh"

namespace movfem {

// ------------------------------------------------------------------------------------------
// node kernel: problem.f90:257-358 per grid node instead of per (element, node)
// ------------------------------------------------------------------------------------------
__global__ void node_kernel(int npt, double omega, const double *__restrict__ zp, const double *__restrict__ mu,
                            const double2 *__restrict__ sigma, NodeRec *__restrict__ out, int *__restrict__ status,
                            int *__restrict__ flags /* [0]: any dmu != 0, [1]: Re sigma changed */, int check_re) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
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
    out[i] = r;
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
    double *Ke, *Me;      // [ne][NP]
    double *be;           // [ne][ME][4]  (re,im) x 2 polarisations
    int *status;
    const int *flags;     // flags[0] any dmu, flags[1] Re sigma changed
    int skip_unless_changed;   // launch is a cache refresh: exit unless flags[1]
};

template <int MN_, int ME_, int NGP_, int GCH_, int EB_, int THREADS_, int MINB_, bool PML_>
struct ElemCfg {
    static constexpr int MN = MN_, ME = ME_, NGP = NGP_, GCH = GCH_, EB = EB_, THREADS = THREADS_, MINB = MINB_;
    static constexpr bool PML = PML_;
    static constexpr int MEP = (ME + 3) / 4 * 4;
    static constexpr int NT = MEP / 4, NTILES = NT * (NT + 1) / 2;
    static constexpr int NP = ME * (ME + 1) / 2;
    static constexpr int KR = PML ? 6 : 3;                 // rows of the curl operator
    static constexpr int NC = 2 * KR + 6;                  // C, DC, V, SV
    static constexpr int NDW = kNodeDoubles + 2;           // node record + x + y
    static constexpr int GEO = PML ? 48 : 34;              // Ji 9 + D (21|6) + S 6 + w*src 12 (+1 pad)
    static constexpr int NCHUNK = NGP / GCH;
    // shared memory: node records [EB][MN][NDW] | geometry [EB][NGP][GEO] | B chunk [EB][GCH][NC][MEP]
    static constexpr size_t SMEM = sizeof(double) * ((size_t)EB * MN * NDW + (size_t)EB * NGP * GEO + (size_t)EB * GCH * NC * MEP) +
                                   sizeof(int) * EB * 4;
    static_assert(NGP % GCH == 0, "chunking");
    static_assert(THREADS >= EB * NTILES && THREADS >= EB * ME, "one tile / one DOF per thread");
};

// B rows are stored with the two 16-byte halves of every 4-DOF group swapped in alternate groups of
// four, so that the 8 distinct groups a warp touches in one LDS.128 fall into distinct banks.
__device__ __forceinline__ int swz(int j) {
    const int grp = j >> 2, pos = j & 3;
    return (grp << 2) + ((((pos >> 1) ^ ((grp >> 2) & 1)) << 1) | (pos & 1));
}

// gpml_h for one axis, boundary_conds.f90:94-123
__device__ __forceinline__ double2 gpml_axis(const PmlParams &p, int flag, int axis, double r, double omega) {
    if (flag == 0) return make_double2(1.0, 0.0);
    const int s = flag < 0 ? 0 : 1;
    const double dw = p.omegar[1] - p.omegar[0], ww_pml = sqrt(dw * dw);
    const double d1 = omega - p.omegar[0], ww = sqrt(d1 * d1);
    double a0 = p.a0, b0 = p.b0;
    if (p.sch == 1) { a0 = 100.0 * (ww / ww_pml); b0 = (1.e6 - 1.e-2) * (ww / ww_pml) + 1.e-2; }
    const double dl = p.b[axis][s] - p.a[axis][s], rr_pml = sqrt(dl * dl);
    const double dr = r - p.a[axis][s], rr = sqrt(dr * dr);
    const double rho = rr / rr_pml;
    const double pw = (p.nn == 2.0) ? rho * rho : (p.nn == 1.0 ? rho : pow(rho, p.nn));
    if (p.sch == 0) {
        const double hx0 = 1.0 + a0 * pw;
        const double sn = sin((kPi / 2.0) * rho);
        const double bx = b0 * (sn * sn);
        return make_double2(hx0 * 1.0, hx0 * f32r(-bx / (omega * kEps0)));
    }
    // Re[(b0*rho^n, 0) / cmplx32(a0, omega)], Smith's division as gfortran emits it
    const double x = b0 * pw, br = f32r(a0), bi = f32r(omega);
    double re;
    if (fabs(br) < fabs(bi)) { const double ratio = br / bi, div = (br * ratio) + bi; re = ((x * ratio) + 0.0) / div; }
    else { const double ratio = bi / br, div = (bi * ratio) + br; re = ((0.0 * ratio) + x) / div; }
    return make_double2(1.0, f32r(re));
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

template <class CFG, bool DO_KM>
__global__ void __launch_bounds__(CFG::THREADS, CFG::MINB) element_kernel(ElemArgs A) {
    constexpr int MN = CFG::MN, ME = CFG::ME, MEP = CFG::MEP, GCH = CFG::GCH, EB = CFG::EB, NC = CFG::NC, NGP = CFG::NGP;
    constexpr int KR = CFG::KR, GEO = CFG::GEO, NDW = CFG::NDW, NTILES = CFG::NTILES, NP = CFG::NP;
    constexpr bool PML = CFG::PML;
    if (A.skip_unless_changed && A.flags[1] == 0) return;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_nodes = reinterpret_cast<double *>(smem_raw);           // [EB][MN][NDW]
    double *s_geo = s_nodes + EB * MN * NDW;                          // [EB][NGP][GEO]
    double *s_B = s_geo + EB * NGP * GEO;                             // [EB][GCH][NC][MEP] (swizzled rows)
    int *s_el = reinterpret_cast<int *>(s_B + EB * GCH * NC * MEP);   // [EB][4]: element id, GPML flags

    const ElemTables &T = *A.T;
    const MeshDims &m = A.m;
    const int tid = threadIdx.x;
    const int first = blockIdx.x * EB;
    const int nb = min(EB, A.nlist - first);

    if (tid < EB) {
        const int e = tid < nb ? A.list[first + tid] : -1;
        s_el[tid * 4] = e;
        int f[3] = {0, 0, 0};
        if (PML && e >= 0) effective_pml(m, A.pml, e, f);
        s_el[tid * 4 + 1] = f[0]; s_el[tid * 4 + 2] = f[1]; s_el[tid * 4 + 3] = f[2];
    }
    if constexpr (MEP > ME) {   // zero the padded DOF columns once
        constexpr int PAD = MEP - ME;
        for (int i = tid; i < EB * GCH * NC * PAD; i += CFG::THREADS) {
            const int row = i / PAD, c = ME + i % PAD;
            s_B[row * MEP + swz(c)] = 0.0;
        }
    }
    __syncthreads();

    // ---- phase A: gather the element's node records (16-byte pieces, coalesced per record) ----
    for (int i = tid; i < nb * MN * (kNodeDoubles / 2 + 1); i += CFG::THREADS) {
        const int part = i % (kNodeDoubles / 2 + 1), sl = i / (kNodeDoubles / 2 + 1);
        const int l = sl % MN, s = sl / MN;
        const int e = s_el[s * 4];
        int ie, je, ke;
        elem_ijk(m, e, ie, je, ke);
        const int g1 = m.nord - 1;
        double2 *dst = reinterpret_cast<double2 *>(s_nodes + (s * MN + l) * NDW);
        if (part < kNodeDoubles / 2) {
            const int64_t id = (int64_t)(ie - 1) * g1 * m.nyz + (int64_t)(je - 1) * g1 * m.nnz + (ke - 1) * g1 + T.node_off[l];
            dst[part] = reinterpret_cast<const double2 *>(A.nodes + id)[part];
        } else {
            dst[part] = make_double2(A.xp[(ie - 1) * g1 + T.node_i[l]], A.yp[(je - 1) * g1 + T.node_j[l]]);
        }
    }
    __syncthreads();

    // ---- phase B: one thread per (element, Gauss point): Jacobian, materials, GPML, source ----
    {
        const int has_dmu = A.flags[0];
        const double w32 = f32r(A.omega);            // cmplx(0.d0,-omega), problem.f90:112
        const double psig = f32r(A.omega * kEps0);   // pset_pmodel, problem.f90:250
        for (int i = tid; i < nb * NGP; i += CFG::THREADS) {
            const int s = i / NGP, g = i % NGP;
            const double *nd = s_nodes + s * MN * NDW;
            double *geo = s_geo + (s * NGP + g) * GEO;
            // nf_jacobian, n_fem.f90:359-366: J(m,n) = sum_l dN_l/dxi_m * r_l(n), l ascending, no FMA
            double J[3][3];
#pragma unroll
            for (int mm = 0; mm < 3; ++mm) {
                double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll 4
                for (int l = 0; l < MN; ++l) {
                    const double dn = T.dN[g][l][mm];
                    sx = sx + dn * nd[l * NDW + kNodeDoubles];
                    sy = sy + dn * nd[l * NDW + kNodeDoubles + 1];
                    sz = sz + dn * nd[l * NDW];
                }
                J[mm][0] = sx; J[mm][1] = sy; J[mm][2] = sz;
            }
            // nf_det, n_fem.f90:393-394 ; wgt, integration.f90:71
            const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
            if (det == 0.0) atomicCAS(A.status, 0, -3);
            const double w = det * T.rw[g][3];
            const double ad = fabs(det);   // Q6
            double Ji[9];
            Ji[0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / ad;
            Ji[1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / ad;
            Ji[2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / ad;
            Ji[3] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / ad;
            Ji[4] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / ad;
            Ji[5] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / ad;
            Ji[6] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / ad;
            Ji[7] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / ad;
            Ji[8] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / ad;
#pragma unroll
            for (int k = 0; k < 9; ++k) geo[k] = Ji[k];
            // p_intmodels, problem.f90:139-142: material tensors at the Gauss point
            double mu[6] = {0, 0, 0, 0, 0, 0}, sr[6] = {0, 0, 0, 0, 0, 0}, si[6] = {0, 0, 0, 0, 0, 0};
            double xg[3] = {0, 0, 0};
            double dm1r[3] = {0, 0, 0}, dm1i[3] = {0, 0, 0}, dm2r[3] = {0, 0, 0}, dm2i[3] = {0, 0, 0};
#pragma unroll 2
            for (int l = 0; l < MN; ++l) {
                const double ln = T.N[g][l];
                const double *r = nd + l * NDW;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    mu[k] = dfma(ln, r[2 + k], mu[k]);
                    sr[k] = dfma(ln, r[8 + k], sr[k]);
                    if (PML) si[k] = dfma(ln, r[14 + k], si[k]);
                }
                if (PML) {   // g_rw, integration.f90:120-125 (reference order, no FMA: feeds the float32-rounded h)
                    xg[0] = xg[0] + ln * r[kNodeDoubles]; xg[1] = xg[1] + ln * r[kNodeDoubles + 1]; xg[2] = xg[2] + ln * r[0];
                }
                // p_dmpf, problem.f90:424-457: ln * (dsigma . Ep); Ep_1 = (0,-e) x^, Ep_2 = (0,+e) y^
                const double le = ln * r[1];
                const double d0 = r[14] - psig, d3 = r[17] - psig;   // Im(dsigma) on the diagonal
                dm1r[0] = dfma(le, d0, dm1r[0]);     dm1i[0] = dfma(le, r[8], dm1i[0]);
                dm1r[1] = dfma(le, r[15], dm1r[1]);  dm1i[1] = dfma(le, r[9], dm1i[1]);
                dm1r[2] = dfma(le, r[16], dm1r[2]);  dm1i[2] = dfma(le, r[10], dm1i[2]);
                dm2r[0] = dfma(le, r[15], dm2r[0]);  dm2i[0] = dfma(le, r[9], dm2i[0]);
                dm2r[1] = dfma(le, d3, dm2r[1]);     dm2i[1] = dfma(le, r[11], dm2i[1]);
                dm2r[2] = dfma(le, r[18], dm2r[2]);  dm2i[2] = dfma(le, r[12], dm2i[2]);
            }
            // signs: pol 1 dmpf = (+Im ds*e, -Re ds*e), pol 2 = (-Im ds*e, +Re ds*e)
#pragma unroll
            for (int k = 0; k < 3; ++k) { dm1i[k] = -dm1i[k]; dm2r[k] = -dm2r[k]; }
            double pc1[3] = {0, 0, 0}, pc2[3] = {0, 0, 0};
            if (has_dmu) {   // p_pcurl, problem.f90:362-374: grad N_l x (mu^-1 dmu Hp)_l
                for (int l = 0; l < MN; ++l) {
                    const double *r = nd + l * NDW;
                    double dn[3];
#pragma unroll
                    for (int mm = 0; mm < 3; ++mm)
                        dn[mm] = Ji[mm * 3] * T.dN[g][l][0] + Ji[mm * 3 + 1] * T.dN[g][l][1] + Ji[mm * 3 + 2] * T.dN[g][l][2];
                    pc1[0] += r[22] * dn[1] - r[21] * dn[2]; pc1[1] += r[20] * dn[2] - r[22] * dn[0]; pc1[2] += r[21] * dn[0] - r[20] * dn[1];
                    pc2[0] += r[25] * dn[1] - r[24] * dn[2]; pc2[1] += r[23] * dn[2] - r[25] * dn[0]; pc2[2] += r[24] * dn[0] - r[23] * dn[1];
                }
            }
            // GPML stretch (boundary_conds.f90:84-186) with the LAGGING flags (Q17)
            double2 hhh = make_double2(1.0, 0.0);
            double G[6] = {1, 0, 0, 1, 0, 1};   // Re G11,G12,G13,G22,G23,G33 ; G_ij = h1h2h3/(h_i h_j)
            if (PML) {
                const double2 h1 = gpml_axis(A.pml, s_el[s * 4 + 1], 0, xg[0], A.omega);
                const double2 h2 = gpml_axis(A.pml, s_el[s * 4 + 2], 1, xg[1], A.omega);
                const double2 h3 = gpml_axis(A.pml, s_el[s * 4 + 3], 2, xg[2], A.omega);
                hhh = cmul(cmul(h1, h2), h3);
                G[0] = cdivf(cmul(h2, h3), h1).x; G[3] = cdivf(cmul(h1, h3), h2).x; G[5] = cdivf(cmul(h1, h2), h3).x;
                G[1] = h3.x; G[2] = h2.x; G[4] = h1.x;
            }
            double *gp = geo + 9;
            if (!PML) {
#pragma unroll
                for (int k = 0; k < 6; ++k) gp[k] = w * mu[k];
                gp += 6;
            } else {
                // D6[a][b] = w * mu^-1[p_a][p_b] * Re G[d_a][d_b]; a = (p,s): (1,1),(1,2),(2,1),(2,2),(3,1),(3,2)
                // derivative axis d(a) = 2,3,3,1,1,2 (integration.f90:171-188)
                const int pa[6] = {0, 0, 1, 1, 2, 2}, da[6] = {1, 2, 2, 0, 0, 1};
                const int sym6[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
                int q = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = a; b < 6; ++b) gp[q++] = w * mu[sym6[pa[a]][pa[b]]] * G[sym6[da[a]][da[b]]];
                gp += 21;
            }
            // mass tensor: w * Re[h1h2h3 * sigma_g]  (integration.f90:228-236, Q3)
#pragma unroll
            for (int k = 0; k < 6; ++k) gp[k] = PML ? w * (hhh.x * sr[k] - hhh.y * si[k]) : w * sr[k];
            gp += 6;
            // p_source, problem.f90:112: (dmpf + pcrl) * cmplx32(0,-omega); then w*[h1h2h3] (integration.f90:258-263)
            const double2 whh = make_double2(w * hhh.x, w * hhh.y);
#pragma unroll
            for (int mm = 0; mm < 3; ++mm) {
                const double2 s1 = make_double2(dm1i[mm] * w32, -((dm1r[mm] + pc1[mm]) * w32));
                const double2 s2 = make_double2(dm2i[mm] * w32, -((dm2r[mm] + pc2[mm]) * w32));
                const double2 a1 = cmul(whh, s1), a2 = cmul(whh, s2);
                gp[mm * 2] = a1.x; gp[mm * 2 + 1] = a1.y; gp[6 + mm * 2] = a2.x; gp[6 + mm * 2 + 1] = a2.y;
            }
        }
    }
    __syncthreads();

    double accK[16], accM[16], bacc[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { accK[i] = 0.0; accM[i] = 0.0; }
    bacc[0] = bacc[1] = bacc[2] = bacc[3] = 0.0;

    // tile owned by this thread in the contraction (lower triangle of 4x4 blocks)
    const int ts = tid / NTILES, tt = tid % NTILES;
    int ti = 0, tj = 0;
    {
        int rem = tt;
        while (rem > ti) { rem -= ti + 1; ++ti; }
        tj = rem;
    }
    // swizzled 16-byte half offsets of this thread's row / column groups (in doubles)
    const int a_lo = 4 * ti + (((ti >> 2) & 1) << 1), a_hi = 4 * ti + ((((ti >> 2) & 1) ^ 1) << 1);
    const int b_lo = 4 * tj + (((tj >> 2) & 1) << 1), b_hi = 4 * tj + ((((tj >> 2) & 1) ^ 1) << 1);
    // DOF owned by this thread in the basis phase
    const int cs = tid / ME, cdof = tid % ME;
    const int cd = T.edir[cdof < ME ? cdof : 0], cpos = swz(cdof);

    for (int chunk = 0; chunk < CFG::NCHUNK; ++chunk) {
        // ---- phase C: one thread per (element, DOF): basis, curl, D*B products, RHS ----
        if (tid < nb * ME) {
#pragma unroll 1
            for (int gc = 0; gc < GCH; ++gc) {
                const int g = chunk * GCH + gc;
                const double *geo = s_geo + (cs * NGP + g) * GEO;
                double *B = s_B + (size_t)(cs * GCH + gc) * NC * MEP + cpos;
                const double phi = T.phi[g][cdof];
                const double dp0 = T.dphi[g][cdof][0], dp1 = T.dphi[g][cdof][1], dp2 = T.dphi[g][cdof][2];
                double vij[3], V[3], dni[3];
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) {
                    vij[mm] = geo[mm * 3 + cd];                     // grad_xi, v_fem.f90:518
                    V[mm] = phi * vij[mm];                          // vf_elem_ve, v_fem.f90:43
                    dni[mm] = geo[mm * 3] * dp0 + geo[mm * 3 + 1] * dp1 + geo[mm * 3 + 2] * dp2;   // mix_grad_ln
                }
                if (DO_KM) {
                    // vf_elem_curl, v_fem.f90:57-59
                    const double c11 = dni[1] * vij[2], c12 = dni[2] * vij[1];
                    const double c21 = dni[2] * vij[0], c22 = dni[0] * vij[2];
                    const double c31 = dni[0] * vij[1], c32 = dni[1] * vij[0];
                    if (!PML) {
                        const double *D = geo + 9;
                        const double C0 = c11 - c12, C1 = c21 - c22, C2 = c31 - c32;
                        B[0 * MEP] = C0; B[1 * MEP] = C1; B[2 * MEP] = C2;
                        B[3 * MEP] = dfma(D[0], C0, dfma(D[1], C1, D[2] * C2));
                        B[4 * MEP] = dfma(D[1], C0, dfma(D[3], C1, D[4] * C2));
                        B[5 * MEP] = dfma(D[2], C0, dfma(D[4], C1, D[5] * C2));
                    } else {
                        const double *D = geo + 9;   // 21 packed upper entries of the symmetric 6x6
                        const double C6[6] = {c11, -c12, c21, -c22, c31, -c32};
                        double DC[6] = {0, 0, 0, 0, 0, 0};
                        int q = 0;
#pragma unroll
                        for (int a = 0; a < 6; ++a)
#pragma unroll
                            for (int b = a; b < 6; ++b) {
                                const double v = D[q++];
                                DC[a] = dfma(v, C6[b], DC[a]);
                                if (b != a) DC[b] = dfma(v, C6[a], DC[b]);
                            }
#pragma unroll
                        for (int a = 0; a < 6; ++a) { B[a * MEP] = C6[a]; B[(6 + a) * MEP] = DC[a]; }
                    }
                    const double *S = geo + 9 + (PML ? 21 : 6);
                    double *BV = B + 2 * KR * MEP;
                    BV[0 * MEP] = V[0]; BV[1 * MEP] = V[1]; BV[2 * MEP] = V[2];
                    BV[3 * MEP] = dfma(S[0], V[0], dfma(S[1], V[1], S[2] * V[2]));
                    BV[4 * MEP] = dfma(S[1], V[0], dfma(S[3], V[1], S[4] * V[2]));
                    BV[5 * MEP] = dfma(S[2], V[0], dfma(S[4], V[1], S[5] * V[2]));
                }
                // blocal / f3, integration.f90:96-104,258-263: sum_g w [h1h2h3] N . src_d
                const double *ws = geo + 9 + (PML ? 21 : 6) + 6;
#pragma unroll
                for (int mm = 0; mm < 3; ++mm) {
                    bacc[0] = dfma(V[mm], ws[mm * 2], bacc[0]);     bacc[1] = dfma(V[mm], ws[mm * 2 + 1], bacc[1]);
                    bacc[2] = dfma(V[mm], ws[6 + mm * 2], bacc[2]); bacc[3] = dfma(V[mm], ws[6 + mm * 2 + 1], bacc[3]);
                }
            }
        }
        if (!DO_KM) continue;
        __syncthreads();

        // ---- phase D: register-tiled lower-triangle contraction over this chunk ----
        if (tid < nb * NTILES) {
            const double *Bs = s_B + (size_t)ts * GCH * NC * MEP;
#pragma unroll 1
            for (int gc = 0; gc < GCH; ++gc) {
                const double *Bg = Bs + gc * NC * MEP;
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    const double2 a0 = *reinterpret_cast<const double2 *>(Bg + (KR + k) * MEP + a_lo);   // D*C rows
                    const double2 a1 = *reinterpret_cast<const double2 *>(Bg + (KR + k) * MEP + a_hi);
                    const double2 b0 = *reinterpret_cast<const double2 *>(Bg + k * MEP + b_lo);          // C cols
                    const double2 b1 = *reinterpret_cast<const double2 *>(Bg + k * MEP + b_hi);
                    const double av[4] = {a0.x, a0.y, a1.x, a1.y}, bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) accK[i * 4 + j] = dfma(av[i], bv[j], accK[i * 4 + j]);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double2 a0 = *reinterpret_cast<const double2 *>(Bg + (2 * KR + 3 + k) * MEP + a_lo);   // S*V rows
                    const double2 a1 = *reinterpret_cast<const double2 *>(Bg + (2 * KR + 3 + k) * MEP + a_hi);
                    const double2 b0 = *reinterpret_cast<const double2 *>(Bg + (2 * KR + k) * MEP + b_lo);       // V cols
                    const double2 b1 = *reinterpret_cast<const double2 *>(Bg + (2 * KR + k) * MEP + b_hi);
                    const double av[4] = {a0.x, a0.y, a1.x, a1.y}, bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) accM[i * 4 + j] = dfma(av[i], bv[j], accM[i * 4 + j]);
                }
            }
        }
        if (chunk + 1 < CFG::NCHUNK) __syncthreads();
    }

    // ---- write-out: element-major, packed lower triangle by local index ----
    if (DO_KM && tid < nb * NTILES) {
        const int64_t e = s_el[ts * 4];
        double *Ko = A.Ke + e * NP, *Mo = A.Me + e * NP;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int im = 4 * ti + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int jm = 4 * tj + j;
                if (im < ME && jm <= im) {
                    const int p = im * (im + 1) / 2 + jm;
                    Ko[p] = accK[i * 4 + j]; Mo[p] = accM[i * 4 + j];
                }
            }
        }
    }
    if (tid < nb * ME) {
        const int64_t e = s_el[cs * 4];
        reinterpret_cast<double4 *>(A.be)[e * ME + cdof] = make_double4(bacc[0], bacc[1], bacc[2], bacc[3]);
    }
}

}  // namespace movfem

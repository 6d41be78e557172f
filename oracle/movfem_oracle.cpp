/*
 * oracle/movfem_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the MoVFEM_3DMT element assembly path, used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * CHECKER and the timed CPU baseline.  It is never linked into, imported by or called from the
 * product (movfem_b200/): the product fails loudly when its CUDA library is missing.
 *
 * PARITY PIN STATUS: **pinned against outputs of the reference itself.**  The reference ships no tests or
 * golden vectors and no Fortran compiler exists in this image (no oracle/_ref), but its own Fortran sources are
 * EXECUTED by tests/golden/f90exec.py (a Fortran-subset executor with Fortran kind/assignment/array semantics;
 * nothing is copied, the sources are read where they lie) through tests/golden/ref_exec.py: init_n_fem,
 * init_v_fem, init_problem, init_integration, ga_init, bd_setmodel, global_vfem / local_vfem (MoVFEM_3DMT.f90:
 * 167-263), find_zeros / rem_zeros, and solution.f90 node_solution.  tests/golden/ref_*.npz hold those outputs for
 * 8/20/27-node elements, GPML Fang / Zhou, Dirichlet boundary models 1-3, two frequencies of the sequential loop;
 * tests/test_reference_vectors.py checks this oracle against them: gne / nne / nnze / IRN / JCN identical, the
 * delivered values (tap T2) and every per-element cache, A_e and b_e BIT FOR BIT equal, RHS bit for bit (<= 2e-16
 * for boundary models 2/3, whose complex sqrt/exp come from different math libraries).  Additional pins:
 * (i) the element-matrix known answers of SURVEY App. B item 4, (ii) the exact nne / nnze counts of SURVEY
 * section 6, (iii) mathematical invariants (App. B items 3 and 5).  See tests/test_oracle_pins.py.
 *
 * Build: g++ -O1 -ffp-contract=off  (mirrors `gfortran -O` on x86-64: no FMA contraction, no
 * reassociation).  Every floating-point expression keeps the Fortran evaluation order.
 *
 * Quirks reproduced on purpose: Q1 float32 Gauss literals, Q2 single
 * precision cmplx(), Q3 real-valued f1/f2, Q4 +i*omega, Q5 8-node dN/dzeta typo, Q6 abs(det),
 * Q7 one-sided GPML, Q8 omega-dependent gpml_h (its omega-dependent part is imaginary and dropped by Q18), Q9 full structural pattern, Q10 float32 scratch +
 * unpermuted second pass in ga_sort_sparse, Q11 zero stripping, Q17 lagging GPML flags,
 * Q18 the GPML stretch is stored in a REAL array (integration.f90:16), i.e. only Re(h) is used.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <time.h>
#include <numeric>
#include <vector>
#include <complex>

#include "../include/movfem_b200.h"
#include "shape.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

// ---- complex arithmetic with gfortran's default rules (-fcx-fortran-rules) -----------------
struct C {
    double re, im;
};
static inline C mk(double r, double i) { return C{r, i}; }
static inline C operator+(C a, C b) { return mk(a.re + b.re, a.im + b.im); }
static inline C operator-(C a, C b) { return mk(a.re - b.re, a.im - b.im); }
static inline C operator*(C a, C b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
// real*complex and complex*real: the real operand is promoted to (r,0); for finite operands the
// full product equals the component-wise one (differences only in the sign of zero).
static inline C operator*(double r, C b) { return mk(r * b.re, r * b.im); }
static inline C operator*(C a, double r) { return mk(a.re * r, a.im * r); }
static inline C operator+(double r, C b) { return mk(r + b.re, b.im); }
static inline C operator-(C a) { return mk(-a.re, -a.im); }
// complex division, Smith's range-reduced algorithm as emitted by GCC for Fortran
static inline C operator/(C a, C b) {
    if (std::fabs(b.re) < std::fabs(b.im)) {
        const double ratio = b.re / b.im, div = (b.re * ratio) + b.im;
        return mk(((a.re * ratio) + a.im) / div, ((a.im * ratio) - a.re) / div);
    }
    const double ratio = b.im / b.re, div = (b.im * ratio) + b.re;
    return mk(((a.im * ratio) + a.re) / div, (a.im - (a.re * ratio)) / div);
}
static inline C operator/(double r, C b) { return mk(r, 0.0) / b; }
// cmplx(x,y) WITHOUT a kind argument returns default (single precision) complex (Q2)
static inline C cmplx32(double x, double y) { return mk((double)(float)x, (double)(float)y); }

static const double PI = 3.1415926535897932384626433;  // geometry.f90:25, boundary_conds.f90:23
static const double EPS0 = 8.854187817e-12;             // geometry.f90:25
static const double B0 = 1.e-9;                         // problem.f90:27

struct Ctx {
    movfem_desc d;
    int nx, ny, nz, ne;          // elements per axis (global_assembly.f90:33-34)
    int nnx, nny, nnz, nyz, npt; // geometry.f90:517-521
    int mn, me, ngp;
    Shape shape;
    int i1[27], j1[27], k1[27];
    int enode[54], edir[54];
    double rw[27][4];            // integration.f90:267-279  i_rw
    std::vector<int> gne;        // gne(ne,me) column-major: gne[(im-1)*ne + (ide-1)]
    int nne;
    // full structural pattern in (row, col) ascending order (see build_pattern)
    std::vector<int> pat_ia, pat_ja;
    std::vector<int64_t> row_ptr;   // size nne+1, into pat_*
    int64_t nnze;
    // boundary_conds.f90 GPML state
    int el_xa[2], el_xb[2], el_ya[2], el_yb[2], el_za[2], el_zb[2];
    double xa[2], xb[2], ya[2], yb[2], za[2], zb[2], omegar[2];
    int in_pml[3];               // SAVEd module variable: persists across elements AND frequencies (Q17)
    char err[256];
    explicit Ctx(int mn_) : shape(mn_) {}
};

// per-element working state = the module variables of n_fem / v_fem / problem / integration
struct Elem {
    const Ctx *c;
    double omega;
    const C *g_sigma;
    int faithful;
    long jac_builds;
    double nf_re[27][3];
    int nf_index[27];            // 1-based grid node ids
    double nf_j[3][3], nf_ji[3][3];
    double jac_key[3]; bool jac_valid;   // memo for faithful==0 (same inputs -> same bits)
    // problem.f90 per-element fields
    double pe_inmu[27][6], pe_dmu[27][6];
    C pe_dsigma[27][6], pe_ep[54][3], pe_hp[54][3];
    C pe_psigma;
    // integration.f90 caches
    double wgt[27];
    double cve1[27 * 54][3], cve2[27 * 54][3], ve[27 * 54][3];
    C mf1[27][6], mf2[27][6], src[54][3], gpml[27][3];
    int in_pml[3];
    int status;
};

// ---- n_fem.f90:66-102 nf_get_r ------------------------------------------------------------
static void nf_get_r(Elem &E, int i, int j, int k, int no) {
    const Ctx &c = *E.c;
    const int g = c.d.nord;
    for (int ni = 0; ni < c.mn; ++ni) {
        const int ii = (i - 1) * (g - 1) + c.i1[ni];
        const int jj = (j - 1) * (g - 1) + c.j1[ni];
        const int id = no + (c.i1[ni] - 1) * c.nyz + (c.j1[ni] - 1) * c.nnz + (c.k1[ni] - 1);
        E.nf_re[ni][0] = c.d.g_xp[ii - 1];
        E.nf_re[ni][1] = c.d.g_yp[jj - 1];
        E.nf_re[ni][2] = c.d.g_zp[id - 1];
        E.nf_index[ni] = id;
    }
    E.jac_valid = false;
}

// ---- n_fem.f90:391-395 --------------------------------------------------------------------
static inline double nf_det(const double a[3][3]) {
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) + a[0][1] * (a[1][2] * a[2][0] - a[1][0] * a[2][2]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}

// ---- n_fem.f90:355-367 nf_jacobian --------------------------------------------------------
static void nf_jacobian(Elem &E, double xi, double eta, double zeta) {
    if (!E.faithful && E.jac_valid && E.jac_key[0] == xi && E.jac_key[1] == eta && E.jac_key[2] == zeta) return;
    const Ctx &c = *E.c;
    ++E.jac_builds;
    for (int m = 0; m < 3; ++m)
        for (int n = 0; n < 3; ++n) {
            double s = 0.0;
            for (int l = 0; l < c.mn; ++l) s = s + c.shape.nf_dln_dxi(m + 1, l + 1, xi, eta, zeta) * E.nf_re[l][n];
            E.nf_j[m][n] = s;
        }
    E.jac_key[0] = xi; E.jac_key[1] = eta; E.jac_key[2] = zeta; E.jac_valid = true;
}

// ---- n_fem.f90:370-388 nf_inv_jac (Q6: divides by dabs(det)) -------------------------------
static void nf_inv_jac(Elem &E) {
    const double(*J)[3] = E.nf_j;
    const double det_j = nf_det(J);
    if (det_j == 0) { E.status = MOVFEM_E_SINGULAR_JAC; return; }
    const double ad = std::fabs(det_j);
    E.nf_ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / ad;
    E.nf_ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / ad;
    E.nf_ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / ad;
    E.nf_ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / ad;
    E.nf_ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / ad;
    E.nf_ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / ad;
    E.nf_ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / ad;
    E.nf_ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / ad;
    E.nf_ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / ad;
}

// ---- n_fem.f90:320-334 nf_grad_ln ---------------------------------------------------------
static void nf_grad_ln(Elem &E, int i, const double rn[3], double d_ne[3]) {
    nf_jacobian(E, rn[0], rn[1], rn[2]);
    nf_inv_jac(E);
    for (int m = 0; m < 3; ++m) {
        double s = 0.0;
        for (int n = 0; n < 3; ++n) s = s + E.nf_ji[m][n] * E.c->shape.nf_dln_dxi(n + 1, i, rn[0], rn[1], rn[2]);
        d_ne[m] = s;
    }
}

// ---- v_fem.f90:470-484 mix_grad_ln --------------------------------------------------------
static void mix_grad_ln(Elem &E, int i, int dir, const double rn[3], double d_ne[3]) {
    nf_jacobian(E, rn[0], rn[1], rn[2]);
    nf_inv_jac(E);
    for (int m = 0; m < 3; ++m) {
        double s = 0.0;
        for (int n = 0; n < 3; ++n) s = s + E.nf_ji[m][n] * E.c->shape.mix_dln_dxi(dir, n + 1, i, rn[0], rn[1], rn[2]);
        d_ne[m] = s;
    }
}

// ---- v_fem.f90:510-519 grad_xi ------------------------------------------------------------
static void grad_xi(Elem &E, int edge, const double r[3], double out[3]) {
    nf_jacobian(E, r[0], r[1], r[2]);
    nf_inv_jac(E);
    const int dcol = E.c->edir[edge - 1] - 1;
    for (int m = 0; m < 3; ++m) out[m] = E.nf_ji[m][dcol];
}

// ---- v_fem.f90:38-44 vf_elem_ve -----------------------------------------------------------
static void vf_elem_ve(Elem &E, int edge, const double r[3], double vf_ve[3]) {
    const Ctx &c = *E.c;
    const double phi = c.shape.mix_ln(c.enode[edge - 1], c.edir[edge - 1], r[0], r[1], r[2]);
    double g[3];
    grad_xi(E, edge, r, g);
    for (int m = 0; m < 3; ++m) vf_ve[m] = phi * g[m];
}

// ---- v_fem.f90:49-60 vf_elem_curl (two halves of each curl component) ----------------------
static void vf_elem_curl(Elem &E, int edge, const double r[3], double cve[3][2]) {
    const Ctx &c = *E.c;
    double vij[3], dni[3];
    grad_xi(E, edge, r, vij);
    mix_grad_ln(E, c.enode[edge - 1], c.edir[edge - 1], r, dni);
    cve[0][0] = dni[1] * vij[2]; cve[0][1] = dni[2] * vij[1];
    cve[1][0] = dni[2] * vij[0]; cve[1][1] = dni[0] * vij[2];
    cve[2][0] = dni[0] * vij[1]; cve[2][1] = dni[1] * vij[0];
}

// ---- problem.f90:301-313 det / cdet -------------------------------------------------------
static inline double det6(const double *a) {  // a[0..5] = 11,12,13,22,23,33
    return a[0] * (a[3] * a[5] - a[4] * a[4]) + a[1] * (a[2] * a[4] - a[1] * a[5]) + a[2] * (a[1] * a[4] - a[3] * a[2]);
}
static inline C cdet6(const C *a) {
    return a[0] * (a[3] * a[5] - a[4] * a[4]) + a[1] * (a[2] * a[4] - a[1] * a[5]) + a[2] * (a[1] * a[4] - a[3] * a[2]);
}

// ---- problem.f90:70-87 p_elem_fields (pinv_geomodel 257-277, pdelta_model 318-335,
//      p_pfields 340-358) -------------------------------------------------------------------
static void p_elem_fields(Elem &E) {
    const Ctx &c = *E.c;
    const int mn = c.mn;
    for (int j = 0; j < 2 * mn; ++j)
        for (int m = 0; m < 3; ++m) { E.pe_ep[j][m] = mk(0, 0); E.pe_hp[j][m] = mk(0, 0); }
    for (int i = 0; i < mn; ++i) {
        const int id = E.nf_index[i] - 1;
        const C *sg = E.g_sigma + (size_t)6 * id;
        const double *mu = c.d.g_mu + (size_t)6 * id;
        // pinv_geomodel: stops on singular tensors; sigma^-1 is unused for pe_sch=1
        const C cd = cdet6(sg);
        if (cd.re == 0.0 && cd.im == 0.0) { E.status = MOVFEM_E_SINGULAR_MODEL; return; }
        const double dm = det6(mu);
        if (dm == 0.) { E.status = MOVFEM_E_SINGULAR_MODEL; return; }
        // inv_tensor, problem.f90:279-288 (each entry divides by a fresh det(a))
        E.pe_inmu[i][0] = (mu[3] * mu[5] - mu[4] * mu[4]) / det6(mu);
        E.pe_inmu[i][1] = (mu[2] * mu[4] - mu[1] * mu[5]) / det6(mu);
        E.pe_inmu[i][2] = (mu[1] * mu[4] - mu[2] * mu[3]) / det6(mu);
        E.pe_inmu[i][3] = (mu[0] * mu[5] - mu[2] * mu[2]) / det6(mu);
        E.pe_inmu[i][4] = (mu[2] * mu[1] - mu[0] * mu[4]) / det6(mu);
        E.pe_inmu[i][5] = (mu[3] * mu[0] - mu[1] * mu[1]) / det6(mu);
        // pdelta_model
        const double pmu = 4 * PI * 1.e-7;   // pset_pmodel, problem.f90:251
        E.pe_dsigma[i][0] = sg[0] - E.pe_psigma; E.pe_dsigma[i][1] = sg[1]; E.pe_dsigma[i][2] = sg[2];
        E.pe_dsigma[i][3] = sg[3] - E.pe_psigma; E.pe_dsigma[i][4] = sg[4]; E.pe_dsigma[i][5] = sg[5] - E.pe_psigma;
        E.pe_dmu[i][0] = mu[0] - pmu; E.pe_dmu[i][1] = mu[1]; E.pe_dmu[i][2] = mu[2];
        E.pe_dmu[i][3] = mu[3] - pmu; E.pe_dmu[i][4] = mu[4]; E.pe_dmu[i][5] = mu[5] - pmu;
        // p_pfields d=1 (Ex,Hy) and d=2 (Ey,Hx); e0 = 0.d0
        const double z = c.d.g_zp[id];
        {
            const C t = cmplx32(0.0, E.omega * B0 * z);
            E.pe_ep[i][0] = mk(0.0 - t.re, -t.im);                          // e0-cmplx(0.d0,omega*b0*z)
            E.pe_hp[i][1] = mk(cmplx32(B0, 0.0).re / (4 * PI * 1.e-7), cmplx32(B0, 0.0).im / (4 * PI * 1.e-7));
        }
        {
            const C t = cmplx32(0.0, E.omega * B0 * z);
            E.pe_ep[i + mn][1] = mk(0.0 + t.re, t.im);                      // e0+cmplx(0.0,omega*b0*z)
            E.pe_hp[i + mn][0] = mk(cmplx32(B0, 0.0).re / (4 * PI * 1.e-7), cmplx32(B0, 0.0).im / (4 * PI * 1.e-7));
        }
    }
}

// ---- problem.f90:376-420 pe_modelcurl (pe_sch=1) ------------------------------------------
static void pe_modelcurl(const Elem &E, int d, int i, C v[3]) {
    const int j = i + (d - 1) * E.c->mn;
    const double *im = E.pe_inmu[i], *dm = E.pe_dmu[i];
    double m[3][3];
    m[0][0] = im[0] * dm[0] + im[1] * dm[1] + im[2] * dm[2];
    m[0][1] = im[0] * dm[1] + im[1] * dm[3] + im[2] * dm[4];
    m[0][2] = im[0] * dm[2] + im[1] * dm[4] + im[2] * dm[5];
    m[1][0] = im[1] * dm[0] + im[3] * dm[1] + im[4] * dm[2];
    m[1][1] = im[1] * dm[1] + im[3] * dm[3] + im[4] * dm[4];
    m[1][2] = im[1] * dm[2] + im[3] * dm[4] + im[4] * dm[5];
    m[2][0] = im[2] * dm[0] + im[4] * dm[1] + im[5] * dm[2];
    m[2][1] = im[2] * dm[1] + im[4] * dm[3] + im[5] * dm[4];
    m[2][2] = im[2] * dm[2] + im[4] * dm[4] + im[5] * dm[5];
    // m is declared complex in the reference: complex*complex products with zero imaginary parts
    for (int r = 0; r < 3; ++r)
        v[r] = mk(m[r][0], 0) * E.pe_hp[j][0] + mk(m[r][1], 0) * E.pe_hp[j][1] + mk(m[r][2], 0) * E.pe_hp[j][2];
}

// ---- problem.f90:434-457 pe_dmodel_pfield (pe_sch=1) ---------------------------------------
static void pe_dmodel_pfield(const Elem &E, int d, int i, C v[3]) {
    const int j = i + (d - 1) * E.c->mn;
    const C *s = E.pe_dsigma[i];
    const C *ep = E.pe_ep[j];
    v[0] = s[0] * ep[0] + s[1] * ep[1] + s[2] * ep[2];
    v[1] = s[1] * ep[0] + s[3] * ep[1] + s[4] * ep[2];
    v[2] = s[2] * ep[0] + s[4] * ep[1] + s[5] * ep[2];
}

// ---- problem.f90:91-128 p_source (pe_sch=1, ndir=2) ----------------------------------------
static void p_source(Elem &E, const double r[3], C out[2][3]) {
    const Ctx &c = *E.c;
    C pcrl[2][3], dmpf[2][3];
    for (int d = 0; d < 2; ++d)
        for (int m = 0; m < 3; ++m) { pcrl[d][m] = mk(0, 0); dmpf[d][m] = mk(0, 0); }
    for (int i = 0; i < c.mn; ++i) {
        for (int d = 1; d <= 2; ++d) {
            // p_pcurl, problem.f90:362-374
            double d_ne[3];
            C v[3];
            nf_grad_ln(E, i + 1, r, d_ne);
            pe_modelcurl(E, d, i, v);
            C pc[3];
            pc[0] = (v[2] * d_ne[1] - v[1] * d_ne[2]);
            pc[1] = (v[0] * d_ne[2] - v[2] * d_ne[0]);
            pc[2] = (v[1] * d_ne[0] - v[0] * d_ne[1]);
            for (int m = 0; m < 3; ++m) pcrl[d - 1][m] = pcrl[d - 1][m] + pc[m];
            // p_dmpf, problem.f90:424-432
            C w[3];
            pe_dmodel_pfield(E, d, i, w);
            const double ln = c.shape.nf_ln(i + 1, r[0], r[1], r[2]);
            for (int m = 0; m < 3; ++m) dmpf[d - 1][m] = dmpf[d - 1][m] + ln * w[m];
        }
    }
    const C miw = cmplx32(0.0, -E.omega);   // cmplx(0.d0,-omega): single precision (Q2)
    for (int d = 0; d < 2; ++d)
        for (int m = 0; m < 3; ++m) out[d][m] = (dmpf[d][m] + pcrl[d][m]) * miw;
}

// ---- boundary_conds.f90:84-186 gpml_h ------------------------------------------------------
static C gpml_h_axis(const Ctx &c, int flag, double r, const double a_[2], const double b_[2], double omega) {
    if (flag == 0) return mk(1.0, 0.0);
    const int s = (flag == -1) ? 0 : 1;
    const double ww_pml = std::sqrt((c.omegar[1] - c.omegar[0]) * (c.omegar[1] - c.omegar[0]));
    const double ww = std::sqrt((omega - c.omegar[0]) * (omega - c.omegar[0]));
    double a0 = c.d.a0, b0 = c.d.b0;
    if (c.d.gpml_sch == 1) { a0 = 100.0 * (ww / ww_pml); b0 = (1.e6 - 1.e-2) * (ww / ww_pml) + 1.e-2; }
    const double rr_pml = std::sqrt((b_[s] - a_[s]) * (b_[s] - a_[s]));
    const double rr = std::sqrt((r - a_[s]) * (r - a_[s]));
    if (c.d.gpml_sch == 0) {
        const double hx0 = 1.0 + a0 * std::pow(rr / rr_pml, c.d.nn);
        const double sn = std::sin((PI / 2.0) * (rr / rr_pml));
        const double bx = b0 * (sn * sn);
        const C t = cmplx32(1.0, -bx / (omega * EPS0));
        return hx0 * t;
    }
    // bx = b0*(rr/rr_pml)**nn / cmplx(a0,omega): real / single-complex, real part kept
    const C q = (b0 * std::pow(rr / rr_pml, c.d.nn)) / cmplx32(a0, omega);
    const double bx = q.re;
    return 1.0 + cmplx32(0.0, bx);
}

// ---- integration.f90:60-74,108-152 int_elem_params / int_param / int_edges -----------------
static void int_elem_params(Elem &E) {
    const Ctx &c = *E.c;
    const int ngp = c.ngp, me = c.me, mn = c.mn;
    for (int id = 0; id < ngp; ++id) {
        const double *r = c.rw[id];
        nf_jacobian(E, r[0], r[1], r[2]);
        E.wgt[id] = (nf_det(E.nf_j)) * r[3];
        // p_intmodels (pe_sch=1), problem.f90:139-142, accumulating into zeroed mf1/mf2 (Q15)
        for (int k = 0; k < 6; ++k) { E.mf1[id][k] = mk(0, 0); E.mf2[id][k] = mk(0, 0); }
        for (int i = 0; i < mn; ++i) {
            const double ln = c.shape.nf_ln(i + 1, r[0], r[1], r[2]);
            const C *sg = E.g_sigma + (size_t)6 * (E.nf_index[i] - 1);
            for (int k = 0; k < 6; ++k) {
                E.mf1[id][k] = mk(E.mf1[id][k].re + ln * E.pe_inmu[i][k], E.mf1[id][k].im);
                E.mf2[id][k] = E.mf2[id][k] + ln * sg[k];
            }
        }
        if (!c.d.dirichlet) {
            double g_rw[3] = {0.0, 0.0, 0.0};
            for (int j = 0; j < mn; ++j) {
                g_rw[0] = g_rw[0] + c.shape.nf_ln(j + 1, r[0], r[1], r[2]) * E.nf_re[j][0];
                g_rw[1] = g_rw[1] + c.shape.nf_ln(j + 1, r[0], r[1], r[2]) * E.nf_re[j][1];
                g_rw[2] = g_rw[2] + c.shape.nf_ln(j + 1, r[0], r[1], r[2]) * E.nf_re[j][2];
            }
            // Q18: integration.f90:16 declares gpml REAL(kind=double), so `gpml(i,:)=gpml_h(g_rw(1:3))`
            // (integration.f90:125) keeps only the real part of the complex stretch; f1/f2/f3 then read it back
            // into a complex h with zero imaginary part (integration.f90:170,225,253).  Found by executing the
            // reference source (tests/golden/f90exec.py); scheme 1 (Zhou) therefore has h = 1 exactly.
            E.gpml[id][0] = mk(gpml_h_axis(c, E.in_pml[0], g_rw[0], c.xa, c.xb, E.omega).re, 0.0);
            E.gpml[id][1] = mk(gpml_h_axis(c, E.in_pml[1], g_rw[1], c.ya, c.yb, E.omega).re, 0.0);
            E.gpml[id][2] = mk(gpml_h_axis(c, E.in_pml[2], g_rw[2], c.za, c.zb, E.omega).re, 0.0);
        }
        C t_src[2][3];
        p_source(E, r, t_src);
        for (int d = 0; d < 2; ++d)
            for (int m = 0; m < 3; ++m) E.src[id + d * ngp][m] = t_src[d][m];
        // int_edges
        for (int j = 1; j <= me; ++j) {
            double cve[3][2], v[3];
            vf_elem_curl(E, j, r, cve);
            vf_elem_ve(E, j, r, v);
            const int row = id + (j - 1) * ngp;
            for (int m = 0; m < 3; ++m) { E.cve1[row][m] = cve[m][0]; E.cve2[row][m] = cve[m][1]; E.ve[row][m] = v[m]; }
        }
    }
}

// ---- integration.f90:154-209 f1 (declared real: only the real part survives, Q3) -----------
static double f1(const Elem &E, int ig, int im, int jm) {
    const Ctx &c = *E.c;
    const int ngp = c.ngp;
    double cv1[3][2], cv2[3][2];
    for (int p = 0; p < 3; ++p) {
        cv1[p][0] = E.cve1[ig + (im - 1) * ngp][p]; cv1[p][1] = E.cve2[ig + (im - 1) * ngp][p];
        cv2[p][0] = E.cve1[ig + (jm - 1) * ngp][p]; cv2[p][1] = E.cve2[ig + (jm - 1) * ngp][p];
    }
    const C *m = E.mf1[ig];   // m(1..6) -> m[0..5]
#define A(p, s) cv1[p - 1][s - 1]
#define B(p, s) cv2[p - 1][s - 1]
    C r;
    if (!c.d.dirichlet) {
        const C h1 = E.gpml[ig][0], h2 = E.gpml[ig][1], h3 = E.gpml[ig][2];
        r = (h1 * h3 / h2) * m[0] * A(1, 1) * B(1, 1) - (h1)*m[0] * A(1, 2) * B(1, 1) -
            (h1)*m[0] * A(1, 1) * B(1, 2) + (h1 * h2 / h3) * m[0] * A(1, 2) * B(1, 2) +
            (h1)*m[1] * A(1, 1) * B(2, 1) - (h1 * h2 / h3) * m[1] * A(1, 2) * B(2, 1) -
            (h3)*m[1] * A(1, 1) * B(2, 2) + (h2)*m[1] * A(1, 2) * B(2, 2) +
            (h3)*m[2] * A(1, 1) * B(3, 1) - (h2)*m[2] * A(1, 2) * B(3, 1) -
            (h1 * h3 / h2) * m[2] * A(1, 1) * B(3, 2) + (h1)*m[2] * A(1, 2) * B(3, 2) +
            (h1)*m[1] * A(2, 1) * B(1, 1) - (h3)*m[1] * A(2, 2) * B(1, 1) -
            (h1 * h2 / h3) * m[1] * A(2, 1) * B(1, 2) + (h2)*m[1] * A(2, 2) * B(1, 2) +
            (h1 * h2 / h3) * m[3] * A(2, 1) * B(2, 1) - h2 * m[3] * A(2, 2) * B(2, 1) -
            h2 * m[3] * A(2, 1) * B(2, 2) + (h2 * h3 / h1) * m[3] * A(2, 2) * B(2, 2) +
            (h2)*m[4] * A(2, 1) * B(3, 1) - (h2 * h3 / h1) * m[4] * A(2, 2) * B(3, 1) -
            h1 * m[4] * A(2, 1) * B(3, 2) + h3 * m[4] * A(2, 2) * B(3, 2) +
            h3 * m[2] * A(3, 1) * B(1, 1) - (h1 * h3 / h2) * m[2] * A(3, 2) * B(1, 1) -
            h2 * m[2] * A(3, 1) * B(1, 2) + h1 * m[2] * A(3, 2) * B(1, 2) +
            h2 * m[4] * A(3, 1) * B(2, 1) - h1 * m[4] * A(3, 2) * B(2, 1) -
            (h2 * h3 / h1) * m[4] * A(3, 1) * B(2, 2) + h3 * m[4] * A(3, 2) * B(2, 2) +
            (h2 * h3 / h1) * m[5] * A(3, 1) * B(3, 1) - h3 * m[5] * A(3, 2) * B(3, 1) -
            h3 * m[5] * A(3, 1) * B(3, 2) + (h1 * h3 / h2) * m[5] * A(3, 2) * B(3, 2);
    } else {
        r = m[0] * A(1, 1) * B(1, 1) - m[0] * A(1, 2) * B(1, 1) -
            m[0] * A(1, 1) * B(1, 2) + m[0] * A(1, 2) * B(1, 2) +
            m[1] * A(1, 1) * B(2, 1) - m[1] * A(1, 2) * B(2, 1) -
            m[1] * A(1, 1) * B(2, 2) + m[1] * A(1, 2) * B(2, 2) +
            m[2] * A(1, 1) * B(3, 1) - m[2] * A(1, 2) * B(3, 1) -
            m[2] * A(1, 1) * B(3, 2) + m[2] * A(1, 2) * B(3, 2) +
            m[1] * A(2, 1) * B(1, 1) - m[1] * A(2, 2) * B(1, 1) -
            m[1] * A(2, 1) * B(1, 2) + m[1] * A(2, 2) * B(1, 2) +
            m[3] * A(2, 1) * B(2, 1) - m[3] * A(2, 2) * B(2, 1) -
            m[3] * A(2, 1) * B(2, 2) + m[3] * A(2, 2) * B(2, 2) +
            m[4] * A(2, 1) * B(3, 1) - m[4] * A(2, 2) * B(3, 1) -
            m[4] * A(2, 1) * B(3, 2) + m[4] * A(2, 2) * B(3, 2) +
            m[2] * A(3, 1) * B(1, 1) - m[2] * A(3, 2) * B(1, 1) -
            m[2] * A(3, 1) * B(1, 2) + m[2] * A(3, 2) * B(1, 2) +
            m[4] * A(3, 1) * B(2, 1) - m[4] * A(3, 2) * B(2, 1) -
            m[4] * A(3, 1) * B(2, 2) + m[4] * A(3, 2) * B(2, 2) +
            m[5] * A(3, 1) * B(3, 1) - m[5] * A(3, 2) * B(3, 1) -
            m[5] * A(3, 1) * B(3, 2) + m[5] * A(3, 2) * B(3, 2);
    }
#undef A
#undef B
    return r.re;
}

// ---- integration.f90:211-238 f2 (declared real, Q3) ----------------------------------------
static double f2(const Elem &E, int ig, int im, int jm) {
    const Ctx &c = *E.c;
    const int ngp = c.ngp;
    const double *cv1 = E.ve[ig + (im - 1) * ngp];
    const double *cv2 = E.ve[ig + (jm - 1) * ngp];
    const C *m = E.mf2[ig];
    C r;
    if (!c.d.dirichlet) {
        const C h1 = E.gpml[ig][0], h2 = E.gpml[ig][1], h3 = E.gpml[ig][2];
        r = h1 * h2 * h3 * cv2[0] * m[0] * cv1[0] + h1 * h2 * h3 * cv2[1] * m[1] * cv1[0] +
            h1 * h2 * h3 * cv2[2] * m[2] * cv1[0] + h1 * h2 * h3 * cv2[0] * m[1] * cv1[1] +
            h1 * h2 * h3 * cv2[1] * m[3] * cv1[1] + h1 * h2 * h3 * cv2[2] * m[4] * cv1[1] +
            h1 * h2 * h3 * cv2[0] * m[2] * cv1[2] + h1 * h2 * h3 * cv2[1] * m[4] * cv1[2] +
            h1 * h2 * h3 * cv2[2] * m[5] * cv1[2];
    } else {
        r = (cv2[0] * m[0] * cv1[0] + cv2[1] * m[1] * cv1[0] + cv2[2] * m[2] * cv1[0]) +
            (cv2[0] * m[1] * cv1[1] + cv2[1] * m[3] * cv1[1] + cv2[2] * m[4] * cv1[1]) +
            (cv2[0] * m[2] * cv1[2] + cv2[1] * m[4] * cv1[2] + cv2[2] * m[5] * cv1[2]);
    }
    return r.re;
}

// ---- integration.f90:240-265 f3 -------------------------------------------------------------
static C f3(const Elem &E, int im, int i, int edir) {
    const Ctx &c = *E.c;
    const int ngp = c.ngp;
    const int j = i + (edir - 1) * ngp;
    const double *v = E.ve[i + (im - 1) * ngp];
    if (!c.d.dirichlet) {
        const C h1 = E.gpml[i][0], h2 = E.gpml[i][1], h3 = E.gpml[i][2];
        return h1 * h2 * h3 * v[0] * E.src[j][0] + h1 * h2 * h3 * v[1] * E.src[j][1] + h1 * h2 * h3 * v[2] * E.src[j][2];
    }
    return v[0] * E.src[j][0] + v[1] * E.src[j][1] + v[2] * E.src[j][2];
}

// ---- integration.f90:76-86 alocal (Q2: cmplx(0.d0,omega) is single precision; Q4: +i*omega) --
static C alocal(const Elem &E, int im, int jm) {
    const Ctx &c = *E.c;
    C a = mk(0, 0);
    const C iw = cmplx32(0.0, E.omega);
    for (int i = 0; i < c.ngp; ++i) {
        const double v1 = f1(E, i, im, jm), v2 = f2(E, i, im, jm);
        const C t = v1 + iw * mk(v2, 0.0);
        a = a + E.wgt[i] * t;
    }
    return a;
}

// ---- integration.f90:88-106 blocal -----------------------------------------------------------
static void blocal(const Elem &E, int im, C b[2]) {
    const Ctx &c = *E.c;
    b[0] = mk(0, 0); b[1] = mk(0, 0);
    for (int i = 0; i < c.ngp; ++i)
        for (int edir = 1; edir <= 2; ++edir) b[edir - 1] = b[edir - 1] + E.wgt[i] * f3(E, im, i, edir);
}

// ---- boundary_conds.f90:262-390 edge_bdary ---------------------------------------------------
static int edge_bdary(const Ctx &c, int ie, int je, int ke, int im) {
    static const int f12[6][4] = {{1, 2, 3, 4}, {1, 5, 6, 9}, {2, 5, 7, 10}, {9, 10, 11, 12}, {4, 7, 8, 12}, {3, 6, 8, 11}};
    static const int f36[6][10] = {{1, 2, 3, 4, 5, 6, 7, 8, 26, 31},       {1, 2, 9, 10, 11, 12, 17, 18, 25, 32},
                                   {3, 4, 9, 10, 13, 14, 19, 20, 27, 33},   {17, 18, 19, 20, 21, 22, 23, 24, 30, 36},
                                   {7, 8, 13, 14, 15, 16, 23, 24, 29, 35},  {5, 6, 11, 12, 15, 16, 21, 22, 28, 34}};
    static const int f54[6][12] = {{1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12},          {1, 2, 13, 14, 15, 21, 22, 29, 30, 31, 37, 38},
                                   {3, 8, 13, 16, 18, 23, 24, 29, 32, 34, 39, 44},   {37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48},
                                   {11, 12, 18, 19, 20, 27, 28, 34, 35, 36, 47, 48}, {5, 10, 15, 17, 20, 25, 26, 31, 33, 36, 41, 46}};
    const int mef = c.me == 12 ? 4 : (c.me == 36 ? 10 : 12);
    const bool on[6] = {ie == 1, je == 1, ke == 1, ie == c.nx, je == c.ny, ke == c.nz};   // order of the reference's if-chain
    for (int f = 0; f < 6; ++f) {
        if (!on[f]) continue;
        const int *lst = c.me == 12 ? f12[f] : (c.me == 36 ? f36[f] : f54[f]);
        for (int i = 0; i < mef; ++i)
            if (lst[i] == im) return f + 1;
    }
    return 0;
}

// ---- global_assembly.f90:231-475 c_gne12 / c_gne36 / c_gne54 ---------------------------------
struct ShareTab { int n; int mine[12]; int theirs[12]; };
static void share_tables(int me, ShareTab &zt, ShareTab &yt, ShareTab &xt) {
    if (me == 12) {
        zt = {4, {2, 5, 7, 10}, {3, 6, 8, 11}};
        yt = {4, {1, 5, 6, 9}, {4, 7, 8, 12}};
        xt = {4, {1, 2, 3, 4}, {9, 10, 11, 12}};
    } else if (me == 36) {
        zt = {10, {3, 4, 9, 10, 13, 14, 19, 20, 27, 33}, {5, 6, 11, 12, 15, 16, 21, 22, 28, 34}};
        yt = {10, {1, 2, 9, 10, 11, 12, 17, 18, 25, 32}, {7, 8, 13, 14, 15, 16, 23, 24, 29, 35}};
        xt = {10, {1, 2, 3, 4, 5, 6, 7, 8, 26, 31}, {17, 18, 19, 20, 21, 22, 23, 24, 30, 36}};
    } else {
        zt = {12, {3, 8, 13, 16, 18, 23, 24, 29, 32, 34, 39, 44}, {5, 10, 15, 17, 20, 25, 26, 31, 33, 36, 41, 46}};
        yt = {12, {1, 2, 13, 14, 15, 21, 22, 29, 30, 31, 37, 38}, {11, 12, 18, 19, 20, 27, 28, 34, 35, 36, 47, 48}};
        xt = {12, {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}, {37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48}};
    }
}

static void build_gne(Ctx &c) {
    const int nx = c.nx, ny = c.ny, nz = c.nz, me = c.me, ne = c.ne;
    c.gne.assign((size_t)ne * me, 0);
    auto G = [&](int ide, int im) -> int & { return c.gne[(size_t)(im - 1) * ne + (ide - 1)]; };
    ShareTab zt, yt, xt;
    share_tables(me, zt, yt, xt);
    int indx = 0;
    std::vector<int> iedge(me + 1);
    for (int ie = 1; ie <= nx; ++ie)
        for (int je = 1; je <= ny; ++je)
            for (int ke = 1; ke <= nz; ++ke) {
                const int ide = (ie - 1) * ny * nz + (je - 1) * nz + ke;
                std::fill(iedge.begin(), iedge.end(), 0);
                if (ke > 1) {
                    const int ide1 = (ie - 1) * ny * nz + (je - 1) * nz + (ke - 1);
                    for (int t = 0; t < zt.n; ++t) { G(ide, zt.mine[t]) = G(ide1, zt.theirs[t]); iedge[zt.mine[t]] = 1; }
                }
                if (je > 1) {
                    const int ide1 = (ie - 1) * ny * nz + (je - 2) * nz + ke;
                    for (int t = 0; t < yt.n; ++t) { G(ide, yt.mine[t]) = G(ide1, yt.theirs[t]); iedge[yt.mine[t]] = 1; }
                }
                if (ie > 1) {
                    const int ide1 = (ie - 2) * ny * nz + (je - 1) * nz + ke;
                    for (int t = 0; t < xt.n; ++t) { G(ide, xt.mine[t]) = G(ide1, xt.theirs[t]); iedge[xt.mine[t]] = 1; }
                }
                if (c.d.dirichlet) {
                    const bool bdaryel = (ie == 1 || ie == nx || je == 1 || je == ny || ke == 1 || ke == nz);
                    for (int im = 1; im <= me; ++im) {
                        if (bdaryel) {
                            const int edgebd = edge_bdary(c, ie, je, ke, im);
                            if (edgebd != 0) G(ide, im) = -edgebd;
                            else if (iedge[im] == 0) { indx = indx + 1; G(ide, im) = indx; }
                        } else if (iedge[im] == 0) { indx = indx + 1; G(ide, im) = indx; }
                    }
                } else {
                    for (int im = 1; im <= me; ++im)
                        if (iedge[im] == 0) { indx = indx + 1; G(ide, im) = indx; }
                }
            }
    c.nne = *std::max_element(c.gne.begin(), c.gne.end());   // nne=maxval(gne), global_assembly.f90:194
}

/*
 * global_assembly.f90:197-229 ga_nzindx + shr_nzindx12/36/54 (488-1906).  The 1400 lines of
 * hand-enumerated neighbour sharing are NOT transliterated: SURVEY Q9 / App. B.1 established
 * (by executing a mechanical translation on meshes up to 5x3x2 / 3x4x5, all three element
 * types, both boundary modes) that their net effect is "slots of different elements that
 * address the same (row,col) share one index; nnze = number of unique (row,col) pairs; rows
 * appear in ascending order".  The order WITHIN a row differs from the reference's, which is
 * immaterial: ga_sort_sparse's first pass is a stable sort on ja, so its output depends only
 * on rows being ascending, and the final triplets are fully sorted by (ia,ja).
 */
static void build_pattern(Ctx &c) {
    const int me = c.me, ne = c.ne, nne = c.nne;
    auto G = [&](int ide, int im) -> int { return c.gne[(size_t)(im - 1) * ne + (ide - 1)]; };
    // per-row column sets via counting + sort/unique (rows have <= 4*me entries)
    std::vector<int64_t> cnt((size_t)nne + 2, 0);
    for (int ide = 1; ide <= ne; ++ide)
        for (int im = 1; im <= me; ++im) {
            const int r = G(ide, im);
            if (r < 0) continue;
            int nv = 0;
            for (int jm = 1; jm <= me; ++jm) nv += (G(ide, jm) >= 0);
            cnt[r] += nv;
        }
    std::vector<int64_t> off((size_t)nne + 2, 0);
    for (int r = 1; r <= nne; ++r) off[r + 1] = off[r] + cnt[r];
    std::vector<int> cols((size_t)off[nne + 1]);
    std::vector<int64_t> fill(off.begin(), off.end());
    for (int ide = 1; ide <= ne; ++ide)
        for (int im = 1; im <= me; ++im) {
            const int r = G(ide, im);
            if (r < 0) continue;
            for (int jm = 1; jm <= me; ++jm) {
                const int cc = G(ide, jm);
                if (cc >= 0) cols[fill[r]++] = cc;
            }
        }
    c.row_ptr.assign((size_t)nne + 1, 0);
    c.pat_ia.clear(); c.pat_ja.clear();
    for (int r = 1; r <= nne; ++r) {
        auto b = cols.begin() + off[r], e = cols.begin() + off[r + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        for (auto it = b; it != e; ++it) { c.pat_ia.push_back(r); c.pat_ja.push_back(*it); }
        c.row_ptr[r] = (int64_t)c.pat_ia.size();
    }
    c.nnze = (int64_t)c.pat_ia.size();
}

// position of (r,cc) in the pattern (the nzindx lookup)
static inline int64_t pat_find(const Ctx &c, int r, int cc) {
    const int *b = c.pat_ja.data() + c.row_ptr[r - 1], *e = c.pat_ja.data() + c.row_ptr[r];
    const int *p = std::lower_bound(b, e, cc);
    return (p != e && *p == cc) ? (int64_t)(p - c.pat_ja.data()) : -1;
}

// ---- boundary_conds.f90:38-70 init_bdary / init_gpml ----------------------------------------
static void init_gpml(Ctx &c) {
    const int nextd = c.d.nextd, g = c.d.nord;
    c.el_xa[0] = nextd; c.el_xa[1] = c.d.g_nx - nextd; c.el_xb[0] = 1; c.el_xb[1] = c.d.g_nx - 1;
    c.el_ya[0] = nextd; c.el_ya[1] = c.d.g_ny - nextd; c.el_yb[0] = 1; c.el_yb[1] = c.d.g_ny - 1;
    c.el_za[0] = nextd; c.el_za[1] = c.d.g_nz - nextd; c.el_zb[0] = 1; c.el_zb[1] = c.d.g_nz - 1;
    const double *xp = c.d.g_xp, *yp = c.d.g_yp, *zp = c.d.g_zp;
    c.xa[0] = xp[nextd * (g - 1) + 1 - 1]; c.xa[1] = xp[c.nnx - nextd * (g - 1) - 1];
    c.xb[0] = xp[0]; c.xb[1] = xp[c.nnx - 1];
    c.ya[0] = yp[nextd * (g - 1) + 1 - 1]; c.ya[1] = yp[c.nny - nextd * (g - 1) - 1];
    c.yb[0] = yp[0]; c.yb[1] = yp[c.nny - 1];
    const int64_t id1 = nextd * (g - 1) + 1;
    const int64_t id2 = (int64_t)(c.nnx - 1) * c.nyz + (int64_t)(c.nny - 1) * c.nnz + c.nnz - c.d.nzl_top * (g - 1);
    c.za[0] = zp[id1 - 1]; c.za[1] = zp[id2 - 1];
    c.zb[0] = zp[0]; c.zb[1] = zp[(int64_t)(c.nnx - 1) * c.nyz + (int64_t)(c.nny - 1) * c.nnz + c.nnz - 1];
    const double f1_ = (double)1.e-5f, f2_ = (double)1.e3f;   // boundary_conds.f90:40, default-real literals
    c.omegar[0] = 2.0 * PI * f1_; c.omegar[1] = 2.0 * PI * f2_;
}

// ---- boundary_conds.f90:72-82 get_pml (Q7: the lower-side tests can never be true) ----------
static void get_pml(const Ctx &c, int i, int j, int k, int in_pml[3]) {
    in_pml[0] = in_pml[1] = in_pml[2] = 0;
    if (i >= c.el_xa[0] && i <= c.el_xb[0]) in_pml[0] = -1;
    if (i >= c.el_xa[1] && i <= c.el_xb[1]) in_pml[0] = 1;
    if (j >= c.el_ya[0] && j <= c.el_yb[0]) in_pml[1] = -1;
    if (j >= c.el_ya[1] && j <= c.el_yb[1]) in_pml[1] = 1;
    if (k >= c.el_za[0] && k <= c.el_zb[0]) in_pml[2] = -1;
    if (k >= c.el_za[1] && k <= c.el_zb[1]) in_pml[2] = 1;
}

// flags the reference actually uses for element ide (Q17): those of its predecessor in loop order
static void effective_pml(const Ctx &c, int ide, const int first_flags[3], int out[3]) {
    if (ide == 1) { out[0] = first_flags[0]; out[1] = first_flags[1]; out[2] = first_flags[2]; return; }
    const int p = ide - 2;   // 0-based predecessor
    const int ke = p % c.nz + 1, je = (p / c.nz) % c.ny + 1, ie = p / (c.nz * c.ny) + 1;
    get_pml(c, ie, je, ke, out);
}

// ---- Dirichlet boundary models 2 / 3: boundary_conds.f90:188-250 f_boundary, :392-430 bd_setmodel (+ bd_updatemodel,
//      :48-50, called every frequency), :436-518 p_pfields (the module's own, NOT problem.f90's), :523-552
//      wait_recursion, :557-598 ezl.  Units and sign conventions are the reference's (depths in km against layer
//      coordinates <= 0, single-precision cmplx() constructors); they are restated, not repaired. ----
struct BdModel {
    int nl;
    C psig[17];              // pe_psigma: (f32(sigma), f32(eps*omega)) per medium / layer
    double pmu;              // pe_pmu(1)
    double dl[16], zl[17];   // pe_dl, pe_zl
    C cz[17], ez[17];        // wait_recursion
};
typedef std::complex<double> Z;
static inline Z zc(C a) { return Z(a.re, a.im); }
static inline C cz_(Z a) { return mk(a.real(), a.imag()); }
static inline C csqrt_(C a) { return cz_(std::sqrt(zc(a))); }
static inline C cexp_(C a) { return cz_(std::exp(zc(a))); }
static const C CI = {0.0, 1.0};

static void bd_wait_recursion(BdModel &B, double omega) {
    const int nl = B.nl;
    for (int l = 0; l < nl; ++l) { B.cz[l] = mk(0, 0); B.ez[l] = mk(0, 0); }
    C alpha = (omega * B.pmu) * B.psig[nl - 1];
    C gamma = csqrt_(CI * alpha);
    B.cz[nl - 1] = 1.0 / gamma;
    for (int l = nl - 1; l >= 1; --l) {
        alpha = (omega * B.pmu) * B.psig[l - 1];
        gamma = csqrt_(CI * alpha);
        const C r = (1.0 + (-(gamma * B.cz[l]))) / (1.0 + gamma * B.cz[l]);
        const C e2 = cexp_((-2.0 * gamma) * B.dl[l - 1]);
        B.cz[l - 1] = (1.0 + (-(r * e2))) / (gamma * (1.0 + r * e2));
    }
    B.ez[0] = mk(1.0, 0.0);
    for (int l = 1; l <= nl - 1; ++l) {
        alpha = (omega * B.pmu) * B.psig[l - 1];
        gamma = csqrt_(CI * alpha);
        B.ez[l] = cexp_((-gamma) * B.dl[l - 1]) * (B.ez[l - 1] * B.cz[l] * (1.0 + B.cz[l - 1] * gamma)) / (B.cz[l - 1] * (1.0 + B.cz[l] * gamma));
    }
}

static void bd_model(const Ctx &c, double omega, BdModel &B) {
    const movfem_desc &d = c.d;
    B.pmu = 4.0 * PI * 1.e-7;
    const double im = (double)(float)(EPS0 * omega);       // bd_updatemodel: real(pe_psigma)+cmplx(0.d0,eps*omega)
    if (d.bd_inimod == 2) {
        B.nl = 2;
        B.psig[0] = mk(0.0, im);                            // air
        B.psig[1] = mk((double)(float)d.bd_hsigma, im);     // cmplx(h_sigma,omega*eps): single precision
    } else {
        B.nl = d.bd_nl;
        for (int l = 0; l < B.nl; ++l) B.psig[l] = mk((double)(float)d.bd_lsigma[l], im);
        for (int l = 0; l < B.nl - 1; ++l) B.dl[l] = d.bd_ldz[l];
        B.zl[0] = 0.0;
        for (int l = 1; l < B.nl; ++l) B.zl[l] = B.zl[l - 1] - B.dl[l - 1];
        bd_wait_recursion(B, omega);
    }
}

static C bd_ezl(const BdModel &B, int d, double z, double omega) {
    C out = mk(0, 0);
    const int nl = B.nl;
    for (int l = 1; l <= nl - 1; ++l) {
        if (z <= B.zl[l - 1] && z > B.zl[l]) {
            const C alpha = (omega * B.pmu) * B.psig[l - 1];
            const C gamma = csqrt_(CI * alpha);
            const C r = (1.0 + (-(gamma * B.cz[l]))) / (1.0 + gamma * B.cz[l]);
            const C e2 = cexp_((-2.0 * gamma) * (z - B.zl[l])), e1 = cexp_((-gamma) * (B.zl[l - 1] - z));
            if (d == 0) out = (B.ez[l - 1] * 0.5) * (1.0 + 1.0 / (B.cz[l - 1] * gamma)) * (1.0 + (-(r * e2))) * e1;
            else out = (-(B.ez[l - 1] * 0.5)) * (gamma + 1.0 / B.cz[l - 1]) * (1.0 + r * e2) * e1;
        }
    }
    if (z <= B.zl[nl - 1]) {
        const C alpha = (omega * B.pmu) * B.psig[nl - 1];
        const C gamma = csqrt_(CI * alpha);
        const C e1 = cexp_((-gamma) * (B.zl[nl - 1] - z));
        out = d == 0 ? B.ez[nl - 1] * e1 : B.ez[nl - 1] * e1 / (-gamma);
    }
    return out;
}

// pe_ep(1,1) and pe_ep(2,2) of boundary_conds.f90's p_pfields at grid node `id` (1-based)
static void bd_pfields(const Ctx &c, const BdModel &B, double omega, int id, C &ex1, C &ey2) {
    const double bb0 = 1.e-9;
    const double zn = c.d.g_zp[id - 1];
    const double z = (c.d.g_ztop - zn) / 1000.0;
    if (c.d.bd_inimod == 2) {
        C fp;
        if (zn > c.d.g_ztop) {
            const C alpha = (omega * B.pmu) * B.psig[0];
            const C sq = csqrt_(CI * alpha);
            fp = (1.0 / sq) * (1.0 + (-(z * sq)));
        } else {
            const C alpha = (omega * B.pmu) * B.psig[1];
            const C sq = csqrt_(CI * alpha);
            fp = (1.0 / sq) * cexp_((-z) * sq);
        }
        ex1 = cmplx32(0.0, -2.0 * omega * bb0) * fp;
        ey2 = cmplx32(0.0, 2.0 * omega * bb0) * fp;
    } else {
        if (zn > c.d.g_ztop) {
            ex1 = cmplx32(0.0, 2.0 * omega * bb0) * (B.cz[0] + mk(-z, 0.0));
            ey2 = cmplx32(0.0, -2.0 * omega * bb0) * (B.cz[0] + mk(-z, 0.0));
        } else {
            const C e = bd_ezl(B, 0, z, omega);
            ex1 = cmplx32(0.0, 2.0 * omega * bb0) * B.cz[0] * e;
            ey2 = cmplx32(0.0, -2.0 * omega * bb0) * B.cz[0] * e;
        }
    }
}

// f_boundary(bd, jm), boundary_conds.f90:188-250: value of the Dirichlet DOF jm lying on face bd
static void f_boundary(Elem &E, const BdModel &B, int bd, int jm, C f[2]) {
    const Ctx &c = *E.c;
    f[0] = mk(0, 0); f[1] = mk(0, 0);
    if (!c.d.dirichlet || c.d.bd_inimod == 1) return;
    if (bd == 3 || bd == 6) return;
    const int i = c.enode[jm - 1], d = c.edir[jm - 1];
    nf_jacobian(E, c.shape.nr[i - 1][0], c.shape.nr[i - 1][1], c.shape.nr[i - 1][2]);
    const double *dne = E.nf_j[d - 1];                     // nf_dr_dxi: row d of the Jacobian
    C ex1, ey2;
    bd_pfields(c, B, E.omega, E.nf_index[i - 1], ex1, ey2);
    // f(1) = pe_ep(1,1)*dne(1)+pe_ep(2,1)*dne(2)+pe_ep(3,1)*dne(3) with pe_ep(2:3,1) = 0 ; f(2) likewise with pe_ep(2,2)
    f[0] = ex1 * dne[0] + mk(0, 0) * dne[1] + mk(0, 0) * dne[2];
    f[1] = mk(0, 0) * dne[0] + ey2 * dne[1] + mk(0, 0) * dne[2];
}

static void elem_setup(Elem &E, const Ctx &c, double omega, const C *g_sigma, int faithful) {
    E.c = &c; E.omega = omega; E.g_sigma = g_sigma; E.faithful = faithful; E.jac_builds = 0;
    E.jac_valid = false; E.status = 0;
    E.pe_psigma = cmplx32(0.0, omega * EPS0);   // pset_pmodel, problem.f90:250
}

// one element up to and including int_elem_params (MoVFEM_3DMT.f90:197-203)
static void elem_compute(Elem &E, int ide, const int pml[3]) {
    const Ctx &c = *E.c;
    const int p = ide - 1;
    const int ke = p % c.nz + 1, je = (p / c.nz) % c.ny + 1, ie = p / (c.nz * c.ny) + 1;
    const int g = c.d.nord;
    const int eno = (ie - 1) * c.nyz * (g - 1) + (je - 1) * c.nnz * (g - 1) + (ke - 1) * (g - 1) + 1;
    E.in_pml[0] = pml[0]; E.in_pml[1] = pml[1]; E.in_pml[2] = pml[2];
    nf_get_r(E, ie, je, ke, eno);
    p_elem_fields(E);
    if (E.status) return;
    int_elem_params(E);
}

}  // namespace oracle

using namespace oracle;

extern "C" {

struct oracle_ctx { Ctx *c; };

int oracle_create(const movfem_desc *d, oracle_ctx **out) {
    if (!d || !out) return MOVFEM_E_BADARG;
    if (!((d->mn == 8 && d->me == 12 && d->nord == 2) || (d->mn == 20 && d->me == 36 && d->nord == 3) ||
          (d->mn == 27 && d->me == 54 && d->nord == 3)))
        return MOVFEM_E_BADARG;
    if (d->ndir != 2 || d->pe_sch != 1 || d->sym != 1) return MOVFEM_E_UNSUPPORTED;
    if (d->dirichlet && (d->bd_inimod < 1 || d->bd_inimod > 3 || (d->bd_inimod == 3 && (d->bd_nl < 1 || d->bd_nl > 16)))) return MOVFEM_E_BADARG;
    Ctx *c = new Ctx(d->mn);
    c->d = *d;
    c->nx = d->g_nx - 1; c->ny = d->g_ny - 1; c->nz = d->g_nz - 1; c->ne = c->nx * c->ny * c->nz;
    c->nnx = c->nx * (d->nord - 1) + 1; c->nny = c->ny * (d->nord - 1) + 1; c->nnz = c->nz * (d->nord - 1) + 1;
    c->nyz = c->nny * c->nnz; c->npt = c->nnx * c->nyz;
    c->mn = d->mn; c->me = d->me;
    node_offsets(c->mn, d->nord, c->i1, c->j1, c->k1);
    edge_dir_table(c->me, c->enode, c->edir);
    double pt[3], wt[3];
    const int n1 = gauss_rule(c->me, pt, wt);
    c->ngp = n1 * n1 * n1;
    for (int a = 0; a < n1; ++a)
        for (int b = 0; b < n1; ++b)
            for (int k = 0; k < n1; ++k) {
                const int id = a * n1 * n1 + b * n1 + k;
                c->rw[id][0] = pt[a]; c->rw[id][1] = pt[b]; c->rw[id][2] = pt[k];
                c->rw[id][3] = wt[a] * wt[b] * wt[k];
            }
    build_gne(*c);
    build_pattern(*c);
    if (!d->dirichlet) init_gpml(*c);
    c->in_pml[0] = c->in_pml[1] = c->in_pml[2] = 0;   // SAVE variable, zero-initialised
    c->err[0] = 0;
    *out = new oracle_ctx{c};
    return 0;
}

void oracle_destroy(oracle_ctx *h) {
    if (h) { delete h->c; delete h; }
}

int oracle_sizes(const oracle_ctx *h, int32_t *nne, int64_t *nnze, int64_t *nz_upper) {
    const Ctx &c = *h->c;
    if (nne) *nne = c.nne;
    if (nnze) *nnze = c.nnze;
    if (nz_upper) {
        int64_t n = 0;
        for (int64_t k = 0; k < c.nnze; ++k) n += (c.pat_ia[k] <= c.pat_ja[k]);
        *nz_upper = n;
    }
    return 0;
}

int oracle_get_gne(const oracle_ctx *h, int32_t *gne) {
    std::memcpy(gne, h->c->gne.data(), sizeof(int) * h->c->gne.size());
    return 0;
}

int oracle_get_pattern(const oracle_ctx *h, int32_t *ia, int32_t *ja) {
    std::memcpy(ia, h->c->pat_ia.data(), sizeof(int) * h->c->nnze);
    std::memcpy(ja, h->c->pat_ja.data(), sizeof(int) * h->c->nnze);
    return 0;
}

// reference tables evaluated at the Gauss points (for bit-level comparison with the product's
// host-side tables): N[g][l], dN[g][l][3], phi[g][e], dphi[g][e][3], rw[g][4]
int oracle_tables(const oracle_ctx *h, double *N, double *dN, double *phi, double *dphi, double *rw) {
    const Ctx &c = *h->c;
    for (int g = 0; g < c.ngp; ++g) {
        const double *r = c.rw[g];
        for (int k = 0; k < 4; ++k) rw[g * 4 + k] = r[k];
        for (int l = 0; l < c.mn; ++l) {
            N[g * c.mn + l] = c.shape.nf_ln(l + 1, r[0], r[1], r[2]);
            for (int d = 0; d < 3; ++d) dN[(g * c.mn + l) * 3 + d] = c.shape.nf_dln_dxi(d + 1, l + 1, r[0], r[1], r[2]);
        }
        for (int e = 0; e < c.me; ++e) {
            phi[g * c.me + e] = c.shape.mix_ln(c.enode[e], c.edir[e], r[0], r[1], r[2]);
            for (int d = 0; d < 3; ++d)
                dphi[(g * c.me + e) * 3 + d] = c.shape.mix_dln_dxi(c.edir[e], d + 1, c.enode[e], r[0], r[1], r[2]);
        }
    }
    return 0;
}

// raw shape-function evaluation at an arbitrary point (derivative / partition-of-unity self-tests)
int oracle_shape_eval(int mn, int me, double xi, double eta, double zeta, double *N, double *dN, double *phi, double *dphi) {
    Shape s(mn);
    int en[54], ed[54];
    edge_dir_table(me, en, ed);
    for (int l = 0; l < mn; ++l) {
        N[l] = s.nf_ln(l + 1, xi, eta, zeta);
        for (int d = 0; d < 3; ++d) dN[l * 3 + d] = s.nf_dln_dxi(d + 1, l + 1, xi, eta, zeta);
    }
    for (int e = 0; e < me; ++e) {
        phi[e] = s.mix_ln(en[e], ed[e], xi, eta, zeta);
        for (int d = 0; d < 3; ++d) dphi[e * 3 + d] = s.mix_dln_dxi(ed[e], d + 1, en[e], xi, eta, zeta);
    }
    return 0;
}

/*
 * One element: full me x me alocal matrix (row-major [im][jm], complex), blocal (me x 2) and the
 * integration caches.  pml = the in_pml flags in force while the element is integrated.
 */
int oracle_element(const oracle_ctx *h, int ide, double omega, const double *g_sigma, const int *pml,
                   double *Ae, double *be, double *wgt, double *cve1, double *cve2, double *ve,
                   double *mf1, double *mf2, double *gpml, double *src) {
    const Ctx &c = *h->c;
    Elem *E = new Elem;
    elem_setup(*E, c, omega, (const C *)g_sigma, 0);
    elem_compute(*E, ide, pml);
    int st = E->status;
    if (!st) {
        const int me = c.me, ngp = c.ngp;
        if (Ae)
            for (int im = 1; im <= me; ++im)
                for (int jm = 1; jm <= me; ++jm) {
                    const C a = alocal(*E, im, jm);
                    Ae[2 * ((im - 1) * me + jm - 1)] = a.re; Ae[2 * ((im - 1) * me + jm - 1) + 1] = a.im;
                }
        if (be)
            for (int im = 1; im <= me; ++im) {
                C b[2];
                blocal(*E, im, b);
                be[4 * (im - 1) + 0] = b[0].re; be[4 * (im - 1) + 1] = b[0].im;
                be[4 * (im - 1) + 2] = b[1].re; be[4 * (im - 1) + 3] = b[1].im;
            }
        if (wgt) std::memcpy(wgt, E->wgt, sizeof(double) * ngp);
        if (cve1) std::memcpy(cve1, E->cve1, sizeof(double) * 3 * ngp * me);
        if (cve2) std::memcpy(cve2, E->cve2, sizeof(double) * 3 * ngp * me);
        if (ve) std::memcpy(ve, E->ve, sizeof(double) * 3 * ngp * me);
        if (mf1) std::memcpy(mf1, E->mf1, sizeof(C) * 6 * ngp);
        if (mf2) std::memcpy(mf2, E->mf2, sizeof(C) * 6 * ngp);
        if (gpml) std::memcpy(gpml, E->gpml, sizeof(C) * 3 * ngp);
        if (src) std::memcpy(src, E->src, sizeof(C) * 3 * 2 * ngp);
    }
    delete E;
    return st;
}

/*
 * Whole assembly for one frequency: MoVFEM_3DMT.f90:167-216 (global_vfem) + 85-97.
 *   faithful = 1 : reference loop structure incl. every redundant Jacobian rebuild (timed baseline)
 *   faithful = 0 : identical arithmetic, Jacobian memoised per Gauss point (same bits, faster)
 *   nthreads > 1 : element matrices computed in parallel, scattered serially in element order
 *                  (same bits as nthreads = 1); only with faithful = 0
 *   ide_lo..ide_hi: element sub-range to assemble (1-based inclusive; 0,0 = all) -- used to time a
 *                  bounded sample of a big mesh; the pattern is still the full one.
 * Outputs (any may be NULL):
 *   a_t1[nnze]       tap T1: values after the element loop, pattern order (row-major sorted)
 *   irn,jcn,a,nz     tap T2: after ga_sort_sparse (Q10) and find_zeros/rem_zeros (Q11)
 *   rhs[2*nne]
 */
int oracle_assemble(oracle_ctx *h, double omega, const double *g_sigma_, int faithful, int nthreads,
                    int ide_lo, int ide_hi, double *a_t1, int32_t *irn, int32_t *jcn, double *a_out,
                    int64_t *nz_out, double *rhs, double *seconds_elements, int64_t *jac_builds) {
    Ctx &c = *h->c;
    const C *g_sigma = (const C *)g_sigma_;
    const int me = c.me, ne = c.ne, nne = c.nne;
    const int64_t nnze = c.nnze;
    auto G = [&](int ide, int im) -> int { return c.gne[(size_t)(im - 1) * ne + (ide - 1)]; };
    if (ide_lo <= 0) { ide_lo = 1; ide_hi = ne; }
    // ga_assemble_nze: ia/ja per pattern slot = pat_ia/pat_ja (consistency checks of
    // global_assembly.f90:53-58,104-112 hold by construction of pat_find)
    std::vector<C> a((size_t)nnze, mk(0, 0));
    std::vector<C> b((size_t)2 * nne, mk(0, 0));
    int status = 0;
    long jb = 0;
    int first_flags[3] = {c.in_pml[0], c.in_pml[1], c.in_pml[2]};
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);

    BdModel BD;
    const bool bd_vals = c.d.dirichlet && c.d.bd_inimod >= 2;
    if (bd_vals) bd_model(c, omega, BD);
    auto scatter = [&](Elem &E, int ide, const C *Al /*me*me or null*/, const C *Bl, const C *Fb /*me*2 or null*/) {
        for (int im = 1; im <= me; ++im) {
            if (G(ide, im) < 0) continue;
            for (int jm = 1; jm <= me; ++jm) {
                if (G(ide, jm) < 0) continue;
                if (c.d.sym && G(ide, im) < G(ide, jm)) continue;
                const int64_t idd = pat_find(c, G(ide, im), G(ide, jm));
                const C v = Al ? Al[(im - 1) * me + jm - 1] : alocal(E, im, jm);
                a[idd] = a[idd] + v;
            }
        }
        for (int im = 1; im <= me; ++im) {
            if (G(ide, im) < 0) continue;
            C bda[2] = {mk(0, 0), mk(0, 0)};
            for (int jm = 1; jm <= me; ++jm) {
                if (G(ide, jm) >= 0) continue;
                // f_boundary == (0,0) for bd_inimod=1 (boundary_conds.f90:200-214); the reference
                // still evaluates alocal(im,jm) here
                C fbv[2] = {mk(0, 0), mk(0, 0)};
                if (bd_vals) {
                    if (Fb) { fbv[0] = Fb[2 * (jm - 1)]; fbv[1] = Fb[2 * (jm - 1) + 1]; }
                    else f_boundary(E, BD, -G(ide, jm), jm, fbv);
                }
                const C al = Al ? Al[(im - 1) * me + jm - 1] : alocal(E, im, jm);
                bda[0] = bda[0] + fbv[0] * al; bda[1] = bda[1] + fbv[1] * al;
            }
            C bl[2];
            if (Bl) { bl[0] = Bl[2 * (im - 1)]; bl[1] = Bl[2 * (im - 1) + 1]; }
            else blocal(E, im, bl);
            for (int d = 0; d < 2; ++d) {
                const size_t idd = (size_t)G(ide, im) - 1 + (size_t)d * nne;
                b[idd] = b[idd] + (bl[d] - bda[d]);
            }
        }
    };

    if (faithful || nthreads <= 1) {
        Elem *E = new Elem;
        elem_setup(*E, c, omega, g_sigma, faithful);
        for (int ide = ide_lo; ide <= ide_hi && !status; ++ide) {
            int pml[3];
            effective_pml(c, ide, first_flags, pml);
            elem_compute(*E, ide, pml);
            if (E->status) { status = E->status; break; }
            scatter(*E, ide, nullptr, nullptr, nullptr);
        }
        jb = E->jac_builds;
        delete E;
    } else {
        const int CH = 256;
        std::vector<C> Al((size_t)CH * me * me), Bl((size_t)CH * me * 2), Fb((size_t)CH * me * 2);
        for (int base = ide_lo; base <= ide_hi && !status; base += CH) {
            const int n = std::min(CH, ide_hi - base + 1);
#pragma omp parallel num_threads(nthreads)
            {
                Elem *E = new Elem;
                elem_setup(*E, c, omega, g_sigma, 0);
#pragma omp for schedule(dynamic, 4)
                for (int t = 0; t < n; ++t) {
                    const int ide = base + t;
                    int pml[3];
                    effective_pml(c, ide, first_flags, pml);
                    elem_compute(*E, ide, pml);
                    if (E->status) {
#pragma omp critical
                        status = E->status;
                        continue;
                    }
                    for (int im = 1; im <= me; ++im) {
                        const bool vi = G(ide, im) >= 0;
                        for (int jm = 1; jm <= me; ++jm) {
                            const bool need = vi && ((G(ide, jm) >= 0 && !(c.d.sym && G(ide, im) < G(ide, jm))) || (bd_vals && G(ide, jm) < 0));
                            Al[((size_t)t * me + im - 1) * me + jm - 1] = need ? alocal(*E, im, jm) : mk(0, 0);
                        }
                        if (vi) blocal(*E, im, &Bl[((size_t)t * me + im - 1) * 2]);
                    }
                    if (bd_vals)
                        for (int jm = 1; jm <= me; ++jm)
                            if (G(ide, jm) < 0) f_boundary(*E, BD, -G(ide, jm), jm, &Fb[((size_t)t * me + jm - 1) * 2]);
                }
                delete E;
            }
            if (status) break;
            Elem dummy;
            dummy.c = &c;
            for (int t = 0; t < n; ++t) scatter(dummy, base + t, &Al[(size_t)t * me * me], &Bl[(size_t)t * me * 2], &Fb[(size_t)t * me * 2]);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds_elements) *seconds_elements = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if (jac_builds) *jac_builds = jb;
    if (status) return status;
    // the element loop ended on the last element: its get_pml call leaves these flags behind (Q17)
    if (!c.d.dirichlet && ide_hi == ne) get_pml(c, c.nx, c.ny, c.nz, c.in_pml);

    if (a_t1) std::memcpy(a_t1, a.data(), sizeof(C) * nnze);
    if (rhs) std::memcpy(rhs, b.data(), sizeof(C) * 2 * nne);

    if (irn && jcn && a_out && nz_out) {
        // ga_sort_sparse, global_assembly.f90:152-181, literally (Q10)
        std::vector<int> ia(c.pat_ia), ja(c.pat_ja);
        std::vector<int64_t> q((size_t)nnze);
        std::iota(q.begin(), q.end(), (int64_t)0);
        std::stable_sort(q.begin(), q.end(), [&](int64_t x, int64_t y) { return ja[x] < ja[y]; });   // merge_sort(ja,...)
        std::vector<int> tempi(ia), tempj(ja);
        std::vector<float> tare((size_t)nnze), taim((size_t)nnze);                                  // complex (single) tempa
        for (int64_t k = 0; k < nnze; ++k) { tare[k] = (float)a[k].re; taim[k] = (float)a[k].im; }
        for (int64_t k = 0; k < nnze; ++k) { ja[k] = tempj[q[k]]; ia[k] = tempi[q[k]]; a[k] = mk((double)tare[q[k]], (double)taim[q[k]]); }
        std::iota(q.begin(), q.end(), (int64_t)0);
        std::stable_sort(q.begin(), q.end(), [&](int64_t x, int64_t y) { return ia[x] < ia[y]; });   // merge_sort(ia,...)
        tempi = ia; tempj = ja;
        for (int64_t k = 0; k < nnze; ++k) { tare[k] = (float)a[k].re; taim[k] = (float)a[k].im; }
        for (int64_t k = 0; k < nnze; ++k) { ja[k] = tempj[q[k]]; ia[k] = tempi[q[k]]; a[k] = mk((double)tare[k], (double)taim[k]); }   // a=tempa: NOT permuted
        // find_zeros / rem_zeros, global_assembly.f90:123-150
        int64_t n = 0;
        for (int64_t k = 0; k < nnze; ++k) {
            if (a[k].re == 0.0 && a[k].im == 0.0) continue;
            irn[n] = ia[k]; jcn[n] = ja[k]; a_out[2 * n] = a[k].re; a_out[2 * n + 1] = a[k].im;
            ++n;
        }
        *nz_out = n;
    }
    return 0;
}

// Q17 state handling for tests: read / set the SAVEd in_pml flags
void oracle_get_in_pml(const oracle_ctx *h, int *f) { f[0] = h->c->in_pml[0]; f[1] = h->c->in_pml[1]; f[2] = h->c->in_pml[2]; }
void oracle_set_in_pml(oracle_ctx *h, const int *f) { h->c->in_pml[0] = f[0]; h->c->in_pml[1] = f[1]; h->c->in_pml[2] = f[2]; }
void oracle_effective_pml(const oracle_ctx *h, int ide, int *f) {
    int first[3] = {h->c->in_pml[0], h->c->in_pml[1], h->c->in_pml[2]};
    effective_pml(*h->c, ide, first, f);
}

/*
 * Post-processing of a solved system (TEST INFRASTRUCTURE for the north_star's end-to-end criterion):
 * solution.f90:18-69 node_solution / z_rho_phi, :207-256 get_elem_sol (pe_sch=1), :304-343 get_er, :370-413 get_hr,
 * :440-505 hor_fields / get_impedance / get_res_phase / inv_matrix.
 *   x[2*nne]        solution of both polarisations (column d at offset (d-1)*nne), complex
 *   esol,hsol[2*npt][3] total E and H at the grid nodes (row idd-1 = node + (edir-1)*npt), complex
 *   z[npt][4] complex, rho[npt][4], phi[npt][4]
 */
int oracle_node_solution(const oracle_ctx *h, double omega, const double *g_sigma_, const double *x_, double *esol_, double *hsol_,
                         double *z_, double *rho_, double *phi_) {
    const Ctx &c = *h->c;
    const C *x = (const C *)x_;
    C *esol = (C *)esol_, *hsol = (C *)hsol_, *z = (C *)z_;
    const int g = c.d.nord, npt = c.npt, nne = c.nne, ne = c.ne, me = c.me, mn = c.mn;
    auto G = [&](int ide, int im) -> int { return c.gne[(size_t)(im - 1) * ne + (ide - 1)]; };
    std::vector<char> valued_e((size_t)2 * npt, 0), valued_h((size_t)2 * npt, 0);
    BdModel BD;
    const bool bd_vals = c.d.dirichlet && c.d.bd_inimod >= 2;
    if (bd_vals) bd_model(c, omega, BD);
    for (size_t i = 0; i < (size_t)6 * npt; ++i) { esol[i] = mk(0, 0); hsol[i] = mk(0, 0); }
    double as[3];
    for (int j = 1; j <= g; ++j) as[j - 1] = -1 + (double)(j - 1) * (2 / (double)(g - 1));   // gqg_nodes, geometry.f90:720-739
    Elem *E = new Elem;
    elem_setup(*E, c, omega, (const C *)g_sigma_, 0);
    for (int ie = 1; ie <= c.nx; ++ie)
        for (int je = 1; je <= c.ny; ++je)
            for (int ke = 1; ke <= c.nz; ++ke) {
                const int eno = (ie - 1) * c.nyz * (g - 1) + (je - 1) * c.nnz * (g - 1) + (ke - 1) * (g - 1) + 1;
                const int ide = (ie - 1) * c.ny * c.nz + (je - 1) * c.nz + ke;
                nf_get_r(*E, ie, je, ke, eno);
                p_elem_fields(*E);
                if (E->status) { const int st = E->status; delete E; return st; }
                for (int edir = 1; edir <= 2; ++edir) {
                    // get_elem_sol, pe_sch = 1: E first (secondary + primary), then H from the nodal E of this element
                    for (int i1 = 1; i1 <= g; ++i1)
                        for (int j1 = 1; j1 <= g; ++j1)
                            for (int k1 = 1; k1 <= g; ++k1) {
                                const int idd = eno + (i1 - 1) * c.nyz + (j1 - 1) * c.nnz + (k1 - 1) + (edir - 1) * npt;
                                if (valued_e[idd - 1]) continue;
                                const double r[3] = {as[i1 - 1], as[j1 - 1], as[k1 - 1]};
                                C ef[3] = {mk(0, 0), mk(0, 0), mk(0, 0)};            // get_er
                                for (int im = 1; im <= me; ++im) {
                                    C f = mk(0, 0);
                                    if (G(ide, im) >= 0) f = x[(size_t)G(ide, im) - 1 + (size_t)(edir - 1) * nne];
                                    else if (bd_vals) { C fbv[2]; f_boundary(*E, BD, -G(ide, im), im, fbv); f = fbv[edir - 1]; }
                                    double ve[3];
                                    vf_elem_ve(*E, im, r, ve);
                                    ef[0] = ef[0] + f * ve[0]; ef[1] = ef[1] + f * ve[1]; ef[2] = ef[2] + f * ve[2];
                                }
                                C ep[3] = {mk(0, 0), mk(0, 0), mk(0, 0)};            // get_ep, problem.f90:151-168
                                for (int i = 1; i <= mn; ++i) {
                                    const int j = i + (edir - 1) * mn;
                                    const double ln = c.shape.nf_ln(i, r[0], r[1], r[2]);
                                    for (int m = 0; m < 3; ++m) ep[m] = ep[m] + E->pe_ep[j - 1][m] * ln;
                                }
                                for (int m = 0; m < 3; ++m) esol[(size_t)(idd - 1) * 3 + m] = ef[m] + ep[m];
                                valued_e[idd - 1] = 1;
                            }
                    for (int i1 = 1; i1 <= g; ++i1)
                        for (int j1 = 1; j1 <= g; ++j1)
                            for (int k1 = 1; k1 <= g; ++k1) {
                                const int idd = eno + (i1 - 1) * c.nyz + (j1 - 1) * c.nnz + (k1 - 1) + (edir - 1) * npt;
                                if (valued_h[idd - 1]) continue;
                                const double r[3] = {as[i1 - 1], as[j1 - 1], as[k1 - 1]};
                                C ef[3] = {mk(0, 0), mk(0, 0), mk(0, 0)};            // get_hr
                                for (int im = 1; im <= mn; ++im) {
                                    const C *f2 = esol + ((size_t)E->nf_index[im - 1] - 1 + (size_t)(edir - 1) * npt) * 3;
                                    double dn[3];
                                    nf_grad_ln(*E, im, r, dn);
                                    ef[0] = ef[0] + (f2[2] * dn[1] - f2[1] * dn[2]);
                                    ef[1] = ef[1] + (f2[0] * dn[2] - f2[2] * dn[0]);
                                    ef[2] = ef[2] + (f2[1] * dn[0] - f2[0] * dn[1]);
                                }
                                C mf1[6];
                                for (int k = 0; k < 6; ++k) mf1[k] = mk(0, 0);
                                for (int i = 1; i <= mn; ++i) {                      // p_intmodels, problem.f90:139-142
                                    const double ln = c.shape.nf_ln(i, r[0], r[1], r[2]);
                                    for (int k = 0; k < 6; ++k) mf1[k] = mf1[k] + mk(ln * E->pe_inmu[i - 1][k], 0.0);
                                }
                                C md[6];
                                for (int k = 0; k < 6; ++k) md[k] = cmplx32(0.0, -1.0 / omega) * mf1[k];   // cmplx() without KIND (Q2)
                                C *ho = hsol + (size_t)(idd - 1) * 3;
                                ho[0] = md[0] * ef[0] + md[1] * ef[1] + md[2] * ef[2];
                                ho[1] = md[1] * ef[0] + md[3] * ef[1] + md[4] * ef[2];
                                ho[2] = md[2] * ef[0] + md[4] * ef[1] + md[5] * ef[2];
                                valued_h[idd - 1] = 1;
                            }
                }
            }
    delete E;
    // z_rho_phi
    const double mu0 = 4 * PI * 1.e-7;
    for (int id = 0; id < npt; ++id) {
        const C e[4] = {esol[(size_t)id * 3], esol[((size_t)id + npt) * 3], esol[(size_t)id * 3 + 1], esol[((size_t)id + npt) * 3 + 1]};
        const C hh[4] = {hsol[(size_t)id * 3], hsol[((size_t)id + npt) * 3], hsol[(size_t)id * 3 + 1], hsol[((size_t)id + npt) * 3 + 1]};
        const C det = hh[0] * hh[3] - hh[1] * hh[2];
        const C ih[4] = {hh[3] / det, -hh[1] / det, -hh[2] / det, hh[0] / det};
        C zz[4];
        zz[0] = ih[0] * e[0] + ih[2] * e[1];
        zz[1] = ih[1] * e[0] + ih[3] * e[1];
        zz[2] = ih[0] * e[2] + ih[2] * e[3];
        zz[3] = ih[1] * e[2] + ih[3] * e[3];
        for (int i = 0; i < 4; ++i) {
            z[(size_t)id * 4 + i] = zz[i];
            double rho = (1.0 / (omega * mu0)) * (zz[i].re * zz[i].re + zz[i].im * zz[i].im), phi = 0.0;
            if (rho < 1.e-2) rho = 0.0;
            else phi = std::atan(zz[i].im / zz[i].re) * (180. / PI);
            rho_[(size_t)id * 4 + i] = rho; phi_[(size_t)id * 4 + i] = phi;
        }
    }
    return 0;
}

}  // extern "C"

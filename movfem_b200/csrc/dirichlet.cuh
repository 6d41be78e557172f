// movfem_b200/csrc/dirichlet.cuh -- non-zero Dirichlet boundary values moved to the right-hand side.
//
// Replaces, for boundary models 2 (homogeneous earth) and 3 (layered earth):
//   MoVFEM_3DMT.f90:252-261      bda = sum_jm f_boundary(-gne(jm),jm) * alocal(im,jm) ;  b += blocal(im) - bda
//   boundary_conds.f90:188-250   f_boundary: primary E at the edge's node dotted with dr/dxi_d (faces 1,2,4,5; 0 on 3,6)
//   boundary_conds.f90:436-518   the module's p_pfields ; :523-552 wait_recursion ; :557-598 ezl
//   boundary_conds.f90:392-430   bd_setmodel (+ bd_updatemodel :48-50 every frequency)
// The per-frequency model constants (pe_psigma, cz, ez: O(layers) scalars) are formed on the host in bd_model_host
// exactly as the oracle does; the per-node fields and the products with A_e = K_e + i*f32(omega)*M_e (already in the
// K/M store for EVERY local pair, Dirichlet columns included) run here, one thread per (boundary element, local DOF).
// Units and sign conventions are the reference's (depths in km against layer coordinates <= 0); restated, not repaired.
#pragma once
#include <complex>

#include "common.cuh"

namespace movfem {

struct BdModelDev {
    int inimod, nl;
    double omega, pmu, g_ztop, w32;
    double2 psig[17], cz[17], ez[17];
    double zl[17];
};

struct BdTables {
    double dNn[kMaxMn][kMaxMn][3];   // dN_l/dxi_m at the reference coordinates of node i: [i][l][m]
    int enode[kMaxMep], edir[kMaxMep];   // v_fem.f90:491-504, 0-based node / direction of each local DOF
    int node_off[kMaxMn], node_i[kMaxMn], node_j[kMaxMn];
};

__host__ __device__ __forceinline__ double2 zmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ double2 zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 zscale(double r, double2 a) { return make_double2(r * a.x, r * a.y); }
__host__ __device__ __forceinline__ double2 zdiv(double2 a, double2 b) {   // Fortran rules (Smith), as the oracle
    if (fabs(b.x) < fabs(b.y)) {
        const double ratio = b.x / b.y, div = (b.x * ratio) + b.y;
        return make_double2(((a.x * ratio) + a.y) / div, ((a.y * ratio) - a.x) / div);
    }
    const double ratio = b.y / b.x, div = (b.y * ratio) + b.x;
    return make_double2(((a.y * ratio) + a.x) / div, (a.y - (a.x * ratio)) / div);
}
__device__ __forceinline__ double2 zsqrt(double2 a) {   // principal square root
    const double r = hypot(a.x, a.y);
    if (r == 0.0) return make_double2(0.0, 0.0);
    double re, im;
    if (a.x >= 0.0) { re = sqrt(0.5 * (r + a.x)); im = a.y / (2.0 * re); }
    else { im = copysign(sqrt(0.5 * (r - a.x)), a.y); re = a.y / (2.0 * im); }
    return make_double2(re, im);
}
__device__ __forceinline__ double2 zexp(double2 a) {
    double s, c;
    sincos(a.y, &s, &c);
    const double e = exp(a.x);
    return make_double2(e * c, e * s);
}

// host: bd_setmodel / bd_updatemodel / wait_recursion with the oracle's operation order
inline void bd_model_host(const movfem_desc &d, double omega, BdModelDev &B) {
    typedef std::complex<double> Z;
    auto zd = [](Z a) { return make_double2(a.real(), a.imag()); };
    auto dz = [](double2 a) { return Z(a.x, a.y); };
    auto cdiv = [&](Z a, Z b) { return dz(zdiv(zd(a), zd(b))); };
    auto cmul = [&](Z a, Z b) { return dz(zmul(zd(a), zd(b))); };
    B.inimod = d.bd_inimod; B.omega = omega; B.pmu = 4.0 * kPi * 1.e-7; B.g_ztop = d.g_ztop; B.w32 = f32r(omega);
    const double im = f32r(kEps0 * omega);
    if (d.bd_inimod == 2) {
        B.nl = 2;
        B.psig[0] = make_double2(0.0, im);
        B.psig[1] = make_double2(f32r(d.bd_hsigma), im);
        return;
    }
    const int nl = B.nl = d.bd_nl;
    double dl[16];
    for (int l = 0; l < nl; ++l) B.psig[l] = make_double2(f32r(d.bd_lsigma[l]), im);
    for (int l = 0; l < nl - 1; ++l) dl[l] = d.bd_ldz[l];
    B.zl[0] = 0.0;
    for (int l = 1; l < nl; ++l) B.zl[l] = B.zl[l - 1] - dl[l - 1];
    const Z I(0.0, 1.0), one(1.0, 0.0);
    std::vector<Z> cz(nl), ez(nl);
    auto gam = [&](int l) { return std::sqrt(cmul(I, (omega * B.pmu) * dz(B.psig[l]))); };
    Z gamma = gam(nl - 1);
    cz[nl - 1] = cdiv(one, gamma);
    for (int l = nl - 1; l >= 1; --l) {
        gamma = gam(l - 1);
        const Z gc = cmul(gamma, cz[l]);
        const Z r = cdiv(one + (-gc), one + gc);
        const Z e2 = std::exp((-2.0 * gamma) * dl[l - 1]);
        const Z re2 = cmul(r, e2);
        cz[l - 1] = cdiv(one + (-re2), cmul(gamma, one + re2));
    }
    ez[0] = one;
    for (int l = 1; l <= nl - 1; ++l) {
        gamma = gam(l - 1);
        const Z num = cmul(cmul(ez[l - 1], cz[l]), one + cmul(cz[l - 1], gamma));
        const Z den = cmul(cz[l - 1], one + cmul(cz[l], gamma));
        ez[l] = cdiv(cmul(std::exp((-gamma) * dl[l - 1]), num), den);
    }
    for (int l = 0; l < nl; ++l) { B.cz[l] = zd(cz[l]); B.ez[l] = zd(ez[l]); }
}

// ezl(0, z), boundary_conds.f90:566-580
__device__ inline double2 bd_ezl0(const BdModelDev &B, double z) {
    double2 out = make_double2(0.0, 0.0);
    const double2 I = make_double2(0.0, 1.0), one = make_double2(1.0, 0.0);
    for (int l = 1; l <= B.nl - 1; ++l) {
        if (z <= B.zl[l - 1] && z > B.zl[l]) {
            const double2 gamma = zsqrt(zmul(I, zscale(B.omega * B.pmu, B.psig[l - 1])));
            const double2 gc = zmul(gamma, B.cz[l]);
            const double2 r = zdiv(make_double2(1.0 - gc.x, -gc.y), zadd(one, gc));
            const double2 e2 = zexp(zscale(z - B.zl[l], zscale(-2.0, gamma))), e1 = zexp(zscale(B.zl[l - 1] - z, make_double2(-gamma.x, -gamma.y)));
            const double2 re2 = zmul(r, e2);
            double2 t = zmul(zscale(0.5, B.ez[l - 1]), zadd(one, zdiv(one, zmul(B.cz[l - 1], gamma))));
            t = zmul(t, make_double2(1.0 - re2.x, -re2.y));
            out = zmul(t, e1);
        }
    }
    if (z <= B.zl[B.nl - 1]) {
        const double2 gamma = zsqrt(zmul(I, zscale(B.omega * B.pmu, B.psig[B.nl - 1])));
        out = zmul(B.ez[B.nl - 1], zexp(zscale(B.zl[B.nl - 1] - z, make_double2(-gamma.x, -gamma.y))));
    }
    return out;
}

// pe_ep(1,1), pe_ep(2,2) at a node of height zn
__device__ inline void bd_pfields(const BdModelDev &B, double zn, double2 &ex1, double2 &ey2) {
    const double bb0 = 1.e-9;
    const double z = (B.g_ztop - zn) / 1000.0;
    const double2 I = make_double2(0.0, 1.0), one = make_double2(1.0, 0.0);
    const double2 cpos = make_double2(0.0, f32r(2.0 * B.omega * bb0)), cneg = make_double2(0.0, f32r(-2.0 * B.omega * bb0));
    if (B.inimod == 2) {
        double2 fp;
        if (zn > B.g_ztop) {
            const double2 sq = zsqrt(zmul(I, zscale(B.omega * B.pmu, B.psig[0])));
            const double2 zs = zscale(z, sq);
            fp = zmul(zdiv(one, sq), make_double2(1.0 - zs.x, -zs.y));
        } else {
            const double2 sq = zsqrt(zmul(I, zscale(B.omega * B.pmu, B.psig[1])));
            fp = zmul(zdiv(one, sq), zexp(zscale(-z, sq)));
        }
        ex1 = zmul(cneg, fp);
        ey2 = zmul(cpos, fp);
    } else {
        if (zn > B.g_ztop) {
            const double2 t = make_double2(B.cz[0].x - z, B.cz[0].y);
            ex1 = zmul(cpos, t);
            ey2 = zmul(cneg, t);
        } else {
            const double2 e = bd_ezl0(B, z);
            ex1 = zmul(zmul(cpos, B.cz[0]), e);
            ey2 = zmul(zmul(cneg, B.cz[0]), e);
        }
    }
}

// one thread per (side-face element of the list, local DOF im): be[e][im] -= sum over Dirichlet columns jm
__global__ void dirichlet_rhs_kernel(int nlist, const int *__restrict__ list, MeshDims m, BdModelDev B, const BdTables *__restrict__ Tp,
                                     const int *__restrict__ gne, const int *__restrict__ kmrow, int e_base, int NP,
                                     const double2 *__restrict__ KM, const double *__restrict__ xp, const double *__restrict__ yp,
                                     const double *__restrict__ zp, double *__restrict__ be) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlist * m.me) return;
    const int e = list[t / m.me], im = t % m.me;
    if (gne[(size_t)im * m.ne + e] < 0) return;
    const BdTables &T = *Tp;
    int ie, je, ke;
    elem_ijk(m, e, ie, je, ke);
    const int g1 = m.nord - 1;
    const int64_t id0 = (int64_t)(ie - 1) * g1 * m.nyz + (int64_t)(je - 1) * g1 * m.nnz + (ke - 1) * g1;
    const int kr = kmrow[e - e_base];
    const double2 *KMe = KM + (((size_t)(kr >> 5) * NP) << 5) + (kr & 31);
    double2 bda0 = make_double2(0.0, 0.0), bda1 = make_double2(0.0, 0.0);
    for (int jm = 0; jm < m.me; ++jm) {
        const int g = gne[(size_t)jm * m.ne + e];
        if (g >= 0 || g == -3 || g == -6) continue;       // f_boundary is zero on the bottom / top faces
        const int i = T.enode[jm], d = T.edir[jm];
        double dne0 = 0.0, dne1 = 0.0;                    // row d of the Jacobian at node i (x and y components are all f needs)
        for (int l = 0; l < m.mn; ++l) {
            const double dn = T.dNn[i][l][d];
            dne0 = dne0 + dn * xp[(ie - 1) * g1 + T.node_i[l]];
            dne1 = dne1 + dn * yp[(je - 1) * g1 + T.node_j[l]];
        }
        double2 ex1, ey2;
        bd_pfields(B, zp[id0 + T.node_off[i]], ex1, ey2);
        const int hi = im > jm ? im : jm, lo = im > jm ? jm : im;
        const double2 km = KMe[(size_t)(hi * (hi + 1) / 2 + lo) << 5];
        const double2 al = make_double2(km.x, B.w32 * km.y);
        bda0 = zadd(bda0, zmul(zscale(dne0, ex1), al));
        bda1 = zadd(bda1, zmul(zscale(dne1, ey2), al));
    }
    double4 *bo = reinterpret_cast<double4 *>(be) + be_index(kr, m.me, im);
    double4 v = *bo;
    v.x -= bda0.x; v.y -= bda0.y; v.z -= bda1.x; v.w -= bda1.y;
    *bo = v;
}

}  // namespace movfem

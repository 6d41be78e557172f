// movfem_b200/csrc/common.cuh -- shared device/host structures of the assembly library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace movfem {

// ---------------------------------------------------------------------------------------------
// Arithmetic policy.  The library is compiled with -fmad=false: every `a*b+c` written with plain
// operators is an IEEE multiply followed by an IEEE add, exactly like the reference built with
// `gfortran -O` on x86-64.  The Jacobian, its inverse, det, the quadrature weight and the GPML
// coordinate are written that way, in the reference's evaluation order, so they carry the
// reference's bits (and with them its exact zeros and round-off residues, which decide what
// rem_zeros strips).  Fused multiply-adds are used only where asked for explicitly (dfma below):
// material interpolation, the per-Gauss-point tensors, the contractions and the RHS -- the
// FP64-throughput-critical part, which agrees with the reference to rounding (<= 1e-12, north_star).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }

// cmplx(x,y) without KIND: single precision
__host__ __device__ __forceinline__ double f32r(double x) { return (double)(float)x; }

constexpr double kPi = 3.1415926535897932384626433;   // geometry.f90:25
constexpr double kEps0 = 8.854187817e-12;              // geometry.f90:25
constexpr double kMu0 = 4.0 * kPi * 1.0e-7;            // geometry.f90:26  mu_0=4.d0*pi*1.d-7
constexpr double kB0 = 1.e-9;                          // problem.f90:27

// Per grid node, rebuilt every frequency by node_kernel (replaces the per-element recomputation
// of problem.f90:70-87 p_elem_fields: each node is shared by up to 8 elements).
struct __align__(16) NodeRec {
    double z;        // g_zp
    double e;        // f32(omega*b0*z): |Ep| of the primary plane wave (problem.f90:352,355)
    double inmu[6];  // mu^-1, cofactor/det (problem.f90:279-288)
    double sre[6];   // Re g_sigma
    double sim[6];   // Im g_sigma
    double vc[6];    // (mu^-1 dmu)(:,2)*Hp and (mu^-1 dmu)(:,1)*Hp  (problem.f90:391-403), pol 1 / pol 2
};
static_assert(sizeof(NodeRec) == 26 * 8, "NodeRec layout");
constexpr int kNodeDoubles = 26;

// Element right-hand sides b_e are stored by K/M row (position in the launch lists, see contract.cuh), 32 rows interleaved:
// [row / 32][local DOF][row % 32] of (pol 1 re, im, pol 2 re, im) -- a warp whose lanes are 32 consecutive rows stores 1 kB runs
__host__ __device__ __forceinline__ size_t be_index(int64_t kr, int me, int dof) { return ((size_t)(kr >> 5) * me + dof) * 32 + (size_t)(kr & 31); }

constexpr int kMaxGp = 27, kMaxMn = 27, kMaxMe = 54, kMaxMep = 56, kMaxSlots = 60;

// Reference-element tables at the Gauss points (global memory, read through L1).
struct ElemTables {
    double rw[kMaxGp][4];                   // integration.f90:267-279 i_rw
    double N[kMaxGp][kMaxMn];               // nf_ln
    double dN[kMaxGp][kMaxMn][3];           // nf_dln_dxi
    double phi[kMaxGp][kMaxMep];            // mix_ln of the DOF's (node,dir)
    double dphi[kMaxGp][kMaxMep][3];        // mix_dln_dxi
    int node_off[kMaxMn];                   // linear grid offset of local node from element base node
    int node_i[kMaxMn], node_j[kMaxMn];     // i1-1, j1-1 (n_fem.f90:36-59)
    int edir[kMaxMep];                      // direction of DOF, 0-based (v_fem.f90:491-504)
    // DOFs permuted into direction-uniform groups of four (padded with -1): the contraction's slot order
    int slot_dof[kMaxSlots];                // slot -> local DOF (0-based) or -1
    int slot_dir[kMaxSlots];                // slot -> direction (0-based), defined for padding slots too
    int nslots;
    double dNt[kMaxMn * 4 * 32];            // [l][4][32]: dN/dxi_m (m=0..2) and N (3) with the Gauss point fastest
};

// DOF sharing tables (global_assembly.f90:242-265,310-351,396-443) and Dirichlet face lists
// (boundary_conds.f90:276-388), 1-based local DOF ids, 0 = none.
struct ShareTables {
    uint8_t back[3][kMaxMe + 1];   // back[a][mine]  = theirs in the -x(0) / -y(1) / -z(2) neighbour
    uint8_t fwd[3][kMaxMe + 1];    // fwd[a][theirs] = mine in the +x / +y / +z neighbour
    uint8_t face[kMaxMe + 1];      // bit f set: DOF lies on face f+1 (reference order ie=1,je=1,ke=1,ie=nx,je=ny,ke=nz)
};

struct MeshDims {
    int nx, ny, nz;            // elements per axis
    int nnx, nny, nnz, nyz;    // node lines, g_nyz
    int nord, mn, me, ngp;
    int ne, npt;
    int dirichlet;
};

// GPML state of boundary_conds.f90:51-70
struct PmlParams {
    int sch;
    double a0, b0, nn;
    double a[3][2], b[3][2];   // xa,ya,za / xb,yb,zb
    double omegar[2];
    int el_a[3][2], el_b[3][2];
    int first[3];              // in_pml flags element (1,1,1) sees
};

__host__ __device__ __forceinline__ void elem_ijk(const MeshDims &m, int e, int &ie, int &je, int &ke) {
    ke = e % m.nz + 1;
    je = (e / m.nz) % m.ny + 1;
    ie = e / (m.nz * m.ny) + 1;
}

// boundary_conds.f90:72-82 get_pml
__host__ __device__ __forceinline__ void get_pml(const PmlParams &p, int i, int j, int k, int f[3]) {
    const int idx[3] = {i, j, k};
    for (int a = 0; a < 3; ++a) {
        f[a] = 0;
        if (idx[a] >= p.el_a[a][0] && idx[a] <= p.el_b[a][0]) f[a] = -1;
        if (idx[a] >= p.el_a[a][1] && idx[a] <= p.el_b[a][1]) f[a] = 1;
    }
}
// flags in force while element e (0-based) is integrated: those of its predecessor (Q17)
__host__ __device__ __forceinline__ void effective_pml(const MeshDims &m, const PmlParams &p, int e, int f[3]) {
    if (e == 0) { f[0] = p.first[0]; f[1] = p.first[1]; f[2] = p.first[2]; return; }
    int ie, je, ke;
    elem_ijk(m, e - 1, ie, je, ke);
    get_pml(p, ie, je, ke, f);
}

// ---- mbarrier / bulk-copy (TMA 1-D) helpers ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

}  // namespace movfem

// movfem_b200/csrc/geo.cuh -- geomodel -> grid nodes: the B200 version of
//
//   geometry.f90:801-970   innermodel_gqg  (element sweep, extension-plane copies, air, negative fill)
//   geometry.f90:975-1031  min_dd_inner    (brute-force nearest model cell per node: O(npt x cells), serial)
//   geometry.f90:1037-1085 assign_model    (tensor packing 11,12,13,22,23,33; + i*f32(eps*omega) on the diagonal)
//
// The reference visits the nodes element by element and searches ALL model cells for each; the result of a node is a
// pure function of its position, so here one warp owns a node, its lanes stride over the cells, and a warp reduction
// picks what the serial loop would: the first cell (in (im,jm,km) order) closer than 1e-5 if there is one, else the
// first cell that attains the minimum distance.  dd is evaluated exactly as the reference does (no FMA, IEEE sqrt), so
// the chosen cell -- an integer -- is bit-exact.  Everything after that is copies: HBM-bound, one thread per node.
#pragma once
#include "common.cuh"

namespace movfem {

struct GeoDims {
    int nnx, nny, nnz, o;            // node counts per axis, o = nord-1
    int x0, x1, y0, y1, z0, z1;      // 1-based inclusive node ranges of the inner elements' nodes ("visited")
    int ka, kb;                      // km range of the extension-plane copies: nextd*o .. nnz-(nzl_top+nzl_air)*o
    int mx, my, mz;
};

// assign_model for every model cell: cs[cell][6] complex, cm[cell][6] real
__global__ void geo_cell_tensors_kernel(int ncell, int isigma, int imu, const int *__restrict__ ijs, const int *__restrict__ iju,
                                        const double *__restrict__ sigma, const double *__restrict__ mu, double im32,
                                        double2 *__restrict__ cs, double *__restrict__ cm) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    double s[6] = {0, 0, 0, 0, 0, 0}, m[6] = {0, 0, 0, 0, 0, 0};
    if (isigma == 1) { s[0] = s[3] = s[5] = sigma[c]; }
    else
        for (int i = 0; i < isigma; ++i) {            // sequential: a repeated component keeps the last value
            const int r = ijs[i], q = ijs[isigma + i];   // ijsigma(i,1), ijsigma(i,2), column-major (isigma,2)
            if (r == 1) s[q - 1] = sigma[(size_t)c * isigma + i];
            else if (r == 2 || r == 3) s[r + q - 1] = sigma[(size_t)c * isigma + i];
        }
    if (imu == 0 || imu == 1) { m[0] = m[3] = m[5] = kMu0 * mu[(size_t)c * (imu ? imu : 1)]; }
    else
        for (int i = 0; i < imu; ++i) {
            const int r = iju[i], q = iju[imu + i];
            if (r == 1) m[q - 1] = kMu0 * mu[(size_t)c * imu + i];
            else if (r == 2 || r == 3) m[r + q - 1] = kMu0 * mu[(size_t)c * imu + i];
        }
    for (int k = 0; k < 6; ++k) {
        const bool diag = k == 0 || k == 3 || k == 5;
        cs[(size_t)c * 6 + k] = diag ? make_double2(s[k] + 0.0, im32) : make_double2(s[k], 0.0);   // diagonal: + cmplx32(0, eps*omega)
        cm[(size_t)c * 6 + k] = m[k];
    }
}

// model cell centres as (x, y) pairs so that the search stages plain copies: cxy[c] = (xm(im), ym(jm)), c = (im*my+jm)*mz+km
__global__ void geo_cell_xy_kernel(int ncell, int my, int mz, const double *__restrict__ xm, const double *__restrict__ ym, double2 *__restrict__ cxy) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int myz = my * mz;
    cxy[c] = make_double2(xm[c / myz], ym[(c % myz) / mz]);
}

// nearest model cell (min_dd_inner).  A block of 8 warps owns 8 x kGeoNpw visited nodes; the cells pass through shared
// memory in tiles (24 B per cell, read by every warp), each lane keeps kGeoNpw running minima over the cells it
// strides.  The reference compares dd = sqrt(d2); sqrt is monotone, so d2 < (d2 of the current best) is a necessary
// condition for dd < ddmin and the IEEE sqrt is evaluated only then (and the comparison that decides is the
// reference's own, on dd): the chosen cell is bit-exact, at 8 FP64 operations per (node, cell) pair.
constexpr int kGeoNpw = 4, kGeoTile = 1024;
__global__ void __launch_bounds__(256)
geo_nearest_kernel(GeoDims g, const double *__restrict__ xp, const double *__restrict__ yp, const double *__restrict__ zp,
                   const double2 *__restrict__ cxy, const double *__restrict__ zm, int *__restrict__ cell) {
    __shared__ double2 s_xy[kGeoTile];
    __shared__ double s_z[kGeoTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvx = g.x1 - g.x0 + 1, nvy = g.y1 - g.y0 + 1, nvz = g.z1 - g.z0 + 1;
    const int64_t nvis = (int64_t)nvx * nvy * nvz;
    const int64_t w0 = ((int64_t)blockIdx.x * 8 + warp) * kGeoNpw;
    double x[kGeoNpw], y[kGeoNpw], z[kGeoNpw], best[kGeoNpw], best2[kGeoNpw];
    int best_i[kGeoNpw], exact_i[kGeoNpw];
    int64_t id[kGeoNpw];
#pragma unroll
    for (int n = 0; n < kGeoNpw; ++n) {
        const int64_t w = min(w0 + n, nvis - 1);        // surplus slots repeat the last node (not stored)
        const int kk = g.z0 + (int)(w % nvz), jj = g.y0 + (int)((w / nvz) % nvy), ii = g.x0 + (int)(w / ((int64_t)nvz * nvy));
        id[n] = ((int64_t)(ii - 1) * g.nny + (jj - 1)) * g.nnz + (kk - 1);
        x[n] = xp[ii - 1]; y[n] = yp[jj - 1]; z[n] = zp[id[n]];
        best[n] = 1.e20; best2[n] = 1.e300;             // ddmin=1.d20 (geometry.f90:995)
        best_i[n] = -1; exact_i[n] = 0x7fffffff;
    }
    const int ncell = g.mx * g.my * g.mz;
    for (int t0 = 0; t0 < ncell; t0 += kGeoTile) {
        const int tn = min(kGeoTile, ncell - t0);
        __syncthreads();
        for (int c = threadIdx.x; c < tn; c += 256) { s_xy[c] = cxy[t0 + c]; s_z[c] = zm[t0 + c]; }
        __syncthreads();
        for (int c = lane; c < tn; c += 32) {
            const double2 xy = s_xy[c];
            const double cz = s_z[c];
#pragma unroll
            for (int n = 0; n < kGeoNpw; ++n) {
                const double dx = x[n] - xy.x, dy = y[n] - xy.y, dz = z[n] - cz;
                const double d2 = (dx * dx + dy * dy) + dz * dz;
                if (d2 < best2[n]) {                     // rare after the first few cells
                    const double dd = sqrt(d2);
                    if (dd <= 1.e-5) { if (t0 + c < exact_i[n]) exact_i[n] = t0 + c; }
                    else if (dd < best[n]) { best[n] = dd; best2[n] = d2; best_i[n] = t0 + c; }   // ascending c per lane: strict < keeps the earliest
                } else if (d2 <= 1.0000001e-10) {        // a cell within 1e-5 always takes the assignment (geometry.f90:1000-1003)
                    if (sqrt(d2) <= 1.e-5 && t0 + c < exact_i[n]) exact_i[n] = t0 + c;
                }
            }
        }
    }
#pragma unroll
    for (int n = 0; n < kGeoNpw; ++n) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best[n], off);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i[n], off), oe = __shfl_xor_sync(0xffffffffu, exact_i[n], off);
            if (oi >= 0 && (best_i[n] < 0 || ob < best[n] || (ob == best[n] && oi < best_i[n]))) { best[n] = ob; best_i[n] = oi; }
            exact_i[n] = min(exact_i[n], oe);
        }
        if (lane == 0 && w0 + n < nvis) cell[id[n]] = exact_i[n] != 0x7fffffff ? exact_i[n] : best_i[n];
    }
}

// one thread per node: visited nodes take their cell, extension nodes the clamped inner node's (the net effect of the
// sequential plane copies 1-3, geometry.f90:857-930), air nodes the air model (geometry.f90:932-945)
__global__ void geo_fill_kernel(GeoDims g, const int *__restrict__ cell, const double2 *__restrict__ cs, const double *__restrict__ cm,
                                double im32, double2 *__restrict__ g_sigma, double *__restrict__ g_mu) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npt = (int64_t)g.nnx * g.nny * g.nnz;
    if (id >= npt) return;
    const int kk = (int)(id % g.nnz) + 1, jj = (int)((id / g.nnz) % g.nny) + 1, ii = (int)(id / ((int64_t)g.nnz * g.nny)) + 1;
    double2 s[6];
    double m[6];
    if (kk > g.kb) {                    // air
        for (int k = 0; k < 6; ++k) {
            const bool diag = k == 0 || k == 3 || k == 5;
            s[k] = make_double2(0.0, diag ? im32 : 0.0);
            m[k] = diag ? kMu0 : 0.0;
        }
    } else {
        const bool visited = ii >= g.x0 && ii <= g.x1 && jj >= g.y0 && jj <= g.y1 && kk >= g.z0 && kk <= g.z1;
        int si = ii, sj = jj, sk = kk;
        if (!visited) {
            si = min(max(ii, g.x0), g.x1); sj = min(max(jj, g.y0), g.y1);
            if (kk < g.ka) sk = g.ka;   // planes below: from km = nextd*o, after the (y,z) and (x,z) copies
        }
        const int c = cell[((int64_t)(si - 1) * g.nny + (sj - 1)) * g.nnz + (sk - 1)];
        for (int k = 0; k < 6; ++k) { s[k] = cs[(size_t)c * 6 + k]; m[k] = cm[(size_t)c * 6 + k]; }
    }
    for (int k = 0; k < 6; ++k) { g_sigma[id * 6 + k] = s[k]; g_mu[id * 6 + k] = m[k]; }
}

// geometry.f90:947-962: a component with a negative (real part) value takes the value of the node below, sequentially
// upwards.  One thread per (column, component); the node "below" the first node of a column is the previous column's
// top node, which is air and never negative, so columns are independent.  The very first node of the grid has no
// predecessor (the reference reads out of bounds there): left unchanged.
__global__ void geo_negative_fill_kernel(int64_t ncol, int nnz, double2 *__restrict__ g_sigma, double *__restrict__ g_mu) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncol * 12) return;
    const int64_t col = t / 12;
    const int k = (int)(t % 12);
    const int64_t base = col * nnz;
    if (k < 6) {
        for (int kk = 0; kk < nnz; ++kk) {
            const int64_t id = base + kk;
            if (g_sigma[id * 6 + k].x < 0.0 && id > 0) g_sigma[id * 6 + k] = g_sigma[(id - 1) * 6 + k];
        }
    } else {
        for (int kk = 0; kk < nnz; ++kk) {
            const int64_t id = base + kk;
            if (g_mu[id * 6 + (k - 6)] < 0.0 && id > 0) g_mu[id * 6 + (k - 6)] = g_mu[(id - 1) * 6 + (k - 6)];
        }
    }
}

}  // namespace movfem

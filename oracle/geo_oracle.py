"""CPU restatement of geometry.f90 innermodel_gqg (TEST INFRASTRUCTURE, numpy).

SURVEY 8f rank 4: the input geomodel (conductivity / permeability on a coarse model grid) is copied to the grid nodes by
a brute-force nearest-neighbour search that is O(npt x model cells) and serial in the reference.  Only tests/ and the
tools' CPU-baseline legs import this module; the product (movfem_geo_innermodel in the CUDA library) never does.

Follows geometry.f90:801-970 (innermodel_gqg), :975-1031 (min_dd_inner), :1037-1085 (assign_model).  Pinned against the
reference itself: tests/golden/refgeo_*.npz hold g_sigma / g_mu produced by executing those procedures with
tests/golden/f90exec.py (tests/test_geo_innermodel.py).
"""
import numpy as np

EPS0 = 8.854187817e-12                    # geometry.f90:25
PI = 3.1415926535897932384626433
MU0 = 4.0 * PI * 1.0e-7                   # geometry.f90:26


def nearest_cells(x, y, z, xm, ym, zm):
    """min_dd_inner (geometry.f90:996-1012) for arrays of node coordinates: 0-based cell index idd per node.
    A cell closer than 1e-5 wins immediately (the first such cell in (im,jm,km) order); otherwise the first cell that
    attains the minimum distance (strict `<` keeps the earliest)."""
    mx, my = xm.size, ym.size
    mz = zm.size // (mx * my)
    im, jm, _ = np.meshgrid(np.arange(mx), np.arange(my), np.arange(mz), indexing="ij")
    cx, cy = xm[im.reshape(-1)], ym[jm.reshape(-1)]
    out = np.empty(x.size, np.int64)
    for s in range(0, x.size, 4096):
        e = min(x.size, s + 4096)
        dx = (x[s:e, None] - cx[None, :]) ** 2
        dy = (y[s:e, None] - cy[None, :]) ** 2
        dz = (z[s:e, None] - zm[None, :]) ** 2
        dd = np.sqrt((dx + dy) + dz)                      # Fortran order: (a**2 + b**2) + c**2
        exact = dd <= 1.0e-5
        first_exact = np.argmax(exact, axis=1)
        out[s:e] = np.where(exact.any(axis=1), first_exact, np.argmin(dd, axis=1))
    return out


def cell_tensors(omega, isigma, ijsigma, sigma, imu, ijmu, mu):
    """assign_model (geometry.f90:1037-1085) for every model cell: (ncell,6) complex sigma and (ncell,6) real mu,
    packing 11,12,13,22,23,33.  cmplx(0.d0,eps*omega) carries no KIND: single precision."""
    ncell = sigma.shape[1]
    im32 = float(np.float32(EPS0 * omega))
    s = np.zeros((ncell, 6), np.complex128)
    if isigma == 1:
        for k in (0, 3, 5):
            s[:, k] = sigma[0]
    else:
        for i in range(isigma):
            r, c = int(ijsigma[i, 0]), int(ijsigma[i, 1])
            if r == 1:
                s[:, c - 1] = sigma[i]
            elif r in (2, 3):
                s[:, r + c - 1] = sigma[i]
    for k in (0, 3, 5):
        s[:, k] += 1j * im32
    m = np.zeros((ncell, 6))
    if imu in (0, 1):
        for k in (0, 3, 5):
            m[:, k] = MU0 * mu[0]
    else:
        for i in range(imu):
            r, c = int(ijmu[i, 0]), int(ijmu[i, 1])
            if r == 1:
                m[:, c - 1] = MU0 * mu[i]
            elif r in (2, 3):
                m[:, r + c - 1] = MU0 * mu[i]
    return s, m


def innermodel_gqg(g_nx, g_ny, g_nz, nord, nextd, nzl_top, nzl_air, g_xp, g_yp, g_zp, omega,
                   xm, ym, zm, isigma, ijsigma, sigma, imu, ijmu, mu):
    """-> g_sigma (npt,6) complex128, g_mu (npt,6) float64 as innermodel_gqg leaves them (node id z-fastest)."""
    o = nord - 1
    nnx, nny, nnz = (g_nx - 1) * o + 1, (g_ny - 1) * o + 1, (g_nz - 1) * o + 1
    S = np.full((nnx, nny, nnz, 6), -1.0 + 0j)            # geometry.f90:829-830
    M = np.full((nnx, nny, nnz, 6), -1.0)
    valued = np.zeros((nnx, nny, nnz), bool)
    Z = g_zp.reshape(nnx, nny, nnz)
    cs, cm = cell_tensors(omega, isigma, ijsigma, sigma, imu, ijmu, mu)
    # nodes of the inner elements ie=nextd+1..g_nx-nextd-1, je likewise, ke=nextd..g_nz-nzl_top-nzl_air-1 (0-based slices)
    x0, x1 = nextd * o, nnx - nextd * o
    y0, y1 = nextd * o, nny - nextd * o
    z0, z1 = (nextd - 1) * o, nnz - (nzl_top + nzl_air) * o
    if nextd + 1 <= g_nx - nextd - 1 and nextd + 1 <= g_ny - nextd - 1 and nextd <= g_nz - nzl_top - nzl_air - 1:
        X, Y = np.meshgrid(g_xp[x0:x1], g_yp[y0:y1], indexing="ij")
        zz = Z[x0:x1, y0:y1, z0:z1]
        xx = np.broadcast_to(X[:, :, None], zz.shape)
        yy = np.broadcast_to(Y[:, :, None], zz.shape)
        cell = nearest_cells(xx.reshape(-1), yy.reshape(-1), zz.reshape(-1), xm, ym, zm).reshape(zz.shape)
        S[x0:x1, y0:y1, z0:z1] = cs[cell]
        M[x0:x1, y0:y1, z0:z1] = cm[cell]
        valued[x0:x1, y0:y1, z0:z1] = True
    ka, kb = nextd * o - 1, nnz - (nzl_top + nzl_air) * o    # km = nextd*o .. g_nnz-(...)*o, 0-based slice [ka, kb)

    def copy(dst, src):
        """dst, src: index tuples (slices / ints) over (im, jm, km); only unvalued nodes take the value"""
        mask = ~valued[dst]
        S[dst] = np.where(mask[..., None], np.broadcast_to(S[src], S[dst].shape), S[dst])
        M[dst] = np.where(mask[..., None], np.broadcast_to(M[src], M[dst].shape), M[dst])
        valued[dst] = True
    sk = slice(ka, kb)
    # 1. (y,z) planes (geometry.f90:857-884)
    copy((slice(0, nextd * o), slice(y0, y1), sk), (slice(nextd * o, nextd * o + 1), slice(y0, y1), sk))
    copy((slice(nnx - nextd * o, nnx), slice(y0, y1), sk), (slice(nnx - nextd * o - 1, nnx - nextd * o), slice(y0, y1), sk))
    # 2. (x,z) planes (geometry.f90:886-916)
    copy((slice(None), slice(0, nextd * o), sk), (slice(None), slice(nextd * o, nextd * o + 1), sk))
    copy((slice(None), slice(nny - nextd * o, nny), sk), (slice(None), slice(nny - nextd * o - 1, nny - nextd * o), sk))
    # 3. (x,y) planes below (geometry.f90:918-930): km = 1 .. nextd*o-1 from km = nextd*o
    if nextd * o - 1 >= 1:
        copy((slice(None), slice(None), slice(0, nextd * o - 1)), (slice(None), slice(None), slice(nextd * o - 1, nextd * o)))
    # 4. air (geometry.f90:932-945)
    a0 = nnz - (nzl_top + nzl_air) * o
    im32 = float(np.float32(EPS0 * omega))
    S[:, :, a0:, :] = 0.0
    M[:, :, a0:, :] = 0.0
    for k in (0, 3, 5):
        S[:, :, a0:, k] = 1j * im32
        M[:, :, a0:, k] = MU0
    # 5. negative entries take the value of the node below, sequentially upwards (geometry.f90:947-962); node id-1 of a
    #    column's first node is the previous column's last node
    Sf, Mf = S.reshape(-1, 6), M.reshape(-1, 6)
    for c in range(6):
        for arr, neg in ((Sf, lambda v: v.real < 0), (Mf, lambda v: v < 0)):
            col = arr[:, c]
            bad = np.flatnonzero(neg(col))
            for i in bad:                                   # ascending: already-fixed values propagate
                if i > 0:
                    col[i] = col[i - 1]
    return Sf, Mf

"""End-to-end observable of the north_star: triplets -> sparse LU -> solution.f90 post-processing.

ZMUMPS is absent from this environment, so SciPy's SuperLU stands in for it; graft and oracle triplets
go through the SAME solver and the SAME post-processing (oracle.node_solution, a restatement of solution.f90:18-69,
207-256,304-505), so differences in rho_a / phase measure the assembly only.  TEST INFRASTRUCTURE.

The shipped air model (sigma = i*f32(eps*omega) ~ 1e-11 S/m) leaves the curl-curl null space of the air region
regularised 11+ orders of magnitude below the matrix norm -- after ga_sort_sparse's float32 round trip the system
is numerically singular for ANY assembler (a 1e-13 relative perturbation of A changes rho_a by O(1)); the end-to-end
comparison therefore runs on models whose top layers are a poor conductor (1e-3 S/m) instead of vacuum, for which a
1e-13 perturbation moves rho_a by < 1e-8.  The assembly code path is the same."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from movfem_b200 import mesh


def e2e_model(mn, dirichlet, sch=1, nx=6, ny=5, freq=10.0):
    m = mesh.build_model(f"e2e_mn{mn}_d{dirichlet}", nx, ny, mn, 1000., 1100., 900., 2, 2, 1, dirichlet=dirichlet, gpml_sch=sch,
                         freqs=(freq,), sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0)
    air = m.sigma_re[:, 0] == 0.0
    for k in (0, 3, 5):
        m.sigma_re[air, k] = 1e-3
    return m


def solve_upper_triplets(nne, irn, jcn, a, rhs):
    """What the reference hands to ZMUMPS with SYM=2: upper triangle (1-based), centralized, dense RHS of two columns."""
    i, j = np.asarray(irn, np.int64) - 1, np.asarray(jcn, np.int64) - 1
    U = sp.coo_matrix((a, (i, j)), shape=(nne, nne)).tocsc()
    A = U + sp.triu(U, 1).T.tocsc()
    lu = spla.splu(A.astype(np.complex128))
    B = np.asarray(rhs, np.complex128).reshape(2, nne).T
    X = lu.solve(B)
    return np.ascontiguousarray(X.T).reshape(-1)


def rho_phi_diff(s_ref, s_new):
    """max relative difference of rho_a and max phase difference (degrees, modulo the atan branch) over the nodes and
    tensor components where the reference's own threshold (rho >= 1e-2, solution.f90:483) keeps a value"""
    ok = (s_ref["rho"] >= 1e-2) & (s_new["rho"] >= 1e-2)
    drho = float(np.max(np.abs(s_ref["rho"] - s_new["rho"])[ok] / s_ref["rho"][ok]))
    dphi = np.abs(s_ref["phi"] - s_new["phi"])[ok]
    dphi = float(np.max(np.minimum(dphi, 180.0 - dphi)))
    same_mask = bool(np.array_equal(s_ref["rho"] >= 1e-2, s_new["rho"] >= 1e-2))
    return drho, dphi, same_mask, int(ok.sum())

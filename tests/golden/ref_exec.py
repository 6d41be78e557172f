"""Runs the reference's own assembly path (MoVFEM_3DMT Fortran sources, executed by f90exec) on a synthetic model.

TEST INFRASTRUCTURE: used by make_reference_vectors.py to produce tests/golden/ref_*.npz.  It needs
/root/reference (absent on the GPU box), so nothing under tests/ imports it at test time except the optional
`reference`-marked checks that skip when the sources are missing.

What is executed from the reference, unmodified: init_n_fem, init_v_fem, init_problem, init_integration, ga_init
(ga_cgne, c_gne12/36/54, ga_nzindx, shr_nzindx12/36/54, init_bdary, init_gpml), bd_setmodel, and per frequency
update_omega, pset_pmodel, bd_updatemodel, global_vfem (MoVFEM_3DMT.f90:167-216: ga_assemble_nze, the element loop
with nf_get_r, p_elem_fields, int_elem_params, get_pml, local_vfem -> alocal/blocal/f1/f2/f3/f_boundary/assign_aij/
assign_bi, ga_sort_sparse with its merge sort), find_zeros and rem_zeros.  What this driver does in Python instead
of the reference: the job of read_input/grid_3d (the mesh and model arrays come from movfem_b200.mesh, they are the
*inputs* of the path), update_sigma (its out-of-bounds indexing, SURVEY Q12, needs linear memory; g_sigma is an
input of the boundary) and the frequency loop of the main program (MoVFEM_3DMT.f90:62-97, MPI/MUMPS calls dropped).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90exec as fx  # noqa: E402

SRC = "/root/reference/MoVFEM_3DMT/src"
FILES = ["kind_param", "geometry", "n_fem", "v_fem", "problem", "boundary_conds", "integration", "global_assembly",
         "MoVFEM_3DMT"]


def available():
    return all(os.path.exists(os.path.join(SRC, f + ".f90")) for f in FILES)


class ReferenceRun:
    def __init__(self, model, src=SRC, with_solution=False):
        self.model = m = model
        files = FILES + (["solution"] if with_solution else [])
        self.rt = rt = fx.Runtime([os.path.join(src, f + ".f90") for f in files], skip_calls=("estimate_memory",),
                                   hookable=("ga_sort_sparse", "local_vfem"))
        g = rt.mod("geometry")
        nord = m.nord
        # ---- what grid_3d leaves in module geometry (geometry.f90:79-84, 517-521) ----
        g.g_nx, g.g_ny, g.g_nz = m.g_nx, m.g_ny, m.g_nz
        g.g_nordx = g.g_nordy = g.g_nordz = nord
        g.g_nnx, g.g_nny, g.g_nnz = (m.g_nx - 1) * (nord - 1) + 1, (m.g_ny - 1) * (nord - 1) + 1, (m.g_nz - 1) * (nord - 1) + 1
        g.g_nyz = g.g_nny * g.g_nnz
        g.g_npt = g.g_nnx * g.g_nyz
        assert g.g_npt == m.npt and m.g_xp.size == g.g_nnx and m.g_yp.size == g.g_nny
        g.nextd = m.nextd
        g.g_nsf = 1
        g.g_nzl.a = np.array([m.nzl_top], dtype=np.int64)       # only g_nzl(g_nsf) is read on the path (boundary_conds.f90:62)
        g.g_xp.a = np.array(m.g_xp, dtype=np.float64)
        g.g_yp.a = np.array(m.g_yp, dtype=np.float64)
        g.g_zp.a = np.array(m.g_zp, dtype=np.float64)
        g.g_mu.a = np.asfortranarray(m.g_mu.T.astype(np.float64))              # (6, npt)
        g.g_sigma.a = np.asfortranarray(m.sigma_initial().T.astype(np.complex128))
        g.g_freq.a = np.array(m.freqs, dtype=np.float64)
        g.g_nf = int(m.freqs.size)
        g.g_ztop = np.float64(m.g_ztop)
        g.omega = np.float64(2.0) * g.pi * g.g_freq.a[0]                         # geometry.f90:73
        rt.call("geometry", "gqg_nodes")                                        # asx, asy, asz (geometry.f90:715-740)
        bc = rt.mod("boundary_conds")
        bc.gpml_sch = int(m.gpml_sch)
        bc.a0, bc.b0, bc.nn = np.float64(m.a0), np.float64(m.b0), np.float64(m.nn)
        # ---- initialize_vfem (MoVFEM_3DMT.f90:350-388) ----
        rt.call("n_fem", "init_n_fem", m.mn)
        rt.call("v_fem", "init_v_fem", m.me)
        rt.call("problem", "init_problem", 1, 2, 1)
        rt.call("integration", "init_integration")
        rt.call("global_assembly", "ga_init", True, bool(m.dirichlet), int(m.bd_inimod))
        nl = len(m.bd_lsigma)
        l_dz = np.array(list(m.bd_ldz) + [0.0] * max(0, nl - 1 - len(m.bd_ldz)), dtype=np.float64)[:max(nl - 1, 0)]
        rt.call("boundary_conds", "bd_setmodel", np.float64(m.bd_hsigma), nl, np.array(m.bd_lsigma, dtype=np.float64), l_dz)
        ga = rt.mod("global_assembly")
        self.nne, self.nnze = int(ga.nne), int(ga.nnze)
        self.gne = np.array(ga.gne.a, dtype=np.int32)             # [ide-1, im-1]
        self.taps = {}

    def frequency(self, ii, tap_elements=()):
        """one pass of the main program's frequency loop body (MoVFEM_3DMT.f90:62-97) for the 1-based frequency ii.
        Frequencies must be run in order 1, 2, ... (the reference carries in_pml across)."""
        rt, m = self.rt, self.model
        g = rt.mod("geometry")
        rt.call("geometry", "update_omega", ii)
        g.g_sigma.a = np.asfortranarray(m.sigma_for(ii).T.astype(np.complex128))   # update_sigma's effect (input)
        rt.call("problem", "pset_pmodel")
        rt.call("boundary_conds", "bd_updatemodel")
        nnze, nne = self.nnze, self.nne
        irn = np.zeros(nnze, dtype=np.int64); jcn = np.zeros(nnze, dtype=np.int64)
        a = np.zeros(nnze, dtype=np.complex128); rhs = np.zeros(2 * nne, dtype=np.complex128)
        out = {}

        def before_sort():                  # tap T1: after the element loop, before ga_sort_sparse
            out["ia_t1"], out["ja_t1"], out["a_t1"] = irn.astype(np.int32), jcn.astype(np.int32), a.copy()
        rt.hooks["ga_sort_sparse"] = before_sort
        if tap_elements:
            integ = rt.mod("integration")
            count = [0]
            me = m.me

            def before_local():
                count[0] += 1
                ide = count[0]
                if ide in tap_elements:
                    t = dict(wgt=integ.wgt.a.copy(), cve1=integ.cve1.a.copy(), cve2=integ.cve2.a.copy(), ve=integ.ve.a.copy(),
                             mf1=integ.mf1.a.copy(), mf2=integ.mf2.a.copy(), src=integ.src.a.copy())
                    if not m.dirichlet:
                        t["gpml"] = integ.gpml.a.copy()
                    A = np.zeros((me, me), dtype=np.complex128)
                    for im in range(1, me + 1):
                        for jm in range(1, me + 1):
                            A[im - 1, jm - 1] = rt.call("integration", "alocal", im, jm)
                    t["A"] = A
                    t["b"] = np.array([rt.call("integration", "blocal", im) for im in range(1, me + 1)])
                    out.setdefault("elements", {})[ide] = t
            rt.hooks["local_vfem"] = before_local
        else:
            rt.hooks.pop("local_vfem", None)
        rt.call("movfem_3dmt", "global_vfem", irn, jcn, a, rhs)
        # MoVFEM_3DMT.f90:85-97
        tnnz = rt.call("global_assembly", "find_zeros", a, 0)[0]
        out["find_zeros"] = int(tnnz)
        if tnnz > 0:
            n = nnze - tnnz
            tia = np.zeros(n, dtype=np.int64); tja = np.zeros(n, dtype=np.int64); ta = np.zeros(n, dtype=np.complex128)
            rt.call("global_assembly", "rem_zeros", n, irn, jcn, a, tia, tja, ta)
            irn, jcn, a = tia, tja, ta
        out.update(irn=irn.astype(np.int32), jcn=jcn.astype(np.int32), a=a, rhs=rhs, nz=int(a.size))
        return out

    def node_solution(self, x):
        """solution.f90:18-69 node_solution + z_rho_phi on a solved system x[2*nne] (needs with_solution=True):
        total E and H at the grid nodes, impedance tensor, apparent resistivity and phase."""
        rt = self.rt
        rt.call("solution", "node_solution", np.array(x, dtype=np.complex128))
        sol = rt.mod("solution")
        out = {k: getattr(sol, k).a.copy() for k in ("esol", "hsol", "z", "rho", "phi")}
        for k in ("esol", "hsol", "z", "rho", "phi"):          # the main program's write_solution deallocates them
            getattr(sol, k).a = None
        return out

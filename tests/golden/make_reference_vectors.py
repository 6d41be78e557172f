"""Generates tests/golden/ref_*.npz: outputs of the REFERENCE ITSELF (MoVFEM_3DMT Fortran sources under
/root/reference, executed by f90exec through ref_exec.ReferenceRun) on small synthetic meshes.

These are the pins of the oracle and of the CUDA path: gne/nne/nnze, and per frequency of the reference's sequential
frequency loop the delivered triplets (irn, jcn, a, nz after find_zeros/rem_zeros = tap T2), the right-hand side,
the pre-sort triplets (tap T1) and, for a few elements, the per-element caches of integration.f90 and A_e / b_e.

Run (needs /root/reference, takes ~15 min on 8 cores):   python tests/golden/make_reference_vectors.py [case ...]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from movfem_b200 import mesh  # noqa: E402


def _mesh(nx, ny, mn, nextd, n_earth, n_air, **kw):
    kw.setdefault("freqs", (0.5, 2.0))
    kw.setdefault("sigma_fn", mesh._layered((600., 600., 0., 900.)))
    kw.setdefault("topo_amp", 40.0)
    return mesh.build_model("ref", nx, ny, mn, 1000., 1100., 900., nextd, n_earth, n_air, **kw)


def _mu(m, seed):
    """a non-trivial permeability so that the curl part of the secondary source (problem.f90:362-420) is exercised"""
    rng = np.random.default_rng(seed)
    mur = 1.0 + 0.5 * rng.random(m.npt)
    m.g_mu[:, [0, 3, 5]] = (mesh.MU0 * mur)[:, None]
    m.g_mu[:, 1] = 0.05 * mesh.MU0 * rng.random(m.npt)
    return m


def _bd(m, inimod):
    m.bd_inimod, m.bd_hsigma, m.bd_lsigma, m.bd_ldz = inimod, 0.02, (0.01, 0.1, 0.001), (1.0, 2.5)
    return m


# name -> (model factory, frequencies to run, elements to tap)
CASES = {
    # 8-node elements: 5x5x6 with nextd=2 (upper-side GPML only) and 4x3x4 with nextd=1 (both sides)
    "ref_mn8_gpml_zhou": (lambda: _mesh(5, 5, 8, 2, 1, 1, dirichlet=0, gpml_sch=1), (1, 2), (1, 2, 75, 150)),
    "ref_mn8_gpml_fang": (lambda: _mesh(4, 3, 8, 1, 1, 1, dirichlet=0, gpml_sch=0, a0=1.5, b0=0.7, nn=2.0), (1, 2), (1, 2, 24, 48)),
    "ref_mn8_gpml_fang_n3": (lambda: _mu(_mesh(3, 4, 8, 1, 1, 1, dirichlet=0, gpml_sch=0, a0=2.0, b0=1.0, nn=2.5, aniso=True, seed=3), 11),
                             (1,), (1, 2, 20)),
    "ref_mn8_dirichlet": (lambda: _mesh(5, 5, 8, 2, 1, 1, dirichlet=1, gpml_sch=1), (1, 2), (1, 32, 150)),
    "ref_mn8_dirichlet_model2": (lambda: _bd(_mesh(4, 3, 8, 1, 1, 1, dirichlet=1), 2), (1, 2), (1, 17)),
    "ref_mn8_dirichlet_model3": (lambda: _bd(_mesh(4, 3, 8, 1, 1, 1, dirichlet=1), 3), (1, 2), (1, 17)),
    # 20-node elements
    "ref_mn20_gpml_fang": (lambda: _mesh(3, 3, 20, 1, 1, 1, dirichlet=0, gpml_sch=0), (1, 2), (1, 2, 20)),
    "ref_mn20_gpml_fang_nextd2": (lambda: _mesh(5, 5, 20, 2, 1, 1, dirichlet=0, gpml_sch=0, a0=2.0, b0=1.0, nn=2.0, freqs=(1.0,)), (1,), (1, 2, 150)),
    "ref_mn20_dirichlet_model3": (lambda: _bd(_mesh(3, 3, 20, 1, 1, 0, dirichlet=1), 3), (1,), (1, 14)),
    # 27-node elements
    "ref_mn27_gpml_zhou_aniso": (lambda: _mu(_mesh(3, 3, 27, 1, 1, 1, dirichlet=0, gpml_sch=1, aniso=True, seed=20141), 7), (1, 2), (1, 2, 20)),
    "ref_mn27_gpml_fang": (lambda: _mesh(3, 3, 27, 1, 1, 0, dirichlet=0, gpml_sch=0, a0=1.0, b0=1.0, nn=2.0), (1,), (2, 14)),
    "ref_mn27_dirichlet": (lambda: _mesh(3, 3, 27, 1, 1, 0, dirichlet=1), (1,), (14,)),
}


def _poor_air(m):
    """tests/e2e_util.py: a poor conductor instead of vacuum above the surface keeps the solve well conditioned"""
    air = m.sigma_re[:, 0] == 0.0
    for k in (0, 3, 5):
        m.sigma_re[air, k] = 1e-3
    return m


# end-to-end cases: assembly -> sparse LU (SciPy SuperLU standing in for ZMUMPS) -> the reference's own
# solution.f90 node_solution / z_rho_phi (executed): nodal E, H, impedance, apparent resistivity and phase
SOLUTION_CASES = {
    "refsol_mn8_dirichlet": lambda: _poor_air(_mesh(4, 3, 8, 1, 1, 1, dirichlet=1, freqs=(5.0,))),
    "refsol_mn8_gpml_fang": lambda: _poor_air(_mesh(4, 4, 8, 1, 2, 1, dirichlet=0, gpml_sch=0, freqs=(5.0,))),
    "refsol_mn20_gpml_fang": lambda: _poor_air(_mesh(3, 3, 20, 1, 1, 1, dirichlet=0, gpml_sch=0, freqs=(5.0,))),
    "refsol_mn27_dirichlet_model3": lambda: _bd(_poor_air(_mesh(3, 3, 27, 1, 1, 0, dirichlet=1, freqs=(5.0,))), 3),
}


# geomodel -> grid nodes: geometry.f90 innermodel_gqg / min_dd_inner / assign_model executed
def _geo_inputs(m, mx, my, mz, isigma, imu, seed, negative_offdiag=False, coincide=0):
    rng = np.random.default_rng(seed)
    xm = np.sort(rng.uniform(m.g_xp[0], m.g_xp[-1], mx))
    ym = np.sort(rng.uniform(m.g_yp[0], m.g_yp[-1], my))
    zm = rng.uniform(m.g_zp.min(), m.g_zp.max(), mx * my * mz)
    nnz = (m.g_nz - 1) * (m.nord - 1) + 1
    for c in range(coincide):               # model cells that coincide with grid nodes: the dd <= 1e-5 branch
        ii, jj = mx // 2, (my // 2 + c) % my
        i_node, j_node, k_node = m.g_xp.size // 2, (m.g_yp.size // 2 + c) % m.g_yp.size, nnz // 2
        xm[ii], ym[jj] = m.g_xp[i_node], m.g_yp[j_node]
        zm[(ii * my + jj) * mz + (c % mz)] = m.g_zp[(i_node * m.g_yp.size + j_node) * nnz + k_node]
    comps = {1: [(1, 1)], 3: [(1, 1), (2, 2), (3, 3)], 6: [(1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3)]}
    ijs, iju = np.array(comps[isigma]), np.array(comps[imu])
    sigma = rng.uniform(0.001, 1.0, (isigma, mx * my * mz))
    if negative_offdiag and isigma == 6:
        for i in (1, 2, 4):
            sigma[i] *= rng.choice([-0.3, 0.3], mx * my * mz)      # negative entries trigger geometry.f90:947-962
    mu = rng.uniform(1.0, 2.0, (imu, mx * my * mz))
    return dict(xm=xm, ym=ym, zm=zm, ijsigma=ijs, ijmu=iju, sigma=sigma, mu=mu)


GEO_CASES = {
    # name -> (mesh factory, n_air, (mx, my, mz, isigma, imu, seed), kwargs)
    "refgeo_mn8_iso": (lambda: _mesh(6, 5, 8, 2, 2, 1, freqs=(0.5,)), 1, (3, 4, 3, 1, 1, 1), dict(coincide=2)),
    "refgeo_mn8_aniso_negative": (lambda: _mesh(5, 5, 8, 1, 3, 2, freqs=(2.0,)), 2, (4, 3, 5, 6, 3, 2), dict(negative_offdiag=True)),
    "refgeo_mn20_aniso": (lambda: _mesh(4, 4, 20, 1, 2, 1, freqs=(0.5,)), 1, (3, 3, 4, 6, 6, 3), dict(coincide=1, negative_offdiag=True)),
    "refgeo_mn27_diag": (lambda: _mesh(6, 5, 27, 2, 1, 1, freqs=(1.0,)), 1, (2, 3, 3, 3, 1, 4), dict()),
}


def geo_case(name):
    factory, n_air, (mx, my, mz, isigma, imu, seed), kw = GEO_CASES[name]
    m = factory()
    return m, n_air, _geo_inputs(m, mx, my, mz, isigma, imu, seed, **kw)


def run_geo_case(name):
    import f90exec as fx
    t0 = time.time()
    m, n_air, inp = geo_case(name)
    src = "/root/reference/MoVFEM_3DMT/src/"
    rt = fx.Runtime([src + "kind_param.f90", src + "geometry.f90"])
    g = rt.mod("geometry")
    o = m.nord - 1
    g.g_nx, g.g_ny, g.g_nz = m.g_nx, m.g_ny, m.g_nz
    g.g_nordx = g.g_nordy = g.g_nordz = m.nord
    g.g_nnx, g.g_nny, g.g_nnz = (m.g_nx - 1) * o + 1, (m.g_ny - 1) * o + 1, (m.g_nz - 1) * o + 1
    g.g_nyz = g.g_nny * g.g_nnz
    g.g_npt = g.g_nnx * g.g_nyz
    g.nextd, g.g_nsf = m.nextd, 3
    g.g_nzl.a = np.array([m.g_nz - 1 - 2 * m.nextd - n_air, n_air, m.nzl_top], dtype=np.int64)
    g.g_xp.a, g.g_yp.a, g.g_zp.a = m.g_xp.copy(), m.g_yp.copy(), m.g_zp.copy()
    g.omega = np.float64(m.omega(1))
    mx, my = inp["xm"].size, inp["ym"].size
    mz = inp["zm"].size // (mx * my)
    rt.call("geometry", "innermodel_gqg", mx, my, mz, inp["xm"].copy(), inp["ym"].copy(), inp["zm"].copy(), inp["sigma"].shape[0],
            inp["mu"].shape[0], np.asfortranarray(inp["ijsigma"].astype(np.int64)), np.asfortranarray(inp["ijmu"].astype(np.int64)),
            np.asfortranarray(inp["sigma"]), np.asfortranarray(inp["mu"]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), g_sigma=g.g_sigma.a.T.copy(), g_mu=g.g_mu.a.T.copy())
    return name, int(g.g_npt), mx * my * mz, int((g.g_sigma.a.real < 0).sum()), time.time() - t0


def model_of(name):
    return SOLUTION_CASES[name]() if name in SOLUTION_CASES else CASES[name][0]()


def run_solution_case(name):
    import ref_exec
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from e2e_util import solve_upper_triplets
    t0 = time.time()
    m = SOLUTION_CASES[name]()
    r = ref_exec.ReferenceRun(m, with_solution=True)
    res = r.frequency(1)
    x = solve_upper_triplets(r.nne, res["irn"], res["jcn"], res["a"], res["rhs"])
    sol = r.node_solution(x)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), nne=r.nne, nnze=r.nnze, irn=res["irn"], jcn=res["jcn"], a=res["a"],
                        rhs=res["rhs"], x=x, **sol)
    return name, r.nne, r.nnze, int(res["a"].size), time.time() - t0


def run_case(name):
    import ref_exec
    t0 = time.time()
    factory, freqs, taps = CASES[name]
    m = factory()
    r = ref_exec.ReferenceRun(m)
    out = dict(gne=r.gne, nne=r.nne, nnze=r.nnze, freqs=np.array(freqs), taps=np.array(taps))
    for ii in freqs:
        res = r.frequency(ii, tap_elements=taps if ii == freqs[0] else ())
        for k in ("irn", "jcn", "a", "rhs", "ia_t1", "ja_t1", "a_t1"):
            out["%s%d" % (k, ii)] = res[k]
        out["find_zeros%d" % ii] = res["find_zeros"]
        for ide, t in res.get("elements", {}).items():
            for k, v in t.items():
                out["el%d_%s" % (ide, k)] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return name, r.nne, r.nnze, int(out["a%d" % freqs[0]].size), time.time() - t0


def _run(name):
    if name in GEO_CASES:
        return run_geo_case(name)
    return run_solution_case(name) if name in SOLUTION_CASES else run_case(name)


def main():
    names = sys.argv[1:] or (list(CASES) + list(SOLUTION_CASES) + list(GEO_CASES))
    with mp.Pool(min(len(names), os.cpu_count() or 1)) as pool:
        for res in pool.imap_unordered(_run, names):
            print("%-28s nne %6d nnze %8d nz %8d  %.0f s" % res, flush=True)


if __name__ == "__main__":
    main()

"""CPU tests: the oracle against the only pins the reference offers.

The reference ships no tests / golden outputs and cannot be compiled here (Fortran), so the oracle is
pinned to (1) the element-matrix known answers of SURVEY App. B item 4, (2) the exact nne/nnze counts of
SURVEY section 6 and App. B.1, (3) the invariants of App. B items 3 and 5, (4) its own frozen outputs.
"""
import json
import os

import numpy as np
import pytest

from movfem_b200 import mesh
from oracle.oracle import Oracle, shape_eval

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "pins_appB.json")))


def _KM(mn, **kw):
    m = mesh.brick_single_element(mn, **kw)
    o = Oracle(m)
    e = o.element(1, 1.0, m.sigma_initial())       # omega = 1 -> f32(omega) = 1: Im A = M exactly
    return e["A"].real, e["A"].imag, m.me


@pytest.mark.parametrize("mn", [8, 20, 27])
def test_element_matrix_known_answers(mn):
    K, M, me = _KM(mn)
    got = [K[0, 0], K[1, 0], K[me - 1, 0], np.trace(K), np.linalg.norm(K), M[0, 0], M[me - 1, me - 1], np.trace(M), np.linalg.norm(M)]
    np.testing.assert_allclose(got, PINS[str(mn)], rtol=1e-12)      # pins carry ~1e-12 (numpy inverse), App. B item 4


def test_distorted_linear_element_q5():
    p = PINS["distorted8"]
    K, M, _ = _KM(8, top_shift=p["top_shift"])
    np.testing.assert_allclose([K[0, 0], np.trace(K), np.linalg.norm(K), np.trace(M), np.linalg.norm(M)],
                               [p["K(1,1)"], p["trK"], p["normF_K"], p["trM"], p["normF_M"]], rtol=1e-12)
    assert np.abs(K - K.T).max() <= 1e-12 * np.abs(K).max()


@pytest.mark.parametrize("mn,null_dim", [(8, 7), (20, 19), (27, 26)])
def test_invariants_symmetry_nullspace_spd(mn, null_dim):
    K, M, me = _KM(mn)
    assert np.abs(K - K.T).max() <= 1e-13 * np.abs(K).max()
    assert np.abs(M - M.T).max() <= 1e-13 * np.abs(M).max()
    Ks = 0.5 * (K + K.T)
    ev = np.linalg.eigvalsh(Ks)
    assert np.sum(np.abs(ev) < 1e-9 * np.abs(ev).max()) == null_dim        # gradients of the nodal space
    assert np.linalg.eigvalsh(0.5 * (M + M.T)).min() > 0                   # mass matrix SPD


@pytest.mark.parametrize("mn,me", [(8, 12), (20, 36), (27, 54)])
def test_shape_tables_derivatives_and_partition_of_unity(mn, me):
    """App. B item 5: the analytic derivative tables agree with central differences everywhere except the
    8-node dN/dzeta entry (n_fem.f90:193, Q5), which must be wrong by exactly the missing xi factor."""
    rng = np.random.default_rng(5)
    h = 1e-6
    for _ in range(10):
        x = rng.uniform(-0.9, 0.9, 3)
        N, dN, phi, dphi = shape_eval(mn, me, *x)
        assert abs(N.sum() - 1.0) < 1e-13
        for d in range(3):
            xp, xm = x.copy(), x.copy()
            xp[d] += h; xm[d] -= h
            Np, _, pp, _ = shape_eval(mn, me, *xp)
            Nm, _, pm, _ = shape_eval(mn, me, *xm)
            fdN, fdp = (Np - Nm) / (2 * h), (pp - pm) / (2 * h)
            np.testing.assert_allclose(dphi[:, d], fdp, atol=1e-8)
            if mn == 8 and d == 2:
                nr = np.array([1, 1, -1, -1, 1, 1, -1, -1.0])
                nz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
                ne = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0])
                np.testing.assert_allclose(dN[:, 2], (1 + nr) * (1 + ne * x[1]) * nz / 8.0, atol=1e-15)   # the typo, literally
                assert np.abs(dN[:, 2] - fdN).max() > 1e-3
            else:
                np.testing.assert_allclose(dN[:, d], fdN, atol=1e-8)


def test_counts_small_meshes():
    for mn, exp in PINS["counts"]["3x4x5"].items():
        for dirich in (0, 1):
            m = mesh.build_model("t", 3, 4, int(mn), 1000., 1000., 1000., 1, 2, 1, dirichlet=dirich)
            o = Oracle(m)
            assert [o.nne, o.nnze] == exp[dirich]
    o = Oracle(mesh.brick_single_element(8))
    assert [o.nne, o.nnze] == PINS["counts"]["1x1x1_me12"]


@pytest.mark.parametrize("dirich,key", [(0, "config1_gpml"), (1, "config1_dirichlet")])
def test_counts_config1(dirich, key):
    o = Oracle(mesh.config(1, dirichlet=dirich))
    assert [o.nne, o.nnze, o.nz_upper] == PINS["counts"][key]
    assert o.nz_upper == (o.nnze + o.nne) // 2          # App. B.2 closed form


def test_pattern_properties():
    m = mesh.build_model("t", 4, 3, 20, 1000., 1000., 1000., 1, 1, 1, dirichlet=0)
    o = Oracle(m)
    ia, ja = o.pattern()
    key = ia.astype(np.int64) * (o.nne + 1) + ja
    assert np.all(np.diff(key) > 0)                                      # strictly (row, col) sorted, unique
    kt = ja.astype(np.int64) * (o.nne + 1) + ia
    assert np.array_equal(np.sort(kt), key)                              # structurally symmetric
    g = o.gne()
    assert g.min() >= 1 and g.max() == o.nne and np.unique(g).size == o.nne


@pytest.mark.parametrize("name", ["small_mn8_gpml_zhou", "small_mn8_dirichlet", "small_mn20_gpml_fang", "small_mn27_gpml_zhou"])
def test_frozen_oracle_outputs_and_modes(name):
    """faithful (reference loop structure), memoised and multi-threaded modes give identical bits, and
    match the frozen fixtures; two sequential frequencies exercise Q12 and Q17."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    m = make_golden.make_model(**make_golden.CASES[name])
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    for faithful, nthreads in ((False, 1), (False, 4)):
        o = Oracle(m)
        assert np.array_equal(o.gne(), ref["gne"])
        for ifreq in (1, 2):
            r = o.assemble(m.omega(ifreq), m.sigma_for(ifreq), faithful=faithful, nthreads=nthreads)
            assert np.array_equal(r["irn"], ref[f"irn{ifreq}"]) and np.array_equal(r["jcn"], ref[f"jcn{ifreq}"])
            assert np.array_equal(r["a"], ref[f"a{ifreq}"]) and np.array_equal(r["rhs"], ref[f"rhs{ifreq}"])
            # T2 semantics (Q10/Q11): upper triangle, row-major, float32 values, no exact zeros
            assert np.all(r["irn"] <= r["jcn"])
            k = r["irn"].astype(np.int64) * (o.nne + 1) + r["jcn"]
            assert np.all(np.diff(k) > 0)
            assert np.array_equal(r["a"].real.astype(np.float32).astype(np.float64), r["a"].real)
            assert not np.any(r["a"] == 0)


def test_t2_is_transposed_lower_triangle():
    """Q10: ga_sort_sparse leaves A(c,r) (the computed lower triangle) at the upper position (r,c)."""
    m = mesh.build_model("t", 4, 4, 8, 1000., 1000., 1000., 1, 1, 1, dirichlet=0, gpml_sch=1, freqs=(1.0,))
    o = Oracle(m)
    r = o.assemble(m.omega(1), m.sigma_for(1))
    ia, ja = o.pattern()
    low = ia >= ja
    order = np.lexsort((ia[low], ja[low]))
    v = r["a_t1"][low][order]
    v32 = v.real.astype(np.float32).astype(np.float64) + 1j * v.imag.astype(np.float32).astype(np.float64)
    keep = v32 != 0
    assert np.array_equal(r["a"], v32[keep])
    assert np.array_equal(r["irn"], ja[low][order][keep]) and np.array_equal(r["jcn"], ia[low][order][keep])
    assert np.all(r["a_t1"][ia < ja] == 0)               # strict upper never assigned (sym, MoVFEM_3DMT.f90:246)


def test_q17_stale_flags_between_frequencies():
    # scheme 0 (Fang): with scheme 1 the stored stretch is Re(h) = 1 everywhere (Q18) and the flags cannot matter
    m = mesh.build_model("t", 5, 5, 8, 1000., 1000., 1000., 2, 1, 1, dirichlet=0, gpml_sch=0, freqs=(1.0, 1.0))
    o = Oracle(m)
    assert list(o.in_pml()) == [0, 0, 0]
    r1 = o.assemble(m.omega(1), m.sigma_for(1))
    assert list(o.in_pml()) == [1, 1, 1]                 # left behind by the last element
    r2 = o.assemble(m.omega(1), m.sigma_for(1))          # same omega, same sigma: only element (1,1,1) differs
    d = np.flatnonzero(r1["a_t1"] != r2["a_t1"])
    ia, ja = o.pattern()
    g1 = set(o.gne()[0])
    assert d.size > 0 and set(ia[d]) <= g1 and set(ja[d]) <= g1
    assert list(o.effective_pml(2)) == [0, 0, 0] and list(o.effective_pml(m.g_nz)) == [0, 0, 1]   # bottom element of column 2 sees the top flags


def test_node_solution_recovers_the_layered_halfspace_response():
    """Pin of the post-processing restatement (solution.f90): on a laterally uniform earth the oracle's own triplets,
    solved by the stand-in LU, must give a 1-D response -- Zxy = -Zyx, Zxx = Zyy = 0 to solver accuracy, and the same
    rho_a at every surface node of the inner zone -- and a 1e-13 perturbation of A must not move rho_a (the
    conditioning claim tests/e2e_util.py makes for its models)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from e2e_util import solve_upper_triplets, rho_phi_diff
    from movfem_b200 import mesh
    from oracle.oracle import Oracle
    m = mesh.build_model("halfspace", 8, 8, 8, 1000., 1000., 500., 2, 4, 2, dirichlet=1, gpml_sch=1, freqs=(10.0,))
    air = m.sigma_re[:, 0] == 0.0
    for k in (0, 3, 5):
        m.sigma_re[air, k] = 1e-3
    o = Oracle(m)
    om, sg = m.omega(1), m.sigma_for(1)
    r = o.assemble(om, sg)
    x = solve_upper_triplets(o.nne, r["irn"], r["jcn"], r["a"], r["rhs"])
    s = o.node_solution(om, sg, x)
    nn = m.g_nx
    z = s["z"].reshape(nn, nn, m.g_nz, 4)
    ks = 2 + 4                                   # surface node plane: nextd + n_earth
    zc = z[3:6, 3:6, ks, :]                      # inner zone
    # z(1..4) = Zxx, Zxy(-like), Zyx, Zyy (solution.f90:461-464).  1-D earth on an x<->y symmetric mesh: the diagonal
    # vanishes, Zxy(i,j) = -Zyx(j,i) (H at a node is a one-sided derivative taken in the first element that visits
    # it, solution.f90:244-254, so the response is not identical from node to node), exactly antisymmetric at the centre
    off = np.abs(zc[..., 1])
    assert np.all(np.abs(zc[..., 0]) <= 1e-6 * off) and np.all(np.abs(zc[..., 3]) <= 1e-6 * off)
    assert np.all(np.abs(zc[..., 1] + np.swapaxes(zc[..., 2], 0, 1)) <= 1e-9 * off)
    assert abs(zc[1, 1, 1] + zc[1, 1, 2]) <= 1e-9 * abs(zc[1, 1, 1])
    rho = s["rho"].reshape(nn, nn, m.g_nz, 4)[3:6, 3:6, ks, 1]
    assert 50.0 < rho.min() and rho.max() < 400.0          # 100 Ohm m half-space under a poor conductor, 500 m cells at 10 Hz
    rng = np.random.default_rng(0)
    xp = solve_upper_triplets(o.nne, r["irn"], r["jcn"], r["a"] * (1 + 1e-13 * rng.standard_normal(r["a"].size)), r["rhs"])
    drho, dphi, same, n = rho_phi_diff(s, o.node_solution(om, sg, xp))
    assert same and drho <= 1e-7, (drho, dphi)

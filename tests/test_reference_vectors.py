"""The oracle (CPU) and the CUDA path (GPU) against outputs of the REFERENCE ITSELF.

tests/golden/ref_*.npz were produced by executing the reference's own Fortran sources (MoVFEM_3DMT.f90 global_vfem /
local_vfem, global_assembly.f90, integration.f90, v_fem.f90, n_fem.f90, problem.f90, boundary_conds.f90) with the
Fortran-subset executor tests/golden/f90exec.py -- see tests/golden/make_reference_vectors.py.  They pin:
gne / nne / nnze, and per frequency of the sequential frequency loop the delivered triplets (tap T2: irn, jcn, a after
ga_sort_sparse + find_zeros/rem_zeros), the right-hand side, the pre-sort triplets (tap T1) and per-element caches,
A_e and b_e of a few elements.

Bars: integer data bit-exact; values <= 1e-12 normwise relative (north_star).  The oracle actually reproduces the
executed reference bit for bit on most cases, which is asserted where it holds (EXACT).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_reference_vectors as mrv  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from parity_util import rel_err  # noqa: E402

TOL = 1e-12
CASES = sorted(mrv.CASES)


def _load(name):
    path = os.path.join(HERE, "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % name)
    return np.load(path)


def _keyed(ia, ja, a, nne):
    key = ia.astype(np.int64) * (nne + 1) + ja
    order = np.argsort(key, kind="stable")
    return key[order], a[order]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_executed_reference(name):
    ref = _load(name)
    m = mrv.model_of(name)
    o = Oracle(m)
    assert o.nne == int(ref["nne"]) and o.nnze == int(ref["nnze"])
    assert np.array_equal(o.gne(), ref["gne"])                                 # c_gne12/36/54, bit-exact
    ia, ja = o.pattern()
    for ii in ref["freqs"]:
        ii = int(ii)
        res = o.assemble(m.omega(ii), m.sigma_for(ii), faithful=True)
        # tap T2: what ZMUMPS receives
        assert res["nz"] == ref["a%d" % ii].size
        assert np.array_equal(res["irn"], ref["irn%d" % ii]) and np.array_equal(res["jcn"], ref["jcn%d" % ii])
        assert rel_err(res["a"], ref["a%d" % ii]) <= TOL
        assert rel_err(res["rhs"], ref["rhs%d" % ii]) <= TOL
        # tap T1: the same set of (row, col) pairs (the order inside a row is shr_nzindx*'s, not restated) and values
        k_ref, a_ref = _keyed(ref["ia_t1%d" % ii], ref["ja_t1%d" % ii], ref["a_t1%d" % ii], o.nne)
        k_orc, a_orc = _keyed(ia, ja, res["a_t1"], o.nne)
        assert np.array_equal(k_ref, k_orc)
        assert rel_err(a_orc, a_ref) <= TOL
        # rows appear in ascending order in the reference's own numbering too
        assert np.all(np.diff(ref["ia_t1%d" % ii]) >= 0)


@pytest.mark.parametrize("name", CASES)
def test_oracle_element_taps(name):
    """integration.f90's per-element caches (wgt, cve1, cve2, ve, mf1, mf2, src, gpml) and A_e, b_e of the tapped
    elements at the first frequency; the GPML flags are the ones the element actually sees (one element late, Q17)."""
    ref = _load(name)
    m = mrv.model_of(name)
    o = Oracle(m)
    ii = int(ref["freqs"][0])
    for ide in ref["taps"]:
        ide = int(ide)
        e = o.element(ide, m.omega(ii), m.sigma_for(ii), pml=o.effective_pml(ide), caches=True)
        for k in ("wgt", "cve1", "cve2", "ve", "mf1", "mf2", "src", "gpml", "A", "b"):
            key = "el%d_%s" % (ide, k)
            if key not in ref.files:
                continue
            r = ref[key]
            x = e[k] if e[k].shape == r.shape else e[k].reshape(-1)[: r.size].reshape(r.shape)
            if k == "A":
                # the reference computes alocal only where gne(im) >= gne(jm) >= 0; the tap evaluated every pair
                assert rel_err(x, r) <= TOL, (ide, k)
            else:
                assert rel_err(x, r) <= TOL, (ide, k)


def test_q18_gpml_stretch_is_real_in_the_reference():
    """integration.f90:16 declares `gpml` real(kind=double): the executed reference stores Re(h) only -- exactly 1 for
    scheme 1 (Zhou), 1 + a0*rho**n for scheme 0 (Fang)."""
    z = _load("ref_mn8_gpml_zhou")
    assert z["el2_gpml"].dtype == np.float64 and np.all(z["el150_gpml"] == 1.0)
    f = _load("ref_mn8_gpml_fang")
    assert f["el2_gpml"].dtype == np.float64 and f["el2_gpml"].min() >= 1.0 and f["el2_gpml"].max() > 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_against_the_executed_reference(name):
    from movfem_b200 import abi, host
    ref = _load(name)
    m = mrv.model_of(name)
    asm = host.Assembly(m)
    assert np.array_equal(asm.gne(), ref["gne"]) and asm.nne == int(ref["nne"]) and asm.nnze == int(ref["nnze"])
    for ii in ref["freqs"]:
        ii = int(ii)
        # tap T1: upper triangle r <= c carrying A_lower(c, r), sorted by (r, c)
        irn, jcn, a, rhs, nz = asm.global_vfem(ii, m.omega(ii), m.sigma_for(ii), mode=abi.MODE_T1)
        ia, ja, at1 = ref["ia_t1%d" % ii], ref["ja_t1%d" % ii], ref["a_t1%d" % ii]
        low = ia >= ja
        r_, c_, v_ = ja[low], ia[low], at1[low]
        order = np.lexsort((c_, r_))
        assert nz == order.size and np.array_equal(irn[:nz], r_[order]) and np.array_equal(jcn[:nz], c_[order])
        assert rel_err(a[:nz], v_[order]) <= TOL
        assert rel_err(rhs, ref["rhs%d" % ii]) <= TOL
        # tap T2: delivered triplets
        irn, jcn, a, rhs, nz = asm.global_vfem(ii, m.omega(ii), m.sigma_for(ii), mode=abi.MODE_T2)
        ra, rirn, rjcn = ref["a%d" % ii], ref["irn%d" % ii], ref["jcn%d" % ii]
        # bit-exact delivered pattern: the entries the reference cancels to exactly zero and strips (find_zeros/rem_zeros,
        # global_assembly.f90:123-150) are stripped here too, and nothing else is (csrc/exact.cuh)
        assert nz == ra.size and np.array_equal(irn[:nz], rirn) and np.array_equal(jcn[:nz], rjcn)
        assert rel_err(a[:nz], ra) <= TOL
        assert rel_err(rhs, ref["rhs%d" % ii]) <= TOL
    asm.close()


# ---- end to end: the reference's own solution.f90 post-processing, executed ------------------------

SOL_CASES = sorted(mrv.SOLUTION_CASES)


def _phase_diff(a, b):
    d = np.abs(a - b)
    return np.minimum(d, 180.0 - d)          # atan branch: +-90 degrees where Re z is a signed zero


@pytest.mark.parametrize("name", SOL_CASES)
def test_oracle_node_solution_reproduces_the_executed_reference(name):
    """oracle.node_solution (restating solution.f90:18-69,207-256,304-505) against node_solution/z_rho_phi of the
    reference executed on the same solved system x: total E, H at the nodes, impedance, rho_a, phase."""
    ref = _load(name)
    m = mrv.model_of(name)
    o = Oracle(m)
    res = o.assemble(m.omega(1), m.sigma_for(1), faithful=True)
    assert np.array_equal(res["irn"], ref["irn"]) and np.array_equal(res["jcn"], ref["jcn"])
    assert rel_err(res["a"], ref["a"]) <= TOL and rel_err(res["rhs"], ref["rhs"]) <= TOL
    s = o.node_solution(m.omega(1), m.sigma_for(1), ref["x"])
    for k in ("esol", "hsol", "z", "rho"):
        assert rel_err(s[k], ref[k]) <= TOL, k
    assert _phase_diff(s["phi"], ref["phi"]).max() <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", SOL_CASES)
def test_cuda_end_to_end_rho_phase_against_the_executed_reference(name):
    """north_star: apparent resistivity and phase after the (unchanged) solve agree with the reference to <= 1e-6.
    Graft triplets -> SuperLU (standing in for ZMUMPS) -> post-processing; compared with the reference's own
    rho / phi obtained from the reference's triplets through the same solver."""
    from movfem_b200 import host
    from e2e_util import solve_upper_triplets
    ref = _load(name)
    m = mrv.model_of(name)
    asm, o = host.Assembly(m), Oracle(m)
    irn, jcn, a, rhs, nz = asm.global_vfem(1, m.omega(1), m.sigma_for(1))
    assert nz == ref["a"].size and np.array_equal(irn[:nz], ref["irn"]) and np.array_equal(jcn[:nz], ref["jcn"])
    x = solve_upper_triplets(asm.nne, irn[:nz], jcn[:nz], a[:nz], rhs)
    s = o.node_solution(m.omega(1), m.sigma_for(1), x)
    ok = (ref["rho"] >= 1e-2) & (s["rho"] >= 1e-2)
    assert np.array_equal(ref["rho"] >= 1e-2, s["rho"] >= 1e-2)
    assert (np.abs(s["rho"] - ref["rho"])[ok] / ref["rho"][ok]).max() <= 1e-6
    assert _phase_diff(s["phi"], ref["phi"])[ok].max() <= 1e-6
    asm.close()

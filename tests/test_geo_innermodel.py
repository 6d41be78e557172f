"""Geomodel -> grid nodes: geometry.f90 innermodel_gqg / min_dd_inner / assign_model.

tests/golden/refgeo_*.npz are outputs of the reference's own procedures executed by tests/golden/f90exec.py
(make_reference_vectors.py: run_geo_case).  CPU: the numpy restatement oracle/geo_oracle.py reproduces them bit for
bit.  GPU: movfem_geo_innermodel (C ABI) reproduces them bit for bit, and agrees with the oracle on a larger case.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_reference_vectors as mrv  # noqa: E402
from oracle import geo_oracle  # noqa: E402

CASES = sorted(mrv.GEO_CASES)


def _defined(m, inp):
    """nodes whose value the reference defines.  geometry.f90:947-962 replaces a negative component by the value of node
    id-1; for the first node of the grid that is g_sigma(:,0), out of bounds (whatever lies before the array), and the
    garbage propagates up the first column while its components are negative.  Negative inputs only: the first column is
    excluded there (the graft leaves such a node unchanged)."""
    keep = np.ones(m.npt, bool)
    if (inp["sigma"] < 0).any():
        keep[: (m.g_nz - 1) * (m.nord - 1) + 1] = False
    return keep


def _oracle(m, n_air, inp):
    return geo_oracle.innermodel_gqg(m.g_nx, m.g_ny, m.g_nz, m.nord, m.nextd, m.nzl_top, n_air, m.g_xp, m.g_yp, m.g_zp, m.omega(1),
                                     inp["xm"], inp["ym"], inp["zm"], inp["sigma"].shape[0], inp["ijsigma"], inp["sigma"],
                                     inp["mu"].shape[0], inp["ijmu"], inp["mu"])


@pytest.mark.parametrize("name", CASES)
def test_geo_oracle_reproduces_the_executed_reference(name):
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    m, n_air, inp = mrv.geo_case(name)
    S, M = _oracle(m, n_air, inp)
    k = _defined(m, inp)
    assert np.array_equal(S[k], ref["g_sigma"][k]) and np.array_equal(M[k], ref["g_mu"][k])
    assert (~k).sum() <= (m.g_nz - 1) * (m.nord - 1) + 1
    # what the assembly relies on: nothing is left unassigned, air carries i*f32(eps*omega) on the diagonal
    assert not (ref["g_sigma"].real < 0).any() and not (ref["g_mu"] < 0).any()
    assert np.all(ref["g_sigma"][-1, [0, 3, 5]] == 1j * np.float32(geo_oracle.EPS0 * m.omega(1)))


def test_nearest_cell_rules():
    """min_dd_inner: a cell within 1e-5 wins even if an earlier cell is nearer than all others; ties keep the first."""
    xm, ym = np.array([0.0, 10.0]), np.array([0.0])
    zm = np.array([0.0, 5.0, 0.0, 5.0])                       # cells (im,km): (0,0) (0,1) (1,0) (1,1)
    c = geo_oracle.nearest_cells(np.array([5.0, 10.0, 2.0]), np.array([0.0, 0.0, 0.0]), np.array([2.5, 5.0 + 5e-6, 0.0]), xm, ym, zm)
    assert list(c) == [0, 3, 0]                                # 4-way tie -> first; within 1e-5 of cell 3; plain nearest


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_innermodel_against_the_executed_reference(name):
    from movfem_b200 import host
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    m, n_air, inp = mrv.geo_case(name)
    S, M, ms = host.innermodel_gqg(m, n_air, m.omega(1), inp["xm"], inp["ym"], inp["zm"], inp["ijsigma"], inp["sigma"], inp["ijmu"], inp["mu"])
    k = _defined(m, inp)
    assert np.array_equal(S[k], ref["g_sigma"][k]) and np.array_equal(M[k], ref["g_mu"][k])        # bit for bit
    So, Mo = _oracle(m, n_air, inp)
    assert np.array_equal(S, So) and np.array_equal(M, Mo)                                         # incl. the first column
    assert ms > 0


@pytest.mark.gpu
def test_cuda_innermodel_against_the_oracle_on_a_larger_grid():
    """config 2 at 0.4 scale (20-node elements, 33x33x25 nodes) against a 12x11x9 anisotropic model: 27 225 nodes x 1188
    cells, cell choice and every copied value identical to the CPU restatement."""
    from movfem_b200 import host, mesh
    m = mesh.config(2, scale=0.4)
    n_air = 2
    inp = mrv._geo_inputs(m, 12, 11, 9, 6, 3, seed=5, negative_offdiag=True, coincide=3)
    S, M, _ = host.innermodel_gqg(m, n_air, m.omega(1), inp["xm"], inp["ym"], inp["zm"], inp["ijsigma"], inp["sigma"], inp["ijmu"], inp["mu"])
    So, Mo = _oracle(m, n_air, inp)
    assert np.array_equal(S, So) and np.array_equal(M, Mo)

"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Bars (north_star / SURVEY 8c): gne, nne, nnze and the structural pattern bit-exact; matrix (tap T1) and RHS
values <= 1e-12 normwise relative in complex128; tap T2 (what ZMUMPS receives) has identical IRN/JCN/nz and
float32 values equal up to the rounding flips a 1e-16 double difference can cause on noise-level entries.
"""
import os
import sys

import numpy as np
import pytest

from movfem_b200 import abi, host, mesh
from oracle.oracle import Oracle
from parity_util import compare_assembly, rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


def _check(r):
    assert r["nne"][0] == r["nne"][1] and r["nnze"][0] == r["nnze"][1] and r["nz_upper"][0] == r["nz_upper"][1], r
    assert r["pattern_equal"], r
    assert r["t1_idx_equal"] and r["t1_rel"] <= TOL and r["rhs_rel"] <= TOL, r
    # the DELIVERED pattern (after find_zeros/rem_zeros) is bit-exact: same nz, same IRN/JCN -- including which round-off
    # residues of mathematically zero entries happen to cancel to exactly zero in the reference (csrc/exact.cuh)
    assert r["t2_idx_equal"], r
    assert r["t2_rel"] <= TOL, r
    assert r["rhs_rel_t2"] <= TOL


def _small(mn, dirichlet, sch, **kw):
    return mesh.build_model(f"small_mn{mn}", 6, 5, mn, 1000., 1100., 900., 2, 2, 1, dirichlet=dirichlet, gpml_sch=sch, freqs=(0.5, 3.0),
                            sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0, **kw)


def test_device_tables_bitwise_equal_oracle():
    """The product's reference-element tables (N, dN, phi, dphi, Gauss points incl. float32 literals and the
    n_fem.f90:193 typo) carry the same bits as the oracle's independent copy."""
    for mn in (8, 20, 27):
        m = _small(mn, 0, 1)
        asm, o = host.Assembly(m), Oracle(m)
        a, b = asm.debug_tables(), o.tables()
        for k in a:
            assert np.array_equal(a[k], b[k]), (mn, k)
        asm.close()


@pytest.mark.parametrize("mn", [8, 20, 27])
@pytest.mark.parametrize("dirichlet,sch", [(1, 1), (0, 0), (0, 1)])
def test_parity_small(mn, dirichlet, sch):
    m = _small(mn, dirichlet, sch)
    asm, o = host.Assembly(m), Oracle(m)
    assert np.array_equal(asm.gne(), o.gne())
    _check(compare_assembly(asm, o, m, ifreq=1, faithful=(mn == 8)))
    # second frequency of the sequential loop: Q12 (partial sigma update), Q17 (element 1 sees stale flags),
    # cached K_e/M_e of the unstretched elements
    _check(compare_assembly(asm, o, m, ifreq=2))
    asm.close()


@pytest.mark.parametrize("mn", [8, 20, 27])
@pytest.mark.parametrize("inimod", [2, 3])
def test_parity_dirichlet_boundary_models(mn, inimod):
    """Dirichlet boundary models 2 (homogeneous earth) and 3 (layered earth, Wait recursion): the primary field on the
    side faces is moved to the right-hand side through the Dirichlet columns of A_e (MoVFEM_3DMT.f90:252-261,
    boundary_conds.f90:188-250,436-598).  The matrix is untouched; the RHS must match the oracle and differ from the
    zero-boundary model's."""
    m = _small(mn, 1, 1)
    m.bd_inimod = inimod
    m.bd_hsigma = 0.02
    m.bd_lsigma, m.bd_ldz = (0.01, 0.1, 0.001), (1.0, 2.5)
    asm, o = host.Assembly(m), Oracle(m)
    for ifreq in (1, 2):
        r = compare_assembly(asm, o, m, ifreq=ifreq)
        _check(r)
    m1 = _small(mn, 1, 1)
    a1 = host.Assembly(m1)
    r2 = asm.global_vfem(1, m.omega(1), m.sigma_for(1), mode=abi.MODE_T1)
    r1 = a1.global_vfem(1, m1.omega(1), m1.sigma_for(1), mode=abi.MODE_T1)
    assert np.array_equal(r1[2], r2[2]) and not np.allclose(r1[3], r2[3])
    asm.close(); a1.close()


def test_parity_chunked_scratch(monkeypatch):
    """The Q|P,T scratch between geometry_kernel and contract_kernel is processed in chunks when an element list
    does not fit it (config 5 sizes).  Force a 1 MiB scratch so a small mesh runs through many chunks, ragged last
    batch included, and must reproduce the oracle (and the unchunked run) exactly as before."""
    m = _small(20, 0, 0)
    asm0 = host.Assembly(m)
    monkeypatch.setenv("MOVFEM_SCRATCH_MB", "1")
    asm1, o = host.Assembly(m), Oracle(m)
    monkeypatch.delenv("MOVFEM_SCRATCH_MB")
    _check(compare_assembly(asm1, o, m, ifreq=1))
    r0 = asm0.global_vfem(1, m.omega(1), m.sigma_for(1), mode=abi.MODE_T1)
    r1 = asm1.global_vfem(1, m.omega(1), m.sigma_for(1), mode=abi.MODE_T1)
    for x, y in zip(r0[:4], r1[:4]):
        assert np.array_equal(x, y)           # chunking changes nothing, bit for bit
    assert asm1.stats()["launches"] > asm0.stats()["launches"]
    asm0.close(); asm1.close()


def test_parity_anisotropic_sigma_and_mu():
    """config 3 (full 6-component sigma) plus a non-trivial permeability: exercises the curl part of the
    secondary source (problem.f90:362-420), which vanishes for mu = mu0."""
    m = mesh.config(3, scale=0.3)
    rng = np.random.default_rng(7)
    mur = 1.0 + 0.5 * rng.random(m.npt)
    m.g_mu[:, [0, 3, 5]] = mesh.MU0 * mur[:, None]
    m.g_mu[:, 1] = 0.05 * mesh.MU0 * rng.standard_normal(m.npt)
    asm, o = host.Assembly(m), Oracle(m)
    _check(compare_assembly(asm, o, m))
    asm.close()


def test_fused12_pipeline_many_batches_per_cta(monkeypatch):
    """fused12_kernel with two CTAs for the whole mesh (test hook): every CTA walks several 32-element batches, i.e. the
    double-buffered mbarrier pipeline (staged / full / empty) wraps around -- on BASELINE-size meshes only config 5 does."""
    monkeypatch.setenv("MOVFEM_TEST_FUSED_GRID", "2")
    m = mesh.build_model("fused_pipeline", 14, 9, 8, 1000., 1100., 900., 2, 2, 1, dirichlet=0, gpml_sch=1, freqs=(0.5, 3.0),
                         sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0)
    asm, o = host.Assembly(m), Oracle(m)
    _check(compare_assembly(asm, o, m, ifreq=1))
    _check(compare_assembly(asm, o, m, ifreq=2))
    asm.close()


def test_parity_linear_elements_unequal_diagonal_sigma():
    """Linear elements, mu = mu0, DIAGONAL sigma with unequal entries: the general variant of fused12_kernel (the isotropic
    one stands down on node_kernel's flag); the unmodified model runs the isotropic variant on frequency 1."""
    m = _small(8, 0, 1)
    m.sigma_re[:, 3] *= 1.5
    m.sigma_re[:, 5] *= 0.25
    asm, o = host.Assembly(m), Oracle(m)
    _check(compare_assembly(asm, o, m, ifreq=1))
    assert asm.stats()["ms_fused"] > 0.0
    _check(compare_assembly(asm, o, m, ifreq=2))
    asm.close()


@pytest.mark.parametrize("name", ["small_mn8_gpml_zhou", "small_mn8_dirichlet", "small_mn20_gpml_fang", "small_mn27_gpml_zhou"])
def test_golden_fixtures(name):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    m = make_golden.make_model(**make_golden.CASES[name])
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    asm = host.Assembly(m)
    assert np.array_equal(asm.gne(), ref["gne"]) and asm.nne == int(ref["nne"]) and asm.nnze == int(ref["nnze"])
    for ifreq in (1, 2):
        irn, jcn, a, rhs, nz = asm.global_vfem(ifreq, m.omega(ifreq), m.sigma_for(ifreq))
        assert nz == ref[f"a{ifreq}"].size
        assert np.array_equal(irn[:nz], ref[f"irn{ifreq}"]) and np.array_equal(jcn[:nz], ref[f"jcn{ifreq}"])
        assert rel_err(a[:nz], ref[f"a{ifreq}"]) <= TOL and rel_err(rhs, ref[f"rhs{ifreq}"]) <= TOL
    asm.close()


@pytest.mark.parametrize("dirichlet", [0, 1])
def test_config1_full_size(dirichlet):
    """BASELINE configs[0]: the shipped 58x58x43 example, GPML as shipped and Dirichlet."""
    m = mesh.config(1, dirichlet=dirichlet)
    asm, o = host.Assembly(m), Oracle(m)
    assert (asm.nne, asm.nnze) == ((450819, 14437635) if not dirichlet else (417411, 13355403))
    assert np.array_equal(asm.gne(), o.gne())
    _check(compare_assembly(asm, o, m))
    asm.close()


@pytest.mark.parametrize("n,scale", [(2, 0.3), (3, 0.35), (5, 0.1)])
def test_configs_scaled(n, scale):
    """configs 2, 3 and 5 (topography, exercises Q5) on sub-meshes the oracle finishes in seconds."""
    m = mesh.config(n, scale=scale)
    asm, o = host.Assembly(m), Oracle(m)
    _check(compare_assembly(asm, o, m))
    asm.close()


def test_config4_sweep_cached_equals_cold_and_oracle():
    """Frequency sweep on a reduced config 4: K_e/M_e cached across frequencies give the same bits
    as a cold assembly, and every frequency matches the oracle's sequential loop."""
    m = mesh.config(4, scale=0.12)
    asm, o = host.Assembly(m), Oracle(m)
    for ifreq in (1, 2, 17, 32):
        om, sg = m.omega(ifreq), m.sigma_for(ifreq)
        warm = asm.global_vfem(ifreq, om, sg)
        asm.reset_cache()
        cold = asm.global_vfem(ifreq, om, sg)
        nz = warm[4]
        assert nz == cold[4]
        for k in range(3):
            assert np.array_equal(warm[k][:nz], cold[k][:nz])
        assert np.array_equal(warm[3], cold[3])
        if ifreq > 1:
            o.set_in_pml((1, 1, 1))          # state the sequential loop would have left behind (Q17)
        r = o.assemble(om, sg)
        assert nz == r["nz"] and np.array_equal(warm[0][:nz], r["irn"]) and np.array_equal(warm[1][:nz], r["jcn"])
        assert rel_err(warm[2][:nz], r["a"]) <= TOL and rel_err(warm[3], r["rhs"]) <= TOL
    asm.close()


def test_sweep_gather_cache_is_bitwise_neutral(monkeypatch):
    """From the third frequency of a sweep on, entries that only unstretched elements touch take their gathered (K, M)
    from a cache instead of re-gathering; the delivered triplets must not change by a bit, also when Re(sigma) changes
    in mid-sweep (the cache is refilled on the device) -- compared with a handle that has the cache disabled."""
    m = mesh.config(4, scale=0.12)
    a1 = host.Assembly(m)
    monkeypatch.setenv("MOVFEM_NO_GATHER_CACHE", "1")
    a0 = host.Assembly(m)
    seq = [1, 2, 3, 4, 5, 6]
    for n, ifreq in enumerate(seq):
        om, sg = m.omega(ifreq), m.sigma_for(ifreq).copy()
        if n >= 4:
            sg[::7] *= 1.25                      # a different earth from the fifth call on: cached K/M must be refreshed
        r1 = a1.global_vfem(ifreq, om, sg)
        r0 = a0.global_vfem(ifreq, om, sg)
        nz = r1[4]
        assert nz == r0[4]
        for k in range(3):
            assert np.array_equal(r1[k][:nz], r0[k][:nz]), (ifreq, k)
        assert np.array_equal(r1[3], r0[3])
        if n in (2, 5):                          # and both equal a cold assembly of the same inputs
            a1.reset_cache()
            rc = a1.global_vfem(ifreq, om, sg)
            assert rc[4] == nz and np.array_equal(rc[2][:nz], r0[2][:nz])
    a1.close(); a0.close()


@pytest.mark.parametrize("n,stripped", [(2, 254), (3, 0)])
def test_full_size_parity_against_the_oracle(n, stripped):
    """BASELINE configs[1] (48000 20-node elements, GPML Fang) and configs[2] (10368 27-node elements, anisotropic sigma,
    GPML Zhou) at FULL size against the oracle (OpenMP over elements, memoised Jacobians: the same bits as its faithful
    mode).  Config 2 is the mesh on which the delivered pattern is decided by round-off: 1.71 M of its 24.7 M entries are
    residues of mathematically zero pairs, and exactly 254 of them cancel to (0,0) in the reference and are stripped."""
    m = mesh.config(n)
    asm, o = host.Assembly(m), Oracle(m)
    assert np.array_equal(asm.gne(), o.gne())
    r = compare_assembly(asm, o, m)
    _check(r)
    assert asm.nz_upper - r["t2_nz"][0] == stripped, r
    st = asm.stats()
    # config 2: ~36 residue pairs in each of the 48000 elements go through the reference-order kernel; config 3 only has
    # them in the air (sigma = 0: M_e vanishes, so the pairs whose K_e is a residue decide on their own)
    assert st["nflagged"] > (1_000_000 if n == 2 else -1), st
    asm.close()


def test_config5_submesh_full_topography():
    """BASELINE configs[4] on the 40x40x20 sub-mesh SURVEY 8d names (the reference cannot index the full 400x400x200,
    Q16) with the FULL 300 m topography amplitude and nextd = 4 (exercises the n_fem.f90:193 typo, Q5)."""
    m = mesh.config5_submesh()
    assert (m.g_nx - 1, m.g_ny - 1, m.g_nz - 1) == (40, 40, 20) and m.nextd == 4
    asm, o = host.Assembly(m), Oracle(m)
    _check(compare_assembly(asm, o, m, faithful=True))
    asm.close()


def test_cross_element_doubt_path(monkeypatch):
    """The second line of defence of the exact-zero set: an entry that cancels ACROSS elements is flagged by the gather
    and its contributions are re-evaluated in the reference's order (api.cu: movfem_device_result loop).  No BASELINE
    mesh has such entries, so the path is driven by a test hook: element-level flagging off, and the gather told to
    doubt every entry below an absolute size.  The delivered pattern must still be the oracle's."""
    m = _small(20, 0, 0)
    monkeypatch.setenv("MOVFEM_TEST_NO_L1", "1")
    monkeypatch.setenv("MOVFEM_TEST_DOUBT_ABS", "1e-6,1e-9")
    asm, o = host.Assembly(m), Oracle(m)
    r = compare_assembly(asm, o, m)
    _check(r)
    assert asm.stats()["nflagged"] == 0          # nothing came from the element-level test ...
    asm.close()
    monkeypatch.delenv("MOVFEM_TEST_NO_L1"); monkeypatch.delenv("MOVFEM_TEST_DOUBT_ABS")
    a2 = host.Assembly(m)
    irn, jcn, a, rhs, nz = a2.global_vfem(1, m.omega(1), m.sigma_for(1))
    assert nz == r["t2_nz"][1] and a2.stats()["nflagged"] > 0    # ... and the normal path flags at the element level
    a2.close()


def test_full_size_properties_config2():
    """BASELINE configs[1] at full size (48000 20-node elements), through size-independent properties:
    determinism (bit-identical reruns), T2 == float32 round trip of T1 minus exact zeros, sorted upper
    triangle, nz == (nnze+nne)/2 when nothing is stripped, K/M cache idempotence."""
    m = mesh.config(2)
    asm = host.Assembly(m)
    om, sg = m.omega(1), m.sigma_for(1)
    i1, j1, a1, b1, n1 = [x.copy() if isinstance(x, np.ndarray) else x for x in asm.global_vfem(1, om, sg, mode=abi.MODE_T1)]
    i2, j2, a2, b2, n2 = asm.global_vfem(1, om, sg, mode=abi.MODE_T2)            # cached K/M
    asm.reset_cache()
    i3, j3, a3, b3, n3 = [x.copy() if isinstance(x, np.ndarray) else x for x in asm.global_vfem(1, om, sg, mode=abi.MODE_T2)]
    assert n1 == asm.nz_upper == (asm.nnze + asm.nne) // 2
    assert n2 == n3 and np.array_equal(a2[:n2], a3[:n3]) and np.array_equal(b2, b3)
    r32 = a1.real.astype(np.float32).astype(np.float64) + 1j * a1.imag.astype(np.float32).astype(np.float64)
    keep = r32 != 0
    assert n2 == int(keep.sum()) and np.array_equal(a2[:n2], r32[keep])
    assert np.array_equal(i2[:n2], i1[keep]) and np.array_equal(j2[:n2], j1[keep])
    key = i1.astype(np.int64) * (asm.nne + 1) + j1
    assert np.all(np.diff(key) > 0) and np.all(i1 <= j1) and i1.min() == 1 and j1.max() == asm.nne
    assert np.all(np.isfinite(a1.view(np.float64))) and np.all(np.isfinite(b1.view(np.float64)))
    asm.close()


def test_error_codes_replace_stops():
    m = _small(8, 1, 1)
    asm = host.Assembly(m)
    sg = m.sigma_for(1)
    bad = sg.copy(); bad[5, :] = 0                      # singular sigma tensor: problem.f90:260-265 'stop'
    with pytest.raises(host.MovfemError) as ei:
        asm.global_vfem(1, m.omega(1), bad)
    assert ei.value.code == abi.MOVFEM_E_SINGULAR_MODEL
    irn, jcn, a, rhs, nz = asm.global_vfem(1, m.omega(1), sg)      # the handle stays usable
    assert nz == asm.nz_upper
    with pytest.raises(host.MovfemError) as ei:
        asm.global_vfem(1, m.omega(1), sg, irn=np.zeros(3, np.int32))
    assert ei.value.code == abi.MOVFEM_E_CAPACITY
    asm.close()
    m2 = _small(8, 1, 1)
    m2.g_zp[:] = 0.0                                    # flat mesh: det J = 0 -> n_fem.f90:374-377 'stop'
    asm = host.Assembly(m2)
    with pytest.raises(host.MovfemError) as ei:
        asm.global_vfem(1, m2.omega(1), m2.sigma_for(1))
    assert ei.value.code == abi.MOVFEM_E_SINGULAR_JAC
    asm.close()


def test_keep_pattern_mode_skips_static_irn_jcn():
    """MOVFEM_MODE_KEEP_PATTERN: when the caller passes the SAME irn/jcn arrays the previous call filled and the delivered set
    is unchanged (same nz, same signature of the stripped entries), they are not sent again; other arrays -- or a first call
    -- get the pattern.  Also on a mesh whose zero strip removes entries (20-node elements, GPML Fang: compacted pattern)."""
    for sch in (1, 0):
        m = _small(20, 0, sch)
        asm = host.Assembly(m)
        full = asm.global_vfem(1, m.omega(1), m.sigma_for(1), mode=abi.MODE_T2)
        nz = full[4]
        irn, jcn = full[0].copy(), full[1].copy()
        a = np.empty(asm.nz_upper, np.complex128); rhs = np.empty(2 * asm.nne, np.complex128)
        r0 = asm.global_vfem(1, m.omega(1), m.sigma_for(1), mode=abi.MODE_T2 | abi.MODE_KEEP_PATTERN, irn=irn, jcn=jcn, a=a, rhs=rhs)
        assert r0[4] == nz and np.array_equal(irn[:nz], full[0][:nz])       # other arrays than the last call's: delivered
        irn[:] = -7; jcn[:] = -7                                            # same arrays again: left alone
        r1 = asm.global_vfem(1, m.omega(1), m.sigma_for(1), mode=abi.MODE_T2 | abi.MODE_KEEP_PATTERN, irn=irn, jcn=jcn, a=a, rhs=rhs)
        assert r1[4] == nz and np.all(irn == -7) and np.all(jcn == -7)
        assert np.array_equal(a[:nz], full[2][:nz]) and np.array_equal(rhs, full[3])
        r2 = asm.global_vfem(1, m.omega(1), m.sigma_for(1), mode=abi.MODE_T2, irn=irn, jcn=jcn, a=a, rhs=rhs)   # without the flag: always delivered
        assert np.array_equal(irn[:nz], full[0][:nz]) and np.array_equal(jcn[:nz], full[1][:nz]) and r2[4] == nz
        asm.close()


def test_pageable_and_pinned_caller_arrays_deliver_the_same_bits():
    """movfem_assemble into pageable arrays (what a Fortran allocate gives: pinned staging ring, complex64 over the link and
    widening on the host) and into pinned arrays (direct copies) must deliver identical triplets."""
    import torch
    m = _small(27, 0, 0)
    asm = host.Assembly(m)
    om, sg = m.omega(1), m.sigma_for(1)
    page = asm.global_vfem(1, om, sg, mode=abi.MODE_T2)
    nz = page[4]
    pin = lambda n, dt: torch.empty(n, dtype=dt, pin_memory=True).numpy()   # noqa: E731
    p_irn, p_jcn = pin(asm.nz_upper, torch.int32), pin(asm.nz_upper, torch.int32)
    p_a, p_rhs = pin(2 * asm.nz_upper, torch.float64).view(np.complex128), pin(4 * asm.nne, torch.float64).view(np.complex128)
    pinned = asm.global_vfem(1, om, sg, mode=abi.MODE_T2, irn=p_irn, jcn=p_jcn, a=p_a, rhs=p_rhs)
    assert pinned[4] == nz
    for k in range(3):
        assert np.array_equal(page[k][:nz], pinned[k][:nz]), k
    assert np.array_equal(page[3], pinned[3])
    t1 = asm.global_vfem(1, om, sg, mode=abi.MODE_T1)                       # double values: never narrowed
    t1p = asm.global_vfem(1, om, sg, mode=abi.MODE_T1, irn=p_irn, jcn=p_jcn, a=p_a, rhs=p_rhs)
    assert np.array_equal(t1[2][: t1[4]], t1p[2][: t1p[4]])
    asm.close()


def test_device_resident_api():
    import torch
    m = _small(20, 0, 0)
    asm = host.Assembly(m)
    stream = torch.cuda.Stream()
    asm.set_stream(stream.cuda_stream)
    sg = m.sigma_for(1)
    d_sigma = torch.from_numpy(sg.view(np.float64).reshape(-1).copy()).cuda()
    asm.assemble_device(1, m.omega(1), d_sigma.data_ptr(), abi.MODE_T2)
    p_irn, p_jcn, p_a, p_rhs, nz = asm.device_result()
    irn, jcn, a, rhs, nz2 = asm.global_vfem(1, m.omega(1), sg)
    assert nz == nz2 and p_a and p_rhs
    import ctypes
    buf = np.empty(nz, np.complex128)
    torch.cuda.synchronize()
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rc = rt.cudaMemcpy(buf.ctypes.data, p_a, nz * 16, 2)
    assert int(rc) == 0 and np.array_equal(buf, a[:nz])
    # CSR hand-off: row pointers of the sorted upper-triangle triplets
    p_rp, nrows = asm.device_csr()
    rp = np.empty(nrows + 1, np.int64)
    assert int(rt.cudaMemcpy(rp.ctypes.data, p_rp, rp.nbytes, 2)) == 0 and nrows == asm.nne
    assert np.array_equal(rp, np.searchsorted(irn[:nz], np.arange(1, nrows + 2), side="left"))
    assert rp[0] == 0 and rp[-1] == nz and np.all(jcn[rp[:-1]] == np.arange(1, nrows + 1))   # every row starts at its diagonal
    # ... and a device consumer of that view: y = A x for the complex symmetric matrix whose upper triangle was delivered
    import scipy.sparse as sp
    rng = np.random.default_rng(7)
    x = rng.standard_normal(asm.nne) + 1j * rng.standard_normal(asm.nne)
    d_x = torch.from_numpy(x.view(np.float64).copy()).cuda()
    d_y = torch.zeros(2 * asm.nne, dtype=torch.float64, device="cuda")
    ms = asm.device_spmv(d_x.data_ptr(), d_y.data_ptr())
    U = sp.coo_matrix((a[:nz], (irn[:nz] - 1, jcn[:nz] - 1)), shape=(asm.nne, asm.nne)).tocsr()
    want = U @ x + U.T @ x - U.diagonal() * x
    got = d_y.cpu().numpy().view(np.complex128)
    assert ms > 0 and rel_err(got, want) <= 1e-13
    asm.close()


@pytest.mark.parametrize("mn,dirichlet", [(8, 0), (20, 1), (27, 0)])
def test_x_slab_handles_concatenate_to_the_full_assembly(mn, dirichlet):
    """SURVEY 8e slab sharding: handles that own ranges of ie (plus a +x halo they compute themselves) deliver
    disjoint, contiguous row ranges; concatenated in slab order they are bit-identical to the single-handle result."""
    import copy
    from movfem_b200.sharding import slab_partition
    m = mesh.build_model("slab", 7, 4, mn, 1000., 1100., 900., 2, 2, 1, dirichlet=dirichlet, gpml_sch=0, freqs=(0.5,),
                         sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=30.0)
    om, sg = m.omega(1), m.sigma_for(1)
    full = host.Assembly(m)
    irn, jcn, a, rhs, nz = full.global_vfem(1, om, sg)
    world = 3
    parts, rhs_acc, row_next = [], np.zeros(2 * full.nne, np.complex128), 1
    for r in range(world):
        ms = copy.copy(m)
        ms.ie_lo, ms.ie_hi = slab_partition(m.g_nx - 1, r, world)
        sl = host.Assembly(ms)
        assert sl.nne == full.nne and sl.row_lo == row_next
        row_next += sl.nrows
        i2, j2, a2, _, n2 = sl.global_vfem(1, om, sg, rhs=rhs_acc)
        assert n2 == sl.nz_upper and (n2 == 0 or (i2[:n2].min() >= sl.row_lo and i2[:n2].max() < sl.row_lo + sl.nrows))
        parts.append((i2[:n2].copy(), j2[:n2].copy(), a2[:n2].copy()))
        sl.close()
    assert row_next == full.nne + 1
    assert np.array_equal(np.concatenate([p[0] for p in parts]), irn[:nz])
    assert np.array_equal(np.concatenate([p[1] for p in parts]), jcn[:nz])
    assert np.array_equal(np.concatenate([p[2] for p in parts]), a[:nz])
    assert np.array_equal(rhs_acc, rhs)
    full.close()


def test_config4_full_size_slabs_and_sweep_properties():
    """BASELINE configs[3] at full size (100x100x60 linear elements, 1.84 M unknowns, 30.8 M entries), where the oracle
    is too slow: size-independent properties.  Two x-slab handles concatenate to the single-handle result bit for bit
    (checksums of IRN/JCN and exact equality of the values), the structural invariants hold (sorted upper triangle,
    every row starts at its diagonal, nz = (nnze + nne)/2), and a cached later frequency equals its cold assembly."""
    import copy
    from movfem_b200.sharding import slab_partition
    m = mesh.config(4)
    full = host.Assembly(m)
    om, sg = m.omega(3), m.sigma_for(3)
    full.global_vfem(1, m.omega(1), m.sigma_for(1))                 # first frequency of the sweep: fills the K/M cache
    irn, jcn, a, rhs, nz = full.global_vfem(3, om, sg)              # a cached frequency
    assert nz == full.nz_upper == (full.nnze + full.nne) // 2
    key = irn[:nz].astype(np.int64) * (full.nne + 1) + jcn[:nz]
    assert np.all(np.diff(key) > 0) and np.all(jcn[:nz] >= irn[:nz])
    first = np.searchsorted(irn[:nz], np.arange(1, full.nne + 1))
    assert np.array_equal(jcn[first], np.arange(1, full.nne + 1))
    full.reset_cache()
    cold = full.global_vfem(3, om, sg)
    assert cold[4] == nz and np.array_equal(cold[2][:nz], a[:nz]) and np.array_equal(cold[3], rhs)
    full.close()
    rhs_acc, parts = np.zeros(2 * (rhs.size // 2), np.complex128), []
    for r in range(2):
        ms = copy.copy(m)
        ms.ie_lo, ms.ie_hi = slab_partition(m.g_nx - 1, r, 2)
        sl = host.Assembly(ms)
        i2, j2, a2, _, n2 = sl.global_vfem(3, om, sg, rhs=rhs_acc)
        parts.append((i2[:n2].copy(), j2[:n2].copy(), a2[:n2].copy()))
        sl.close()
    assert sum(p[0].size for p in parts) == nz
    assert np.array_equal(np.concatenate([p[0] for p in parts]), irn[:nz]) and np.array_equal(np.concatenate([p[1] for p in parts]), jcn[:nz])
    assert np.array_equal(np.concatenate([p[2] for p in parts]), a[:nz]) and np.array_equal(rhs_acc, rhs)


@pytest.mark.parametrize("mn,dirichlet,inimod", [(8, 0, 1), (8, 1, 1), (20, 1, 1), (27, 0, 1), (8, 1, 3), (20, 1, 2)])
def test_end_to_end_apparent_resistivity_and_phase(mn, dirichlet, inimod):
    """north_star: 'End-to-end apparent resistivity and phase after the unchanged solve must agree to <= 1e-6 relative.'
    Graft and oracle triplets (tap T2 = what ZMUMPS receives, and the double-precision tap T1) go through the same
    stand-in sparse LU and the same solution.f90 post-processing (tests/e2e_util.py)."""
    from e2e_util import e2e_model, rho_phi_diff, solve_upper_triplets
    m = e2e_model(mn, dirichlet)
    m.bd_inimod, m.bd_hsigma, m.bd_lsigma, m.bd_ldz = inimod, 0.01, (0.01, 0.1), (1.5,)
    asm, o = host.Assembly(m), Oracle(m)
    om, sg = m.omega(1), m.sigma_for(1)
    ro = o.assemble(om, sg)
    # T2
    irn, jcn, a, rhs, nz = asm.global_vfem(1, om, sg, mode=abi.MODE_T2)
    assert nz == ro["nz"] and np.array_equal(irn[:nz], ro["irn"]) and np.array_equal(jcn[:nz], ro["jcn"])
    xg = solve_upper_triplets(asm.nne, irn[:nz], jcn[:nz], a[:nz], rhs)
    xo = solve_upper_triplets(o.nne, ro["irn"], ro["jcn"], ro["a"], ro["rhs"])
    sg_, so_ = o.node_solution(om, sg, xg), o.node_solution(om, sg, xo)
    drho, dphi, same, n = rho_phi_diff(so_, sg_)
    assert same and n > 100 and drho <= 1e-6 and dphi <= 1e-6 * 90.0, (drho, dphi, same, n)
    # T1 (double values, upper triangle in delivery order) against the oracle's T1 (lower triangle of the full pattern)
    irn1, jcn1, a1, rhs1, nz1 = asm.global_vfem(1, om, sg, mode=abi.MODE_T1)
    ia, ja = o.pattern()
    keep = ia >= ja
    xg1 = solve_upper_triplets(asm.nne, irn1[:nz1], jcn1[:nz1], a1[:nz1], rhs1)
    xo1 = solve_upper_triplets(o.nne, ja[keep], ia[keep], ro["a_t1"][keep], ro["rhs"])
    drho1, dphi1, same1, n1 = rho_phi_diff(o.node_solution(om, sg, xo1), o.node_solution(om, sg, xg1))
    assert same1 and drho1 <= 1e-6 and dphi1 <= 1e-6 * 90.0, (drho1, dphi1)
    assert np.linalg.norm(xg1 - xo1) <= 1e-9 * np.linalg.norm(xo1)
    asm.close()


@pytest.mark.parametrize("mn", [8, 20, 27])
def test_fang_sweep_keeps_the_stretched_elements_cached(mn):
    """GPML scheme 0 over four frequencies of the sequential loop.  The reference stores Re(h) = 1 + a0*rho^n (Q18), which
    does not depend on omega, so K_e/M_e of the stretched elements are cached like the others: frequency 1 is a cold
    pass, frequency 2 recomputes once because element (1,1,1) now sees the flags of the last element (Q17), later
    frequencies form only the right-hand sides -- and every frequency must match the oracle's sequential run."""
    m = mesh.build_model(f"fang_sweep_mn{mn}", 6, 5, mn, 1000., 1100., 900., 2, 2, 1, dirichlet=0, gpml_sch=0, a0=1.5, b0=0.8, nn=2.0,
                         freqs=(0.5, 3.0, 7.0, 20.0), sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0)
    asm, o = host.Assembly(m), Oracle(m)
    for ifreq in (1, 2, 3, 4):
        _check(compare_assembly(asm, o, m, ifreq=ifreq))
    # the cached pass of frequency 4 equals a cold pass of a fresh handle at the same position of the loop, bit for bit
    cached = asm.global_vfem(4, m.omega(4), m.sigma_for(4), mode=abi.MODE_T1)
    fresh = host.Assembly(m)
    cold = fresh.global_vfem(4, m.omega(4), m.sigma_for(4), mode=abi.MODE_T1)
    for x, y in zip(cached[:4], cold[:4]):
        assert np.array_equal(x, y)
    asm.close(); fresh.close()

"""CPU tests: the C-ABI library loads and exports what include/movfem_b200.h declares; host-side logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from movfem_b200 import abi, host, mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "movfem_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(movfem_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(host.build())
    names = _header_functions()
    assert "movfem_create" in names and "movfem_assemble" in names and len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/movfem_b200.h but not exported"
    assert set(host.EXPORTED_SYMBOLS) <= set(names)
    assert b"sm_100a" in host.lib().movfem_version()
    if not os.environ.get("MOVFEM_B200_LIB"):     # the in-tree library is the default configuration, never an A/B variant
        assert b"A/B" not in host.lib().movfem_version()


def test_desc_layout_matches_header():
    """ctypes mirror vs the C struct: field order from the header, size from the compiler's rules."""
    txt = open(os.path.join(ROOT, "include", "movfem_b200.h")).read()
    body = re.search(r"typedef struct movfem_desc \{(.*?)\} movfem_desc;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.replace("const double *", "").replace("int32_t", "").replace("double", "").split(",")
        fields += [re.sub(r"\[\d+\]", "", n).strip().lstrip("*") for n in names]
    assert fields == [f for f, _ in abi.MovfemDesc._fields_]
    assert C.sizeof(abi.MovfemDesc) == 14 * 4 + 3 * 8 + 4 * 8 + 2 * 4 + 2 * 8 + 2 * 4 + 32 * 8


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(host.MovfemError) as ei:
        host.Assembly(mesh.config(1, scale=0.2))
    assert ei.value.code == abi.MOVFEM_E_NOGPU


def test_bad_descriptors_are_rejected_before_touching_cuda():
    m = mesh.config(1, scale=0.2)
    d = m.desc()
    h = C.c_void_p()
    d.me = 13
    assert host.lib().movfem_create(C.byref(d), 0, C.byref(h)) == abi.MOVFEM_E_BADARG
    d = m.desc(); d.bd_inimod = 4; d.dirichlet = 1
    assert host.lib().movfem_create(C.byref(d), 0, C.byref(h)) == abi.MOVFEM_E_BADARG
    d = m.desc(); d.bd_inimod = 3; d.dirichlet = 1; d.bd_nl = 17          # more layers than the descriptor holds
    assert host.lib().movfem_create(C.byref(d), 0, C.byref(h)) == abi.MOVFEM_E_BADARG
    d = m.desc(); d.ndir = 1
    assert host.lib().movfem_create(C.byref(d), 0, C.byref(h)) == abi.MOVFEM_E_UNSUPPORTED
    assert host.lib().movfem_create(None, 0, C.byref(h)) == abi.MOVFEM_E_BADARG


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under movfem_b200/ may import, link or call it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|oracle_[a-z]+\s*\(|#include\s+\"[^\"]*oracle", re.M)
    for dp, _, files in os.walk(os.path.join(ROOT, "movfem_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".f90")) or f == "Makefile":
                s = open(os.path.join(dp, f), errors="ignore").read()
                assert not pat.search(s), f


def test_config_sizes_match_survey():
    for n, (nx, ny, nz, mn) in {1: (58, 58, 43, 8), 2: (40, 40, 30, 20), 3: (24, 24, 18, 27), 4: (100, 100, 60, 8)}.items():
        m = mesh.config(n)
        assert (m.g_nx - 1, m.g_ny - 1, m.g_nz - 1, m.mn) == (nx, ny, nz, mn)
        assert m.g_zp.size == m.g_xp.size * m.g_yp.size * ((m.g_nz - 1) * (m.nord - 1) + 1)
    m = mesh.config(1)
    dx = np.diff(m.g_xp)
    np.testing.assert_allclose(dx[:4], 1.3 * 1990 * np.array([4, 3, 2, 1]))     # geometry.f90:268-317
    np.testing.assert_allclose(dx[4:-4], 1990.0)
    assert mesh.config(4).freqs.size == 32


def test_update_sigma_reproduces_q12():
    """geometry.f90:144-153 indexes g_sigma(i,j), i<=npt, j<=6 on a (6,npt) array: only the first npt+30
    linear entries are refreshed; real parts never change."""
    m = mesh.config(4, scale=0.12)
    s1, s5 = m.sigma_for(1), m.sigma_for(5)
    assert np.array_equal(s1.real, s5.real)
    flat1, flat5 = s1.reshape(-1), s5.reshape(-1)
    n = m.npt + 30
    changed = np.flatnonzero(flat1.imag != flat5.imag)
    assert changed.size > 0 and changed.max() < n
    w5 = np.float64(np.float32(mesh.EPS0 * m.omega(5)))
    assert np.all(flat5.imag[:n][flat5.imag[:n] != 0] == w5)
    assert np.all(flat5.imag[n:][flat5.imag[n:] != 0] == np.float64(np.float32(mesh.EPS0 * m.omega(1))))


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """bench.py --impl reference (the reference's CPU algorithm through the oracle port) runs without a GPU and prints
    exactly one JSON line with the keys the driver reads."""
    import json, subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-elements", "40"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_step_roofs_uses_the_survey_figures():
    """bench.py's roofline table (both fractions for the step and for its kernels) is a pure function: check it
    on round-1's measured phase times of config 2 -- the contraction must reproduce that round's headline FP64 fraction."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    ph = {"ms_node": 0.0436, "ms_geometry": 0.2878, "ms_contract": 0.4337, "ms_fused": 0.0, "ms_exact": 0.0, "ms_gather": 0.3344}
    r = b.step_roofs(36, 48000, 24676330, ph, 1.169, 34.87, 6556.5)
    assert abs(r["flop_kernels"]["tflops"] - 250776 * 48000 / 0.4337e-3 * 1e-12) < 1e-9
    assert 0.79 < r["flop_kernels"]["frac_fp64"] < 0.80 and r["flop_kernels"]["kernels"] == "contract_kernel"
    assert abs(r["gather"]["gbs"] - 24676330 * 32 / 0.3344e-3 * 1e-9) < 1e-6
    assert abs(r["step"]["tflops"] - 250776 * 48000 / 1.169e-3 * 1e-12) < 1e-9 and 0.29 < r["step"]["frac_fp64"] < 0.30
    assert abs(r["step"]["gbs"] - 9900 * 48000 / 1.169e-3 * 1e-9) < 1e-6
    assert abs(sum(r["share_of_step"].values()) - (0.0436 + 0.2878 + 0.4337 + 0.3344) / 1.169) < 1e-12
    # linear elements: the fused kernel carries the flops
    r12 = b.step_roofs(12, 1000, 51000, {"ms_fused": 0.5, "ms_contract": 0.1, "ms_gather": 0.2}, 1.0, 30.0, 6000.0)
    assert r12["flop_kernels"]["kernels"].startswith("fused12_kernel") and abs(r12["flop_kernels"]["ms"] - 0.6) < 1e-12
    # the top-level `roofline` repeats the object of the kernels with the larger share of the step
    assert b.dominant_roofline(17.7, 22.6) == "roofline_hbm" and b.dominant_roofline(0.44, 0.33) == "roofline_fp64"
    assert b.dominant_roofline(None, 1.0) == "roofline_hbm" and b.dominant_roofline(1.0, None) == "roofline_fp64"


def test_global_vfem_refuses_arrays_that_are_not_the_callers_fortran_types():
    """host.Assembly.global_vfem hands raw pointers to the library: wrong dtype, strided or read-only arrays and short
    arrays are rejected in Python, before any pointer crosses the ABI (no handle, no GPU needed to see that)."""
    m = mesh.config(1, scale=0.2)
    asm = host.Assembly.__new__(host.Assembly)          # no movfem_create: only the argument checks run
    asm.model, asm.nne, asm.nz_upper, asm._h = m, 10, 20, None
    sg = m.sigma_for(1)
    good = dict(irn=np.zeros(20, np.int32), jcn=np.zeros(20, np.int32), a=np.zeros(20, np.complex128), rhs=np.zeros(20, np.complex128))
    ro = np.zeros(20, np.int32); ro.flags.writeable = False
    for key, bad in (("irn", np.zeros(20, np.int64)), ("jcn", np.zeros(40, np.int32)[::2]), ("a", np.zeros(40, np.float64)),
                     ("rhs", np.zeros(20, np.complex64)), ("irn", ro), ("a", [0j] * 20)):
        with pytest.raises(host.MovfemError) as ei:
            asm.global_vfem(1, 1.0, sg, **{**good, key: bad})
        assert ei.value.code == abi.MOVFEM_E_BADARG, key
    with pytest.raises(host.MovfemError) as ei:
        asm.global_vfem(1, 1.0, sg, **{**good, "rhs": np.zeros(19, np.complex128)})
    assert ei.value.code == abi.MOVFEM_E_CAPACITY
    with pytest.raises(host.MovfemError) as ei:
        asm.global_vfem(1, 1.0, sg[:-1], **good)
    assert ei.value.code == abi.MOVFEM_E_BADARG


def test_sass_of_the_library_uses_the_sm100_paths_the_design_names():
    """Static check on the built library (cuobjdump -sass, no GPU): the hardware paths DESIGN.md claims are in the code that
    ships -- TMA bulk copies + mbarriers in the generic element kernels, the FP64 tensor-core path in the interpolation phase of
    geometry_kernel, cp.async + mbarriers and no local-memory spills in the linear-element kernel, cp.async in the gather."""
    import shutil, subprocess, sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_opmix.py")], capture_output=True, text=True, timeout=300).stdout
    cols, rows = None, {}
    for line in out.splitlines():
        cells = [c.strip() for c in line.strip().strip("|").split("|")]
        if cells and cells[0] == "kernel":
            cols = cells
        elif cols and len(cells) == len(cols) and cells[0].startswith("`"):
            rows[cells[0].strip("`")] = {c: int(v) if v else 0 for c, v in zip(cols[1:], cells[1:])}
    def pick(sub):
        hit = [v for k, v in rows.items() if sub in k]
        assert hit, (sub, sorted(rows))
        return hit
    for r in pick("contract_kernel"):
        assert r["UBLKCP"] > 0 and r["SYNCS"] > 0 and r["DFMA"] > 100 and r["STL"] <= 2, r      # (the GPML variant spills one register)
    assert any(r["DMMA"] > 0 and r["UBLKCP"] > 0 for r in pick("geometry_kernel"))
    fused = pick("fused12_kernel<true, true>")[0]
    assert fused["LDGSTS"] > 0 and fused["SYNCS"] > 0 and fused["DFMA"] > 500 and fused["STL"] == 0 and fused["LDL"] == 0, fused
    assert all(r["LDGSTS"] > 0 for r in pick("gather_finalize_kernel"))

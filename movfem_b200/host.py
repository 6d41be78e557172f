"""Host-side mirror of the reference interface for the assembly path, over the C ABI.

The reference is Fortran and no Fortran toolchain exists in this image, so the host side above
``include/movfem_b200.h`` is written in Python/ctypes with the reference's conventions (1-based
index values, column-major ``gne(ne,me)``, caller-owned ``irn/jcn/a/rhs``) and the reference's
names and error behaviour:

    ga_init        (global_assembly.f90:26-41)   -> :class:`Assembly` constructor: ``nne``, ``nnze``, ``gne``
    global_vfem    (MoVFEM_3DMT.f90:167-216)     -> :meth:`Assembly.global_vfem`
    find_zeros / rem_zeros (MoVFEM_3DMT.f90:85-97) are folded into the same call (``nz`` is returned)

The reference ``stop``s on inconsistencies; here they raise :class:`MovfemError` carrying the
C-ABI code.  There is no CPU fallback: if ``libmovfem_b200.so`` is missing or no GPU is visible,
construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import abi
from .abi import MovfemDesc, MovfemStats

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("MOVFEM_B200_LIB") or os.path.join(_HERE, "libmovfem_b200.so")   # override: A/B builds of the same ABI
_LIB = None


class MovfemError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"movfem_b200 error {code} ({abi.ERROR_NAMES.get(code, '?')}) {msg}")


def build(force: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)] + [os.path.join(_HERE, "..", "include", "movfem_b200.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", src_dir, "-s"])
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise MovfemError(abi.MOVFEM_E_NOGPU, f"{_SO} not built: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(_SO)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.movfem_create.argtypes = [C.POINTER(MovfemDesc), C.c_int, C.POINTER(vp)]
        L.movfem_destroy.argtypes = [vp]
        L.movfem_destroy.restype = None
        L.movfem_sizes.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
        L.movfem_slab_rows.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        L.movfem_get_gne.argtypes = [vp, vp]
        L.movfem_get_pattern.argtypes = [vp, vp, vp]
        L.movfem_assemble.argtypes = [vp, i32, dbl, vp, vp, vp, vp, vp, C.POINTER(i64), i32]
        L.movfem_assemble_device.argtypes = [vp, i32, dbl, vp, i32]
        L.movfem_device_result.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
        L.movfem_device_csr.argtypes = [vp, C.POINTER(vp), C.POINTER(i32)]
        L.movfem_device_spmv.argtypes = [vp, vp, vp, C.POINTER(dbl)]
        L.movfem_set_stream.argtypes = [vp, vp]
        L.movfem_get_stats.argtypes = [vp, C.POINTER(MovfemStats)]
        L.movfem_last_error.argtypes = [vp]
        L.movfem_last_error.restype = C.c_char_p
        L.movfem_version.restype = C.c_char_p
        L.movfem_reset_cache.argtypes = [vp]
        L.movfem_fp64_peak.argtypes = [C.c_int, C.POINTER(dbl)]
        L.movfem_debug_element.argtypes = [vp, i32, vp, vp, vp]
        L.movfem_debug_tables.argtypes = [vp, vp, vp, vp, vp, vp]
        L.movfem_geo_innermodel.argtypes = [C.POINTER(MovfemDesc), i32, C.POINTER(abi.MovfemGeomodel), dbl, vp, vp, C.POINTER(dbl)]
        _LIB = L
    return _LIB


EXPORTED_SYMBOLS = [
    "movfem_create", "movfem_destroy", "movfem_sizes", "movfem_get_gne", "movfem_get_pattern", "movfem_assemble",
    "movfem_assemble_device", "movfem_device_result", "movfem_device_csr", "movfem_device_spmv", "movfem_set_stream", "movfem_get_stats", "movfem_last_error",
    "movfem_version", "movfem_debug_element", "movfem_reset_cache", "movfem_fp64_peak", "movfem_slab_rows",
    "movfem_geo_innermodel",
]


def fp64_peak_tflops(device: int = 0) -> float:
    """Measured FP64 FMA-loop throughput (the FP64 roofline denominator)."""
    v = C.c_double(0)
    rc = lib().movfem_fp64_peak(device, C.byref(v))
    if rc:
        raise MovfemError(rc)
    return v.value


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def innermodel_gqg(model, nzl_air, omega, xm, ym, zm, ijsigma, sigma, ijmu, mu, device: int = 0):
    """geometry.f90 innermodel_gqg on the GPU: copies the input geomodel to the grid nodes.
    model: mesh.Model (its g_nx.., nord, nextd, nzl_top, g_xp, g_yp, g_zp are read); xm(mx), ym(my), zm(mx*my*mz);
    sigma (isigma, ncell) and mu (imu, ncell) as the Fortran arrays (column-major); ijsigma/ijmu (n, 2).
    Returns g_sigma (npt, 6) complex128, g_mu (npt, 6) float64 and the device time in ms."""
    xm, ym, zm = (np.ascontiguousarray(v, np.float64) for v in (xm, ym, zm))
    sigma = np.asfortranarray(sigma, np.float64)
    mu = np.asfortranarray(mu, np.float64)
    gm = abi.MovfemGeomodel()
    gm.mx, gm.my, gm.mz = xm.size, ym.size, zm.size // (xm.size * ym.size)
    gm.isigma, gm.imu, gm.nzl_air = sigma.shape[0], mu.shape[0], nzl_air
    for i, (r, c) in enumerate(np.asarray(ijsigma).reshape(-1, 2)):
        gm.ijsigma[i][0], gm.ijsigma[i][1] = int(r), int(c)
    for i, (r, c) in enumerate(np.asarray(ijmu).reshape(-1, 2)):
        gm.ijmu[i][0], gm.ijmu[i][1] = int(r), int(c)
    gm.xm, gm.ym, gm.zm, gm.sigma, gm.mu = _p(xm), _p(ym), _p(zm), _p(sigma), _p(mu)
    d = model.desc()
    g_sigma = np.zeros((model.npt, 6), np.complex128)
    g_mu = np.zeros((model.npt, 6), np.float64)
    ms = C.c_double(0)
    rc = lib().movfem_geo_innermodel(C.byref(d), device, C.byref(gm), float(omega), _p(g_sigma), _p(g_mu), C.byref(ms))
    if rc:
        raise MovfemError(rc)
    return g_sigma, g_mu, ms.value


class Assembly:
    """One mesh on one GPU: the graft's replacement for ga_init + global_vfem + zero stripping."""

    def __init__(self, model, device: int = 0):
        self.model = model
        self._desc = model.desc()
        self._h = C.c_void_p()
        rc = lib().movfem_create(C.byref(self._desc), device, C.byref(self._h))
        if rc:
            msg = lib().movfem_last_error(self._h).decode() if self._h else ""
            if self._h:
                lib().movfem_destroy(self._h)
                self._h = C.c_void_p()
            raise MovfemError(rc, msg)
        nne, nnze, nzu = C.c_int32(), C.c_int64(), C.c_int64()
        lib().movfem_sizes(self._h, C.byref(nne), C.byref(nnze), C.byref(nzu))
        self.nne, self.nnze, self.nz_upper = nne.value, nnze.value, nzu.value
        self.me, self.mn, self.ne = model.me, model.mn, model.ne
        lo, n = C.c_int32(), C.c_int32()
        lib().movfem_slab_rows(self._h, C.byref(lo), C.byref(n))
        self.row_lo, self.nrows = lo.value, n.value          # 1-based first owned row, count (x-slab handles)

    def close(self):
        if getattr(self, "_h", None):
            lib().movfem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise MovfemError(rc, lib().movfem_last_error(self._h).decode())

    # -- ga_cgne ---------------------------------------------------------------------------------
    def gne(self) -> np.ndarray:
        """gne[ide-1, im-1] (global_assembly.f90:183-195), Fortran values."""
        g = np.zeros((self.me, self.ne), np.int32)      # memory layout of Fortran gne(ne,me)
        self._check(lib().movfem_get_gne(self._h, _p(g)))
        return g.T

    def pattern(self):
        irn = np.zeros(self.nz_upper, np.int32); jcn = np.zeros(self.nz_upper, np.int32)
        self._check(lib().movfem_get_pattern(self._h, _p(irn), _p(jcn)))
        return irn, jcn

    # -- global_vfem + find_zeros/rem_zeros --------------------------------------------------------
    def global_vfem(self, freq_index: int, omega: float, g_sigma: np.ndarray, *, mode: int = abi.MODE_T2,
                    irn=None, jcn=None, a=None, rhs=None):
        """Assemble one frequency into caller-owned arrays (allocated here when not given, sized as
        the graft needs: ``nz_upper``; the reference allocates ``nnze``).  Returns (irn, jcn, a, rhs, nz)."""
        g_sigma = np.ascontiguousarray(g_sigma, np.complex128)
        if g_sigma.size != 6 * self.model.npt:
            raise MovfemError(abi.MOVFEM_E_BADARG, "g_sigma must be (6, npt)")
        irn = np.empty(self.nz_upper, np.int32) if irn is None else irn
        jcn = np.empty(self.nz_upper, np.int32) if jcn is None else jcn
        a = np.empty(self.nz_upper, np.complex128) if a is None else a
        rhs = np.empty(2 * self.nne, np.complex128) if rhs is None else rhs
        # the library writes through raw pointers: refuse anything that is not the Fortran caller's array types
        for name, arr, dt in (("irn", irn, np.int32), ("jcn", jcn, np.int32), ("a", a, np.complex128), ("rhs", rhs, np.complex128)):
            if not isinstance(arr, np.ndarray) or arr.dtype != dt or not arr.flags.c_contiguous or not arr.flags.writeable:
                raise MovfemError(abi.MOVFEM_E_BADARG, f"{name} must be a writable contiguous numpy array of {np.dtype(dt).name}")
        if min(irn.size, jcn.size, a.size) < self.nz_upper or rhs.size < 2 * self.nne:
            raise MovfemError(abi.MOVFEM_E_CAPACITY)
        nz = C.c_int64(0)
        self._check(lib().movfem_assemble(self._h, freq_index, omega, _p(g_sigma), _p(irn), _p(jcn), _p(a), _p(rhs),
                                          C.byref(nz), mode))
        return irn, jcn, a, rhs, nz.value

    # -- device-resident path (kernel-only timing, device consumers) ------------------------------
    def set_stream(self, cuda_stream_ptr: int):
        self._check(lib().movfem_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def assemble_device(self, freq_index: int, omega: float, sigma_dev_ptr: int, mode: int = abi.MODE_T2):
        self._check(lib().movfem_assemble_device(self._h, freq_index, omega, C.c_void_p(sigma_dev_ptr), mode))

    def device_result(self):
        irn, jcn, a, rhs = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nz = C.c_int64()
        self._check(lib().movfem_device_result(self._h, C.byref(irn), C.byref(jcn), C.byref(a), C.byref(rhs), C.byref(nz)))
        return irn.value, jcn.value, a.value, rhs.value, nz.value

    def device_csr(self):
        """(device pointer to rowptr[nrows+1], nrows): CSR row pointers of the last device result."""
        rp, n = C.c_void_p(), C.c_int32()
        self._check(lib().movfem_device_csr(self._h, C.byref(rp), C.byref(n)))
        return rp.value, n.value

    def device_spmv(self, x_dev_ptr: int, y_dev_ptr: int) -> float:
        """y = A x on the device from the last device result (complex128 device vectors of nne entries); returns the device ms."""
        ms = C.c_double(0)
        self._check(lib().movfem_device_spmv(self._h, C.c_void_p(x_dev_ptr), C.c_void_p(y_dev_ptr), C.byref(ms)))
        return ms.value

    def reset_cache(self):
        """Next call recomputes every element's K_e/M_e (cold single-frequency assembly)."""
        self._check(lib().movfem_reset_cache(self._h))

    def stats(self) -> dict:
        s = MovfemStats()
        self._check(lib().movfem_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in MovfemStats._fields_}

    # -- parity taps --------------------------------------------------------------------------------
    def debug_element(self, ide: int):
        NP = self.me * (self.me + 1) // 2
        K = np.zeros(NP); M = np.zeros(NP); b = np.zeros((self.me, 2), np.complex128)
        self._check(lib().movfem_debug_element(self._h, ide, _p(K), _p(M), _p(b)))
        Kf = np.zeros((self.me, self.me)); Mf = np.zeros((self.me, self.me))
        il = np.tril_indices(self.me)
        Kf[il] = K; Mf[il] = M
        Kf = Kf + np.tril(Kf, -1).T; Mf = Mf + np.tril(Mf, -1).T
        return Kf, Mf, b

    def debug_tables(self):
        g = 8 if self.me == 12 else 27
        N = np.zeros((g, self.mn)); dN = np.zeros((g, self.mn, 3)); phi = np.zeros((g, self.me)); dphi = np.zeros((g, self.me, 3)); rw = np.zeros((g, 4))
        self._check(lib().movfem_debug_tables(self._h, _p(N), _p(dN), _p(phi), _p(dphi), _p(rw)))
        return dict(N=N, dN=dN, phi=phi, dphi=dphi, rw=rw)

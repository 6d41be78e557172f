"""ctypes wrapper of the CPU oracle (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It restates MoVFEM_3DMT's assembly path on the CPU (oracle/movfem_oracle.cpp)
and is the checker the CUDA path is compared against; the product never calls it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from movfem_b200.abi import MovfemDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("movfem_oracle.cpp", "shape.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_create.argtypes = [C.POINTER(MovfemDesc), C.POINTER(C.c_void_p)]
        _LIB.oracle_destroy.argtypes = [C.c_void_p]
        _LIB.oracle_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _LIB.oracle_get_gne.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.oracle_get_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.oracle_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _LIB.oracle_shape_eval.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 4
        _LIB.oracle_element.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p] + [C.c_void_p] * 10
        _LIB.oracle_assemble.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64),
                                         C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        _LIB.oracle_get_in_pml.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.oracle_set_in_pml.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.oracle_effective_pml.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _LIB.oracle_node_solution.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 7
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def shape_eval(mn, me, xi, eta, zeta):
    N = np.zeros(mn); dN = np.zeros((mn, 3)); phi = np.zeros(me); dphi = np.zeros((me, 3))
    lib().oracle_shape_eval(mn, me, xi, eta, zeta, _p(N), _p(dN), _p(phi), _p(dphi))
    return N, dN, phi, dphi


class Oracle:
    """CPU restatement of ga_init + global_vfem + find/rem_zeros for one mesh."""

    def __init__(self, model):
        self.model = model
        self._desc = model.desc()
        h = C.c_void_p()
        rc = lib().oracle_create(C.byref(self._desc), C.byref(h))
        if rc:
            raise RuntimeError(f"oracle_create failed: {rc}")
        self._h = h
        nne, nnze, nzu = C.c_int32(), C.c_int64(), C.c_int64()
        lib().oracle_sizes(h, C.byref(nne), C.byref(nnze), C.byref(nzu))
        self.nne, self.nnze, self.nz_upper = nne.value, nnze.value, nzu.value
        self.me, self.mn, self.ne = model.me, model.mn, model.ne
        self.ngp = 8 if self.me == 12 else 27

    def __del__(self):
        try:
            lib().oracle_destroy(self._h)
        except Exception:
            pass

    def gne(self):
        g = np.zeros((self.me, self.ne), np.int32)       # Fortran gne(ne,me) column-major
        lib().oracle_get_gne(self._h, _p(g))
        return g.T                                        # [ide-1, im-1]

    def pattern(self):
        ia = np.zeros(self.nnze, np.int32); ja = np.zeros(self.nnze, np.int32)
        lib().oracle_get_pattern(self._h, _p(ia), _p(ja))
        return ia, ja

    def tables(self):
        g, mn, me = self.ngp, self.mn, self.me
        N = np.zeros((g, mn)); dN = np.zeros((g, mn, 3)); phi = np.zeros((g, me)); dphi = np.zeros((g, me, 3)); rw = np.zeros((g, 4))
        lib().oracle_tables(self._h, _p(N), _p(dN), _p(phi), _p(dphi), _p(rw))
        return dict(N=N, dN=dN, phi=phi, dphi=dphi, rw=rw)

    def in_pml(self):
        f = np.zeros(3, np.int32); lib().oracle_get_in_pml(self._h, _p(f)); return f

    def set_in_pml(self, f):
        f = np.ascontiguousarray(f, np.int32); lib().oracle_set_in_pml(self._h, _p(f))

    def effective_pml(self, ide):
        f = np.zeros(3, np.int32); lib().oracle_effective_pml(self._h, ide, _p(f)); return f

    def element(self, ide, omega, sigma, pml=(0, 0, 0), caches=False):
        me, g = self.me, self.ngp
        sigma = np.ascontiguousarray(sigma, np.complex128)
        pml = np.ascontiguousarray(pml, np.int32)
        Ae = np.zeros((me, me), np.complex128); be = np.zeros((me, 2), np.complex128)
        out = dict(A=Ae, b=be)
        extra = [None] * 8
        if caches:
            out.update(wgt=np.zeros(g), cve1=np.zeros((27 * 54, 3)), cve2=np.zeros((27 * 54, 3)), ve=np.zeros((27 * 54, 3)),
                       mf1=np.zeros((27, 6), np.complex128), mf2=np.zeros((27, 6), np.complex128),
                       gpml=np.zeros((27, 3), np.complex128), src=np.zeros((54, 3), np.complex128))
            extra = [out[k] for k in ("wgt", "cve1", "cve2", "ve", "mf1", "mf2", "gpml", "src")]
        rc = lib().oracle_element(self._h, ide, omega, _p(sigma), _p(pml), _p(Ae), _p(be), *[_p(x) for x in extra])
        if rc:
            raise RuntimeError(f"oracle_element failed: {rc}")
        if caches:   # the C side copies only the used prefix: row g+(e-1)*ngp
            for k in ("cve1", "cve2", "ve"):
                out[k] = out[k][: g * me]
            out["mf1"] = out["mf1"][:g]; out["mf2"] = out["mf2"][:g]; out["gpml"] = out["gpml"][:g]; out["src"] = out["src"][: 2 * g]
        return out

    def assemble(self, omega, sigma, *, faithful=False, nthreads=None, want_t1=True, want_t2=True,
                 ide_range=(0, 0)):
        """Returns dict(a_t1, irn, jcn, a, nz, rhs, seconds, jac_builds)."""
        sigma = np.ascontiguousarray(sigma, np.complex128)
        if nthreads is None:
            nthreads = 1 if faithful else (os.cpu_count() or 1)
        a_t1 = np.zeros(self.nnze, np.complex128) if want_t1 else None
        irn = np.zeros(self.nnze, np.int32) if want_t2 else None
        jcn = np.zeros(self.nnze, np.int32) if want_t2 else None
        a = np.zeros(self.nnze, np.complex128) if want_t2 else None
        rhs = np.zeros(2 * self.nne, np.complex128)
        nz = C.c_int64(0); secs = C.c_double(0); jb = C.c_int64(0)
        rc = lib().oracle_assemble(self._h, omega, _p(sigma), int(faithful), int(nthreads), ide_range[0], ide_range[1],
                                   _p(a_t1), _p(irn), _p(jcn), _p(a), C.byref(nz), _p(rhs), C.byref(secs), C.byref(jb))
        if rc:
            raise RuntimeError(f"oracle_assemble failed: {rc}")
        n = nz.value
        return dict(a_t1=a_t1, irn=None if irn is None else irn[:n], jcn=None if jcn is None else jcn[:n],
                    a=None if a is None else a[:n], nz=n, rhs=rhs, seconds=secs.value, jac_builds=jb.value)

    def node_solution(self, omega, sigma, x):
        """solution.f90 post-processing of a solved system x[2*nne]: total E, H at the grid nodes, impedance,
        apparent resistivity and phase (the north_star's end-to-end observable).  Returns dict(esol, hsol, z, rho, phi)."""
        npt = self.model.npt
        sigma = np.ascontiguousarray(sigma, np.complex128)
        x = np.ascontiguousarray(x, np.complex128)
        assert x.size == 2 * self.nne
        esol = np.zeros((2 * npt, 3), np.complex128); hsol = np.zeros((2 * npt, 3), np.complex128)
        z = np.zeros((npt, 4), np.complex128); rho = np.zeros((npt, 4)); phi = np.zeros((npt, 4))
        rc = lib().oracle_node_solution(self._h, omega, _p(sigma), _p(x), _p(esol), _p(hsol), _p(z), _p(rho), _p(phi))
        if rc:
            raise RuntimeError(f"oracle_node_solution failed: {rc}")
        return dict(esol=esol, hsol=hsol, z=z, rho=rho, phi=phi)

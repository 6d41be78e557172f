"""ctypes mirror of include/movfem_b200.h (the C-ABI drop-in boundary).

Field order and types must match ``struct movfem_desc`` exactly.  Every field is the
reference module variable of the same name (geometry.f90:17-26, boundary_conds.f90:15-25,
problem.f90:18, v_fem.f90:16, n_fem.f90:14).
"""
import ctypes as C

MOVFEM_OK = 0
MOVFEM_E_BADARG = -1
MOVFEM_E_CUDA = -2
MOVFEM_E_SINGULAR_JAC = -3
MOVFEM_E_SINGULAR_MODEL = -4
MOVFEM_E_NOGPU = -5
MOVFEM_E_CAPACITY = -6
MOVFEM_E_UNSUPPORTED = -7

MODE_T2 = 0
MODE_T1 = 1
MODE_KEEP_PATTERN = 0x100

ERROR_NAMES = {
    MOVFEM_E_BADARG: "bad argument",
    MOVFEM_E_CUDA: "CUDA error",
    MOVFEM_E_SINGULAR_JAC: "no transformation!! nf_det=0!! (n_fem.f90:374-377)",
    MOVFEM_E_SINGULAR_MODEL: "no sigma/mu inversion on node (problem.f90:260-271)",
    MOVFEM_E_NOGPU: "no CUDA device (there is no CPU fallback)",
    MOVFEM_E_CAPACITY: "caller array too small",
    MOVFEM_E_UNSUPPORTED: "unsupported configuration",
}


class MovfemDesc(C.Structure):
    _fields_ = [
        ("g_nx", C.c_int32), ("g_ny", C.c_int32), ("g_nz", C.c_int32),
        ("nord", C.c_int32), ("mn", C.c_int32), ("me", C.c_int32),
        ("nextd", C.c_int32), ("nzl_top", C.c_int32),
        ("dirichlet", C.c_int32), ("bd_inimod", C.c_int32), ("gpml_sch", C.c_int32),
        ("sym", C.c_int32), ("ndir", C.c_int32), ("pe_sch", C.c_int32),
        ("a0", C.c_double), ("b0", C.c_double), ("nn", C.c_double),
        ("g_xp", C.c_void_p), ("g_yp", C.c_void_p), ("g_zp", C.c_void_p), ("g_mu", C.c_void_p),
        ("ie_lo", C.c_int32), ("ie_hi", C.c_int32),
        ("g_ztop", C.c_double), ("bd_hsigma", C.c_double), ("bd_nl", C.c_int32), ("bd_pad", C.c_int32),
        ("bd_lsigma", C.c_double * 16), ("bd_ldz", C.c_double * 16),
    ]


class MovfemStats(C.Structure):
    _fields_ = [
        ("ms_h2d", C.c_double), ("ms_node", C.c_double), ("ms_element", C.c_double),
        ("ms_gather", C.c_double), ("ms_finalize", C.c_double), ("ms_d2h", C.c_double),
        ("ms_total", C.c_double), ("nz", C.c_int64), ("launches", C.c_int64),
        ("ms_geometry", C.c_double), ("ms_contract", C.c_double), ("ms_exact", C.c_double), ("nflagged", C.c_int64), ("ms_fused", C.c_double),
    ]


class MovfemGeomodel(C.Structure):
    """struct movfem_geomodel: arguments of geometry.f90 innermodel_gqg"""
    _fields_ = [
        ("mx", C.c_int32), ("my", C.c_int32), ("mz", C.c_int32),
        ("isigma", C.c_int32), ("imu", C.c_int32), ("nzl_air", C.c_int32),
        ("ijsigma", (C.c_int32 * 2) * 9), ("ijmu", (C.c_int32 * 2) * 9),
        ("xm", C.c_void_p), ("ym", C.c_void_p), ("zm", C.c_void_p), ("sigma", C.c_void_p), ("mu", C.c_void_p),
    ]

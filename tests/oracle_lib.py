"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the reference's own
levmar compiled unmodified (oracle/_ref/liblevmar_ref.so).  Test infrastructure only: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "liblevmar_ref.so")

FUNC_T = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p)
dptr = C.POINTER(C.c_double)
iptr = C.POINTER(C.c_int)


class ExtraData(C.Structure):
    """struct extraData of brdfdata.cpp:962-966"""
    _fields_ = [("angles", dptr), ("modelInfo", C.c_int)]


def build_oracle(force=False):
    if force or not os.path.exists(ORACLE_SO) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(ORACLE_SO)
            for f in ("lm_oracle.c", "brdf_oracle.c", "gather_oracle.c", "oracle.h")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


def build_ref():
    """Compile the reference's levmar from /root/reference when that tree exists (dev container)."""
    if os.path.exists(REF_SO):
        return REF_SO
    if os.path.isdir("/root/reference/levmar"):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "ref"])
        return REF_SO
    return None


_oracle = None
_ref = None


def _decl_solver(lib, prefix):
    bc_der = getattr(lib, prefix + "dlevmar_bc_der")
    bc_der.restype = C.c_int
    bc_der.argtypes = [FUNC_T, FUNC_T, dptr, dptr, C.c_int, C.c_int, dptr, dptr, dptr, C.c_int, dptr, dptr,
                       dptr, dptr, C.c_void_p]
    bc_dif = getattr(lib, prefix + "dlevmar_bc_dif")
    bc_dif.restype = C.c_int
    bc_dif.argtypes = [FUNC_T, dptr, dptr, C.c_int, C.c_int, dptr, dptr, dptr, C.c_int, dptr, dptr,
                       dptr, dptr, C.c_void_p]
    der = getattr(lib, prefix + "dlevmar_der")
    der.restype = C.c_int
    der.argtypes = [FUNC_T, FUNC_T, dptr, dptr, C.c_int, C.c_int, C.c_int, dptr, dptr, dptr, dptr, C.c_void_p]
    dif = getattr(lib, prefix + "dlevmar_dif")
    dif.restype = C.c_int
    dif.argtypes = [FUNC_T, dptr, dptr, C.c_int, C.c_int, C.c_int, dptr, dptr, dptr, dptr, C.c_void_p]


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(ORACLE_SO)
        _decl_solver(lib, "oracle_")
        lib.oracle_L2nrmxmy.restype = C.c_double
        lib.oracle_L2nrmxmy.argtypes = [dptr, dptr, dptr, C.c_int]
        lib.oracle_Ax_eq_b_LU.restype = C.c_int
        lib.oracle_Ax_eq_b_LU.argtypes = [dptr, dptr, dptr, C.c_int]
        lib.oracle_BRDFFunc.restype = None
        lib.oracle_BRDFFunc.argtypes = [dptr, dptr, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_BRDFJac.restype = None
        lib.oracle_BRDFJac.argtypes = [dptr, dptr, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_solve_equation.restype = C.c_int
        lib.oracle_solve_equation.argtypes = [dptr, dptr, dptr, dptr, C.c_int, C.c_int, dptr, dptr]
        lib.oracle_solve_equation_single.restype = C.c_int
        lib.oracle_solve_equation_single.argtypes = [dptr, dptr, dptr, dptr, C.c_long, C.c_int, dptr, dptr]
        lib.oracle_led_table.argtypes = [dptr]
        lib.oracle_face_normals.argtypes = [dptr, iptr, C.c_int, dptr]
        lib.oracle_subtract_ambient.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        lib.oracle_calc_pixel2surface.restype = C.c_int
        lib.oracle_calc_pixel2surface.argtypes = [dptr, iptr, C.c_int, dptr, C.c_int, C.c_int, iptr]
        lib.oracle_gather.restype = C.c_int
        lib.oracle_gather.argtypes = [dptr, iptr, C.c_int, dptr, dptr, C.POINTER(C.c_void_p), C.c_int,
                                      C.c_int, C.c_int, iptr, iptr, iptr, dptr, dptr, dptr, dptr]
        lib.oracle_set_dot_order.restype = None
        lib.oracle_set_dot_order.argtypes = [C.c_int]
        _oracle = lib
    return _oracle


DOT_EIGEN33, DOT_SEQUENTIAL = 0, 1


def set_dot_order(order):
    """Evaluation order of the final dots of GetCosLN / GetCosNH in the gather oracle (gather_oracle.c header)."""
    oracle().oracle_set_dot_order(int(order))


def ref():
    """The reference's levmar (None when oracle/_ref was never built, e.g. a box without it)."""
    global _ref
    if _ref is None:
        so = build_ref()
        if so is None or not os.path.exists(so):
            return None
        lib = C.CDLL(so)
        _decl_solver(lib, "")
        lib.dlevmar_L2nrmxmy.restype = C.c_double
        lib.dlevmar_L2nrmxmy.argtypes = [dptr, dptr, dptr, C.c_int]
        lib.dAx_eq_b_LU_noLapack.restype = C.c_int
        lib.dAx_eq_b_LU_noLapack.argtypes = [dptr, dptr, dptr, C.c_int]
        _ref = lib
    return _ref


def as_d(a):
    return None if a is None else a.ctypes.data_as(dptr)


def as_i(a):
    return a.ctypes.data_as(iptr)


def brdf_callback():
    """The oracle's BRDFFunc as a levmar callback pointer (usable with the reference levmar too)."""
    return C.cast(oracle().oracle_BRDFFunc, FUNC_T)


def brdf_jac_callback():
    return C.cast(oracle().oracle_BRDFJac, FUNC_T)


def make_extra(angles, model):
    ed = ExtraData(as_d(angles), int(model))
    return ed


def levmar_bc_dif(lib, prefix, func, p0, x, lb, ub, itmax, opts, adata=None, dscl=None, want_covar=False):
    """Call <prefix>dlevmar_bc_dif on copies; returns (ret, p, info, covar)."""
    p = np.array(p0, dtype=np.float64)
    m, n = p.size, (x.size if x is not None else 0)
    info = np.zeros(10)
    lb = None if lb is None else np.array(lb, dtype=np.float64)
    ub = None if ub is None else np.array(ub, dtype=np.float64)
    dscl = None if dscl is None else np.array(dscl, dtype=np.float64)
    opts = None if opts is None else np.array(opts, dtype=np.float64)
    covar = np.zeros((m, m)) if want_covar else None
    fn = getattr(lib, prefix + "dlevmar_bc_dif")
    ret = fn(func, as_d(p), as_d(x), m, n, as_d(lb), as_d(ub), as_d(dscl), int(itmax), as_d(opts), as_d(info),
             None, as_d(covar), C.cast(C.pointer(adata), C.c_void_p) if adata is not None else None)
    return ret, p, info, covar


def levmar_bc_der(lib, prefix, func, jacf, p0, x, lb, ub, itmax, opts, adata=None, dscl=None, want_covar=False):
    p = np.array(p0, dtype=np.float64)
    m, n = p.size, x.size
    info = np.zeros(10)
    lb = None if lb is None else np.array(lb, dtype=np.float64)
    ub = None if ub is None else np.array(ub, dtype=np.float64)
    dscl = None if dscl is None else np.array(dscl, dtype=np.float64)
    opts = None if opts is None else np.array(opts, dtype=np.float64)
    covar = np.zeros((m, m)) if want_covar else None
    fn = getattr(lib, prefix + "dlevmar_bc_der")
    ret = fn(func, jacf, as_d(p), as_d(x), m, n, as_d(lb), as_d(ub), as_d(dscl), int(itmax), as_d(opts),
             as_d(info), None, as_d(covar),
             C.cast(C.pointer(adata), C.c_void_p) if adata is not None else None)
    return ret, p, info, covar


def levmar_der(lib, prefix, func, jacf, p0, x, itmax, opts, adata=None, want_covar=False):
    p = np.array(p0, dtype=np.float64)
    m, n = p.size, x.size
    info = np.zeros(10)
    opts = None if opts is None else np.array(opts, dtype=np.float64)
    covar = np.zeros((m, m)) if want_covar else None
    fn = getattr(lib, prefix + "dlevmar_der")
    ret = fn(func, jacf, as_d(p), as_d(x), m, n, int(itmax), as_d(opts), as_d(info), None, as_d(covar),
             C.cast(C.pointer(adata), C.c_void_p) if adata is not None else None)
    return ret, p, info, covar


def levmar_dif(lib, prefix, func, p0, x, itmax, opts, adata=None, want_covar=False):
    p = np.array(p0, dtype=np.float64)
    m, n = p.size, x.size
    info = np.zeros(10)
    opts = None if opts is None else np.array(opts, dtype=np.float64)
    covar = np.zeros((m, m)) if want_covar else None
    fn = getattr(lib, prefix + "dlevmar_dif")
    ret = fn(func, as_d(p), as_d(x), m, n, int(itmax), as_d(opts), as_d(info), None, as_d(covar),
             C.cast(C.pointer(adata), C.c_void_p) if adata is not None else None)
    return ret, p, info, covar


# ---- the reference's two option presets (brdfdata.cpp:1002,1046-1058 and :1085,1107-1119) ----
REF_GLOBAL = dict(p0=(0.0, 0.0, 0.0), itmax=2000, opts=(1e-3, 1e-15, 1e-10, 1e-50, 1.0),
                  lb=(0.0, 0.0, 0.0), ub=(100.0, 100.0, 100.0))
REF_PERFACE = dict(p0=(0.5, 1.0, 1.0), itmax=100, opts=(1e-3, 1e-15, 1e-15, 1e-20, 1e-6),
                   lb=(0.0, 0.0, 0.0), ub=(100.0, 100.0, 100.0))


def brdf_fit(lib, prefix, cosphi, costd, costheta, x, model, preset):
    """One dlevmar_bc_dif BRDF fit through the oracle's BRDFFunc callback."""
    n = x.size
    angles = np.empty(3 * n)
    angles[:n] = cosphi
    angles[n:2 * n] = costd
    angles[2 * n:] = 0.0 if costheta is None else costheta
    ed = make_extra(angles, model)
    xx = np.ascontiguousarray(x, dtype=np.float64)
    ret, p, info, _ = levmar_bc_dif(lib, prefix, brdf_callback(), preset["p0"], xx, preset["lb"], preset["ub"],
                                    preset["itmax"], preset["opts"], adata=ed)
    return ret, p, info

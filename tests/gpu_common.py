"""Shared helpers of the -m gpu parity tests: oracle-side quantities for the same inputs."""
import ctypes as C

import numpy as np

import oracle_lib as O


def oracle_predict(p, cosphi, td, th, model):
    n = cosphi.size
    angles = np.concatenate([cosphi, td, th if th is not None else np.zeros(n)])
    ed = O.make_extra(angles, model)
    hx = np.zeros(n)
    pp = np.array(p, dtype=np.float64)
    O.oracle().oracle_BRDFFunc(O.as_d(pp), O.as_d(hx), 3, n, C.cast(C.pointer(ed), C.c_void_p))
    return hx


def oracle_fd_jacobian(p, cosphi, td, th, model, delta):
    """levmar's forward (delta >= 0) or central (delta < 0) difference Jacobian, misc_core.c:137-211."""
    n = cosphi.size
    angles = np.concatenate([cosphi, td, th if th is not None else np.zeros(n)])
    ed = O.make_extra(angles, model)
    pp = np.array(p, dtype=np.float64)
    hx, hxx, jac = np.zeros(n), np.zeros(n), np.zeros((n, 3))
    lib = O.oracle()
    cb = O.brdf_callback()
    ad = C.cast(C.pointer(ed), C.c_void_p)
    lib.oracle_fdif_forw_jac.argtypes = [O.FUNC_T, O.dptr, O.dptr, O.dptr, C.c_double, O.dptr, C.c_int, C.c_int, C.c_void_p]
    lib.oracle_fdif_cent_jac.argtypes = lib.oracle_fdif_forw_jac.argtypes
    if delta >= 0:
        lib.oracle_BRDFFunc(O.as_d(pp), O.as_d(hx), 3, n, ad)
        lib.oracle_fdif_forw_jac(cb, O.as_d(pp), O.as_d(hx), O.as_d(hxx), float(delta), O.as_d(jac), 3, n, ad)
    else:
        lib.oracle_fdif_cent_jac(cb, O.as_d(pp), O.as_d(hx), O.as_d(hxx), float(-delta), O.as_d(jac), 3, n, ad)
    return jac


def oracle_analytic_jacobian(p, cosphi, td, th, model):
    n = cosphi.size
    angles = np.concatenate([cosphi, td, th if th is not None else np.zeros(n)])
    ed = O.make_extra(angles, model)
    pp = np.array(p, dtype=np.float64)
    jac = np.zeros((n, 3))
    O.oracle().oracle_BRDFJac(O.as_d(pp), O.as_d(jac), 3, n, C.cast(C.pointer(ed), C.c_void_p))
    return jac


def normal_eq_from(jac, e):
    """[JtJ 00 01 02 11 12 22, Jte 0..2, ||e||^2] in float64 with numpy (pairwise) sums."""
    a = jac.T @ jac
    g = jac.T @ e
    return np.array([a[0, 0], a[0, 1], a[0, 2], a[1, 1], a[1, 2], a[2, 2], g[0], g[1], g[2], float(e @ e)])


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1e-300)
    return np.abs(a - b) / scale

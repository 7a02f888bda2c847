"""Host logic of the product: the LM control loop of brdf_b200/csrc/lm_engine.cuh, instantiated on
the host behind brdfgpu_lm_bc_reduced / brdfgpu_lm_unc_reduced (it consumes only J^T J, J^T e and
||e||^2), against the reference's dlevmar_bc_der / dlevmar_der (lmbc_core.c:369-1022,
lm_core.c:64-432) on the levmar demo problems.  The reduced sums are formed here in levmar's own
small-problem order (lmbc_core.c:592-616, misc_core.c:721-807 through the oracle), so the whole
trajectory -- p, info[0..9] -- must agree bit for bit.  No GPU involved."""
import ctypes as C

import numpy as np
import pytest

import kat_problems as K
import oracle_lib as O
from brdf_b200 import api as A


def _reduced_callbacks(prob):
    f, jf = prob["f"], prob["j"]
    m, n = prob["m"], prob["n"]
    x = np.array(prob["x"], dtype=np.float64)
    lib = O.oracle()

    def jac_cb(p, mm, JtJ, Jte, _):
        pv = [p[i] for i in range(m)]
        hx, jac = np.zeros(n), np.zeros(n * m)
        f(pv, hx, m, n)
        jf(pv, jac, m, n)
        e = np.zeros(n)
        lib.oracle_L2nrmxmy(O.as_d(e), O.as_d(x), O.as_d(hx), n)
        a = np.zeros((m, m)); g = np.zeros(m)
        for l in range(n - 1, -1, -1):          # lmbc_core.c:603-613
            row = jac[l * m:(l + 1) * m]
            for i in range(m - 1, -1, -1):
                alpha = row[i]
                for j in range(i, -1, -1):
                    a[i, j] += row[j] * alpha
                g[i] += alpha * e[l]
        for i in range(m):
            for j in range(i + 1, m):
                a[i, j] = a[j, i]
        for i in range(m):
            Jte[i] = g[i]
            for j in range(m):
                JtJ[i * m + j] = a[i, j]

    def cost_cb(p, mm, nonfinite, _):
        pv = [p[i] for i in range(m)]
        hx, e = np.zeros(n), np.zeros(n)
        f(pv, hx, m, n)
        s = lib.oracle_L2nrmxmy(O.as_d(e), O.as_d(x), O.as_d(hx), n)
        nonfinite[0] = float(np.count_nonzero(~np.isfinite(e)))
        return s

    return jac_cb, cost_cb


BC = [p for p in K.PROBLEMS if p["driver"] == "bc_der"]
DER = [p for p in K.PROBLEMS if p["driver"] == "der"]


def _ref_or_oracle():
    ref = O.ref()
    return (ref, "") if ref is not None else (O.oracle(), "oracle_")


@pytest.mark.parametrize("prob", BC, ids=[p["name"] for p in BC])
def test_bc_engine_matches_levmar(prob):
    lib, prefix = _ref_or_oracle()
    f, j = K.callbacks(prob)
    x = np.array(prob["x"], dtype=np.float64)
    r_ret, r_p, r_info, _ = O.levmar_bc_der(lib, prefix, f, j, prob["p0"], x, prob["lb"], prob["ub"], prob["itmax"], K.OPTS)
    jac_cb, cost_cb = _reduced_callbacks(prob)
    ret, p, info, _ = A.lm_bc_reduced(jac_cb, cost_cb, prob["p0"], prob["n"], prob["lb"], prob["ub"], prob["itmax"], K.OPTS)
    assert ret == r_ret
    assert p.tobytes() == r_p.tobytes()
    assert info.tobytes() == r_info.tobytes()
    if prob["info"] is not None:
        assert [int(v) for v in info[5:10]] == prob["info"]


@pytest.mark.parametrize("prob", DER, ids=[p["name"] for p in DER])
def test_unconstrained_engine_matches_levmar(prob):
    lib, prefix = _ref_or_oracle()
    f, j = K.callbacks(prob)
    x = np.array(prob["x"], dtype=np.float64)
    r_ret, r_p, r_info, r_cov = O.levmar_der(lib, prefix, f, j, prob["p0"], x, prob["itmax"], K.OPTS, want_covar=True)
    jac_cb, cost_cb = _reduced_callbacks(prob)
    ret, p, info, cov = A.lm_unc_reduced(jac_cb, cost_cb, prob["p0"], prob["n"], prob["itmax"], K.OPTS, want_covar=True)
    assert ret == r_ret
    assert p.tobytes() == r_p.tobytes()
    assert info.tobytes() == r_info.tobytes()
    assert cov.tobytes() == r_cov.tobytes()


def test_bc_engine_with_diagonal_scaling():
    """dscl path (lmbc_core.c:536-570): scaled variables, scaled bounds, covariance rescaled."""
    prob = next(p for p in K.PROBLEMS if p["name"] == "combust")
    lib, prefix = _ref_or_oracle()
    f, j = K.callbacks(prob)
    x = np.array(prob["x"], dtype=np.float64)
    dscl = [1.0, 2.0, 0.5, 4.0, 1.0]
    r_ret, r_p, r_info, _ = O.levmar_bc_der(lib, prefix, f, j, prob["p0"], x, prob["lb"], prob["ub"], prob["itmax"], K.OPTS,
                                            dscl=dscl)
    jac_cb, cost_cb = _reduced_callbacks(prob)
    ret, p, info, _ = A.lm_bc_reduced(jac_cb, cost_cb, prob["p0"], prob["n"], prob["lb"], prob["ub"], prob["itmax"], K.OPTS,
                                      dscl=dscl)
    # J D is formed from the unscaled sums here (d_i d_j JtJ_ij) instead of scaling J's rows first,
    # so rounding may differ in the last bits: same stop reason, same answer to 1e-9
    assert ret >= 0 and r_ret >= 0
    assert int(info[6]) == int(r_info[6])
    np.testing.assert_allclose(p, r_p, rtol=1e-9, atol=1e-12)


def test_lu_solver_matches_levmar():
    rng = np.random.default_rng(11)
    lib = O.oracle()
    for m in (1, 2, 3, 4, 8):
        for _ in range(20):
            a = rng.standard_normal((m, m)); b = rng.standard_normal(m)
            x0, x1 = np.zeros(m), np.zeros(m)
            r0 = lib.oracle_Ax_eq_b_LU(O.as_d(a.copy()), O.as_d(b.copy()), O.as_d(x0), m)
            r1 = A.lib().brdfgpu_Ax_eq_b_LU(A._d(a), A._d(b), A._d(x1), m)
            assert r0 == r1 and x0.tobytes() == x1.tobytes()
    z = np.zeros((3, 3)); z[0, 0] = 1.0
    assert A.lib().brdfgpu_Ax_eq_b_LU(A._d(z), A._d(np.ones(3)), A._d(np.zeros(3)), 3) == 0   # all-zero row


@pytest.mark.parametrize("prob", BC, ids=[p["name"] for p in BC])
def test_bc_engine_batched_projected_gradient_walk(prob):
    """The persistent fit kernel receives the candidates of the projected-gradient walk
    (lmbc_core.c:885-934) eight at a time and discards the ones past the stopping point: the
    trajectory, p and info[0..9] (nfev included) must still be levmar's, bit for bit."""
    lib, prefix = _ref_or_oracle()
    f, j = K.callbacks(prob)
    x = np.array(prob["x"], dtype=np.float64)
    r_ret, r_p, r_info, _ = O.levmar_bc_der(lib, prefix, f, j, prob["p0"], x, prob["lb"], prob["ub"], prob["itmax"], K.OPTS)
    jac_cb, cost_cb = _reduced_callbacks(prob)
    A.lib().brdfgpu_lm_reduced_batching(1)
    try:
        ret, p, info, _ = A.lm_bc_reduced(jac_cb, cost_cb, prob["p0"], prob["n"], prob["lb"], prob["ub"], prob["itmax"], K.OPTS)
    finally:
        biggest = A.lib().brdfgpu_lm_reduced_batching(0)
    assert ret == r_ret
    assert p.tobytes() == r_p.tobytes()
    assert info.tobytes() == r_info.tobytes()
    if prob["name"] == "hatfldb":
        assert biggest >= 2   # this problem does walk the projected gradient (batches grow 1, 2, 4, 8)

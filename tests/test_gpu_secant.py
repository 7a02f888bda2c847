"""-m gpu parity of brdfgpu_dlevmar_dif (secant LM: difference Jacobian rebuilt only now and then,
Broyden rank-one updates in between, levmar/lm_core.c:438-842 -- the alternative the reference keeps
commented out at brdfdata.cpp:1059,1120) against the reference's dlevmar_dif on identical inputs."""
import numpy as np
import pytest

import oracle_lib as O
import synth
from brdf_b200 import api as A

pytestmark = pytest.mark.gpu
PAR_RTOL, COST_RTOL = 1e-4, 1e-6
CASES = [
    # (n, seed, p0, opts, itmax)
    (30000, 11, (0.5, 1.0, 1.0), (1e-3, 1e-15, 1e-15, 1e-20, 1e-6), 200),      # the per-face option set, forward differences
    (30001, 12, (0.4, 0.5, 8.0), (1e-3, 1e-15, 1e-15, 1e-20, -1e-6), 200),     # central differences, odd n
    (200000, 13, (0.3, 0.3, 5.0), (1e-3, 1e-12, 1e-12, 1e-20, 1e-6), 100),
    (5000, 14, (0.5, 1.0, 1.0), None, 150),                                    # levmar's default options
]


def _ref():
    return (O.ref(), "") if O.ref() is not None else (O.oracle(), "oracle_")


@pytest.mark.parametrize("case", CASES, ids=["fwd", "central-odd", "n2e5", "default-opts"])
def test_dlevmar_dif_matches_levmar(case):
    n, seed, p0, opts, itmax = case
    c, td, th, x = synth.samples(n, seed=seed)
    angles = np.concatenate([c, td, th])
    lib, prefix = _ref()
    want = O.levmar_dif(lib, prefix, O.brdf_callback(), p0, x, itmax, opts, adata=O.make_extra(angles, 1), want_covar=True)
    extra, _keep = A.make_extra(c, td, th, 1)
    if opts is None:
        info = np.zeros(10); p = np.array(p0, dtype=np.float64); covar = np.zeros((3, 3))
        import ctypes as C
        ret = A.lib().brdfgpu_dlevmar_dif(A.func_address("brdfgpu_BRDFFunc"), A._d(p), A._d(x), 3, n, itmax, None, A._d(info), None,
                                          A._d(covar), C.cast(C.pointer(extra), C.c_void_p))
    else:
        ret, p, info, covar = A.dlevmar_dif(p0, x, itmax, opts, extra, want_covar=True)
    w_ret, w_p, w_info, w_cov = want
    assert (ret >= 0) == (w_ret >= 0)
    np.testing.assert_allclose(info[0], w_info[0], rtol=1e-12)          # ||e||^2 at the start
    np.testing.assert_allclose(info[1], w_info[1], rtol=COST_RTOL)      # final cost
    np.testing.assert_allclose(p, w_p, rtol=PAR_RTOL)
    if int(info[6]) not in (3, 5) and int(w_info[6]) not in (3, 5):
        assert int(info[6]) == int(w_info[6])
    np.testing.assert_allclose(covar, w_cov, rtol=2e-3, atol=1e-12)
    assert info[8] >= 1 and info[7] >= info[8] * 3 + 1                   # Jacobian rebuilds charged m evaluations each
    assert info[8] < info[5] or info[5] <= 2                             # ... and most iterations did NOT rebuild it


def test_secant_state_survives_a_bigger_refit():
    """the stored Jacobian grows with the sample set it belongs to"""
    ctx = A.Context()
    opts = (1e-3, 1e-15, 1e-15, 1e-20, 1e-6)
    c, td, th, x = synth.samples(4000, seed=3)
    s = ctx.upload(c, td, x, 1)
    a = ctx.fit_global_unc(s, (0.5, 1.0, 1.0), 100, opts)
    b = ctx.fit_global_unc(s, (0.5, 1.0, 1.0), 100, opts)
    assert a[0] == b[0] and a[1].tobytes() == b[1].tobytes() and a[2].tobytes() == b[2].tobytes()   # deterministic
    s.free()
    ctx.close()

"""The oracle's LM drivers against the known answers of the reference's levmar demo
(levmar/lmdemo.c:859-1111; expectations in kat_problems.PROBLEMS) and,
when oracle/_ref is present, bit-for-bit against the reference's own levmar on the same problems."""
import numpy as np
import pytest

import kat_problems as K
import oracle_lib as O


def _run(lib, prefix, prob):
    f, j = K.callbacks(prob)
    x = np.array(prob["x"], dtype=np.float64)
    drv = prob["driver"]
    want_cov = "covar_row0" in prob
    if drv == "der":
        return O.levmar_der(lib, prefix, f, j, prob["p0"], x, prob["itmax"], K.OPTS, want_covar=want_cov)
    if drv == "dif":
        return O.levmar_dif(lib, prefix, f, prob["p0"], x, prob["itmax"], K.OPTS, want_covar=want_cov)
    if drv == "bc_der":
        return O.levmar_bc_der(lib, prefix, f, j, prob["p0"], x, prob["lb"], prob["ub"], prob["itmax"], K.OPTS)
    raise AssertionError(drv)


@pytest.mark.parametrize("prob", K.PROBLEMS, ids=[p["name"] for p in K.PROBLEMS])
def test_oracle_known_answers(prob):
    ret, p, info, covar = _run(O.oracle(), "oracle_", prob)
    if prob["sol"] is not None:
        for got, want in zip(p, prob["sol"]):
            assert float("%.7g" % got) == pytest.approx(want, rel=1e-6, abs=1e-12)
    if prob["info"] is not None:
        assert [int(v) for v in info[5:10]] == prob["info"]
    if "info1" in prob:
        assert info[1] == pytest.approx(prob["info1"], rel=1e-5)
    if "covar_row0" in prob:
        assert covar[0] == pytest.approx(prob["covar_row0"], rel=1e-5)


@pytest.mark.parametrize("prob", K.PROBLEMS, ids=[p["name"] for p in K.PROBLEMS])
def test_oracle_bit_exact_vs_reference(prob):
    ref = O.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r0 = _run(ref, "", prob)
    r1 = _run(O.oracle(), "oracle_", prob)
    assert r0[0] == r1[0]
    assert r0[1].tobytes() == r1[1].tobytes()
    assert r0[2].tobytes() == r1[2].tobytes()
    if r0[3] is not None:
        assert r0[3].tobytes() == r1[3].tobytes()

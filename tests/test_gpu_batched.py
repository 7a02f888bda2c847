"""-m gpu parity tests of the batched mode (K4: one lane group per independent fit) against the
reference's per-face driver SolveEquation -> dlevmar_bc_dif (brdfdata.cpp:1077-1136)."""
import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O
import synth
from brdf_b200 import api as A

pytestmark = pytest.mark.gpu
PAR_RTOL, COST_RTOL = 1e-4, 1e-6
GOLD = G.load()
CONVERGED = (1, 2, 6)   # info[6]: small gradient / small step / small error


@pytest.fixture(scope="module")
def ctx():
    c = A.Context()
    yield c
    c.close()


def _compare(p, info, ret, want_p, want_info, want_ret, min_agree, determined=True):
    """Per-fit parity classes.
    strict : parameters within 1e-4 and final cost within 1e-6 relative of levmar's -- required for
             at least `min_agree` of the fits;
    loose  : every fit must at least end at levmar's cost to 1e-3 relative (or lower): fits that
             hit itmax=100, stop with 'no further reduction', or crawl along an active bound with
             ks = 0 (flat in n) end wherever rounding takes them -- on the CPU as well when the
             summation order changes."""
    nfit = len(want_ret)
    strict = 0
    for f in range(nfit):
        ok_p = np.allclose(p[f], want_p[f], rtol=PAR_RTOL, atol=1e-9)
        ok_c = np.isclose(info[f][1], want_info[f][1], rtol=COST_RTOL, atol=1e-18)
        strict += bool(ok_c and (ok_p or not determined))
        assert (ret[f] >= 0) == (want_ret[f] >= 0)
        assert info[f][1] <= want_info[f][1] * (1 + 1e-3) + 1e-15, (f, p[f], want_p[f], info[f], want_info[f])
        assert np.isclose(info[f][0], want_info[f][0], rtol=1e-9)
    assert strict >= min_agree * nfit, (strict, nfit)
    return strict


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("case", GOLD["batch"], ids=[c["name"] for c in GOLD["batch"]])
def test_batched_fits_match_reference_golden(ctx, case):
    """The reference's driver (brdfgpu_solve_equation_batch = the per-pixel loop of CalcBRDFEquation) runs levmar-exact:
    EQUAL to the golden values oracle/_ref produced -- parameters, all of info[], return value, every fit."""
    c, td, th, x, _ = G.batch_inputs(case)
    p, info, ret = ctx.solve_equation_batch(c, td, th, x, case["model"])
    assert c.shape[1] <= 128
    assert np.array_equal(ret, np.asarray(case["ret"])) and _same(p, case["p"]) and _same(info, case["info"])
    # the fast kernel on the same problems: the parity classes
    b = ctx.batch_upload(c, td if case["model"] == 1 else th, x, case["model"])
    b.fit(A.REF_PERFACE, jac_mode=A.JAC_FD)
    pf, infof, retf = b.results()
    b.free()
    _compare(pf, infof, retf, case["p"], case["info"], case["ret"], 0.85)


@pytest.mark.parametrize("nper", [3, 16, 17, 32, 64, 100, 200])
def test_batched_fits_match_oracle_fresh(ctx, nper):
    nfit = 96
    c, td, th, x, _ = synth.batched(nfit, nper, seed=900 + nper)
    b = ctx.batch_upload(c, td, x, 1)
    b.fit(A.REF_PERFACE, jac_mode=A.JAC_FD)
    p, info, ret = b.results()
    b.free()
    wp, wi, wr = np.zeros((nfit, 3)), np.zeros((nfit, 10)), np.zeros(nfit, dtype=int)
    for f in range(nfit):
        wr[f], wp[f], wi[f] = O.brdf_fit(O.oracle(), "oracle_", c[f], td[f], th[f], x[f], 1, O.REF_PERFACE)
    # nper == 3 == m is an exactly determined system: the minimiser is not unique to 1e-4 when a bound
    # is active (cost ~1e-20 on both sides), so only the cost is compared there
    _compare(p, info, ret, wp, wi, wr, 0.85 if nper >= 16 else 0.5, determined=nper > 3)


def test_solve_equation_single_fit_entry():
    """brdfgpu_solve_equation == CBRDFdata::SolveEquation for one face (16 LEDs)."""
    case = GOLD["batch"][0]
    c, td, th, x, _ = G.batch_inputs(case)
    for f in (1, 2, 3, 4):
        ret, p, info = A.solve_equation(c[f], td[f], th[f], x[f], 1)
        assert ret == case["ret"][f] and _same(p, case["p"][f]) and _same(info, case["info"][f])   # levmar-exact


def test_solve_equation_single_global_entry():
    g = GOLD["global"][0]
    c, td, th, x = G.global_inputs(g)
    ret, p, info = A.solve_equation_single(c, td, th, x, 1)
    assert ret >= 0 and int(info[6]) == int(g["info"][6])
    np.testing.assert_allclose(p, g["p"], rtol=PAR_RTOL)
    np.testing.assert_allclose(info[1], g["info"][1], rtol=COST_RTOL)


def test_batched_negative_cosines(ctx):
    """NaN-producing samples: same per-fit outcome class as levmar (no LM_ERROR)."""
    nfit, nper = 64, 16
    c, td, th, x, _ = synth.batched(nfit, nper, seed=31)
    td = td.copy(); td[:, 3] *= -1.0
    for mode in (A.JAC_FD_EXACT, A.JAC_FD):
        b = ctx.batch_upload(c, td, x, 1)
        b.fit(A.REF_PERFACE, jac_mode=mode)
        p, info, ret = b.results()
        b.free()
        for f in range(nfit):
            w = O.brdf_fit(O.oracle(), "oracle_", c[f], td[f], th[f], x[f], 1, O.REF_PERFACE)
            assert (ret[f] >= 0) == (w[0] >= 0)
            assert int(info[f][6]) == int(w[2][6])
            np.testing.assert_allclose(p[f], w[1], rtol=PAR_RTOL)
            if mode == A.JAC_FD_EXACT:
                assert ret[f] == w[0] and _same(p[f], w[1]) and _same(info[f], w[2])


def test_full_size_batch_properties(ctx):
    """BASELINE config 4: 65,536 fits x 64 samples.  Properties: every fit returns, the device
    generator equals the numpy recipe on a slice, the spot-checked fits equal the oracle, fit order
    does not matter (fits are independent), and the bulk recovers its generating parameters."""
    nfit, nper = 65536, 64
    b = ctx.batch_synth(nfit, nper, seed=2026)
    b.fit(A.REF_PERFACE)
    p, info, ret = b.results()
    assert np.all(ret >= 0)
    assert np.all(np.isfinite(p)) and np.all(p >= 0.0) and np.all(p <= 100.0)
    c, td, th, x, truth = synth.batched(256, nper, seed=2026)
    conv = 0
    for f in range(0, 256, 8):
        w = O.brdf_fit(O.oracle(), "oracle_", c[f], td[f], th[f], x[f], 1, O.REF_PERFACE)
        if int(w[2][6]) in CONVERGED and int(info[f][6]) in CONVERGED:
            np.testing.assert_allclose(p[f], w[1], rtol=PAR_RTOL, atol=1e-9)
            np.testing.assert_allclose(info[f][1], w[2][1], rtol=COST_RTOL)
            conv += 1
    assert conv >= 20
    # a shard of the same problem set (fits 1000..1999) gives the same answers: no cross-fit coupling
    b2 = ctx.batch_synth(1000, nper, seed=2026, first_fit=1000)
    b2.fit(A.REF_PERFACE)
    p2, info2, _ = b2.results()
    assert np.array_equal(p2, p[1000:2000]) and np.array_equal(info2, info[1000:2000])
    good = np.isin(info[:, 6].astype(int), CONVERGED)
    assert good.mean() > 0.9

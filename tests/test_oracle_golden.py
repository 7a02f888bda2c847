"""The oracle's BRDF fits (restated dlevmar_bc_dif + BRDFFunc, brdfdata.cpp:1058,1119) must
reproduce the committed golden vectors -- produced by the reference's own levmar -- bit for bit,
and agree with oracle/_ref live when that library is present."""
import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O

GOLD = G.load()


@pytest.mark.parametrize("case", GOLD["global"], ids=[c["name"] for c in GOLD["global"]])
def test_global_fit_matches_golden(case):
    c, td, th, x = G.global_inputs(case)
    ret, p, info = O.brdf_fit(O.oracle(), "oracle_", c, td, th, x, case["model"], getattr(O, case["preset"]))
    assert ret == case["ret"]
    assert p.tobytes() == case["p"].tobytes()
    assert info.tobytes() == case["info"].tobytes()


@pytest.mark.parametrize("case", GOLD["batch"], ids=[c["name"] for c in GOLD["batch"]])
def test_batched_fits_match_golden(case):
    c, td, th, x, _ = G.batch_inputs(case)
    for f in range(case["nfit"]):
        ret, p, info = O.brdf_fit(O.oracle(), "oracle_", c[f], td[f], th[f], x[f], case["model"], O.REF_PERFACE)
        assert ret == case["ret"][f]
        assert p.tobytes() == case["p"][f].tobytes()
        assert info.tobytes() == case["info"][f].tobytes()


def test_solve_equation_presets_match_brdf_fit():
    """oracle_solve_equation / _single are the two reference presets (brdfdata.cpp:1077-1136, 991-1075)."""
    lib = O.oracle()
    case = GOLD["batch"][0]
    c, td, th, x, _ = G.batch_inputs(case)
    p = np.zeros(3); info = np.zeros(10)
    ret = lib.oracle_solve_equation(O.as_d(c[3]), O.as_d(td[3]), O.as_d(th[3]), O.as_d(x[3]), case["nper"], 1,
                                    O.as_d(p), O.as_d(info))
    assert ret == case["ret"][3] and p.tobytes() == case["p"][3].tobytes()
    g = GOLD["global"][0]
    c, td, th, x = G.global_inputs(g)
    ret = lib.oracle_solve_equation_single(O.as_d(c), O.as_d(td), O.as_d(th), O.as_d(x), g["n"], 1, O.as_d(p),
                                           O.as_d(info))
    assert ret == g["ret"] and p.tobytes() == g["p"].tobytes() and info.tobytes() == g["info"].tobytes()


def test_live_reference_on_fresh_seed():
    ref = O.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    import synth
    c, td, th, x = synth.samples(4000, seed=31337)
    for preset in (O.REF_GLOBAL, O.REF_PERFACE):
        a = O.brdf_fit(ref, "", c, td, th, x, 1, preset)
        b = O.brdf_fit(O.oracle(), "oracle_", c, td, th, x, 1, preset)
        assert a[0] == b[0] and a[1].tobytes() == b[1].tobytes() and a[2].tobytes() == b[2].tobytes()


def test_primitives_bit_exact_vs_reference():
    ref = O.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    lib = O.oracle()
    for n in (1, 7, 8, 9, 64, 1001):   # L2nrmxmy unrolls by 8 (misc_core.c:721-807)
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        e0, e1 = np.zeros(n), np.zeros(n)
        a = ref.dlevmar_L2nrmxmy(O.as_d(e0), O.as_d(x), O.as_d(y), n)
        b = lib.oracle_L2nrmxmy(O.as_d(e1), O.as_d(x), O.as_d(y), n)
        assert a == b and e0.tobytes() == e1.tobytes()
    for m in (1, 2, 3, 5, 8):          # Axb_core.c:1140-1277
        A = rng.standard_normal((m, m)); B = rng.standard_normal(m)
        x0, x1 = np.zeros(m), np.zeros(m)
        r0 = ref.dAx_eq_b_LU_noLapack(O.as_d(A.copy()), O.as_d(B.copy()), O.as_d(x0), m)
        r1 = lib.oracle_Ax_eq_b_LU(O.as_d(A.copy()), O.as_d(B.copy()), O.as_d(x1), m)
        assert r0 == r1 and x0.tobytes() == x1.tobytes()
    ref.dAx_eq_b_LU_noLapack(None, None, None, 0)

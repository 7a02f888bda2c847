"""-m gpu parity tests of the global-fit path (K2/K3 kernels + LM loop) through the C-ABI, against
the CPU oracle on the same seeded inputs and against the committed golden vectors.

Tolerances are the ones BASELINE.json's north_star states: per-sample residuals 1e-6 relative,
fitted parameters 1e-4 relative, final cost 1e-6 relative (fp64)."""
import numpy as np
import pytest

import golden_lib as G
import gpu_common as GC
import oracle_lib as O
import synth
from brdf_b200 import api as A

pytestmark = pytest.mark.gpu

RES_RTOL, PAR_RTOL, COST_RTOL = 1e-6, 1e-4, 1e-6
GOLD = G.load()


@pytest.fixture(scope="module")
def ctx():
    c = A.Context()
    yield c
    c.close()


def _t(td, th, model):
    return td if model == 1 else th


@pytest.mark.parametrize("model", [1, 0])
def test_brdffunc_matches_reference_callback(model):
    """brdfgpu_BRDFFunc == BRDFFunc (brdfdata.cpp:969-989), including pow() special cases."""
    c, td, th, _ = synth.samples(4097, model_id=model, seed=5)
    td = td.copy(); th = th.copy()
    for arr in (td, th):
        arr[::13] *= -1.0
        arr[7] = 0.0
        arr[8] = 1.0
    extra, _keep = A.make_extra(c, td, th, model)
    for p in ([0.6, 0.35, 12.0], [0.5, 1.0, 1.0], [0.2, 0.3, 0.0], [0.1, 0.2, 2.0], [0.3, 0.1, 3.5]):
        got = A.BRDFFunc(p, extra, c.size)
        want = GC.oracle_predict(p, c, td, th, model)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        np.testing.assert_allclose(got[ok], want[ok], rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("model", [1, 0])
def test_analytic_jacobian_entry(model):
    c, td, th, _ = synth.samples(1000, model_id=model, seed=6)
    extra, _keep = A.make_extra(c, td, th, model)
    p = [0.6, 0.35, 7.5]
    np.testing.assert_allclose(A.BRDFJac(p, extra, c.size), GC.oracle_analytic_jacobian(p, c, td, th, model), rtol=1e-12,
                               atol=1e-300)


@pytest.mark.parametrize("model", [1, 0])
@pytest.mark.parametrize("n", [3, 4, 255, 256, 1001, 65537])
def test_residuals_match_oracle(ctx, model, n):
    c, td, th, x = synth.samples(n, model_id=model, seed=100 + n)
    s = ctx.upload(c, _t(td, th, model), x, model)
    for p in ([0.6, 0.35, 12.0], [0.0, 0.0, 0.0], [0.5, 1.0, 1.0], [1.0, 2.0, 55.5]):
        e = ctx.residuals(s, p)
        want = x - GC.oracle_predict(p, c, td, th, model)
        # relative to the prediction scale: e itself crosses zero
        hx = np.abs(x - want) + np.abs(x)
        assert np.max(np.abs(e - want) / np.maximum(hx, 1e-300)) < RES_RTOL


def test_residuals_nan_semantics(ctx):
    """pow(negative, non-integer) = NaN, pow(negative, integer) finite."""
    c, td, th, x = synth.samples(2000, seed=9)
    td = td.copy(); td[::7] *= -1.0
    s = ctx.upload(c, td, x, 1)
    for p in ([0.5, 1.0, 1.0], [0.5, 1.0, 2.0], [0.5, 1.0, 1.5], [0.5, 1.0, 0.0]):
        e = ctx.residuals(s, p)
        want = x - GC.oracle_predict(p, c, td, None, 1)
        assert np.array_equal(np.isnan(e), np.isnan(want))
        ok = ~np.isnan(want)
        np.testing.assert_allclose(e[ok], want[ok], rtol=1e-9, atol=1e-12)
        cost = ctx.cost(s, p)
        assert cost[1] == np.count_nonzero(~np.isfinite(want))


@pytest.mark.parametrize("model", [1, 0])
@pytest.mark.parametrize("delta", [1.0, 1e-6, -1e-3])
def test_normal_equations_match_levmar_definition(ctx, model, delta):
    """K2 == func + fdif_forw/cent_jac_approx + J^T J, J^T e, ||e||^2 (misc_core.c:137-211, lmbc_core.c:573-632)."""
    n = 50001
    c, td, th, x = synth.samples(n, model_id=model, seed=77)
    s = ctx.upload(c, _t(td, th, model), x, model)
    for p in ([0.6, 0.35, 12.0], [0.0, 0.0, 0.0], [0.5, 1.0, 1.0]):
        got = ctx.normal_eq(s, p, delta, A.JAC_FD)
        jac = GC.oracle_fd_jacobian(p, c, td, th, model, delta)
        e = x - GC.oracle_predict(p, c, td, th, model)
        want = GC.normal_eq_from(jac, e)
        scale = np.array([np.sqrt(want[0] * want[0]), np.sqrt(want[0] * want[3]), np.sqrt(want[0] * want[5]), want[3],
                          np.sqrt(want[3] * want[5]), want[5], 0, 0, 0, want[9]])
        scale[6:9] = np.sqrt(np.array([want[0], want[3], want[5]]) * want[9])
        # small steps amplify the rounding of f(p+d)-f(p) by 1/d on both sides
        tol = 1e-9 if abs(delta) >= 1e-3 else 1e-6
        assert np.all(np.abs(got[:10] - want) <= tol * np.maximum(scale, 1e-300)), (got[:10], want)
        assert got[10] == 0.0


def test_analytic_normal_equations(ctx):
    n = 20000
    c, td, th, x = synth.samples(n, seed=78)
    s = ctx.upload(c, td, x, 1)
    p = [0.55, 0.4, 9.0]
    got = ctx.normal_eq(s, p, 1.0, A.JAC_ANALYTIC)
    want = GC.normal_eq_from(GC.oracle_analytic_jacobian(p, c, td, None, 1), x - GC.oracle_predict(p, c, td, None, 1))
    np.testing.assert_allclose(got[:10], want, rtol=1e-9)


def test_device_synth_matches_numpy_recipe(ctx):
    for model in (1, 0):
        n = 100003
        s = ctx.synth(n, seed=4711, start=12345, model=model)
        c, t, x = s.download()
        rc, rtd, rth, rx = synth.samples(n, model_id=model, seed=4711, start=12345)
        assert np.array_equal(c, rc)
        assert np.array_equal(t, rtd if model == 1 else rth)
        # quantisation boundaries may flip for a handful of samples (device pow vs libm pow, 1 ulp)
        assert np.count_nonzero(x != rx) <= 2
        assert np.max(np.abs(x - rx)) <= 1.0 / 255.0 + 1e-15


def _check_fit(case, ret, p, info):
    if case["ret"] < 0:
        assert ret == A.LM_ERROR
        assert int(info[6]) == int(case["info"][6])
        return
    assert ret >= 0
    assert int(info[6]) == int(case["info"][6]), (info, case["info"])
    np.testing.assert_allclose(p, case["p"], rtol=PAR_RTOL)
    np.testing.assert_allclose(info[1], case["info"][1], rtol=COST_RTOL)
    np.testing.assert_allclose(info[0], case["info"][0], rtol=COST_RTOL)


@pytest.mark.parametrize("drive", [A.DRIVE_HOST, A.DRIVE_PERSISTENT], ids=["host", "persistent"])
@pytest.mark.parametrize("case", GOLD["global"], ids=[c["name"] for c in GOLD["global"]])
def test_global_fit_matches_reference_golden(ctx, case, drive):
    """Same (x, angles, p0, lb, ub, opts, itmax) in -> same p, info out as the reference's
    dlevmar_bc_dif (brdfdata.cpp:1058), golden vectors from oracle/_ref."""
    c, td, th, x = G.global_inputs(case)
    if case["name"] == "global_perface_1k_negcos":
        pytest.skip("trajectory through NaN Jacobians: checked in test_nan_jacobian_case")
    s = ctx.upload(c, _t(td, th, case["model"]), x, case["model"])
    ret, p, info = ctx.fit_global(s, getattr(A, case["preset"]), drive=drive)
    _check_fit(case, ret, p, info)


def test_nan_jacobian_case(ctx):
    """Real scenes have negative cosines.  With the per-face preset the start has an
    integer exponent (finite residuals) but the difference step does not: levmar then walks through
    NaN Jacobians and returns without an error code.  Same outcome required."""
    case = next(c for c in GOLD["global"] if c["name"] == "global_perface_1k_negcos")
    c, td, th, x = G.global_inputs(case)
    s = ctx.upload(c, td, x, 1)
    for drive in (A.DRIVE_HOST, A.DRIVE_PERSISTENT):
        ret, p, info = ctx.fit_global(s, A.REF_PERFACE, drive=drive)
        assert ret == case["ret"]
        assert int(info[6]) == int(case["info"][6])
        np.testing.assert_allclose(p, case["p"], rtol=PAR_RTOL)
        np.testing.assert_allclose(info[1], case["info"][1], rtol=COST_RTOL)


def test_levmar_signature_entry_point(ctx):
    """brdfgpu_dlevmar_bc_dif called exactly like brdfdata.cpp:1058 / :1119."""
    c, td, th, x = synth.samples(30000, seed=2468)
    for preset in (O.REF_GLOBAL, O.REF_PERFACE):
        want = O.brdf_fit(O.oracle(), "oracle_", c, td, th, x, 1, preset)
        extra, _keep = A.make_extra(c, td, th, 1)
        ret, p, info, covar = A.dlevmar_bc_dif(preset["p0"], x, preset["lb"], preset["ub"], preset["itmax"], preset["opts"],
                                               extra, want_covar=True)
        assert ret >= 0 and want[0] >= 0
        # stop reasons 3 (itmax) / 5 (no further reduction) flip with the summation order even on the
        # CPU; the converged ones must agree
        if int(info[6]) not in (3, 5) and int(want[2][6]) not in (3, 5):
            assert int(info[6]) == int(want[2][6])
        np.testing.assert_allclose(p, want[1], rtol=PAR_RTOL)
        np.testing.assert_allclose(info[1], want[2][1], rtol=COST_RTOL)
        assert np.all(np.isfinite(covar)) and np.allclose(covar, covar.T, rtol=1e-9)


def test_covariance_matches_levmar(ctx):
    c, td, th, x = synth.samples(5000, seed=1357)
    angles = np.concatenate([c, td, th])
    ed = O.make_extra(angles, 1)
    pr = O.REF_PERFACE
    want = O.levmar_bc_dif(O.oracle(), "oracle_", O.brdf_callback(), pr["p0"], x, pr["lb"], pr["ub"], pr["itmax"], pr["opts"],
                           adata=ed, want_covar=True)
    s = ctx.upload(c, td, x, 1)
    ret, p, info, covar = ctx.fit_global(s, A.REF_PERFACE, want_covar=True)
    np.testing.assert_allclose(p, want[1], rtol=PAR_RTOL)
    np.testing.assert_allclose(covar, want[3], rtol=1e-3)


def test_analytic_mode_agrees_with_small_step_fd(ctx):
    """exact partials are parity-valid where delta <= 1e-6 (per-face preset)."""
    c, td, th, x = synth.samples(40000, seed=8642)
    s = ctx.upload(c, td, x, 1)
    a = ctx.fit_global(s, A.REF_PERFACE, jac_mode=A.JAC_FD)
    b = ctx.fit_global(s, A.REF_PERFACE, jac_mode=A.JAC_ANALYTIC)
    np.testing.assert_allclose(a[1], b[1], rtol=PAR_RTOL)
    np.testing.assert_allclose(a[2][1], b[2][1], rtol=COST_RTOL)


def test_edge_cases(ctx):
    # n < m: LM_ERROR like lmbc_core.c:440-443
    c, td, th, x = synth.samples(2, seed=1)
    s = ctx.upload(c, td, x, 1)
    assert ctx.fit_global(s, A.REF_PERFACE)[0] == A.LM_ERROR
    # lb > ub: LM_ERROR (lmbc_core.c:451-454)
    c, td, th, x = synth.samples(64, seed=1)
    s = ctx.upload(c, td, x, 1)
    bad = dict(A.REF_PERFACE, lb=(0, 2, 0), ub=(1, 1, 1))
    assert ctx.fit_global(s, bad)[0] == A.LM_ERROR
    # x == NULL means zeros (lmbc_core.c:373)
    s0 = ctx.upload(c, td, None, 1)
    e = ctx.residuals(s0, [0.5, 1.0, 1.0])
    np.testing.assert_allclose(e, -GC.oracle_predict([0.5, 1.0, 1.0], c, td, None, 1), rtol=1e-12)
    # infeasible start is projected (lmbc_core.c:514-520)
    ret, p, info = ctx.fit_global(s, A.REF_PERFACE, p0=(-1.0, 500.0, 1.0))
    want = O.brdf_fit(O.oracle(), "oracle_", c, td, th, x, 1, dict(O.REF_PERFACE, p0=(-1.0, 500.0, 1.0)))
    assert (ret >= 0) == (want[0] >= 0)
    if int(want[2][6]) == 2:
        np.testing.assert_allclose(p, want[1], rtol=PAR_RTOL)


def test_full_size_properties(ctx):
    """BASELINE config 2/5 sizes (10^6 .. 10^7): size-independent properties.
    (a) the sums are additive over shards (checksum of checksums), (b) deterministic run to run,
    (c) host-driven and persistent drivers agree, (d) the fit recovers the generating parameters."""
    n = 10_000_000
    whole = ctx.synth(n, seed=synth.DEFAULT_SEED)
    p = [0.58, 0.36, 11.0]
    w = ctx.normal_eq(whole, p, 1.0)
    parts = np.zeros(11)
    for k in range(4):
        lo, hi = k * n // 4, (k + 1) * n // 4
        parts += ctx.normal_eq(ctx.synth(hi - lo, seed=synth.DEFAULT_SEED, start=lo), p, 1.0)
    np.testing.assert_allclose(parts[:10], w[:10], rtol=1e-11)
    assert np.array_equal(ctx.normal_eq(whole, p, 1.0), w)
    np.testing.assert_allclose(ctx.cost(whole, p)[0], w[9], rtol=1e-12)
    del whole
    s = ctx.synth(1_000_000, seed=synth.DEFAULT_SEED)
    a = ctx.fit_global(s, A.REF_GLOBAL, drive=A.DRIVE_HOST)
    b = ctx.fit_global(s, A.REF_GLOBAL, drive=A.DRIVE_PERSISTENT)
    assert a[0] >= 0 and b[0] >= 0
    np.testing.assert_allclose(a[1], b[1], rtol=PAR_RTOL)
    np.testing.assert_allclose(a[2][1], b[2][1], rtol=COST_RTOL)
    np.testing.assert_allclose(b[1], [0.6, 0.35, 12.0], rtol=2e-2)
    # survey probe of the reference on this exact generator: cost 10.4517173244 --
    # generator details (noise stream) differ from that probe, so only the scale is asserted
    assert 5.0 < b[2][1] < 20.0


@pytest.mark.parametrize("drive", [A.DRIVE_PERSISTENT, A.DRIVE_HOST], ids=["persistent", "host"])
def test_diagonal_scaling_matches_levmar(ctx, drive):
    """dscl (lmbc_core.c:360-366, 536-570): scaled variables and bounds inside, caller coordinates outside,
    covariance rescaled; both drivers, incl. the batched projected-gradient walk of the persistent kernel."""
    c, td, th, x = synth.samples(20000, seed=4242)
    angles = np.concatenate([c, td, th])
    dscl = (0.5, 2.0, 10.0)
    for pr in (O.REF_PERFACE, O.REF_GLOBAL):
        want = O.levmar_bc_dif(O.oracle(), "oracle_", O.brdf_callback(), pr["p0"], x, pr["lb"], pr["ub"], pr["itmax"], pr["opts"],
                               adata=O.make_extra(angles, 1), dscl=dscl, want_covar=True)
        s = ctx.upload(c, td, x, 1)
        ret, p, info, covar = ctx.fit_global(s, dict(pr), drive=drive, dscl=dscl, want_covar=True)
        s.free()
        assert (ret >= 0) == (want[0] >= 0)
        np.testing.assert_allclose(p, want[1], rtol=PAR_RTOL)
        np.testing.assert_allclose(info[1], want[2][1], rtol=COST_RTOL)
        np.testing.assert_allclose(covar, want[3], rtol=5e-3, atol=1e-14)


@pytest.mark.parametrize("tma", ["0", "1"], ids=["register-prefetch", "tma-ring"])
def test_sums_are_additive_over_shards(ctx, tma, monkeypatch):
    """Size-independent property the multi-GPU mode rests on: J^T J, J^T e and ||e||^2 of a
    sample set are the sums over any partition of it -- checked at 4e6 samples (ragged, odd shards) for
    both streaming implementations of the Jacobian pass, against each other and against the oracle on a
    prefix the CPU finishes in a second."""
    monkeypatch.setenv("BRDFGPU_TMA", tma)
    c2 = A.Context()          # BRDFGPU_TMA is read once per context
    n = 4_000_001
    p, delta = (0.55, 0.4, 9.5), 1.0
    whole = c2.synth(n, seed=909)
    total = c2.normal_eq(whole, p, delta)[:10]
    c, t, x = whole.download()
    cuts = [0, 1_000_003, 2_777_777, n]
    parts = np.zeros(10)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        s = c2.upload(c[lo:hi], t[lo:hi], x[lo:hi], 1)
        parts += c2.normal_eq(s, p, delta)[:10]
        s.free()
    np.testing.assert_allclose(parts, total, rtol=1e-12)
    assert np.isclose(c2.cost(whole, p)[0], total[9], rtol=1e-12)
    m = 200_000
    jac = GC.oracle_fd_jacobian(p, c[:m], t[:m], np.zeros(m), 1, delta)
    e = x[:m] - GC.oracle_predict(p, c[:m], t[:m], np.zeros(m), 1)
    s = c2.upload(c[:m], t[:m], x[:m], 1)
    np.testing.assert_allclose(c2.normal_eq(s, p, delta)[:10], GC.normal_eq_from(jac, e), rtol=1e-9)
    s.free(); whole.free(); c2.close()


@pytest.mark.parametrize("case", ["ref_global", "ref_global_small", "ref_perface", "analytic", "phong", "dscl", "negative_cosines",
                                  "streamed", "tiny_cosines", "tiny_cosines_streamed"])
def test_speculative_jacobians_change_nothing(ctx, case, monkeypatch):
    """The persistent fit answers cost requests at likely next iterates with Jacobian sweeps, fuses the
    last line-search probe with the Jacobian at the first projected-gradient candidate and starts walks
    at the last walk's batch width.  All of it is speculation about WHICH sweeps to run: p and every
    entry of info[] (including the nfev / njev counters) must equal the unspeculated run bit for bit,
    while fewer sweeps are made."""
    n, model, preset, kw = 300_001, 1, A.REF_GLOBAL, {}
    if case == "ref_global_small":
        n = 20_001
    elif case == "ref_perface":
        preset = A.REF_PERFACE
    elif case == "analytic":
        kw = {"jac_mode": A.JAC_ANALYTIC}
    elif case == "phong":
        model = 0
    elif case == "dscl":
        kw = {"dscl": (1.0, 0.5, 10.0)}
    elif case == "streamed":
        n = 2_500_001   # beyond on-chip residency: part of the samples goes through the TMA ring every sweep
    c, td, th, x = synth.samples(n, model_id=model, seed=77)
    t = (td if model == 1 else th).copy()
    if case == "negative_cosines":   # careful path (libm pow semantics) inside both kinds of sweep
        t[::97] *= -1.0
        t[5] = 0.0
    if case.startswith("tiny_cosines"):
        # cosines so small that t**(n+d) leaves the fast exponential's range while t**n does not (they occur in the
        # photographed scenes): the Jacobian pass redoes such a sample through pow(), a cost pass does not
        if case.endswith("streamed"):
            c, td, th, x = synth.samples(2_500_001, model_id=model, seed=78)
            t = td.copy()
        t[::53] = np.exp(-np.linspace(20.0, 80.0, t[::53].size))
    s = ctx.upload(c, t, x, model)
    runs = {}
    for mask in ("0", "1", "3", "7"):
        monkeypatch.setenv("BRDFGPU_SPEC_JAC", mask)
        ret, p, info = ctx.fit_global(s, preset, **kw)
        runs[mask] = (ret, p.tobytes(), info.tobytes(), ctx.fit_stats())
    base = runs["0"]
    assert base[3]["spec_jac_issued"] == 0 and base[3]["spec_jac_hits"] == 0
    for mask in ("1", "3", "7"):
        r = runs[mask]
        assert r[0] == base[0] and r[1] == base[1] and r[2] == base[2], mask
    full = runs["7"][3]
    assert full["spec_jac_hits"] <= full["spec_jac_issued"]
    if case not in ("negative_cosines",):
        assert full["spec_jac_hits"] > 0
        assert full["jac_passes"] + full["cost_passes"] < base[3]["jac_passes"] + base[3]["cost_passes"]
    s.free()


@pytest.mark.parametrize("n", [200_001, 2_600_001])
def test_every_kind_of_sweep_gives_the_same_cost_bits(ctx, n, monkeypatch):
    """What the speculation rests on: ||x - f(p)||^2 of one point has the same bits whether the persistent kernel
    gets it from a cost sweep, a Jacobian sweep, a Jacobian sweep with extra cost points, a batch of candidates or a
    wide batch of the lane-parallel walk (BRDFGPU_SPEC_JAC=8 self-check: info[0..5]) -- on chip (n small) and with a
    streamed part, with negative, zero and tiny cosines in the data."""
    c, td, th, x = synth.samples(n, seed=91)
    t = td.copy()
    t[::97] *= -1.0
    t[5] = 0.0
    t[::53] = np.exp(-np.linspace(20.0, 80.0, t[::53].size))
    s = ctx.upload(c, t, x, A.BLINN_PHONG)
    monkeypatch.setenv("BRDFGPU_SPEC_JAC", "8")
    for p0 in ((0.6, 0.35, 12.0), (0.0625, 0.196, 0.0), (0.3, 0.2, 2.0), (0.06, 0.2, 17.0), (0.5, 1.0, 1.0), (0.2, 0.4, 33.0)):
        preset = dict(A.REF_GLOBAL)
        preset["p0"] = p0
        ret, p, info = ctx.fit_global(s, preset)
        # (resident shards: a wide batch of 13 or 32 candidates that gave the point another value put it into info[5])
        v = info[:6]
        assert np.isfinite(v).all(), (p0, v)
        assert all(q.tobytes() == v[0].tobytes() for q in v), (p0, [repr(float(q)) for q in v])
    s.free()


def test_solve_equation_single_colmajor_is_the_reference_flattening():
    """brdfdata.cpp:1008-1042: x row-major, angles through Eigen's column-major linear index -- against
    levmar on arrays flattened the same way in numpy."""
    rows, nimg = 700, 16
    c, td, th, x = synth.samples(rows * nimg, seed=4321)
    phi, tdm, thm, I = (a.reshape(rows, nimg) for a in (c, td, th, x))
    ret, p, info = A.solve_equation_single_colmajor(phi, tdm, thm, I, 1)
    k = np.arange(rows * nimg)
    lin = lambda m: m[k % rows, k // rows]          # Eigen: m(k) on a column-major rows x nimg matrix
    wret, wp, winfo = O.brdf_fit(O.oracle(), "oracle_", lin(phi), lin(tdm), lin(thm), I.ravel(), 1, O.REF_GLOBAL)
    assert (ret >= 0) == (wret >= 0)
    np.testing.assert_allclose(p, wp, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(info[1], winfo[1], rtol=1e-6)
    # the aligned order on the same data is a different (well-posed) problem with a different answer
    ret2, p2, _ = A.solve_equation_single(c, td, th, x, 1)
    assert not np.allclose(p, p2, rtol=1e-3)


@pytest.mark.parametrize("n, resident", [(1_250_000, True), (1_300_000, False)])
def test_shard_residency_boundary(ctx, n, resident):
    """persistent_plan(): a shard of 1.25e6 samples (BASELINE's 10^7 strong-scaled over 8 GPUs) still lives entirely in
    shared memory -- it takes the room of half the exchange's transposition area -- and 1.3e6 streams its tail through
    the TMA ring; both agree with the kernel-per-evaluation driver within the parity tolerances."""
    s = ctx.synth(n, seed=synth.DEFAULT_SEED)
    a = ctx.fit_global(s, A.REF_GLOBAL, drive=A.DRIVE_PERSISTENT)
    st = ctx.fit_stats()
    assert (st["resident_samples"] >= n - 1) == resident, st
    assert st["ctas"] > 0 and st["cyc_total"] > 0          # the persistent kernel ran (no silent change of driver)
    b = ctx.fit_global(s, A.REF_GLOBAL, drive=A.DRIVE_HOST)
    assert a[0] >= 0 and b[0] >= 0
    np.testing.assert_allclose(a[1], b[1], rtol=PAR_RTOL)
    np.testing.assert_allclose(a[2][1], b[2][1], rtol=COST_RTOL)
    s.free()


@pytest.mark.parametrize("n, preset", [(1_000_000, "REF_GLOBAL"), (300_001, "REF_GLOBAL"), (200_000, "REF_PERFACE")])
def test_wide_walk_batches_change_nothing(ctx, n, preset, monkeypatch):
    """A resident shard lets the projected-gradient walk evaluate up to 32 candidates per sweep (the futile 393-candidate
    walk of the 10^6 case: 17 sweeps instead of 52); candidates past the deciding one are discarded uncounted, so p and
    every entry of info[] equal the run held to 8 candidates per sweep (BRDFGPU_SPEC_JAC bit 64) bit for bit."""
    s = ctx.synth(n, seed=synth.DEFAULT_SEED)
    runs = {}
    for mask in ("7", "71"):
        monkeypatch.setenv("BRDFGPU_SPEC_JAC", mask)
        ret, p, info = ctx.fit_global(s, getattr(A, preset))
        st = ctx.fit_stats()
        runs[mask] = (ret, p.tobytes(), info.tobytes(), st["jac_passes"] + st["cost_passes"], st["cost_points"])
    assert runs["7"][:3] == runs["71"][:3]
    assert runs["7"][3] <= runs["71"][3]          # never more sweeps ...
    if n == 1_000_000:
        assert runs["7"][3] < runs["71"][3]       # ... and fewer where a long walk occurs
    s.free()

"""-m gpu: the levmar-exact batched mode (BRDFGPU_JAC_FD_EXACT, csrc/batched_exact.cu) EQUALS the reference's levmar --
parameters, all ten info[] entries and the return value of every fit, not a tolerance -- on every fit of BASELINE
configs[3] (65 536 x 64 samples) and of the per-face path on img/cup (every mapped face x 3 channels, CalcBRDFEquation
brdfdata.cpp:1188-1227), including the 86 % of the cup fits that levmar abandons at itmax.  Reference side: the stored
results of the reference's own levmar (tests/golden/*_full_*.npz; parameters there are float32, so equality of p is
checked after the same rounding, and exactly -- float64, all info[] -- against live reference runs on a sample)."""
import numpy as np
import pytest

import oracle_lib as O
import parity_lib as P
import real_scenes as R
import synth
from brdf_b200 import api as A

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = A.Context()
    yield c
    c.close()


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _against_stored(p, info, ret, ref):
    assert np.array_equal(ret, ref["ret"].astype(np.int32))
    assert np.array_equal(info[:, 6].astype(int), ref["reason"].astype(int))
    assert np.array_equal(info[:, 5].astype(int), ref["iters"].astype(int))
    assert np.array_equal(info[:, 7].astype(int), ref["nfev"].astype(int))
    assert _same(info[:, 1], ref["cost"])
    assert _same(p.astype(np.float32), ref["p"])


def _against_live(c, td, x, p, info, ret, picks):
    lib, prefix = (O.ref(), "") if O.ref() is not None else (O.oracle(), "oracle_")
    for f in picks:
        wret, wp, winfo = O.brdf_fit(lib, prefix, c[f], td[f], None, x[f], 1, O.REF_PERFACE)
        assert wret == ret[f] and _same(wp, p[f]) and _same(winfo, info[f]), (f, wp, p[f], winfo, info[f])


def test_configs3_every_fit_equals_the_reference(ctx):
    ref = P.load_full("batched_full_cfg3.npz")
    nfit, nper = ref["p"].shape[0], int(ref["nper"])
    b = ctx.batch_synth(nfit, nper, seed=int(ref["seed"]))
    b.fit(A.REF_PERFACE, jac_mode=A.JAC_FD_EXACT)
    p, info, ret = b.results()
    b.free()
    _against_stored(p, info, ret, ref)
    c, td, th, x, _ = synth.batched(512, nper, seed=int(ref["seed"]))
    _against_live(c, td, x, p, info, ret, range(0, 512, 4))


def test_cup_every_per_face_fit_equals_the_reference(ctx):
    sc = R.load("cup")
    if sc is None:
        pytest.skip("tests/_scenes/cup.npz absent")
    ref = P.load_full("perface_full_cup.npz")
    nfit = int(ref["nfit"])
    scene = ctx.scene(sc["V"], sc["F"], sc["imgs"], dark=sc["dark"])
    g = scene.gather(sc["cams"][:1])
    ps, infos, rets = [], [], []
    for ch in range(3):
        _, b, n = scene.gather_resident(sc["cams"][:1], model=A.BLINN_PHONG, channel=ch, want_global=False, want_batch=True)
        assert n == nfit
        b.fit(A.REF_PERFACE, jac_mode=A.JAC_FD_EXACT)
        p, info, ret = b.results()
        ps.append(p); infos.append(info); rets.append(ret)
        b.free()
    p, info, ret = np.concatenate(ps), np.concatenate(infos), np.concatenate(rets)
    _against_stored(p, info, ret, ref)
    # live, in float64: a spread over the faces, among them faces that see an LED from behind (pow(negative, n) = NaN)
    # the public driver stores the same parameters per face and channel (SaveValuesToSurface, brdfdata.cpp:368-377)
    n2, surf = scene.calc_brdf_equation(sc["cams"][0])
    assert n2 == nfit
    for ch in range(3):
        assert np.array_equal(surf[ref["fit_face"], ch], p[ch * nfit:(ch + 1) * nfit])
    neg = np.flatnonzero((g["thetaDash"] < 0).any(axis=1))[:40]
    picks = sorted(set(range(0, nfit, 311)) | set(neg.tolist()))
    _against_live(g["phi"], g["thetaDash"], g["I"][1], p[nfit:2 * nfit], info[nfit:2 * nfit], ret[nfit:2 * nfit], picks)
    scene.free()


@pytest.mark.parametrize("nper", [3, 7, 16, 17, 33, 64, 100, 128])
def test_exact_mode_other_sizes(ctx, nper):
    """Ragged sizes walk every branch of dlevmar_L2nrmxmy's remainder switch and both lane-group widths."""
    nfit = 64
    c, td, th, x, _ = synth.batched(nfit, nper, seed=1200 + nper)
    b = ctx.batch_upload(c, td, x, A.BLINN_PHONG)
    b.fit(A.REF_PERFACE, jac_mode=A.JAC_FD_EXACT)
    p, info, ret = b.results()
    b.free()
    _against_live(c, td, x, p, info, ret, range(nfit))


def test_exact_mode_phong_and_limits(ctx):
    nfit, nper = 48, 16
    c, td, th, x, _ = synth.batched(nfit, nper, model_id=0, seed=77)
    b = ctx.batch_upload(c, th, x, A.PHONG)
    b.fit(A.REF_PERFACE, jac_mode=A.JAC_FD_EXACT)
    p, info, ret = b.results()
    b.free()
    lib, prefix = (O.ref(), "") if O.ref() is not None else (O.oracle(), "oracle_")
    for f in range(nfit):
        wret, wp, winfo = O.brdf_fit(lib, prefix, c[f], td[f], th[f], x[f], 0, O.REF_PERFACE)
        assert wret == ret[f] and _same(wp, p[f]) and _same(winfo, info[f]), f
    big = ctx.batch_synth(4, 200, seed=5)
    with pytest.raises(A.BrdfGpuError):
        big.fit(A.REF_PERFACE, jac_mode=A.JAC_FD_EXACT)     # beyond levmar's small-problem branch
    central = dict(A.REF_PERFACE, opts=(1e-3, 1e-15, 1e-15, 1e-20, -1e-6))
    small = ctx.batch_synth(4, 16, seed=5)
    with pytest.raises(A.BrdfGpuError):
        small.fit(central, jac_mode=A.JAC_FD_EXACT)
    big.free(); small.free()

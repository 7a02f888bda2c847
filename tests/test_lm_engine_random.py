"""Randomised differential test of the product's control loops on the host (no GPU): small BRDF fits
with random truth / noise / starts, bounds as in the reference (brdfdata.cpp:1112-1113), solved by
  * the reference's dlevmar_bc_der (oracle/_ref, else the oracle port) with Python callbacks,
  * lm_engine.cuh (brdfgpu_lm_bc_reduced), with and without the batched projected-gradient walk,
all fed by the same callbacks with the normal equations formed in levmar's own order -- so every
trajectory must be bit-identical: p and all of info[0..9].  These fits hit everything the BRDF path
exercises on real data: active bounds, rejected steps, line searches, long projected-gradient walks,
itmax, and pow() of negative cosines (NaN residuals, stop reason 7)."""
import math

import numpy as np
import pytest

import kat_problems as K
import oracle_lib as O
from brdf_b200 import api as A
from test_lm_engine_host import _reduced_callbacks, _ref_or_oracle

OPTS = (1e-3, 1e-15, 1e-15, 1e-20, 1e-6)


def _problem(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(8, 40))
    c = rng.uniform(0.0, 1.0, n)
    t = rng.uniform(0.0, 1.0, n)
    if seed % 7 == 3:
        t[rng.integers(0, n)] = -0.2          # a face seen from behind: pow(negative, non-integer) = NaN
    truth = (rng.uniform(0.1, 0.9), rng.uniform(0.05, 0.8), rng.uniform(1.0, 50.0))
    x = np.clip(np.floor(255.0 * (truth[0] * c + truth[1] * np.abs(t) ** truth[2] + rng.uniform(-0.005, 0.005, n))), 0, 255) / 255.0
    cl, tl = [float(v) for v in c], [float(v) for v in t]

    def f(p, hx, m, nn):                       # BRDFFunc, model 1 (brdfdata.cpp:985-986)
        for i in range(nn):
            try:
                pw = math.pow(tl[i], p[2])
            except (ValueError, OverflowError):
                pw = float("nan")
            hx[i] = p[0] * cl[i] + p[1] * pw

    def j(p, jac, m, nn):                      # exact partials
        for i in range(nn):
            try:
                pw = math.pow(tl[i], p[2])
                lg = math.log(tl[i]) if tl[i] > 0 else float("nan")
            except (ValueError, OverflowError):
                pw, lg = float("nan"), float("nan")
            jac[3 * i] = cl[i]
            jac[3 * i + 1] = pw
            jac[3 * i + 2] = p[1] * pw * lg

    p0 = (0.5, 1.0, 1.0) if seed % 3 else tuple(rng.uniform(0.0, 2.0, 3))
    return dict(name="brdf%d" % seed, driver="bc_der", m=3, n=n, p0=p0, x=[float(v) for v in x], lb=(0.0, 0.0, 0.0),
                ub=(100.0, 100.0, 100.0), itmax=60, f=f, j=j, info=None)


@pytest.mark.parametrize("seed", range(24))
def test_random_brdf_fits_bit_identical(seed):
    prob = _problem(seed)
    lib, prefix = _ref_or_oracle()
    fc, jc = K.callbacks(prob)
    x = np.array(prob["x"], dtype=np.float64)
    r_ret, r_p, r_info, _ = O.levmar_bc_der(lib, prefix, fc, jc, prob["p0"], x, prob["lb"], prob["ub"], prob["itmax"], OPTS)
    jac_cb, cost_cb = _reduced_callbacks(prob)
    outs = {}
    outs["engine"] = A.lm_bc_reduced(jac_cb, cost_cb, prob["p0"], prob["n"], prob["lb"], prob["ub"], prob["itmax"], OPTS)[:3]
    A.lib().brdfgpu_lm_reduced_batching(1)
    try:
        outs["engine, batched walk"] = A.lm_bc_reduced(jac_cb, cost_cb, prob["p0"], prob["n"], prob["lb"], prob["ub"], prob["itmax"], OPTS)[:3]
    finally:
        A.lib().brdfgpu_lm_reduced_batching(0)
    for name, (ret, p, info) in outs.items():
        assert ret == r_ret, name
        assert p.tobytes() == r_p.tobytes(), (name, p, r_p)
        if r_info[8] == 0:   # stopped before the first Jacobian: levmar derives info[4] from an uninitialised J^T J
            info, want_info = np.delete(info, 4), np.delete(r_info, 4)
        else:
            want_info = r_info
        assert info.tobytes() == want_info.tobytes(), (name, info, r_info)

"""BASELINE.json configs[4]: scaling sweep 10^4 .. 10^8 samples (global mode) and 10^3 .. 10^6 fits
(batched mode) on one GPU, against the HBM roofline.  Prints a markdown table (profiles/r01_sweep.md).
    python profiles/sweep.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from brdf_b200 import api as A  # noqa: E402

PEAK = 6537.3
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ctx = A.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    out = fn()
    e1.record(stream)
    ctx.synchronize()
    return e0.elapsed_time(e1), out


print("### global fit (REF_GLOBAL preset, forward differences delta = 1), one B200\n")
print("| samples | driver | ms per fit | iterations | nfev | sweeps | sample-evals/s | GB/s (24 B x samples x sweeps) | frac of %.0f |" % PEAK)
print("|---|---|---|---|---|---|---|---|---|")
for n in (10**4, 10**5, 10**6, 10**7, 10**8):
    s = ctx.synth(n, 88172645463325252)
    ctx.fit_global(s, A.REF_GLOBAL)
    ms, (ret, p, info) = timed(lambda: ctx.fit_global(s, A.REF_GLOBAL))
    st = ctx.fit_stats()
    sweeps = st["jac_passes"] + st["cost_passes"]
    gbs = 24.0 * n * sweeps / (ms * 1e-3) / 1e9
    where = ("persistent, %d%% of the samples on chip" % round(100.0 * min(1.0, st["resident_samples"] / n))) if st["ctas"] else "kernel per evaluation"
    print("| %.0e | %s | %.3f | %d | %d | %d | %.3g | %.0f | %.2f |" % (n, where, ms, info[5], info[7], sweeps, info[7] * n / (ms * 1e-3), gbs, gbs / PEAK))
    s.free()

print("\n### batched per-fit mode (REF_PERFACE preset), one B200\n")
print("| fits | samples per fit | ms | fits/s | mean iterations | mean nfev | sample-evals/s | converged |")
print("|---|---|---|---|---|---|---|---|")
for nfit, nper in ((10**3, 64), (10**4, 64), (65536, 64), (10**6, 64), (38342 * 3, 16), (10**6, 16)):
    b = ctx.batch_synth(nfit, nper, seed=2026)
    b.fit(A.REF_PERFACE)
    ctx.synchronize()
    ms, _ = timed(lambda: b.fit(A.REF_PERFACE))
    pp, info, ret = b.results()
    print("| %d | %d | %.2f | %.3g | %.1f | %.1f | %.3g | %.3f |" % (nfit, nper, ms, nfit / (ms * 1e-3), info[:, 5].mean(), info[:, 7].mean(),
                                                              info[:, 7].sum() * nper / (ms * 1e-3), np.isin(info[:, 6].astype(int), (1, 2, 6)).mean()))
    b.free()
ctx.close()

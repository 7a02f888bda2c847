"""Batched mode: default kernel vs the levmar-exact kernel (BRDFGPU_JAC_FD_EXACT), one B200.
    python profiles/exact_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from brdf_b200 import api as A  # noqa: E402

ctx = A.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)


def timed(fn, reps=3):
    fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.synchronize()
    return e0.elapsed_time(e1) / reps


print("| fits | samples per fit | mode | ms | fits/s | mean nfev |")
print("|---|---|---|---|---|---|")
for nfit, nper in ((65536, 64), (113007, 16), (8192, 64), (1000000, 16)):
    b = ctx.batch_synth(nfit, nper, seed=2026)
    for name, mode in (("default (exp, butterfly sums)", A.JAC_FD), ("levmar-exact (glibc pow, levmar's orders)", A.JAC_FD_EXACT)):
        ms = timed(lambda: b.fit(A.REF_PERFACE, jac_mode=mode))
        p, info, ret = b.results()
        print("| %d | %d | %s | %.2f | %.3g | %.1f |" % (nfit, nper, name, ms, nfit / (ms * 1e-3), info[:, 7].mean()))
    b.free()
ctx.close()

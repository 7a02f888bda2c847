"""Per-fit duration histogram of the batched kernels (debug variant: make VARIANT=_cyc EXTRA=-DBG_FIT_CYCLES, the
kernels then return SM cycles / 64 per fit in place of levmar's return value).
    make -C brdf_b200/csrc VARIANT=_cyc EXTRA=-DBG_FIT_CYCLES && BRDFGPU_LIB=$PWD/brdf_b200/libbrdfgpu_cyc.so python profiles/fit_cycles.py
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
print("| fits | kernel | launch ms | fit cycles: mean | median | p90 | p99 | max | max fit / launch | sum of fit time / (launch x resident warps) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for nfit in (65536, 8192):
    b = ctx.batch_synth(nfit, 64, seed=2026)
    for name, mode, warps_per_sm in (("levmar-exact", A.JAC_FD_EXACT, 16), ("fast", A.JAC_FD, 24)):
        b.fit(A.REF_PERFACE, jac_mode=mode)
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        b.fit(A.REF_PERFACE, jac_mode=mode)
        e1.record(stream)
        ctx.synchronize()
        ms = e0.elapsed_time(e1)
        _, info, cyc = b.results()
        cyc = cyc.astype(np.float64) * 64.0
        launch_cycles = ms * 1e-3 * 1.965e9
        busy = cyc.sum() / (launch_cycles * 148 * warps_per_sm)
        print("| %d | %s | %.2f | %.3g | %.3g | %.3g | %.3g | %.3g | %.2f | %.2f |" % (
            nfit, name, ms, cyc.mean(), np.median(cyc), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max(),
            cyc.max() / launch_cycles, busy))
        # correlation of duration with levmar's evaluation count
        print("|  |  | cycles per evaluation (median) %.0f; corr(cycles, nfev) %.3f |" % (np.median(cyc / info[:, 7]), np.corrcoef(cyc, info[:, 7])[0, 1]))
    b.free()
ctx.close()

"""Resident configs[1] fit through the persistent kernel: ms per fit and the in-kernel cycle accounting.
    [BRDFGPU_LIB=...variant.so] python profiles/persist_bench.py [n]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from brdf_b200 import api as A  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = A.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
s = ctx.synth(n, 88172645463325252)
for _ in range(5):
    ret, p, info = ctx.fit_global(s, A.REF_GLOBAL)
ctx.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(100):
    ret, p, info = ctx.fit_global(s, A.REF_GLOBAL)
e1.record(stream)
ctx.synchronize()
st = ctx.fit_stats()
sweeps = st["jac_passes"] + st["cost_passes"]
ctl = st["cyc_total"] - st["cyc_sweep"] - st["cyc_exchange"]
print("n=%d: %.4f ms per fit, %d iterations, %d sweeps (%d points); cycles/sweep: sweep %.0f exchange %.0f %s control %.0f; p=%s cost=%.15g nfev=%d" % (
    n, e0.elapsed_time(e1) / 100, info[5], sweeps, st["cost_points"], st["cyc_sweep"] / sweeps, st["cyc_exchange"] / sweeps,
    [int(v / sweeps) for v in st["cyc_exchange_phases"]], ctl / sweeps, p, info[1], info[7]))
ctx.close()

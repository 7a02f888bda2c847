"""A/B of the speculative Jacobians of the persistent fit (BRDFGPU_SPEC_JAC=0 switches them off):
results must be bit-identical, only the number of sweeps and the time change."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from brdf_b200 import api as A

ctx = A.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
SEED = 88172645463325252


def run(s, preset, reps=5, **kw):
    for _ in range(2):
        ctx.fit_global(s, preset, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ret, p, info = ctx.fit_global(s, preset, **kw)
    e1.record(stream); ctx.synchronize()
    return e0.elapsed_time(e1) / reps, ret, p, info, ctx.fit_stats()


cases = [("REF_GLOBAL 1e6", 1_000_000, 0, A.REF_GLOBAL, {}),
         ("REF_GLOBAL 1e6 off1", 1_000_000, 1_000_000, A.REF_GLOBAL, {}),
         ("REF_GLOBAL 2e5", 200_000, 0, A.REF_GLOBAL, {}),
         ("REF_PERFACE 1e6", 1_000_000, 0, A.REF_PERFACE, {}),
         ("REF_GLOBAL analytic 1e6", 1_000_000, 0, A.REF_GLOBAL, {"jac_mode": A.JAC_ANALYTIC}),
         ("REF_GLOBAL dscl 1e6", 1_000_000, 0, A.REF_GLOBAL, {"dscl": (1.0, 0.5, 10.0)}),
         ("REF_GLOBAL 3e6 (streamed)", 3_000_000, 0, A.REF_GLOBAL, {}),
         ("REF_GLOBAL 1e7 (streamed)", 10_000_000, 0, A.REF_GLOBAL, {})]
for name, n, start, preset, kw in cases:
    s = ctx.synth(n, SEED, start=start)
    out = {}
    for mode in ("0", "1", "3", "5", "7"):
        os.environ["BRDFGPU_SPEC_JAC"] = mode
        out[mode] = run(s, preset, **kw)
        st = out[mode][4]
        sw = st["jac_passes"] + st["cost_passes"]
        ctl = st["cyc_total"] - st["cyc_sweep"] - st["cyc_exchange"]
        print("   mode %s: %.3f ms, %d sweeps, per sweep: %d cyc = sweep %d + exchange %d + control %d; control by next: %s"
              % (mode, out[mode][0], sw, st["cyc_total"] // sw, st["cyc_sweep"] // sw, st["cyc_exchange"] // sw, ctl // sw,
                 {k: v // 1000 for k, v in st["cyc_control_by_next_sweep"].items() if v}))
    a, b = out["0"], out["7"]
    for mode in ("1", "3", "5"):
        m = out[mode]
        assert m[2].tobytes() == a[2].tobytes() and m[3].tobytes() == a[3].tobytes(), mode
    same = a[2].tobytes() == b[2].tobytes() and a[3].tobytes() == b[3].tobytes() and a[1] == b[1]
    sa, sb = a[4], b[4]
    print("%-28s off %.3f ms (%d jac + %d cost sweeps) | on %.3f ms (%d jac + %d cost sweeps, %d speculated, %d hits, %d creep-fused) | it %d nfev %d stop %d | identical=%s"
          % (name, a[0], sa["jac_passes"], sa["cost_passes"], b[0], sb["jac_passes"], sb["cost_passes"], sb["spec_jac_issued"],
             sb["spec_jac_hits"], sb["creep_fused"], b[3][5], b[3][7], b[3][6], same))
    if not same:
        print("   off:", a[2], a[3]); print("   on: ", b[2], b[3])
    s.free()

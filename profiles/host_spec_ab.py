"""A/B of the speculative Jacobians in the kernel-per-evaluation driver (DRIVE_HOST, used beyond 4e7 samples)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from brdf_b200 import api as A
ctx = A.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
for n in (10_000_000, 100_000_000):
    s = ctx.synth(n, 88172645463325252)
    for mode in ("0", "1"):
        os.environ["BRDFGPU_SPEC_JAC"] = mode
        ctx.fit_global(s, A.REF_GLOBAL, drive=A.DRIVE_HOST)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ret, p, info = ctx.fit_global(s, A.REF_GLOBAL, drive=A.DRIVE_HOST)
        e1.record(stream); ctx.synchronize()
        st = ctx.fit_stats()
        print("n=%d spec=%s: %.2f ms, %d K2 + %d K3 passes, %d speculated, %d hits, it %d nfev %d stop %d p=%s cost=%.12g"
              % (n, mode, e0.elapsed_time(e1), st["jac_passes"], st["cost_passes"], st["spec_jac_issued"], st["spec_jac_hits"],
                 info[5], info[7], info[6], p, info[1]), flush=True)
    s.free()

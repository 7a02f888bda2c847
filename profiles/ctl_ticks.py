"""Control-code cycles of the persistent fit by segment (debug variant: make -C brdf_b200/csrc VARIANT=_ticks
EXTRA=-DBG_CTL_TICKS; BRDFGPU_LIB=brdf_b200/libbrdfgpu_ticks.so python profiles/ctl_ticks.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from brdf_b200 import api as A
ctx = A.Context(0)
for n in (200_000, 1_000_000):
    s = ctx.synth(n, 88172645463325252)
    ctx.fit_global(s, A.REF_GLOBAL)
    print("n =", n, flush=True)
    ret, p, info = ctx.fit_global(s, A.REF_GLOBAL)
    ctx.synchronize()
    st = ctx.fit_stats()
    print("   sweeps", st["jac_passes"] + st["cost_passes"], "cyc total", st["cyc_total"], "sweep", st["cyc_sweep"], "exchange", st["cyc_exchange"], flush=True)
    s.free()

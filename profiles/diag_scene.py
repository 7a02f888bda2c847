import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
if os.environ.get("TORCH_FIRST"):
    import torch
    torch.cuda.set_device(0); torch.zeros(8, device="cuda"); torch.cuda.synchronize()
import real_scenes as R
from brdf_b200 import api as A
ctx=A.Context(0)
sc=R.load("cup")
def T(name, fn):
    ctx.synchronize(); t0=time.perf_counter(); out=fn(); ctx.synchronize(); print("%-40s %.2f ms"%(name,(time.perf_counter()-t0)*1e3), flush=True); return out
scene=T("scene", lambda: ctx.scene(sc["V"], sc["F"], sc["imgs"], dark=sc["dark"]))
for rep in range(3):
    T("gather host", lambda: scene.gather(sc["cams"][:1]))
    T("gather resident", lambda: scene.gather_resident(sc["cams"][:1], want_global=True, want_batch=True))
    T("calc_brdf_equation", lambda: scene.calc_brdf_equation(sc["cams"][0]))
    T("calc_single", lambda: scene.calc_brdf_equation_single(sc["cams"][0]))
    T("pixel2surface", lambda: scene.calc_pixel2surface(sc["cams"][0]))

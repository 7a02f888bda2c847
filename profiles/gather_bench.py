"""Device time of one resident gather call (projection + compaction + sample kernel) on the reference's scenes.
    python profiles/gather_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import real_scenes as R  # noqa: E402
from brdf_b200 import api as A  # noqa: E402

ctx = A.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
for name in ("cup", "bunny"):
    sc = R.load(name)
    cams = sc["cams"][:1] if name == "cup" else sc["cams"]
    scene = ctx.scene(sc["V"], sc["F"], sc["imgs"], dark=sc["dark"])

    def call():
        s, _, nfit = scene.gather_resident(cams, model=A.BLINN_PHONG, channel=0, want_global=True, want_batch=False)
        s.free()
        return nfit
    for _ in range(3):
        nfit = call()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        call()
    e1.record(stream)
    ctx.synchronize()
    print("%s: %d views, %d fits, %.1f us per resident gather call (device)" % (name, len(cams), nfit, 1e3 * e0.elapsed_time(e1) / 20))
    scene.free()
ctx.close()

"""Timings on the reference's photographed scenes (tests/_scenes, BASELINE configs[0] and [2]):
gather (K1), the per-face fits of CalcBRDFEquation and the per-channel global fits, one B200.
    python profiles/scene_bench.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import real_scenes as R  # noqa: E402
from brdf_b200 import api as A  # noqa: E402

ctx = A.Context(0)


def wall(fn, reps=3):
    fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


print("| scene | step | ms (wall, host arrays in and out) | work |")
print("|---|---|---|---|")
for name in ("cup", "bunny"):
    sc = R.load(name)
    if sc is None:
        print("| %s | (tests/_scenes/%s.npz absent) | | |" % (name, name))
        continue
    ms, scene = wall(lambda: ctx.scene(sc["V"], sc["F"], sc["imgs"], dark=sc["dark"]), 1)
    print("| %s | scene upload + ambient subtraction + face normals | %.2f | %d faces, 16 photos 800x600 |" % (name, ms, sc["F"].shape[0]))
    cams = sc["cams"][:1] if name == "cup" else sc["cams"]
    ms, g = wall(lambda: scene.gather(cams))
    print("| %s | gather, %d view(s) (pixel map + cosines + intensities to the host) | %.2f | %d fits, %d samples per channel |" %
          (name, len(cams), ms, g["nfit"], g["phi"].size))
    def resident():
        s, b, nfit = scene.gather_resident(cams, model=A.BLINN_PHONG, channel=0, want_global=True, want_batch=True)
        s.free(); b.free()      # (handles left to the garbage collector pile up device memory and skew later timings)
        return nfit
    ms, nfit = wall(resident)
    print("| %s | gather, results resident on the device | %.2f | |" % (name, ms))
    ms, (nf, surf) = wall(lambda: scene.calc_brdf_equation(sc["cams"][0]))
    print("| %s | CalcBRDFEquation: gather + %d per-face fits (3 channels) | %.2f | %.3g fits/s |" % (name, 3 * nf, ms, 3 * nf / ms * 1e3))
    ms, out = wall(lambda: scene.calc_brdf_equation_single(sc["cams"][0]))
    print("| %s | CalcBRDFEquation_SingleBRDF: gather + 3 global fits | %.2f | stop reasons %s |" % (name, ms, [int(v[6]) for v in out[2]]))
    scene.free()
ctx.close()

"""Small driver for ncu captures: a few launches of each hot kernel at a chosen size.
    python profiles/prof_kernels.py k2 100000000      K2 fused residual+Jacobian+JtJ pass (+K3 cost pass)
    python profiles/prof_kernels.py fit 1000000       the persistent global fit
    python profiles/prof_kernels.py batch 65536 64 [exact]   the batched per-face fits (fast kernel / levmar-exact kernel)
    python profiles/prof_kernels.py gather            the gather on a synthetic scene
    python profiles/prof_kernels.py gather_scene      the gather on the reference's bunny scene (13 views), tests/_scenes
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from brdf_b200 import api as A  # noqa: E402

what = sys.argv[1]
ctx = A.Context(0)
if what == "k2":
    n = int(sys.argv[2])
    s = ctx.synth(n, 88172645463325252)
    p = [0.6, 0.35, 12.0]
    ctx.repeat(s, p, 1.0, 0, 4)
    ctx.repeat(s, p, 1.0, 1, 4)
    ctx.synchronize()
elif what == "fit":
    n = int(sys.argv[2])
    s = ctx.synth(n, 88172645463325252)
    for _ in range(3):
        r = ctx.fit_global(s, A.REF_GLOBAL)
    print(r)
elif what == "batch":
    nfit, nper = int(sys.argv[2]), int(sys.argv[3])
    mode = A.JAC_FD_EXACT if len(sys.argv) > 4 and sys.argv[4] == "exact" else A.JAC_FD
    b = ctx.batch_synth(nfit, nper, seed=2026)
    for _ in range(2):
        b.fit(A.REF_PERFACE, jac_mode=mode)
    ctx.synchronize()
elif what == "micro":
    # per-launch device times of K2 / K3 at several sizes + the two global-fit drivers
    import torch
    stream = torch.cuda.ExternalStream(ctx.stream)

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(reps); e1.record(stream); ctx.synchronize()
        return e0.elapsed_time(e1) / reps
    p = [0.6, 0.35, 12.0]
    for n in (10**4, 10**5, 10**6, 10**7, 10**8):
        s = ctx.synth(n, 88172645463325252)
        out = []
        for kind in (0, 1):
            ctx.repeat(s, p, 1.0, kind, 5); ctx.synchronize()
            us = 1e3 * timed(lambda r: ctx.repeat(s, p, 1.0, kind, r), 50 if n < 10**8 else 10)
            out.append("%s %.2f us %.0f GB/s" % ("K2" if kind == 0 else "K3", us, 24.0 * n / us / 1e3))
        line = "n=%d  %s" % (n, "  ".join(out))
        if n <= 10**8:
            for drive, name in ((A.DRIVE_HOST, "host"), (A.DRIVE_PERSISTENT, "persistent")):
                ctx.fit_global(s, A.REF_GLOBAL, drive=drive)
                import time
                t0 = time.perf_counter(); r = ctx.fit_global(s, A.REF_GLOBAL, drive=drive); dt = time.perf_counter() - t0
                st = ctx.fit_stats()
                passes = st["jac_passes"] + st["cost_passes"]
                line += "  | %s fit %.2f ms, %d passes, %.2f us/pass" % (name, dt * 1e3, passes, dt * 1e6 / passes)
                if st["cyc_total"]:
                    line += " [kernel Mcyc: sweep %.2f exchange %.2f total %.2f, points %d]" % (
                        st["cyc_sweep"] / 1e6, st["cyc_exchange"] / 1e6, st["cyc_total"] / 1e6, st["cost_points"])
                    line += " x-phases/pass %s" % [int(v / passes) for v in st["cyc_exchange_phases"]]
                    line += " ctl Mcyc %s jac=%d cost+many=%d" % ({k: round(v / 1e6, 2) for k, v in st["cyc_control_by_next_sweep"].items() if v},
                                                                 st["jac_passes"], st["cost_passes"])
        print(line, flush=True)
        del s
elif what == "gather":
    import scene_lib as S
    V, F = S.height_field(200, 150, seed=3)
    imgs, dark = S.random_images(16, 800, 600, seed=4)
    cams = [S.look_at_camera((60.0, 40.0, 260.0), (0.0, 0.0, 0.0)), S.look_at_camera((-90.0, 10.0, 230.0), (5.0, -5.0, 0.0))]
    sc = ctx.scene(V, F, imgs, dark=dark)
    for _ in range(2):
        g = sc.gather(cams)
    print(g["nfit"])
elif what == "gather_scene":
    import real_scenes as R
    sc = R.load("bunny")
    scene = ctx.scene(sc["V"], sc["F"], sc["imgs"], dark=sc["dark"])
    for _ in range(2):
        s, b, nfit = scene.gather_resident(sc["cams"], want_global=True, want_batch=False)
        s.free()
    print(nfit)
ctx.close()

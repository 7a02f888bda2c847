"""Run a fixed set of global fits and dump (ret, p, info, sweeps) -- used to compare library variants bit for bit.
    BRDFGPU_LIB=... python profiles/dump_fits.py out.npy"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from brdf_b200 import api as A
ctx = A.Context(0)
rows = []
for n, seed in ((10**4, 1), (10**5, 2), (10**6, 88172645463325252), (3 * 10**6, 4), (200001, 77)):
    s = ctx.synth(n, seed)
    for preset in (A.REF_GLOBAL, A.REF_PERFACE):
        ret, p, info = ctx.fit_global(s, preset)
        st = ctx.fit_stats()
        rows.append(np.concatenate([[ret], p, info, [st["jac_passes"], st["cost_passes"], st["cost_points"]]]))
        print(n, ret, p, info[5:8], st["cost_passes"], st["cost_points"], "control Mcyc", {k: round(v / 1e6, 2) for k, v in st["cyc_control_by_next_sweep"].items() if v}, "total", round(st["cyc_total"] / 1e6, 2), flush=True)
    s.free()
np.save(sys.argv[1], np.array(rows))

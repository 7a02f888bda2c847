"""Run fixed batches and dump (p, info, ret) -- to compare library variants bit for bit.
    BRDFGPU_LIB=... python profiles/dump_batch.py out.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from brdf_b200 import api as A
ctx = A.Context(0)
out = {}
for nfit, nper in ((4096, 64), (4096, 16), (2048, 40)):
    b = ctx.batch_synth(nfit, nper, seed=31)
    b.fit(A.REF_PERFACE)
    p, info, ret = b.results()
    out["p_%d" % nper] = p; out["info_%d" % nper] = info; out["ret_%d" % nper] = ret
np.savez(sys.argv[1], **out)

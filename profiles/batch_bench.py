import sys, os, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from brdf_b200 import api as A
ctx=A.Context(0)
stream=torch.cuda.ExternalStream(ctx.stream)
for nfit,nper in ((65536,64),(65536,16),(262144,64)):
    b=ctx.batch_synth(nfit,nper,seed=2026)
    for _ in range(2): b.fit(A.REF_PERFACE)
    ctx.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3): b.fit(A.REF_PERFACE)
    e1.record(stream); ctx.synchronize()
    ms=e0.elapsed_time(e1)/3
    pp,info,ret=b.results()
    print(os.environ.get("BRDFGPU_LIB","base"), "G=%s" % os.environ.get("BRDFGPU_BATCH_G","auto"), nfit,nper,"%.2f ms  %.3g fits/s  mean iters %.1f nfev %.1f conv %.3f"%(ms,nfit/ms*1e3,info[:,5].mean(),info[:,7].mean(),np.isin(info[:,6].astype(int),(1,2,6)).mean()),flush=True)
    b.free()

"""brdfgpu_dlevmar_dif (secant LM, Jacobian resident in HBM) at 10^7 and 10^8 samples: ms and GB/s of the bytes its passes move.
    python profiles/secant_bench.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


class Args:
    gpus, steps, warmup, quick, no_cpu = 1, 1, 1, True, True


rig = bench.Rig(Args())
for r in bench.secant_leg(rig):
    print(json.dumps({k: v for k, v in r.items() if k != "p"}))
rig.ctx.close()

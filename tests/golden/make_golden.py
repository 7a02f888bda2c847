"""Generates tests/golden/brdf_fits.json with the REFERENCE's own levmar (oracle/_ref/liblevmar_ref.so,
compiled unmodified from /root/reference/levmar by oracle/Makefile) driving the BRDFFunc callback
(brdfdata.cpp:969-989 as restated in oracle/brdf_oracle.c -- brdfdata.cpp itself needs OpenCV/Eigen/
libigl/GL and cannot be built here).  Run in the development container only:

    python tests/golden/make_golden.py

Inputs are regenerated from (seed, n, ...) by tests/synth.py, so only outputs are stored (as C99 hex
floats: bit-exact)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import synth  # noqa: E402

GLOBAL_CASES = [
    # name, n, model, preset, seed, negate_every (0 = all cosines in [0,1))
    ("global_ref_2k_blinn", 2000, 1, "REF_GLOBAL", synth.DEFAULT_SEED, 0),
    ("global_ref_20k_blinn", 20000, 1, "REF_GLOBAL", 12345, 0),
    ("global_perface_opts_5k_blinn", 5000, 1, "REF_PERFACE", 777, 0),
    ("global_ref_3k_phong", 3000, 0, "REF_GLOBAL", 4242, 0),
    ("global_perface_opts_3k_phong", 3000, 0, "REF_PERFACE", 4243, 0),
    ("global_ref_1k_negcos", 1000, 1, "REF_GLOBAL", 99, 17),       # NaN -> ret -1, info[6]=7
    ("global_perface_1k_negcos", 1000, 1, "REF_PERFACE", 99, 17),
]
BATCH_CASES = [
    # name, nfit, nper, model, seed
    ("batch_32x16_blinn", 32, 16, 1, 2024),
    ("batch_32x64_blinn", 32, 64, 1, 2025),
    ("batch_16x16_phong", 16, 16, 0, 2026),
]


def hexes(a):
    return [float(v).hex() for v in np.asarray(a).ravel()]


def global_inputs(n, model, seed, negate_every):
    c, td, th, x = synth.samples(n, model_id=model, seed=seed)
    if negate_every:
        td = td.copy(); th = th.copy()
        td[::negate_every] *= -1.0
        th[::negate_every] *= -1.0
    return c, td, th, x


def main():
    ref = O.ref()
    assert ref is not None, "oracle/_ref/liblevmar_ref.so missing: run `make -C oracle ref` where /root/reference exists"
    out = {"generator": "tests/golden/make_golden.py", "solver": "reference levmar 2.6 (oracle/_ref)",
           "global": [], "batch": []}
    for name, n, model, preset, seed, neg in GLOBAL_CASES:
        c, td, th, x = global_inputs(n, model, seed, neg)
        ret, p, info = O.brdf_fit(ref, "", c, td, th, x, model, getattr(O, preset))
        out["global"].append(dict(name=name, n=n, model=model, preset=preset, seed=seed, negate_every=neg,
                                  ret=int(ret), p=hexes(p), info=hexes(info)))
        print(name, ret, p, info[1], info[5:10])
    for name, nfit, nper, model, seed in BATCH_CASES:
        c, td, th, x, truth = synth.batched(nfit, nper, model_id=model, seed=seed)
        rets, ps, infos = [], [], []
        for f in range(nfit):
            ret, p, info = O.brdf_fit(ref, "", c[f], td[f], th[f], x[f], model, O.REF_PERFACE)
            rets.append(int(ret)); ps += hexes(p); infos += hexes(info)
        out["batch"].append(dict(name=name, nfit=nfit, nper=nper, model=model, seed=seed, ret=rets, p=ps, info=infos))
        print(name, rets)
    with open(os.path.join(HERE, "brdf_fits.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()

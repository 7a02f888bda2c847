"""Reference results for EVERY fit of the two batched workloads (BASELINE.json configs[3] and the per-face
path of configs[0]), so the GPU tests can report the true agreement rate instead of sampling:

    batched_full_cfg3.npz   65 536 fits x 64 samples, tests/synth.py batched(seed=2026), REF_PERFACE preset
    perface_full_cup.npz    every mapped face of img/cup x 3 colour channels (CalcBRDFEquation,
                            brdfdata.cpp:1188-1227) from the oracle's gather of tests/_scenes/cup.npz

computed by the reference's own levmar (oracle/_ref, compiled unmodified) through the oracle's BRDFFunc, one
dlevmar_bc_dif call per fit as brdfdata.cpp:1119 makes them.  Stored compactly: p (float32: the gate is 1e-4
relative), final cost info[1] (float64: the gate is 1e-6), stop reason, iterations, nfev, return value.
    python tests/real_scenes.py && python tests/golden/make_batched_full.py        (about 2 minutes on 8 cores)
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O      # noqa: E402
import synth                # noqa: E402

CFG3 = dict(nfit=65536, nper=64, seed=2026)


def _fit_range(args):
    c, td, x, lo, hi = args
    lib, prefix = (O.ref(), "") if O.ref() is not None else (O.oracle(), "oracle_")
    n = hi - lo
    p, cost = np.zeros((n, 3), np.float32), np.zeros(n)
    reason, iters, nfev, ret = np.zeros(n, np.int8), np.zeros(n, np.int16), np.zeros(n, np.int32), np.zeros(n, np.int16)
    for k in range(n):
        r, pp, info = O.brdf_fit(lib, prefix, c[lo + k], td[lo + k], None, x[lo + k], 1, O.REF_PERFACE)
        p[k], cost[k], reason[k], iters[k], nfev[k], ret[k] = pp, info[1], int(info[6]), int(info[5]), int(info[7]), r
    return p, cost, reason, iters, nfev, ret


def fit_all(c, td, x, workers=None):
    nfit = c.shape[0]
    workers = workers or os.cpu_count()
    step = (nfit + 8 * workers - 1) // (8 * workers)
    jobs = [(c, td, x, lo, min(nfit, lo + step)) for lo in range(0, nfit, step)]
    with mp.Pool(workers) as pool:
        parts = pool.map(_fit_range, jobs)
    return [np.concatenate([q[i] for q in parts]) for i in range(6)]


def save(path, parts, **meta):
    p, cost, reason, iters, nfev, ret = parts
    np.savez_compressed(path, p=p, cost=cost, reason=reason, iters=iters, nfev=nfev, ret=ret,
                        solver=np.array("reference levmar (oracle/_ref)" if O.ref() is not None else "oracle port"), **meta)
    u, n = np.unique(reason, return_counts=True)
    print(os.path.basename(path), p.shape[0], "fits; stop reasons", dict(zip(u.tolist(), n.tolist())), "%.1f KB" % (os.path.getsize(path) / 1e3))


def main():
    which = sys.argv[1:] or ["cfg3", "cup"]
    if "cfg3" in which:
        c, td, th, x, _ = synth.batched(CFG3["nfit"], CFG3["nper"], seed=CFG3["seed"])
        save(os.path.join(HERE, "batched_full_cfg3.npz"), fit_all(c, td, x), nper=np.array(CFG3["nper"]), seed=np.array(CFG3["seed"]))
    if "cup" in which:
        import real_scenes as R
        import scene_lib as S
        sc = R.load("cup")
        H, W = sc["imgs"][0].shape[:2]
        clean = []
        for im in sc["imgs"]:
            w = im.copy()
            O.oracle().oracle_subtract_ambient(w.ctypes.data, sc["dark"].ctypes.data, w.size)
            clean.append(w)
        g = S.oracle_gather(sc["V"], sc["F"], sc["cams"][0], S.led_table(), clean, W, H)
        nfit = g["nfit"]
        # fit ch * nfit + f = face fit_face[f], channel ch: the layout of brdfgpu_calc_brdf_equation's batch
        c = np.concatenate([g["phi"]] * 3)
        td = np.concatenate([g["thetaDash"]] * 3)
        x = np.concatenate([g["I"][ch] for ch in range(3)])
        save(os.path.join(HERE, "perface_full_cup.npz"), fit_all(np.ascontiguousarray(c), np.ascontiguousarray(td), np.ascontiguousarray(x)),
             nfit=np.array(nfit), fit_face=g["fit_face"].astype(np.int32))


if __name__ == "__main__":
    main()

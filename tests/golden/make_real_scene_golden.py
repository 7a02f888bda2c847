"""Golden vectors for the reference's photographed scenes (BASELINE.json configs[0] cup, configs[2]
bunny x all Camera Calibrations): what the CPU oracle (oracle/gather_oracle.c + the reference's own
levmar in oracle/_ref when built) produces on tests/_scenes/*.npz.  Stored: per-view fit counts,
SHA-256 of the pixel maps / cosines / intensities, the three per-channel global fits, and a spread
of per-face fits.  Run here (needs /root/reference for the cache):
    python tests/real_scenes.py && python tests/golden/make_real_scene_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O      # noqa: E402
import real_scenes as R     # noqa: E402
import scene_lib as S       # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def clean_images(sc):
    out = []
    for im in sc["imgs"]:
        w = im.copy()
        O.oracle().oracle_subtract_ambient(w.ctypes.data, sc["dark"].ctypes.data, w.size)   # brdfdata.cpp:130-147
        out.append(w)
    return out


def fit_record(w):
    ret, p, info = w
    return dict(ret=int(ret), p=[float(v) for v in p], info=[float(v) for v in info])


def main():
    lib, prefix = (O.ref(), "") if O.ref() is not None else (O.oracle(), "oracle_")
    led = S.led_table()
    gold = {"solver": "reference levmar (oracle/_ref)" if prefix == "" else "oracle port"}
    for name in ("cup", "bunny"):
        sc = R.load(name)
        H, W = sc["imgs"][0].shape[:2]
        clean = clean_images(sc)
        views = []
        allphi, alltd, allth, allI = [], [], [], []
        ncam = 1 if name == "cup" else len(sc["cams"])
        for v in range(ncam):
            g = S.oracle_gather(sc["V"], sc["F"], sc["cams"][v], led, clean, W, H)
            views.append(dict(cal=sc["cam_names"][v], nfit=int(g["nfit"]), map=sha(g["map"]), fit_face=sha(g["fit_face"]),
                              phi=sha(g["phi"]), thetaDash=sha(g["thetaDash"]), theta=sha(g["theta"]), I=sha(g["I"])))
            allphi.append(g["phi"]); alltd.append(g["thetaDash"]); allth.append(g["theta"]); allI.append(g["I"])
            if v == 0:
                g0 = g
        phi, td, th = (np.concatenate(a).ravel() for a in (allphi, alltd, allth))
        inten = np.concatenate(allI, axis=1)
        rec = dict(views=views, total_fits=int(sum(v["nfit"] for v in views)), samples_per_channel=int(phi.size))
        rec["faces_with_negative_costhetadash"] = int(np.count_nonzero((np.concatenate(alltd) < 0).any(axis=1)))
        rec["global"] = [fit_record(O.brdf_fit(lib, prefix, phi, td, th, inten[ch].ravel(), 1, O.REF_GLOBAL)) for ch in range(3)]
        # per-face fits (first view): every 997th fit, all three channels
        per = []
        for k in range(0, g0["nfit"], 997):
            for ch in range(3):
                w = O.brdf_fit(lib, prefix, g0["phi"][k], g0["thetaDash"][k], g0["theta"][k], g0["I"][ch][k], 1, O.REF_PERFACE)
                r = fit_record(w); r.update(k=int(k), ch=ch)
                per.append(r)
        rec["per_face"] = per
        gold[name] = rec
        print(name, rec["total_fits"], rec["samples_per_channel"], [(g["ret"], g["info"][6], g["info"][5]) for g in rec["global"]])
    with open(os.path.join(HERE, "real_scenes.json"), "w") as f:
        json.dump(gold, f, indent=1)


if __name__ == "__main__":
    main()

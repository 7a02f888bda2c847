"""The reference's photographed scenes as raw arrays (BASELINE.json configs[0] and configs[2]).

`build_cache()` decodes what the reference's loaders read -- img/<scene>/{1..16,dark}.png with the
same libpng decode cv::imread uses (brdfdata.cpp:34-61,117-128), <scene>.obj as igl::readOBJ
(:289-312; `f v/vt v/vt v/vt`, 1-based), <scene>.cal and every file of "Camera Calibrations/" with
the fields CBRDFdata::WriteValue keeps (:195-247) -- into tests/_scenes/<scene>.npz.  That folder is
a DERIVED artefact like oracle/_ref: git-ignored, not gpurun-ignored, so the GPU box gets it while
/root/reference stays behind.  No reference source is copied.  Tests skip when the cache is absent.
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.path.join(HERE, "_scenes")
CAL_FIELDS = ("cx", "cy", "f", "sx", "nx", "ny", "nz", "ox", "oy", "oz", "ax", "ay", "az", "px", "py", "pz")


def parse_cal(path):
    text = open(path, "r", errors="replace").read()
    vals = []
    for k in CAL_FIELDS:
        m = re.search(r"<%s>([^<]*)</%s>" % (k, k), text)
        vals.append(float(m.group(1)))      # atof, brdfdata.cpp:197
    return np.array(vals)


def parse_obj(path):
    V, F = [], []
    for line in open(path, "r", errors="replace"):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            V.append([float(t[1]), float(t[2]), float(t[3])])
        elif t[0] == "f":
            F.append([int(tok.split("/")[0]) - 1 for tok in t[1:4]])
    return np.array(V, dtype=np.float64), np.array(F, dtype=np.int32)


def build_cache(reference="/root/reference", scenes=("cup", "bunny")):
    import cv2
    os.makedirs(CACHE, exist_ok=True)
    cal_dir = os.path.join(reference, "Camera Calibrations")
    extra = sorted(os.listdir(cal_dir))
    for name in scenes:
        out = os.path.join(CACHE, name + ".npz")
        if os.path.exists(out):
            continue
        d = os.path.join(reference, "img", name)
        V, F = parse_obj(os.path.join(d, name + ".obj"))
        imgs = np.stack([cv2.imread(os.path.join(d, "%d.png" % k), cv2.IMREAD_COLOR) for k in range(1, 17)])
        dark = cv2.imread(os.path.join(d, "dark.png"), cv2.IMREAD_COLOR)
        cams = [parse_cal(os.path.join(d, name + ".cal"))] + [parse_cal(os.path.join(cal_dir, f)) for f in extra]
        np.savez_compressed(out, V=V, F=F, imgs=imgs, dark=dark, cams=np.array(cams),
                            cam_names=np.array([name + ".cal"] + extra))


def load(name):
    path = os.path.join(CACHE, name + ".npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    return dict(V=np.ascontiguousarray(z["V"]), F=np.ascontiguousarray(z["F"]),
                imgs=[np.ascontiguousarray(im) for im in z["imgs"]], dark=np.ascontiguousarray(z["dark"]),
                cams=np.ascontiguousarray(z["cams"]), cam_names=[str(s) for s in z["cam_names"]])


if __name__ == "__main__":
    build_cache()
    for n in ("cup", "bunny"):
        s = load(n)
        print(n, s["V"].shape, s["F"].shape, len(s["imgs"]), s["imgs"][0].shape, s["cams"].shape, s["cam_names"])

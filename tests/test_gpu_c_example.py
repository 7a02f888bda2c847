"""The plain-C host program examples/solve_equation_c.c (the reference's two levmar call sites,
brdfdata.cpp:1058 and :1119, with only the brdfgpu_ prefix added) built with gcc against
include/brdfgpu.h and run on the GPU; its printed results against the CPU oracle on the same data."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "solve_equation_c.c")


def build(tmp_path, src=SRC):
    exe = str(tmp_path / os.path.splitext(os.path.basename(src))[0])
    lib_dir = os.path.join(ROOT, "brdf_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-std=c99", "-I" + os.path.join(ROOT, "include"), src, "-L" + lib_dir,
                           "-lbrdfgpu", "-Wl,-rpath," + lib_dir, "-lm", "-o", exe])
    return exe


PIPELINE = os.path.join(ROOT, "examples", "scene_pipeline_c.c")


def test_scene_pipeline_compiles_as_c99_and_checks_its_arguments(tmp_path):
    """no GPU needed: main.cpp's flow in plain C links, and without arguments it answers as the reference does"""
    exe = build(tmp_path, PIPELINE)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode != 0 and "required command line arguments" in out.stdout


@pytest.mark.gpu
def test_scene_pipeline_matches_python_mirror(tmp_path):
    """examples/scene_pipeline_c.c on a scene written in the reference's file formats (.obj, 1..16.png, dark.png,
    .cal) against the same calls through the ctypes mirror on the in-memory arrays."""
    cv2 = pytest.importorskip("cv2")
    import scene_lib as S
    from brdf_b200 import api as A
    W, H = 320, 240
    V, F = S.height_field(40, 30, seed=3)
    imgs, dark = S.random_images(16, W, H, seed=4)
    cam = S.look_at_camera((30.0, 20.0, 260.0), (0.0, 0.0, 0.0), f=400.0, cx=160.0, cy=120.0)
    g = S.oracle_gather(V, F, cam, S.led_table(), imgs, W, H)
    imgs = S.paint_model_radiance(imgs, g)          # so that the fits are well posed
    dark = np.zeros_like(dark)
    folder = str(tmp_path) + "/"
    with open(folder + "m.obj", "w") as f:
        for v in V:
            f.write("v %r %r %r\n" % tuple(float(x) for x in v))
        for a, b, c in F:
            f.write("f %d %d %d\n" % (a + 1, b + 1, c + 1))
    for k, im in enumerate(imgs):
        assert cv2.imwrite(folder + "%d.png" % (k + 1), im)
    assert cv2.imwrite(folder + "dark.png", dark)
    names = ("cx", "cy", "f", "sx", "nx", "ny", "nz", "ox", "oy", "oz", "ax", "ay", "az", "px", "py", "pz")
    with open(folder + "c.cal", "w") as f:
        for n, v in zip(names, cam):
            f.write("<%s>%r</%s>\n" % (n, float(v), n))
    out = subprocess.run([build(tmp_path, PIPELINE), folder, folder + "m.obj", folder + "c.cal"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    ctx = A.Context()
    sc = ctx.scene(V, F, imgs, dark)
    nfit, surf = sc.calc_brdf_equation(cam)
    surf = np.nan_to_num(surf, nan=0.0)     # the C program starts from zeros where no pixel maps to the face
    n1, single, info, ret = sc.calc_brdf_equation_single(cam)
    eye = cam[13:16]
    bgr = sc.shade_faces(eye, eye + cam[10:13], surf, A.BLINN_PHONG, True)
    text = out.stdout
    assert "per-face: %d faces fitted" % nfit in text
    for ch in range(3):
        m = re.search(r"single\[%d\]: ret=(-?\d+) p=(\S+) (\S+) (\S+) reason=(\d+)" % ch, text)
        assert int(m.group(1)) == int(ret[ch]) and int(m.group(5)) == int(info[ch][6])
        assert [float(m.group(i)) for i in (2, 3, 4)] == [float(v) for v in single[ch]]
    m = re.search(r"shaded: (\d+) finite colour values, sum=(\S+)", text)
    ok = np.isfinite(bgr)
    assert int(m.group(1)) == int(ok.sum())
    np.testing.assert_allclose(float(m.group(2)), bgr[ok].sum(), rtol=1e-9)
    sc.free()
    ctx.close()


def test_c_example_compiles_as_c99(tmp_path):
    """no GPU needed: the header is plain C and every symbol the example uses links"""
    assert os.path.exists(build(tmp_path))


def _splitmix(count, state=88172645463325252):
    k = np.arange(1, count + 1, dtype=np.uint64)
    z = np.uint64(state) + k * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


@pytest.mark.gpu
def test_c_example_matches_oracle(tmp_path):
    nfaces = 1500
    out = subprocess.run([build(tmp_path), str(nfaces)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    n = nfaces * 16
    with np.errstate(over="ignore"):
        u = _splitmix(4 * n).reshape(n, 4)
    c, td, th = u[:, 0].copy(), u[:, 1].copy(), u[:, 2].copy()
    x = np.clip(np.floor(255.0 * (0.6 * c + 0.35 * td ** 12.0 + (u[:, 3] - 0.5) * 0.01)), 0, 255) / 255.0

    def parse(line):
        ret = int(re.search(r"ret=(-?\d+)", line).group(1))
        p = [float(v) for v in re.search(r"p=(\S+) (\S+) (\S+)", line).groups()]
        sumsq = float(re.search(r"sumsq=(\S+)", line).group(1))
        return ret, np.array(p), sumsq

    ret, p, sumsq = parse(lines[0])
    w = O.brdf_fit(O.oracle(), "oracle_", c, td, th, x, 1, O.REF_GLOBAL)
    assert ret >= 0 and w[0] >= 0
    np.testing.assert_allclose(p, w[1], rtol=1e-4)
    np.testing.assert_allclose(sumsq, w[2][1], rtol=1e-6)
    for f in range(4):
        ret, p, sumsq = parse(lines[1 + f])
        sl = slice(16 * f, 16 * f + 16)
        w = O.brdf_fit(O.oracle(), "oracle_", c[sl], td[sl], th[sl], x[sl], 1, O.REF_PERFACE)
        assert (ret >= 0) == (w[0] >= 0)
        if int(w[2][6]) in (1, 2, 6):
            np.testing.assert_allclose(p, w[1], rtol=1e-4, atol=1e-8)
            np.testing.assert_allclose(sumsq, w[2][1], rtol=1e-6, atol=1e-18)
    hx = [float(v) for v in lines[5].split("=")[1].split()]
    np.testing.assert_allclose(hx, 0.6 * c[:4] + 0.35 * td[:4] ** 12.0, rtol=1e-13)

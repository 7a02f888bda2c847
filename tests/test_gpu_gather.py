"""-m gpu parity tests of the sample gather (K1) against the CPU oracle: bit-exact pixel indices,
face ids, cosines and intensities (brdfdata.cpp:629-681, 799-960, 314-330, 130-147)."""
import numpy as np
import pytest

import oracle_lib as O
import scene_lib as S
from brdf_b200 import api as A

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = A.Context()
    yield c
    c.close()


@pytest.fixture(scope="module")
def scene_inputs():
    W, H = 800, 600
    V, F = S.height_field(90, 70, seed=3)
    imgs, dark = S.random_images(16, W, H, seed=4)
    cams = [S.look_at_camera((60.0, 40.0, 260.0), (0.0, 0.0, 0.0)),
            S.look_at_camera((-90.0, 10.0, 230.0), (5.0, -5.0, 0.0), f=700.0),
            S.look_at_camera((0.0, -20.0, 150.0), (0.0, 0.0, 0.0), f=900.0),      # partly outside the image
            S.look_at_camera((0.0, 0.0, -200.0), (0.0, 0.0, -400.0))]             # looks away: nothing visible
    return V, F, imgs, dark, np.array(cams), W, H


def test_led_table_matches_reference():
    led = np.zeros((16, 3))
    A.lib().brdfgpu_led_table(A._d(led))
    assert np.array_equal(led, S.led_table())
    assert led[0].tolist() == [303.5, -2.3, 555.3] and led[15][0] == 303.5


def test_face_normals_and_ambient_bit_exact(ctx, scene_inputs):
    V, F, imgs, dark, cams, W, H = scene_inputs
    sc = ctx.scene(V, F, imgs, dark=dark)
    fn = np.zeros((F.shape[0], 3))
    O.oracle().oracle_face_normals(O.as_d(V), O.as_i(F), F.shape[0], O.as_d(fn))
    assert np.array_equal(sc.face_normals(), fn)
    for k in (0, 7, 15):
        want = imgs[k].copy()
        O.oracle().oracle_subtract_ambient(want.ctypes.data, dark.ctypes.data, want.size)
        assert np.array_equal(sc.image(k), want)


def test_pixel_map_bit_exact(ctx, scene_inputs):
    V, F, imgs, dark, cams, W, H = scene_inputs
    sc = ctx.scene(V, F, imgs)
    for cam in cams:
        want = np.empty((H, W), dtype=np.int32)
        hits = O.oracle().oracle_calc_pixel2surface(O.as_d(V), O.as_i(F), F.shape[0], O.as_d(np.ascontiguousarray(cam)), W, H,
                                                    O.as_i(want))
        got = sc.calc_pixel2surface(cam)
        assert np.array_equal(got, want)
        assert (hits == 0) == np.all(got == -1)


def test_gather_bit_exact_multi_view(ctx, scene_inputs):
    V, F, imgs, dark, cams, W, H = scene_inputs
    sc = ctx.scene(V, F, imgs, dark=dark)
    clean = []
    for im in imgs:
        w = im.copy()
        O.oracle().oracle_subtract_ambient(w.ctypes.data, dark.ctypes.data, w.size)
        clean.append(w)
    led = S.led_table()
    g = sc.gather(cams)
    first = 0
    total = 0
    for v, cam in enumerate(cams):
        want = S.oracle_gather(V, F, cam, led, clean, W, H)
        n = want["nfit"]
        assert g["nfit_cam"][v] == n
        sl = slice(first, first + n)
        assert np.array_equal(g["maps"][v], want["map"])
        assert np.array_equal(g["fit_face"][sl], want["fit_face"])
        assert np.array_equal(g["fit_pixel"][sl], want["fit_pixel"])
        for key in ("phi", "thetaDash", "theta"):
            assert g[key][sl].tobytes() == want[key].tobytes(), key
        assert g["I"][:, sl].tobytes() == np.ascontiguousarray(want["I"]).tobytes()
        first += n
        total += n
    assert g["nfit"] == total and total > 5000
    assert g["nfit_cam"][3] == 0


def test_last_face_wins_collisions(ctx):
    """Many faces per pixel: the survivor is the highest face id (brdfdata.cpp:676-677 walks faces
    in ascending order and overwrites)."""
    W, H = 64, 48
    V, F = S.height_field(120, 100, seed=8, size=60.0)
    imgs, _ = S.random_images(16, W, H, seed=9)
    cam = S.look_at_camera((0.0, 0.0, 300.0), (0.0, 0.0, 0.0), f=60.0, cx=32.0, cy=24.0)
    sc = ctx.scene(V, F, imgs)
    want = S.oracle_gather(V, F, cam, S.led_table(), imgs, W, H)
    g = sc.gather(cam)
    assert want["nfit"] < F.shape[0] / 4          # heavy collisions
    assert np.array_equal(g["maps"][0], want["map"])
    assert np.array_equal(g["fit_face"], want["fit_face"])
    assert g["phi"].tobytes() == want["phi"].tobytes()


def test_gather_resident_feeds_the_fits(ctx, scene_inputs):
    """Gather -> resident samples -> fits without leaving the device equals host-side plumbing."""
    V, F, imgs, dark, cams, W, H = scene_inputs
    sc = ctx.scene(V, F, imgs, dark=dark)
    g = sc.gather(cams[:1])
    s, b, nfit = sc.gather_resident(cams[:1], model=A.BLINN_PHONG, channel=1, want_global=True, want_batch=True)
    assert nfit == g["nfit"] and len(s) == nfit * 16 and len(b) == nfit
    c, t, x = s.download()
    assert np.array_equal(c, g["phi"].ravel()) and np.array_equal(t, g["thetaDash"].ravel())
    assert np.array_equal(x, g["I"][1].ravel())
    # global fit on the gathered samples equals the oracle's levmar on the same arrays
    want = O.brdf_fit(O.oracle(), "oracle_", g["phi"].ravel(), g["thetaDash"].ravel(), g["theta"].ravel(), g["I"][1].ravel(), 1,
                      O.REF_GLOBAL)
    ret, p, info = ctx.fit_global(s, A.REF_GLOBAL)
    assert (ret >= 0) == (want[0] >= 0) and int(info[6]) == int(want[2][6])
    if want[0] >= 0:
        np.testing.assert_allclose(info[1], want[2][1], rtol=1e-6)
        np.testing.assert_allclose(p, want[1], rtol=1e-4, atol=1e-7)


def test_calc_brdf_equation_drivers(ctx, scene_inputs):
    """CalcBRDFEquation / CalcBRDFEquation_SingleBRDF (brdfdata.cpp:1188-1227, 1138-1186)."""
    V, F, imgs, dark, cams, W, H = scene_inputs
    # photographs that actually show a Blinn-Phong surface under the 16 LEDs
    imgs = S.paint_model_radiance(imgs, S.oracle_gather(V, F, cams[1], S.led_table(), imgs, W, H))
    sc = ctx.scene(V, F, imgs)
    g = sc.gather(cams[1:2])
    nfit, surf = sc.calc_brdf_equation(cams[1])
    assert nfit == g["nfit"]
    touched = ~np.isnan(surf[:, 0, 0])
    assert touched.sum() == nfit and np.array_equal(np.nonzero(touched)[0], np.sort(g["fit_face"]))
    conv = tried = 0
    for k in range(0, nfit, max(1, nfit // 40)):
        face = g["fit_face"][k]
        for ch in range(3):
            w = O.brdf_fit(O.oracle(), "oracle_", g["phi"][k], g["thetaDash"][k], g["theta"][k], g["I"][ch][k], 1, O.REF_PERFACE)
            if int(w[2][6]) in (1, 2, 6):
                conv += np.allclose(surf[face, ch], w[1], rtol=1e-4, atol=1e-7)
    assert conv >= 30
    n2, p, info, ret = sc.calc_brdf_equation_single(cams[1])
    assert n2 == nfit
    for ch in range(3):
        w = O.brdf_fit(O.oracle(), "oracle_", g["phi"].ravel(), g["thetaDash"].ravel(), g["theta"].ravel(), g["I"][ch].ravel(), 1,
                       O.REF_GLOBAL)
        assert (ret[ch] >= 0) == (w[0] >= 0) and int(info[ch][6]) == int(w[2][6])
        if w[0] >= 0:
            np.testing.assert_allclose(info[ch][1], w[2][1], rtol=1e-6)


def test_scene_load_from_files_equals_scene_from_arrays(ctx, scene_inputs, tmp_path):
    """main.cpp:41-59 from the reference's file formats (brdfgpu_scene_load: .obj, 1..16.png, dark.png, .cal)
    gives the very scene the array entry point builds: ambient-subtracted photographs, face normals and the
    pixel map compared as bytes."""
    cv2 = pytest.importorskip("cv2")
    V, F, imgs, dark, cams, W, H = scene_inputs
    folder = str(tmp_path) + "/"
    with open(folder + "mesh.obj", "w") as f:
        f.write("# Mesh file written by the test\n")
        for v in V:
            f.write("v %r %r %r\n" % tuple(float(x) for x in v))
        for a, b, c in F:
            f.write("f %d/%d %d/%d %d/%d\n" % (a + 1, a + 1, b + 1, b + 1, c + 1, c + 1))
    for k, im in enumerate(imgs):
        assert cv2.imwrite(folder + "%d.png" % (k + 1), im)
    assert cv2.imwrite(folder + "dark.png", dark)
    names = ("cx", "cy", "f", "sx", "nx", "ny", "nz", "ox", "oy", "oz", "ax", "ay", "az", "px", "py", "pz")
    with open(folder + "cam.cal", "w") as f:
        f.write("<camera_model>CameraTsai</camera_model>\n")
        for n, v in zip(names, cams[0]):
            f.write("<%s>%r</%s>\n" % (n, float(v), n))
    sc_f, cam = ctx.scene_load(folder, folder + "mesh.obj", folder + "cam.cal")
    sc_a = ctx.scene(V, F, imgs, dark)
    assert (sc_f.nF, sc_f.nimg, sc_f.W, sc_f.H) == (sc_a.nF, sc_a.nimg, sc_a.W, sc_a.H)
    assert cam.tobytes() == np.asarray(cams[0]).tobytes()
    assert sc_f.face_normals().tobytes() == sc_a.face_normals().tobytes()
    for k in (0, 7, 15):
        assert sc_f.image(k).tobytes() == sc_a.image(k).tobytes()
    assert sc_f.calc_pixel2surface(cam).tobytes() == sc_a.calc_pixel2surface(cams[0]).tobytes()
    # without a dark frame the reference carries on with the raw photographs (brdfdata.cpp:134-138)
    import os
    os.remove(folder + "dark.png")
    sc_n, _ = ctx.scene_load(folder, folder + "mesh.obj")
    assert sc_n.image(3).tobytes() == imgs[3].tobytes()
    for sc in (sc_f, sc_a, sc_n):
        sc.free()


@pytest.mark.parametrize("model", [A.BLINN_PHONG, A.PHONG])
@pytest.mark.parametrize("literal", [True, False])
def test_shade_faces_matches_oracle(ctx, scene_inputs, model, literal):
    """BRDF-shaded preview colours (glutcallbacks.cpp:346-445): everything but pow() is the oracle's single
    roundings, so the colours agree to a few ulp of the power (1e-12 relative), NaNs in the same places."""
    V, F, imgs, dark, cams, W, H = scene_inputs
    sc = ctx.scene(V, F, imgs, dark)
    rng = np.random.default_rng(21)
    eye, center = np.array([60.0, 40.0, 260.0]), np.array([0.0, 0.0, 0.0])
    per_face = np.stack([rng.uniform(0, 1, (F.shape[0], 3)), rng.uniform(0, 1, (F.shape[0], 3)),
                         rng.uniform(0.5, 40, (F.shape[0], 3))], axis=2)
    for brdf in (np.array([[0.55, 0.30, 8.0], [0.45, 0.35, 14.0], [0.35, 0.25, 20.5]]), per_face):
        got = sc.shade_faces(eye, center, brdf, model, literal)
        want = S.oracle_shade_faces(V, F, eye, center, brdf, model, literal)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12, atol=1e-300)
    sc.free()


@pytest.mark.parametrize("flags", [1, 2, 3, 4, 7])
def test_gather_options_bit_exact(ctx, flags):
    """Options beyond the reference (depth test, back-face culling, Tsai kappa1) against the
    oracle's statement of the same rules: pixel maps, owners, cosines and intensities as bytes.  A scene with a
    hidden second sheet, so the depth test has something to decide; flags = 0 afterwards is the reference again."""
    W, H = 320, 240
    V, F = S.height_field(40, 30, seed=3)
    V2, F2 = S.height_field(40, 30, seed=4, z0=-40.0)
    F = np.ascontiguousarray(np.concatenate([F, F2 + V.shape[0]]).astype(np.int32))
    V = np.ascontiguousarray(np.concatenate([V, V2]))
    imgs, dark = S.random_images(16, W, H, seed=5)
    cams = np.array([S.look_at_camera((30.0, 20.0, 260.0), (0.0, 0.0, 0.0), f=400.0, cx=160.0, cy=120.0),
                     S.look_at_camera((-60.0, 10.0, 240.0), (5.0, -5.0, 0.0), f=380.0, cx=150.0, cy=125.0)])
    kappa = np.array([2.5e-6, -1.0e-6])
    sc = ctx.scene(V, F, imgs)
    sc.set_gather_options(flags, kappa if flags & A.GATHER_KAPPA1 else None)
    g = sc.gather(cams)
    off = np.concatenate([[0], np.cumsum(g["nfit_cam"])])
    for v in range(2):
        w = S.oracle_gather_opts(V, F, cams[v], S.led_table(), imgs, W, H, flags, kappa[v])
        assert int(g["nfit_cam"][v]) == w["nfit"]
        assert g["maps"][v].tobytes() == w["map"].tobytes()
        sl = slice(off[v], off[v + 1])
        for k in ("fit_face", "fit_pixel", "phi", "thetaDash", "theta"):
            assert np.ascontiguousarray(g[k][sl]).tobytes() == w[k].tobytes(), (v, k)
        assert np.ascontiguousarray(g["I"][:, sl]).tobytes() == np.ascontiguousarray(w["I"]).tobytes()
        assert sc.calc_pixel2surface(cams[v]).tobytes() == w["map"].tobytes() if not (flags & A.GATHER_KAPPA1) else True
    sc.set_gather_options(0)
    w0 = S.oracle_gather(V, F, cams[0], S.led_table(), imgs, W, H)
    assert sc.calc_pixel2surface(cams[0]).tobytes() == w0["map"].tobytes()
    with pytest.raises(A.BrdfGpuError):
        sc.set_gather_options(A.GATHER_KAPPA1)          # kappa1 values missing
    sc.free()


def test_dot_order_switch_and_partial_outputs(ctx, scene_inputs):
    """The default is Eigen 3.3's order for the final dots of GetCosLN / GetCosNH (oracle/gather_oracle.c);
    GATHER_SEQ_DOT gives the left-to-right order -- each against the oracle in the same mode, as bytes.  The resident
    hand-over (fit-stage arrays written by the gather kernel itself, theta not computed for Blinn-Phong) carries
    the same bits as the plain arrays."""
    V, F, imgs, dark, cams, W, H = scene_inputs
    sc = ctx.scene(V, F, imgs)
    led = S.led_table()
    try:
        for flag, order in ((0, O.DOT_EIGEN33), (A.GATHER_SEQ_DOT, O.DOT_SEQUENTIAL)):
            sc.set_gather_options(flag)
            O.set_dot_order(order)
            g = sc.gather(cams[:2])
            off = np.concatenate([[0], np.cumsum(g["nfit_cam"])])
            for v in range(2):
                want = S.oracle_gather(V, F, cams[v], led, imgs, W, H)
                sl = slice(off[v], off[v + 1])
                for key in ("phi", "thetaDash", "theta"):
                    assert np.ascontiguousarray(g[key][sl]).tobytes() == want[key].tobytes(), (flag, key)
            s, b, nfit = sc.gather_resident(cams[:2], model=A.BLINN_PHONG, channel=1, want_global=True, want_batch=True)
            assert nfit == g["nfit"] and len(s) == nfit * 16 and len(b) == nfit
            c_dev, t_dev, x_dev = s.download()
            assert c_dev.tobytes() == g["phi"].tobytes() and t_dev.tobytes() == g["thetaDash"].tobytes()
            assert x_dev.tobytes() == np.ascontiguousarray(g["I"][1]).tobytes()
            s.free(); b.free()
            s, _, _ = sc.gather_resident(cams[:2], model=A.PHONG, channel=2, want_global=True, want_batch=False)
            c_dev, t_dev, x_dev = s.download()
            assert t_dev.tobytes() == g["theta"].tobytes() and x_dev.tobytes() == np.ascontiguousarray(g["I"][2]).tobytes()
            s.free()
    finally:
        O.set_dot_order(O.DOT_EIGEN33)
        sc.set_gather_options(0)
    sc.free()


def test_literal_gl_projection_bit_exact(ctx):
    """The reference's literal mapping -- gluProject through the GL matrices the caller read back (brdfdata.cpp:662-677),
    bottom-up rows, radiance from image row H-1-y (:955) -- against the oracle's restatement of libGLU: maps, owners,
    cosines and intensities as bytes; then back to the Tsai camera."""
    W, H = 320, 240
    V, F = S.height_field(60, 45, seed=11)
    imgs, dark = S.random_images(16, W, H, seed=12)
    mv, proj, vp = S.gl_matrices_over(V, W, H)
    cam = S.look_at_camera((10.0, -5.0, 400.0), (0.0, 0.0, 0.0), f=400.0, cx=160.0, cy=120.0)
    sc = ctx.scene(V, F, imgs)
    sc.set_gl_projection(mv, proj, vp)
    g = sc.gather(cam)
    want = S.oracle_gather_gl(V, F, cam, mv, proj, vp, S.led_table(), imgs, W, H)
    assert g["nfit"] == want["nfit"] > 2000
    assert g["maps"][0].tobytes() == want["map"].tobytes()
    assert sc.calc_pixel2surface(cam).tobytes() == want["map"].tobytes()
    for k in ("fit_face", "fit_pixel", "phi", "thetaDash", "theta"):
        assert np.ascontiguousarray(g[k]).tobytes() == want[k].tobytes(), k
    assert np.ascontiguousarray(g["I"]).tobytes() == np.ascontiguousarray(want["I"]).tobytes()
    with pytest.raises(A.BrdfGpuError):
        sc.gather(np.stack([cam, cam]))          # one set of GL matrices = one view
    sc.set_gl_projection()
    w0 = S.oracle_gather(V, F, cam, S.led_table(), imgs, W, H)
    assert sc.gather(cam)["maps"][0].tobytes() == w0["map"].tobytes()
    sc.free()

"""The reference's LITERAL pixel<->face mapping (brdfdata.cpp:662-677: gluProject through the GL matrices) in the
oracle: against a numpy restatement of libGLU's gluProject, and the survey's finding that with the matrices the
reference's own Display_ sets up, no face centroid of the shipped scenes lands inside the photographs."""
import numpy as np
import pytest

import oracle_lib as O
import real_scenes as R
import scene_lib as S
from brdf_b200 import api as A


def _glu_project(c, mv, proj, vp):
    M, P = mv.reshape(4, 4).T, proj.reshape(4, 4).T       # column-major storage
    out = M @ np.array([c[0], c[1], c[2], 1.0])
    clip = P @ out
    if clip[3] == 0.0:
        return None
    ndc = clip[:3] / clip[3]
    return (ndc[0] * 0.5 + 0.5) * vp[2] + vp[0], (ndc[1] * 0.5 + 0.5) * vp[3] + vp[1]


def test_oracle_gl_mapping_matches_numpy_gluproject():
    W, H = 160, 120
    V, F = S.height_field(24, 18, seed=5)
    imgs, _ = S.random_images(16, W, H, seed=6)
    mv, proj, vp = S.gl_matrices_over(V, W, H)
    cam = S.look_at_camera((0.0, 0.0, 400.0), (0.0, 0.0, 0.0))
    g = S.oracle_gather_gl(V, F, cam, mv, proj, vp, S.led_table(), imgs, W, H)
    assert g["nfit"] > 300
    want = np.full((H, W), -1, dtype=np.int32)
    for i in range(F.shape[0]):
        c = V[F[i]].sum(axis=0) / 3.0
        w = _glu_project(c, mv, proj, vp)
        if w is not None and w[0] >= 0 and w[1] >= 0 and w[0] < W and w[1] < H:
            want[int(w[1]), int(w[0])] = i
    # (numpy's matrix products may round differently in the last bit: allow a handful of centroids on a pixel border)
    assert np.mean(want != g["map"]) < 1e-3
    # intensities come from image row H-1-y (brdfdata.cpp:955)
    k = 7
    row, col = divmod(int(g["fit_pixel"][k]), W)
    for ch in range(3):
        assert g["I"][ch][k][3] == imgs[3][H - 1 - row, col, ch] / 255.0


def test_reference_matrices_equal_between_library_and_oracle():
    mv_o, pr_o = S.oracle_reference_gl_matrices(388.3, 266.8)
    mv, pr = A.reference_gl_matrices(388.3, 266.8)
    assert np.array_equal(mv, mv_o) and np.array_equal(pr, pr_o)
    assert mv[14] == -50.0 and pr[11] == -1.0


@pytest.mark.parametrize("name", ["cup", "bunny"])
def test_literal_mapping_misses_the_photographs(name):
    """with the reference's own matrices (constant fields of view, camera at (0,0,50), window 1920 x 1080)
    not one face centroid of the shipped meshes falls inside the 800 x 600 photographs."""
    sc = R.load(name)
    if sc is None:
        pytest.skip("tests/_scenes absent")
    cam = sc["cams"][0]
    mv, pr = S.oracle_reference_gl_matrices(cam[0], cam[1])
    m = np.empty((600, 800), dtype=np.int32)
    lib = O.oracle()
    vp = np.array([0, 0, 1920, 1080], dtype=np.int32)
    hits = lib.oracle_calc_pixel2surface_gl(O.as_d(sc["V"]), O.as_i(sc["F"]), sc["F"].shape[0], O.as_d(mv), O.as_d(pr), O.as_i(vp), 800, 600, O.as_i(m))
    assert hits == 0 and np.all(m == -1)

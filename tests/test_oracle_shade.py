"""The oracle's restatement of the BRDF-shaded preview (glutcallbacks.cpp:346-445) against an
independent numpy spelling of the same formulas, including the reference's quirks: cosLN indexes the
normal with the truncated dot product (:385) and Phong's cosine is rounded to float (:420)."""
import numpy as np
import pytest

import scene_lib as S


def _numpy_shade(V, F, eye, center, brdf, model, literal):
    tri = V[F]
    c = (tri[:, 0] + tri[:, 1] + tri[:, 2]) / 3.0
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    l = eye - c
    l /= np.linalg.norm(l, axis=1, keepdims=True)
    v = (eye - center) / np.linalg.norm(eye - center)
    nl = np.einsum("ij,ij->i", n, l)
    cos_ln = n[:, 0] if literal else nl
    if model == 1:
        h = l + v
        h /= np.linalg.norm(h, axis=1, keepdims=True)
        t = np.einsum("ij,ij->i", n, h)
    else:
        r = l - 2.0 * (-nl)[:, None] * n
        t = (r @ v).astype(np.float32).astype(np.float64)
    brdf = np.broadcast_to(np.asarray(brdf, dtype=np.float64).reshape(-1, 3, 3), (F.shape[0], 3, 3))
    kd, ks, nn = brdf[:, :, 0], brdf[:, :, 1], brdf[:, :, 2]
    with np.errstate(invalid="ignore"):
        pw = np.power(t[:, None], nn)
    coef = 1.0 if model == 1 else (nn + 2.0) / (2.0 * np.pi)
    return kd * cos_ln[:, None] + ks * coef * pw


@pytest.mark.parametrize("model", [1, 0])
@pytest.mark.parametrize("literal", [True, False])
@pytest.mark.parametrize("single", [True, False])
def test_oracle_shade_matches_numpy(model, literal, single):
    V, F = S.height_field(30, 20, seed=11)
    rng = np.random.default_rng(12)
    eye, center = np.array([40.0, -30.0, 220.0]), np.array([1.0, 2.0, 0.5])
    brdf = np.array([[0.5, 0.3, 8.0], [0.4, 0.35, 14.0], [0.3, 0.2, 3.0]]) if single else \
        np.stack([rng.uniform(0, 1, (F.shape[0], 3)), rng.uniform(0, 1, (F.shape[0], 3)), rng.integers(1, 30, (F.shape[0], 3))], axis=2)
    got = S.oracle_shade_faces(V, F, eye, center, brdf, model, literal)
    want = _numpy_shade(V, F, eye, center, brdf, model, literal)
    ok = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), ok)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-5 if model == 0 else 1e-11, atol=1e-13)
    assert ok.mean() > 0.9


def test_literal_cosln_is_the_normal_x_component():
    V, F = S.height_field(12, 9, seed=13)
    eye, center = np.array([0.0, 0.0, 300.0]), np.zeros(3)
    brdf = np.array([[1.0, 0.0, 1.0]] * 3)      # colour == cosLN
    import ctypes as C
    import oracle_lib as O
    FN = np.empty((F.shape[0], 3))
    O.oracle().oracle_face_normals(O.as_d(V), O.as_i(F), F.shape[0], O.as_d(FN))
    lit = S.oracle_shade_faces(V, F, eye, center, brdf, 1, True)
    assert lit[:, 0].tobytes() == np.ascontiguousarray(FN[:, 0]).tobytes()
    phys = S.oracle_shade_faces(V, F, eye, center, brdf, 1, False)
    assert np.all(np.abs(phys[:, 0]) > 0.5) and not np.allclose(lit, phys)

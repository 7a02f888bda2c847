"""The oracle's gather options beyond the reference (depth test, back-face culling,
Tsai kappa1) against a brute-force numpy statement of the same rules; flags = 0 must be the reference's
gather exactly."""
import numpy as np
import pytest

import scene_lib as S

W, H = 320, 240


def _scene(seed=3, two_layers=True):
    V, F = S.height_field(40, 30, seed=seed)
    if two_layers:      # a second sheet behind the first: every pixel is contested by a hidden face
        V2, F2 = S.height_field(40, 30, seed=seed + 1, z0=-40.0)
        F = np.concatenate([F, F2 + V.shape[0]]).astype(np.int32)
        V = np.concatenate([V, V2])
    imgs, dark = S.random_images(16, W, H, seed=seed + 2)
    cam = S.look_at_camera((30.0, 20.0, 260.0), (0.0, 0.0, 0.0), f=400.0, cx=160.0, cy=120.0)
    return np.ascontiguousarray(V), np.ascontiguousarray(F), imgs, cam


def _brute(V, F, cam, flags, kappa1):
    tri = V[F]
    c = ((0.0 + tri[:, 0]) + tri[:, 1] + tri[:, 2]) / 3.0
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    cx, cy, f, sx = cam[:4]
    nn, oo, aa, pp = cam[4:7], cam[7:10], cam[10:13], cam[13:16]
    d = c - pp
    xc, yc, zc = d @ nn, d @ oo, d @ aa
    ok = zc > 0
    if flags & S.CULL_BACKFACES:
        ok &= np.einsum("ij,ij->i", n, -d) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        if flags & S.KAPPA1:
            xu, yu = f * xc / zc, f * yc / zc
            xd, yd = xu.copy(), yu.copy()
            for _ in range(5):
                s = 1.0 + kappa1 * (xd * xd + yd * yd)
                xd, yd = xu / s, yu / s
            u, v = cx + sx * xd, cy + yd
        else:
            u, v = cx + sx * f * xc / zc, cy + f * yc / zc
    ok &= (u >= 0) & (v >= 0) & (u < W) & (v < H)
    m = np.full(H * W, -1, dtype=np.int32)
    best = np.full(H * W, np.inf)
    for i in np.nonzero(ok)[0]:
        px = int(v[i]) * W + int(u[i])
        if flags & S.DEPTH_TEST:
            if m[px] >= 0 and best[px] < zc[i]:
                continue
            best[px] = zc[i]
        m[px] = i
    return m.reshape(H, W), int(ok.sum())


def test_flags_zero_is_the_reference_gather():
    V, F, imgs, cam = _scene()
    a = S.oracle_gather(V, F, cam, S.led_table(), imgs, W, H)
    b = S.oracle_gather_opts(V, F, cam, S.led_table(), imgs, W, H, 0)
    assert a["nfit"] == b["nfit"] and np.array_equal(a["map"], b["map"])
    for k in ("fit_face", "fit_pixel", "phi", "thetaDash", "theta"):
        assert a[k].tobytes() == b[k].tobytes()
    assert np.ascontiguousarray(a["I"]).tobytes() == np.ascontiguousarray(b["I"]).tobytes()


@pytest.mark.parametrize("flags", [1, 2, 3, 4, 5, 7])
def test_options_match_brute_force(flags):
    V, F, imgs, cam = _scene()
    kappa1 = 2.5e-6      # strong enough to move edge pixels by several columns
    g = S.oracle_gather_opts(V, F, cam, S.led_table(), imgs, W, H, flags, kappa1)
    want, _ = _brute(V, F, cam, flags, kappa1)
    # the brute force uses numpy dot products (different rounding of the last bit): allow a handful of
    # centroids that sit on a pixel boundary to differ, nothing else
    diff = np.count_nonzero(g["map"] != want)
    assert diff <= 4, diff
    owners = np.unique(g["map"][g["map"] >= 0])
    assert np.array_equal(np.sort(g["fit_face"]), owners)
    ref = S.oracle_gather(V, F, cam, S.led_table(), imgs, W, H)
    if flags == S.DEPTH_TEST:
        # the hidden sheet (second half of the faces) comes later in face order and wins every contested pixel in
        # the reference; with the depth test the front sheet owns them: wherever the two maps differ, the owner
        # changed from a back face to a front face
        nfront = F.shape[0] // 2
        changed = g["map"] != ref["map"]
        assert changed.sum() > 100
        assert np.all(ref["map"][changed] >= nfront) and np.all(g["map"][changed] < nfront)
    if flags == S.KAPPA1:
        assert np.count_nonzero(g["map"] != ref["map"]) > 100

"""The evaluation order of the Eigen reductions in the gather oracle (oracle/gather_oracle.c header).

The reference's Makefile pins Eigen 3.3.x built for baseline x86-64 (SSE2, no FMA); Eigen 3.3 sums the final dot
of GetCosLN / GetCosNH (brdfdata.cpp:893, 937: fixed-size vector times a row of a column-major MatrixXd) as
a0*b0 + (a1*b1 + a2*b2) and everything else left to right.  These tests restate that with numpy scalars, check
the oracle against the restatement bit for bit, and pin how far the two candidate orders are apart."""
import numpy as np

import oracle_lib as O
import scene_lib as S


def _scene():
    V, F = S.height_field(24, 18, seed=5)
    imgs, _ = S.random_images(16, 160, 120, seed=6)
    cam = S.look_at_camera((20.0, -15.0, 240.0), (0.0, 0.0, 0.0), f=300.0, cx=80.0, cy=60.0)
    return V, F, imgs, cam


def _normalize(v):
    z = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]      # packet of two, then the third (vectorised squaredNorm)
    return v / np.sqrt(z) if z > 0 else v               # true division per component (Dot.h, 3.3)


def _restated(V, F, cam, led, face, eigen33):
    v = [V[F[face, j]] for j in range(3)]
    C = np.array([(((0.0 + v[0][k]) + v[1][k]) + v[2][k]) / 3.0 for k in range(3)])
    e1, e2 = v[1] - v[0], v[2] - v[0]
    N = _normalize(np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]]))
    P = cam[13:16]
    phi, nh = [], []
    for k in range(16):
        l = _normalize(led[k] - C)
        h = _normalize(led[k] - 2 * C + P)
        for vec, out in ((l, phi), (h, nh)):
            p = vec * N
            out.append(p[0] + (p[1] + p[2]) if eigen33 else (p[0] + p[1]) + p[2])
    return np.array(phi), np.array(nh)


def test_oracle_orders_match_their_restatement():
    V, F, imgs, cam = _scene()
    led = S.led_table()
    try:
        for order, eigen33 in ((O.DOT_EIGEN33, True), (O.DOT_SEQUENTIAL, False)):
            O.set_dot_order(order)
            g = S.oracle_gather(V, F, cam, led, imgs, 160, 120)
            assert g["nfit"] > 100
            for k in range(0, g["nfit"], 17):
                phi, nh = _restated(V, F, cam, led, g["fit_face"][k], eigen33)
                assert phi.tobytes() == g["phi"][k].tobytes()
                assert nh.tobytes() == g["thetaDash"][k].tobytes()
    finally:
        O.set_dot_order(O.DOT_EIGEN33)


def test_orders_differ_only_in_the_last_bit_of_cosln_and_cosnh():
    V, F, imgs, cam = _scene()
    led = S.led_table()
    try:
        O.set_dot_order(O.DOT_EIGEN33)
        a = S.oracle_gather(V, F, cam, led, imgs, 160, 120)
        O.set_dot_order(O.DOT_SEQUENTIAL)
        b = S.oracle_gather(V, F, cam, led, imgs, 160, 120)
    finally:
        O.set_dot_order(O.DOT_EIGEN33)
    # the pixel map, the face list, the literal cos(theta) and the intensities do not depend on the switch
    assert a["map"].tobytes() == b["map"].tobytes() and a["fit_face"].tobytes() == b["fit_face"].tobytes()
    assert a["theta"].tobytes() == b["theta"].tobytes() and a["I"].tobytes() == b["I"].tobytes()
    for key in ("phi", "thetaDash"):
        differ = a[key] != b[key]
        assert 0.02 < differ.mean() < 0.9, (key, differ.mean())    # a real difference, not a no-op switch
        assert np.max(np.abs(a[key] - b[key])) <= 2.3e-16           # one unit in the last place of a cosine

"""Synthetic scenes for the gather tests: a bumpy height-field mesh in front of a Tsai camera and
random 8-bit photographs, plus the oracle-side gather on the same inputs."""
import ctypes as C

import numpy as np

import oracle_lib as O


def look_at_camera(pos, target, f=660.0, sx=1.007, cx=388.3, cy=266.8):
    """Tsai camera {cx, cy, f, sx, n, o, a, p} with a = viewing direction, n x o = a."""
    pos, target = np.asarray(pos, float), np.asarray(target, float)
    a = target - pos
    a /= np.linalg.norm(a)
    up = np.array([0.0, 1.0, 0.0])
    n = np.cross(up, a)
    n /= np.linalg.norm(n)
    o = np.cross(a, n)
    return np.concatenate([[cx, cy, f, sx], n, o, a, pos])


def height_field(nx, ny, seed, size=120.0, z0=0.0):
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.linspace(-size / 2, size / 2, nx), np.linspace(-size / 2, size / 2, ny))
    zs = z0 + 6.0 * np.sin(xs / 17.0) * np.cos(ys / 11.0) + rng.normal(0, 0.4, xs.shape)
    V = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], axis=1)
    faces = []
    for j in range(ny - 1):
        for i in range(nx - 1):
            a = j * nx + i
            faces.append((a, a + 1, a + nx))
            faces.append((a + 1, a + nx + 1, a + nx))
    F = np.array(faces, dtype=np.int32)
    rng.shuffle(F)   # face order is arbitrary in scanned meshes; "last face wins" must still hold
    return np.ascontiguousarray(V), np.ascontiguousarray(F)


def random_images(nimg, W, H, seed):
    rng = np.random.default_rng(seed)
    imgs = [rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8) for _ in range(nimg)]
    dark = rng.integers(0, 40, size=(H, W, 3), dtype=np.uint8)
    return imgs, dark


def oracle_gather(V, F, cam, led, images, W, H):
    lib = O.oracle()
    nF, nimg = F.shape[0], len(images)
    ptrs = (C.c_void_p * nimg)(*[im.ctypes.data for im in images])
    m = np.empty((H, W), dtype=np.int32)
    fit_face, fit_pixel = np.empty(nF, dtype=np.int32), np.empty(nF, dtype=np.int32)
    phi, td, th = (np.empty((nF, nimg)) for _ in range(3))
    inten = np.empty((3, nF, nimg))
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    led = np.ascontiguousarray(led, dtype=np.float64)
    n = lib.oracle_gather(O.as_d(V), O.as_i(F), nF, O.as_d(cam), O.as_d(led), ptrs, nimg, W, H, O.as_i(m),
                          O.as_i(fit_face), O.as_i(fit_pixel), O.as_d(phi), O.as_d(td), O.as_d(th), O.as_d(inten))
    return dict(nfit=n, map=m, fit_face=fit_face[:n], fit_pixel=fit_pixel[:n], phi=phi[:n], thetaDash=td[:n],
                theta=th[:n], I=inten[:, :n])


def led_table():
    led = np.zeros((16, 3))
    O.oracle().oracle_led_table(O.as_d(led))
    return led


def paint_model_radiance(images, gathered, params_bgr=((0.55, 0.30, 8.0), (0.45, 0.35, 14.0), (0.35, 0.25, 20.0)), seed=0):
    """Overwrite the pixels that own a face with 8-bit Blinn-Phong radiance computed from that face's
    gathered cosines (+- 1 level of noise), so that per-face and global fits on the scene are well posed."""
    rng = np.random.default_rng(seed)
    out = [im.copy() for im in images]
    pix = gathered["fit_pixel"]
    W = images[0].shape[1]
    rows, cols = pix // W, pix % W
    for k in range(len(images)):
        for ch, (kd, ks, n) in enumerate(params_bgr):
            val = kd * np.clip(gathered["phi"][:, k], 0, 1) + ks * np.clip(gathered["thetaDash"][:, k], 0, 1) ** n
            q = np.clip(np.floor(255.0 * val + rng.uniform(-1, 1, val.shape)), 0, 255).astype(np.uint8)
            out[k][rows, cols, ch] = q
    return out


def oracle_shade_faces(V, F, eye, center, brdf, model=1, literal=True):
    lib = O.oracle()
    nF = F.shape[0]
    FN = np.empty((nF, 3))
    lib.oracle_face_normals(O.as_d(V), O.as_i(F), nF, O.as_d(FN))
    brdf = np.ascontiguousarray(brdf, dtype=np.float64)
    out = np.empty((nF, 3))
    eye, center = np.ascontiguousarray(eye, dtype=np.float64), np.ascontiguousarray(center, dtype=np.float64)
    lib.oracle_shade_faces.restype = None
    lib.oracle_shade_faces.argtypes = [O.dptr, O.iptr, O.dptr, C.c_int, O.dptr, O.dptr, C.c_int, C.c_int, O.dptr, C.c_int, O.dptr]
    lib.oracle_shade_faces(O.as_d(V), O.as_i(F), O.as_d(FN), nF, O.as_d(eye), O.as_d(center), model, int(brdf.size == 9),
                           O.as_d(brdf), int(bool(literal)), O.as_d(out))
    return out


DEPTH_TEST, CULL_BACKFACES, KAPPA1 = 1, 2, 4


def oracle_gather_opts(V, F, cam, led, images, W, H, flags, kappa1=0.0):
    """oracle_gather with the options beyond the reference (depth test / back-face culling / Tsai kappa1)."""
    lib = O.oracle()
    nF, nimg = F.shape[0], len(images)
    ptrs = (C.c_void_p * nimg)(*[im.ctypes.data for im in images])
    m = np.empty((H, W), dtype=np.int32)
    fit_face, fit_pixel = np.empty(nF, dtype=np.int32), np.empty(nF, dtype=np.int32)
    phi, td, th = (np.empty((nF, nimg)) for _ in range(3))
    inten = np.empty((3, nF, nimg))
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    led = np.ascontiguousarray(led, dtype=np.float64)
    lib.oracle_gather_opts.restype = C.c_int
    lib.oracle_gather_opts.argtypes = [O.dptr, O.iptr, C.c_int, O.dptr, C.c_double, C.c_int, O.dptr, C.POINTER(C.c_void_p), C.c_int,
                                       C.c_int, C.c_int, O.iptr, O.iptr, O.iptr, O.dptr, O.dptr, O.dptr, O.dptr]
    n = lib.oracle_gather_opts(O.as_d(V), O.as_i(F), nF, O.as_d(cam), float(kappa1), int(flags), O.as_d(led), ptrs, nimg, W, H,
                               O.as_i(m), O.as_i(fit_face), O.as_i(fit_pixel), O.as_d(phi), O.as_d(td), O.as_d(th), O.as_d(inten))
    return dict(nfit=n, map=m, fit_face=fit_face[:n], fit_pixel=fit_pixel[:n], phi=phi[:n], thetaDash=td[:n],
                theta=th[:n], I=inten[:, :n])


def gl_matrices_over(V, W, H, margin=1.08):
    """A MODELVIEW / PROJECTION / viewport triple (column-major, as glGetDoublev returns them) that looks down -z onto the
    xy extent of V and fills a W x H viewport: the literal mapping of the reference lands inside the images with it."""
    lo, hi = V.min(axis=0), V.max(axis=0)
    half = 0.5 * margin * max((hi[0] - lo[0]) / W, (hi[1] - lo[1]) / H)
    l, r, b, t = -half * W, half * W, -half * H, half * H
    n, f = 10.0, 2000.0
    dist = 400.0
    # perspective frustum whose near-plane window, scaled to the object distance, covers the mesh
    k = n / dist
    proj = np.zeros(16)
    proj[0] = 2 * n / ((r - l) * k); proj[5] = 2 * n / ((t - b) * k)
    proj[10] = -(f + n) / (f - n); proj[11] = -1.0; proj[14] = -2 * f * n / (f - n)
    mv = np.eye(4)
    mv[0, 3], mv[1, 3], mv[2, 3] = -0.5 * (lo[0] + hi[0]), -0.5 * (lo[1] + hi[1]), -dist - hi[2]
    mv = mv.T.ravel().copy()    # column-major
    f32 = lambda a: a.astype(np.float32).astype(np.float64)    # GL keeps float32
    return f32(mv), f32(proj), np.array([0, 0, W, H], dtype=np.int32)


def oracle_gather_gl(V, F, cam, mv, proj, viewport, led, images, W, H):
    lib = O.oracle()
    nF, nimg = F.shape[0], len(images)
    ptrs = (C.c_void_p * nimg)(*[im.ctypes.data for im in images])
    m = np.empty((H, W), dtype=np.int32)
    fit_face, fit_pixel = np.empty(nF, dtype=np.int32), np.empty(nF, dtype=np.int32)
    phi, td, th = (np.empty((nF, nimg)) for _ in range(3))
    inten = np.empty((3, nF, nimg))
    cam, led = np.ascontiguousarray(cam, dtype=np.float64), np.ascontiguousarray(led, dtype=np.float64)
    mv, proj = np.ascontiguousarray(mv, dtype=np.float64), np.ascontiguousarray(proj, dtype=np.float64)
    vp = np.ascontiguousarray(viewport, dtype=np.int32)
    lib.oracle_gather_gl.restype = C.c_int
    lib.oracle_gather_gl.argtypes = [O.dptr, O.iptr, C.c_int, O.dptr, O.dptr, O.dptr, O.iptr, O.dptr, C.POINTER(C.c_void_p), C.c_int,
                                     C.c_int, C.c_int, O.iptr, O.iptr, O.iptr, O.dptr, O.dptr, O.dptr, O.dptr]
    n = lib.oracle_gather_gl(O.as_d(V), O.as_i(F), nF, O.as_d(cam), O.as_d(mv), O.as_d(proj), O.as_i(vp), O.as_d(led), ptrs, nimg, W, H,
                             O.as_i(m), O.as_i(fit_face), O.as_i(fit_pixel), O.as_d(phi), O.as_d(td), O.as_d(th), O.as_d(inten))
    return dict(nfit=n, map=m, fit_face=fit_face[:n], fit_pixel=fit_pixel[:n], phi=phi[:n], thetaDash=td[:n],
                theta=th[:n], I=inten[:, :n])


def oracle_reference_gl_matrices(cx, cy, win_w=1920, win_h=1080):
    lib = O.oracle()
    mv, pr = np.zeros(16), np.zeros(16)
    lib.oracle_reference_gl_matrices.restype = None
    lib.oracle_reference_gl_matrices.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, O.dptr, O.dptr]
    lib.oracle_reference_gl_matrices(float(cx), float(cy), int(win_w), int(win_h), O.as_d(mv), O.as_d(pr))
    return mv, pr

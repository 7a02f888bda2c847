"""ctypes mirror of include/brdfgpu.h.

Names follow the reference: ``dlevmar_bc_dif`` / ``dlevmar_dif`` (levmar/levmar.h:106-127) called
with the ``BRDFFunc`` callback and an ``extraData`` (brdfdata.cpp:962-989), ``solve_equation`` /
``solve_equation_single`` (brdfdata.cpp:1077-1136, 991-1075), ``calc_pixel2surface``
(brdfdata.cpp:629-681) and so on.  Everything computes on the GPU through libbrdfgpu.so; importing
this module without the built library raises, there is no fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")

PHONG, BLINN_PHONG = 0, 1
DRIVE_HOST, DRIVE_PERSISTENT = 0, 1
JAC_FD, JAC_ANALYTIC, JAC_FD_EXACT = 0, 1, 2
LM_ERROR = -1
IPC_HANDLE_BYTES = 72   # BRDFGPU_IPC_HANDLE_BYTES

# the reference's two option presets (brdfdata.cpp:1002,1046-1058 and :1085,1107-1119)
REF_GLOBAL = dict(p0=(0.0, 0.0, 0.0), itmax=2000, opts=(1e-3, 1e-15, 1e-10, 1e-50, 1.0),
                  lb=(0.0, 0.0, 0.0), ub=(100.0, 100.0, 100.0))
REF_PERFACE = dict(p0=(0.5, 1.0, 1.0), itmax=100, opts=(1e-3, 1e-15, 1e-15, 1e-20, 1e-6),
                   lb=(0.0, 0.0, 0.0), ub=(100.0, 100.0, 100.0))

dptr = C.POINTER(C.c_double)
iptr = C.POINTER(C.c_int)
lptr = C.POINTER(C.c_long)
FUNC_T = C.CFUNCTYPE(None, dptr, dptr, C.c_int, C.c_int, C.c_void_p)


class BrdfGpuError(RuntimeError):
    pass


class ExtraData(C.Structure):
    """struct extraData, brdfdata.cpp:962-966"""
    _fields_ = [("angles", dptr), ("modelInfo", C.c_int)]


def lib_path():
    # BRDFGPU_LIB: a tuning variant built with `make VARIANT=... EXTRA=...` (experiments only)
    return os.environ.get("BRDFGPU_LIB") or os.path.join(HERE, "libbrdfgpu.so")


def build(force=False):
    """Compile libbrdfgpu.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "-s", "clean"])
    subprocess.check_call(["make", "-C", CSRC, "-s", "-j8"])
    return lib_path()


# every symbol include/brdfgpu.h declares: (restype, argtypes)
_V = C.c_void_p
SIGNATURES = {
    "brdfgpu_BRDFFunc": (None, [dptr, dptr, C.c_int, C.c_int, _V]),
    "brdfgpu_BRDFJac": (None, [dptr, dptr, C.c_int, C.c_int, _V]),
    "brdfgpu_dlevmar_bc_dif": (C.c_int, [_V, dptr, dptr, C.c_int, C.c_int, dptr, dptr, dptr, C.c_int, dptr, dptr, dptr, dptr, _V]),
    "brdfgpu_dlevmar_bc_der": (C.c_int, [_V, _V, dptr, dptr, C.c_int, C.c_int, dptr, dptr, dptr, C.c_int, dptr, dptr, dptr, dptr, _V]),
    "brdfgpu_dlevmar_dif": (C.c_int, [_V, dptr, dptr, C.c_int, C.c_int, C.c_int, dptr, dptr, dptr, dptr, _V]),
    "brdfgpu_dlevmar_der": (C.c_int, [_V, _V, dptr, dptr, C.c_int, C.c_int, C.c_int, dptr, dptr, dptr, dptr, _V]),
    "brdfgpu_create": (C.c_int, [C.c_int, C.POINTER(_V)]),
    "brdfgpu_destroy": (None, [_V]),
    "brdfgpu_last_error": (C.c_char_p, [_V]),
    "brdfgpu_launch_count": (C.c_ulonglong, [_V]),
    "brdfgpu_fit_stats": (C.c_int, [_V, C.POINTER(C.c_ulonglong), C.c_int]),
    "brdfgpu_lm_reduced_batching": (C.c_int, [C.c_int]),
    "brdfgpu_stream": (_V, [_V]),
    "brdfgpu_synchronize": (C.c_int, [_V]),
    "brdfgpu_samples_upload": (C.c_int, [_V, C.c_long, dptr, dptr, dptr, C.c_int, C.POINTER(_V)]),
    "brdfgpu_samples_reload": (C.c_int, [_V, _V, dptr, dptr, dptr]),
    "brdfgpu_samples_from_device": (C.c_int, [_V, C.c_long, _V, _V, _V, C.c_int, C.POINTER(_V)]),
    "brdfgpu_samples_synth": (C.c_int, [_V, C.c_long, C.c_ulonglong, C.c_long, dptr, C.c_int, C.POINTER(_V)]),
    "brdfgpu_samples_count": (C.c_long, [_V]),
    "brdfgpu_samples_download": (C.c_int, [_V, _V, dptr, dptr, dptr]),
    "brdfgpu_samples_free": (None, [_V, _V]),
    "brdfgpu_fit_global": (C.c_int, [_V, _V, dptr, C.c_int, dptr, dptr, dptr, C.c_int, dptr, dptr, dptr, C.c_int, C.c_int]),
    "brdfgpu_fit_global_unc": (C.c_int, [_V, _V, dptr, C.c_int, C.c_int, dptr, dptr, dptr, C.c_int]),
    "brdfgpu_eval_residuals": (C.c_int, [_V, _V, dptr, dptr]),
    "brdfgpu_eval_normal_eq": (C.c_int, [_V, _V, dptr, C.c_double, C.c_int, dptr]),
    "brdfgpu_eval_cost": (C.c_int, [_V, _V, dptr, dptr]),
    "brdfgpu_eval_repeat": (C.c_int, [_V, _V, dptr, C.c_double, C.c_int, C.c_int]),
    "brdfgpu_solve_equation": (C.c_int, [dptr, dptr, dptr, dptr, C.c_int, C.c_int, dptr, dptr]),
    "brdfgpu_solve_equation_single": (C.c_int, [dptr, dptr, dptr, dptr, C.c_long, C.c_int, dptr, dptr]),
    "brdfgpu_solve_equation_single_colmajor": (C.c_int, [dptr, dptr, dptr, dptr, C.c_long, C.c_int, C.c_int, dptr, dptr]),
    "brdfgpu_solve_equation_batch": (C.c_int, [_V, C.c_long, C.c_int, dptr, dptr, dptr, dptr, C.c_int, dptr, dptr, iptr]),
    "brdfgpu_batch_upload": (C.c_int, [_V, C.c_long, C.c_int, dptr, dptr, dptr, C.c_int, C.POINTER(_V)]),
    "brdfgpu_batch_synth": (C.c_int, [_V, C.c_long, C.c_int, C.c_ulonglong, C.c_long, C.c_int, C.POINTER(_V)]),
    "brdfgpu_batch_fit": (C.c_int, [_V, _V, dptr, dptr, dptr, C.c_int, dptr, C.c_int]),
    "brdfgpu_batch_results": (C.c_int, [_V, _V, dptr, dptr, iptr]),
    "brdfgpu_batch_count": (C.c_long, [_V]),
    "brdfgpu_batch_free": (None, [_V, _V]),
    "brdfgpu_led_table": (None, [dptr]),
    "brdfgpu_read_cal": (C.c_int, [C.c_char_p, dptr]),
    "brdfgpu_read_obj": (C.c_int, [C.c_char_p, dptr, iptr, iptr, iptr]),
    "brdfgpu_read_png": (C.c_int, [C.c_char_p, _V, iptr, iptr]),
    "brdfgpu_scene_load": (C.c_int, [_V, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(_V), dptr]),
    "brdfgpu_scene_dims": (C.c_int, [_V, iptr]),
    "brdfgpu_scene_set_gather_options": (C.c_int, [_V, _V, C.c_int, dptr, C.c_int]),
    "brdfgpu_read_cal_kappa1": (C.c_int, [C.c_char_p, dptr]),
    "brdfgpu_scene_set_gl_projection": (C.c_int, [_V, _V, dptr, dptr, iptr]),
    "brdfgpu_reference_gl_matrices": (None, [C.c_double, C.c_double, C.c_int, C.c_int, dptr, dptr]),
    "brdfgpu_shade_faces": (C.c_int, [_V, _V, dptr, dptr, C.c_int, C.c_int, dptr, C.c_int, dptr]),
    "brdfgpu_scene_create": (C.c_int, [_V, dptr, C.c_int, iptr, C.c_int, C.POINTER(_V), C.c_int, C.c_int, C.c_int, _V, dptr, C.POINTER(_V)]),
    "brdfgpu_scene_free": (None, [_V, _V]),
    "brdfgpu_scene_face_normals": (C.c_int, [_V, _V, dptr]),
    "brdfgpu_scene_image": (C.c_int, [_V, _V, C.c_int, _V]),
    "brdfgpu_calc_pixel2surface": (C.c_int, [_V, _V, dptr, iptr]),
    "brdfgpu_gather": (C.c_long, [_V, _V, dptr, C.c_int, C.c_long, iptr, lptr, iptr, iptr, dptr, dptr, dptr, dptr]),
    "brdfgpu_gather_resident": (C.c_int, [_V, _V, dptr, C.c_int, C.c_int, C.c_int, C.POINTER(_V), C.POINTER(_V), lptr]),
    "brdfgpu_calc_brdf_equation": (C.c_long, [_V, _V, dptr, C.c_int, dptr]),
    "brdfgpu_calc_brdf_equation_single": (C.c_long, [_V, _V, dptr, C.c_int, dptr, dptr, iptr]),
    "brdfgpu_comm_unique_id": (C.c_int, [C.c_char_p]),
    "brdfgpu_comm_init": (C.c_int, [_V, C.c_char_p, C.c_int, C.c_int]),
    "brdfgpu_comm_destroy": (None, [_V]),
    "brdfgpu_comm_allreduce": (C.c_int, [_V, dptr, C.c_int]),
    "brdfgpu_peer_export": (C.c_int, [_V, C.c_char_p]),
    "brdfgpu_peer_attach": (C.c_int, [_V, C.c_char_p, C.c_int, C.c_int]),
    "brdfgpu_peer_detach": (None, [_V]),
    "brdfgpu_lm_bc_reduced": (C.c_int, [_V, _V, _V, dptr, C.c_int, C.c_long, dptr, dptr, dptr, C.c_int, dptr, dptr, dptr]),
    "brdfgpu_lm_unc_reduced": (C.c_int, [_V, _V, _V, dptr, C.c_int, C.c_long, C.c_int, dptr, dptr, dptr]),
    "brdfgpu_Ax_eq_b_LU": (C.c_int, [dptr, dptr, dptr, C.c_int]),
    "brdfgpu_version": (C.c_char_p, []),
}

REDUCED_JAC_T = C.CFUNCTYPE(None, dptr, C.c_int, dptr, dptr, C.c_void_p)
REDUCED_COST_T = C.CFUNCTYPE(C.c_double, dptr, C.c_int, dptr, C.c_void_p)

_lib = None


def lib():
    """The loaded C-ABI library.  Raises when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise BrdfGpuError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a).  brdf_b200 has no CPU implementation." % path)
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(dptr)


def _arr(a, n=None):
    if a is None:
        return None
    out = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and out.size != n:
        raise ValueError("expected %d values, got %d" % (n, out.size))
    return out


def func_address(name):
    return C.cast(getattr(lib(), name), C.c_void_p)


class Samples:
    """Device-resident sample set of a global fit (brdfgpu_samples)."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    def __len__(self):
        return lib().brdfgpu_samples_count(self.handle)

    def download(self):
        n = len(self)
        c, t, x = np.empty(n), np.empty(n), np.empty(n)
        self.ctx._ok(lib().brdfgpu_samples_download(self.ctx.handle, self.handle, _d(c), _d(t), _d(x)))
        return c, t, x

    def free(self):
        if self.handle:
            lib().brdfgpu_samples_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):   # at interpreter shutdown module globals may already be gone
        try:
            self.free()
        except Exception:
            pass


class Batch:
    """Device-resident set of independent small fits (brdfgpu_batch)."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    def __len__(self):
        return lib().brdfgpu_batch_count(self.handle)

    def fit(self, preset=REF_PERFACE, jac_mode=JAC_FD):
        p0, lb, ub, opts = _arr(preset["p0"], 3), _arr(preset.get("lb")), _arr(preset.get("ub")), _arr(preset.get("opts"))
        self.ctx._ok(lib().brdfgpu_batch_fit(self.ctx.handle, self.handle, _d(p0), _d(lb), _d(ub), int(preset["itmax"]),
                                             _d(opts), jac_mode))

    def results(self):
        n = len(self)
        p, info, ret = np.empty((n, 3)), np.empty((n, 10)), np.empty(n, dtype=np.int32)
        self.ctx._ok(lib().brdfgpu_batch_results(self.ctx.handle, self.handle, _d(p), _d(info), ret.ctypes.data_as(iptr)))
        return p, info, ret

    def free(self):
        if self.handle:
            lib().brdfgpu_batch_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):   # at interpreter shutdown module globals may already be gone
        try:
            self.free()
        except Exception:
            pass


class Scene:
    """Mesh + photographs resident on the device (brdfgpu_scene)."""

    def __init__(self, ctx, handle, nF, nimg, W, H):
        self.ctx, self.handle, self.nF, self.nimg, self.W, self.H = ctx, handle, nF, nimg, W, H

    def face_normals(self):
        fn = np.empty((self.nF, 3))
        self.ctx._ok(lib().brdfgpu_scene_face_normals(self.ctx.handle, self.handle, _d(fn)))
        return fn

    def image(self, k):
        out = np.empty((self.H, self.W, 3), dtype=np.uint8)
        self.ctx._ok(lib().brdfgpu_scene_image(self.ctx.handle, self.handle, k, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_gather_options(self, flags, kappa1=None):
        """Options for later gathers: GATHER_DEPTH_TEST | GATHER_CULL_BACKFACES | GATHER_KAPPA1 (beyond the reference),
        GATHER_SEQ_DOT (left-to-right dot products in GetCosLN / GetCosNH instead of Eigen 3.3's order)."""
        k = _arr(kappa1)
        self.ctx._ok(lib().brdfgpu_scene_set_gather_options(self.ctx.handle, self.handle, int(flags), _d(k), 0 if k is None else k.size))

    def set_gl_projection(self, model_view=None, projection=None, viewport=None):
        """The reference's literal gluProject mapping (brdfdata.cpp:662-677) for later gathers; no arguments: Tsai again."""
        if model_view is None:
            self.ctx._ok(lib().brdfgpu_scene_set_gl_projection(self.ctx.handle, self.handle, None, None, None))
            return
        mv, pr = _arr(model_view, 16), _arr(projection, 16)
        vp = np.ascontiguousarray(viewport, dtype=np.int32)
        self.ctx._ok(lib().brdfgpu_scene_set_gl_projection(self.ctx.handle, self.handle, _d(mv), _d(pr), vp.ctypes.data_as(iptr)))

    def shade_faces(self, eye, center, brdf, model=BLINN_PHONG, literal_cosln=True):
        """Per-face (B, G, R) of the BRDF-shaded preview, glutcallbacks.cpp:346-445.  brdf: (3, 3) single or (nF, 3, 3)."""
        brdf = _arr(brdf)
        single = brdf.size == 9
        assert single or brdf.size == 9 * self.nF
        out = np.empty((self.nF, 3))
        self.ctx._ok(lib().brdfgpu_shade_faces(self.ctx.handle, self.handle, _d(_arr(eye, 3)), _d(_arr(center, 3)), model,
                                               int(single), _d(brdf), int(bool(literal_cosln)), _d(out)))
        return out

    def calc_pixel2surface(self, cam):
        cam = _arr(cam, 16)
        m = np.empty((self.H, self.W), dtype=np.int32)
        self.ctx._ok(lib().brdfgpu_calc_pixel2surface(self.ctx.handle, self.handle, _d(cam), m.ctypes.data_as(iptr)))
        return m

    def gather(self, cams):
        cams = _arr(cams).reshape(-1, 16)
        ncam = cams.shape[0]
        cap = ncam * self.nF
        maps = np.empty((ncam, self.H, self.W), dtype=np.int32)
        nfit_cam = np.zeros(ncam, dtype=np.int64)
        fit_face = np.empty(cap, dtype=np.int32)
        fit_pixel = np.empty(cap, dtype=np.int32)
        phi, td, th = (np.empty((cap, self.nimg)) for _ in range(3))
        inten = np.empty((3, cap, self.nimg))
        n = lib().brdfgpu_gather(self.ctx.handle, self.handle, _d(cams), ncam, cap, maps.ctypes.data_as(iptr),
                                 nfit_cam.ctypes.data_as(lptr), fit_face.ctypes.data_as(iptr),
                                 fit_pixel.ctypes.data_as(iptr), _d(phi), _d(td), _d(th), _d(inten))
        if n < 0:
            self.ctx._ok(-1)
        return dict(nfit=int(n), maps=maps, nfit_cam=nfit_cam, fit_face=fit_face[:n], fit_pixel=fit_pixel[:n],
                    phi=phi[:n], thetaDash=td[:n], theta=th[:n], I=inten[:, :n])

    def gather_resident(self, cams, model=BLINN_PHONG, channel=0, want_global=True, want_batch=False):
        cams = _arr(cams).reshape(-1, 16)
        g, b, nfit = C.c_void_p(), C.c_void_p(), C.c_long()
        self.ctx._ok(lib().brdfgpu_gather_resident(self.ctx.handle, self.handle, _d(cams), cams.shape[0], model, channel,
                                                   C.byref(g) if want_global else None,
                                                   C.byref(b) if want_batch else None, C.byref(nfit)))
        return (Samples(self.ctx, g) if want_global else None, Batch(self.ctx, b) if want_batch else None, nfit.value)

    def calc_brdf_equation(self, cam, model=BLINN_PHONG):
        cam = _arr(cam, 16)
        out = np.full((self.nF, 3, 3), np.nan)
        n = lib().brdfgpu_calc_brdf_equation(self.ctx.handle, self.handle, _d(cam), model, _d(out))
        if n < 0:
            self.ctx._ok(-1)
        return int(n), out

    def calc_brdf_equation_single(self, cam, model=BLINN_PHONG):
        cam = _arr(cam, 16)
        p, info, ret = np.zeros((3, 3)), np.zeros((3, 10)), np.zeros(3, dtype=np.int32)
        n = lib().brdfgpu_calc_brdf_equation_single(self.ctx.handle, self.handle, _d(cam), model, _d(p), _d(info),
                                                    ret.ctypes.data_as(iptr))
        if n < 0:
            self.ctx._ok(-1)
        return int(n), p, info, ret

    def free(self):
        if self.handle:
            lib().brdfgpu_scene_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):   # at interpreter shutdown module globals may already be gone
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU context (brdfgpu_ctx): a device, a stream and the reduction scratch."""

    def __init__(self, device=-1):
        h = C.c_void_p()
        if lib().brdfgpu_create(device, C.byref(h)) != 0:
            raise BrdfGpuError("brdfgpu_create failed: %s" % (lib().brdfgpu_last_error(None) or b"").decode())
        self.handle = h

    def _ok(self, rc):
        if rc != 0:
            raise BrdfGpuError((lib().brdfgpu_last_error(self.handle) or b"").decode() or "brdfgpu call failed")

    @property
    def launches(self):
        return int(lib().brdfgpu_launch_count(self.handle))

    @property
    def stream(self):
        return lib().brdfgpu_stream(self.handle)

    def fit_stats(self):
        """dict(jac_passes, cost_passes, cost_points, resident_samples, ctas) of the last global fit"""
        buf = (C.c_ulonglong * 24)()
        self._ok(lib().brdfgpu_fit_stats(self.handle, buf, 24))
        return dict(jac_passes=int(buf[0]), cost_passes=int(buf[1]), cost_points=int(buf[2]), resident_samples=int(buf[3]),
                    ctas=int(buf[4]), cyc_sweep=int(buf[5]), cyc_exchange=int(buf[6]), cyc_total=int(buf[7]),
                    cyc_exchange_phases=[int(buf[8 + i]) for i in range(4)],
                    cyc_control_by_next_sweep=dict(zip(("quit", "jac_fwd", "jac_central", "jac_analytic", "cost", "many", "bad"),
                                                       (int(buf[12 + i]) for i in range(7)))),
                    spec_jac_issued=int(buf[19]), spec_jac_hits=int(buf[20]), creep_fused=int(buf[21]),
                    secant_updates=int(buf[22]), secant_updates_accepted=int(buf[23]))

    def synchronize(self):
        self._ok(lib().brdfgpu_synchronize(self.handle))

    # ---- resident sample sets ----
    def upload(self, cosphi, t, x, model=BLINN_PHONG):
        c, t = _arr(cosphi), _arr(t)
        x = _arr(x)
        h = C.c_void_p()
        self._ok(lib().brdfgpu_samples_upload(self.handle, c.size, _d(c), _d(t), _d(x), model, C.byref(h)))
        return Samples(self, h)

    def from_device(self, n, d_cosphi, d_t, d_x, model=BLINN_PHONG):
        h = C.c_void_p()
        self._ok(lib().brdfgpu_samples_from_device(self.handle, n, d_cosphi, d_t, d_x, model, C.byref(h)))
        return Samples(self, h)

    def synth(self, n, seed, start=0, truth=(0.6, 0.35, 12.0), model=BLINN_PHONG):
        tr = _arr(truth, 3)
        h = C.c_void_p()
        self._ok(lib().brdfgpu_samples_synth(self.handle, n, seed, start, _d(tr), model, C.byref(h)))
        return Samples(self, h)

    # ---- global fits ----
    def fit_global(self, samples, preset=REF_GLOBAL, drive=DRIVE_PERSISTENT, jac_mode=JAC_FD, dscl=None,
                   want_covar=False, p0=None):
        p = _arr(preset["p0"] if p0 is None else p0, 3).copy()
        lb, ub, opts, dscl = _arr(preset.get("lb")), _arr(preset.get("ub")), _arr(preset.get("opts")), _arr(dscl)
        info = np.zeros(10)
        covar = np.zeros((3, 3)) if want_covar else None
        ret = lib().brdfgpu_fit_global(self.handle, samples.handle, _d(p), 3, _d(lb), _d(ub), _d(dscl),
                                       int(preset["itmax"]), _d(opts), _d(info), _d(covar), drive, jac_mode)
        return (ret, p, info, covar) if want_covar else (ret, p, info)

    def fit_global_unc(self, samples, p0, itmax, opts, jac_mode=JAC_FD, want_covar=False):
        p = _arr(p0, 3).copy()
        opts = _arr(opts)
        info = np.zeros(10)
        covar = np.zeros((3, 3)) if want_covar else None
        ret = lib().brdfgpu_fit_global_unc(self.handle, samples.handle, _d(p), 3, int(itmax), _d(opts), _d(info),
                                           _d(covar), jac_mode)
        return (ret, p, info, covar) if want_covar else (ret, p, info)

    def residuals(self, samples, p):
        p = _arr(p, 3)
        e = np.empty(len(samples))
        self._ok(lib().brdfgpu_eval_residuals(self.handle, samples.handle, _d(p), _d(e)))
        return e

    def normal_eq(self, samples, p, delta, jac_mode=JAC_FD):
        p = _arr(p, 3)
        out = np.empty(11)
        self._ok(lib().brdfgpu_eval_normal_eq(self.handle, samples.handle, _d(p), float(delta), jac_mode, _d(out)))
        return out

    def cost(self, samples, p):
        p = _arr(p, 3)
        out = np.empty(2)
        self._ok(lib().brdfgpu_eval_cost(self.handle, samples.handle, _d(p), _d(out)))
        return out

    def repeat(self, samples, p, delta, kind, reps):
        p = _arr(p, 3)
        self._ok(lib().brdfgpu_eval_repeat(self.handle, samples.handle, _d(p), float(delta), kind, reps))

    # ---- batched fits ----
    def batch_upload(self, cosphi, t, x, model=BLINN_PHONG):
        c, t, x = _arr(cosphi), _arr(t), _arr(x)
        nfit, nper = np.asarray(cosphi).shape
        h = C.c_void_p()
        self._ok(lib().brdfgpu_batch_upload(self.handle, nfit, nper, _d(c), _d(t), _d(x), model, C.byref(h)))
        return Batch(self, h)

    def batch_synth(self, nfit, nper, seed, first_fit=0, model=BLINN_PHONG):
        h = C.c_void_p()
        self._ok(lib().brdfgpu_batch_synth(self.handle, nfit, nper, seed, first_fit, model, C.byref(h)))
        return Batch(self, h)

    def solve_equation_batch(self, phi, thetaDash, theta, inten, model=BLINN_PHONG):
        """The per-pixel loop of CalcBRDFEquation (brdfdata.cpp:1195-1221) in one launch."""
        phi, td, th, inten = _arr(phi), _arr(thetaDash), _arr(theta), _arr(inten)
        nfit, nper = np.asarray(phi).shape
        p, info, ret = np.empty((nfit, 3)), np.empty((nfit, 10)), np.empty(nfit, dtype=np.int32)
        self._ok(lib().brdfgpu_solve_equation_batch(self.handle, nfit, nper, _d(phi), _d(td), _d(th), _d(inten), model,
                                                    _d(p), _d(info), ret.ctypes.data_as(iptr)))
        return p, info, ret

    # ---- gather ----
    def scene(self, V, F, images, dark=None, led=None):
        V = _arr(V).reshape(-1, 3)
        F = np.ascontiguousarray(F, dtype=np.int32).reshape(-1, 3)
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        H, W = imgs[0].shape[:2]
        ptrs = (C.c_void_p * len(imgs))(*[im.ctypes.data for im in imgs])
        dark = None if dark is None else np.ascontiguousarray(dark, dtype=np.uint8)
        led = _arr(led)
        h = C.c_void_p()
        self._ok(lib().brdfgpu_scene_create(self.handle, _d(V), V.shape[0], F.ctypes.data_as(iptr), F.shape[0], ptrs,
                                            len(imgs), W, H, None if dark is None else dark.ctypes.data_as(C.c_void_p),
                                            _d(led), C.byref(h)))
        return Scene(self, h, F.shape[0], len(imgs), W, H)

    def scene_load(self, image_folder, obj_path, cal_path=None, nimg=16):
        """main.cpp:41-59: LoadModel, LoadImages, SubtractAmbientLight, LoadCameraParameters, InitLEDs -- from the
        reference's own files.  Returns (Scene, cam16 or None)."""
        h = C.c_void_p()
        cam = np.zeros(16)
        self._ok(lib().brdfgpu_scene_load(self.handle, os.fsencode(image_folder), os.fsencode(obj_path),
                                          None if cal_path is None else os.fsencode(cal_path), nimg, C.byref(h), _d(cam)))
        dims = np.zeros(5, dtype=np.int32)
        self._ok(lib().brdfgpu_scene_dims(h, dims.ctypes.data_as(iptr)))
        return Scene(self, h, int(dims[1]), int(dims[2]), int(dims[3]), int(dims[4])), (None if cal_path is None else cam)

    # ---- multi-GPU ----
    def comm_init(self, unique_id, rank, nranks):
        self._ok(lib().brdfgpu_comm_init(self.handle, unique_id, rank, nranks))

    def peer_export(self):
        """This rank's 72-byte export record (CUDA-IPC handle + the exchange tag reached); export again before every
        (re-)attach"""
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._ok(lib().brdfgpu_peer_export(self.handle, buf))
        return buf.raw

    def peer_attach(self, handles, rank, nranks):
        """handles: the export records of all ranks in rank order"""
        self._ok(lib().brdfgpu_peer_attach(self.handle, b"".join(handles), rank, nranks))

    def peer_detach(self):
        lib().brdfgpu_peer_detach(self.handle)

    def allreduce(self, values):
        buf = _arr(values).copy()
        self._ok(lib().brdfgpu_comm_allreduce(self.handle, _d(buf), buf.size))
        return buf

    def close(self):
        if self.handle:
            lib().brdfgpu_peer_detach(self.handle)
            lib().brdfgpu_destroy(self.handle)
            self.handle = None


def comm_unique_id():
    buf = C.create_string_buffer(128)
    if lib().brdfgpu_comm_unique_id(buf) != 0:
        raise BrdfGpuError("ncclGetUniqueId failed")
    return buf.raw


# ---- the levmar-signature calls exactly as the reference makes them (brdfdata.cpp:1058, 1119) ----
def make_extra(cosphi, costhetadash, costheta, model):
    n = np.asarray(cosphi).size
    angles = np.empty(3 * n)
    angles[:n] = cosphi
    angles[n:2 * n] = costhetadash
    angles[2 * n:] = 0.0 if costheta is None else costheta
    return ExtraData(angles.ctypes.data_as(dptr), int(model)), angles


def dlevmar_bc_dif(p0, x, lb, ub, itmax, opts, extra, dscl=None, want_covar=False, func=None):
    p = _arr(p0).copy()
    x = _arr(x)
    m, n = p.size, x.size
    lb, ub, opts, dscl = _arr(lb), _arr(ub), _arr(opts), _arr(dscl)
    info = np.zeros(10)
    covar = np.zeros((m, m)) if want_covar else None
    f = func_address("brdfgpu_BRDFFunc") if func is None else func
    ret = lib().brdfgpu_dlevmar_bc_dif(f, _d(p), _d(x), m, n, _d(lb), _d(ub), _d(dscl), int(itmax), _d(opts), _d(info),
                                       None, _d(covar), C.cast(C.pointer(extra), C.c_void_p))
    return ret, p, info, covar


def dlevmar_bc_der(p0, x, lb, ub, itmax, opts, extra, dscl=None):
    p = _arr(p0).copy()
    x = _arr(x)
    m, n = p.size, x.size
    lb, ub, opts, dscl = _arr(lb), _arr(ub), _arr(opts), _arr(dscl)
    info = np.zeros(10)
    ret = lib().brdfgpu_dlevmar_bc_der(func_address("brdfgpu_BRDFFunc"), func_address("brdfgpu_BRDFJac"), _d(p), _d(x), m, n,
                                       _d(lb), _d(ub), _d(dscl), int(itmax), _d(opts), _d(info), None, None,
                                       C.cast(C.pointer(extra), C.c_void_p))
    return ret, p, info


def dlevmar_dif(p0, x, itmax, opts, extra, want_covar=False):
    p = _arr(p0).copy()
    x = _arr(x)
    m, n = p.size, x.size
    opts = _arr(opts)
    info = np.zeros(10)
    covar = np.zeros((m, m)) if want_covar else None
    ret = lib().brdfgpu_dlevmar_dif(func_address("brdfgpu_BRDFFunc"), _d(p), _d(x), m, n, int(itmax), _d(opts), _d(info),
                                    None, _d(covar), C.cast(C.pointer(extra), C.c_void_p))
    return ret, p, info, covar


def dlevmar_der(p0, x, itmax, opts, extra):
    p = _arr(p0).copy()
    x = _arr(x)
    m, n = p.size, x.size
    opts = _arr(opts)
    info = np.zeros(10)
    ret = lib().brdfgpu_dlevmar_der(func_address("brdfgpu_BRDFFunc"), func_address("brdfgpu_BRDFJac"), _d(p), _d(x), m, n,
                                    int(itmax), _d(opts), _d(info), None, None, C.cast(C.pointer(extra), C.c_void_p))
    return ret, p, info


def BRDFFunc(p, extra, n):
    p = _arr(p, 3).copy()
    hx = np.zeros(n)
    lib().brdfgpu_BRDFFunc(_d(p), _d(hx), 3, n, C.cast(C.pointer(extra), C.c_void_p))
    return hx


def BRDFJac(p, extra, n, m=3):
    p = _arr(p, 3).copy()
    jac = np.zeros((n, m))
    lib().brdfgpu_BRDFJac(_d(p), _d(jac), m, n, C.cast(C.pointer(extra), C.c_void_p))
    return jac


def solve_equation(phi, thetaDash, theta, inten, model=BLINN_PHONG):
    phi, td, th, inten = _arr(phi), _arr(thetaDash), _arr(theta), _arr(inten)
    p, info = np.zeros(3), np.zeros(10)
    ret = lib().brdfgpu_solve_equation(_d(phi), _d(td), _d(th), _d(inten), phi.size, model, _d(p), _d(info))
    return ret, p, info


def solve_equation_single(phi, thetaDash, theta, inten, model=BLINN_PHONG):
    phi, td, th, inten = _arr(phi), _arr(thetaDash), _arr(theta), _arr(inten)
    p, info = np.zeros(3), np.zeros(10)
    ret = lib().brdfgpu_solve_equation_single(_d(phi), _d(td), _d(th), _d(inten), phi.size, model, _d(p), _d(info))
    return ret, p, info


def solve_equation_single_colmajor(phi, thetaDash, theta, inten, model=BLINN_PHONG):
    """SolveEquation_SingleBRDF with the reference's literal flattening (brdfdata.cpp:1008-1042); inputs rows x nimg"""
    rows, nimg = np.asarray(phi).shape
    phi, td, th, inten = _arr(phi), _arr(thetaDash), _arr(theta), _arr(inten)
    p, info = np.zeros(3), np.zeros(10)
    ret = lib().brdfgpu_solve_equation_single_colmajor(_d(phi), _d(td), _d(th), _d(inten), rows, nimg, model, _d(p), _d(info))
    return ret, p, info


GATHER_DEPTH_TEST, GATHER_CULL_BACKFACES, GATHER_KAPPA1, GATHER_SEQ_DOT = 1, 2, 4, 8


# ---- the reference's input files (host code) ----
def read_cal_kappa1(path):
    k = C.c_double(0.0)
    has = lib().brdfgpu_read_cal_kappa1(os.fsencode(path), C.byref(k))
    if has < 0:
        raise BrdfGpuError(lib().brdfgpu_last_error(None).decode())
    return (k.value if has else None)


def reference_gl_matrices(cx, cy, window_width=1920, window_height=1080):
    """(model_view16, projection16) as the reference's Display_ sets them up (glutcallbacks.cpp:626-642, 672-689)"""
    mv, pr = np.zeros(16), np.zeros(16)
    lib().brdfgpu_reference_gl_matrices(float(cx), float(cy), int(window_width), int(window_height), _d(mv), _d(pr))
    return mv, pr


def read_cal(path):
    """(cam16, mask of the fields present) -- CBRDFdata::LoadCameraParameters, brdfdata.cpp:149-247"""
    cam = np.zeros(16)
    mask = lib().brdfgpu_read_cal(os.fsencode(path), _d(cam))
    if mask < 0:
        raise BrdfGpuError(lib().brdfgpu_last_error(None).decode())
    return cam, mask


def read_obj(path):
    """(V nV x 3 float64, F nF x 3 int32) -- igl::readOBJ as CBRDFdata::LoadModel uses it, brdfdata.cpp:289-312"""
    nV, nF = C.c_int(0), C.c_int(0)
    if lib().brdfgpu_read_obj(os.fsencode(path), None, None, C.byref(nV), C.byref(nF)) != 0:
        raise BrdfGpuError(lib().brdfgpu_last_error(None).decode())
    V, F = np.empty((nV.value, 3)), np.empty((nF.value, 3), dtype=np.int32)
    if lib().brdfgpu_read_obj(os.fsencode(path), _d(V), F.ctypes.data_as(iptr), C.byref(nV), C.byref(nF)) != 0:
        raise BrdfGpuError(lib().brdfgpu_last_error(None).decode())
    return V, F


def read_png(path):
    """H x W x 3 uint8 BGR -- cv::imread(path, IMREAD_COLOR) for 8-bit PNG files"""
    W, H = C.c_int(0), C.c_int(0)
    if lib().brdfgpu_read_png(os.fsencode(path), None, C.byref(W), C.byref(H)) != 0:
        raise BrdfGpuError(lib().brdfgpu_last_error(None).decode())
    out = np.empty((H.value, W.value, 3), dtype=np.uint8)
    if lib().brdfgpu_read_png(os.fsencode(path), out.ctypes.data_as(C.c_void_p), C.byref(W), C.byref(H)) != 0:
        raise BrdfGpuError(lib().brdfgpu_last_error(None).decode())
    return out


def lm_bc_reduced(jac_cb, cost_cb, p0, n, lb, ub, itmax, opts, dscl=None, want_covar=False):
    """Host instantiation of the product's LM control loop on caller-supplied reduced sums."""
    p = _arr(p0).copy()
    m = p.size
    lb, ub, opts, dscl = _arr(lb), _arr(ub), _arr(opts), _arr(dscl)
    info = np.zeros(10)
    covar = np.zeros((m, m)) if want_covar else None
    jc, cc = REDUCED_JAC_T(jac_cb), REDUCED_COST_T(cost_cb)
    ret = lib().brdfgpu_lm_bc_reduced(C.cast(jc, C.c_void_p), C.cast(cc, C.c_void_p), None, _d(p), m, n, _d(lb), _d(ub),
                                      _d(dscl), int(itmax), _d(opts), _d(info), _d(covar))
    return ret, p, info, covar


def lm_unc_reduced(jac_cb, cost_cb, p0, n, itmax, opts, want_covar=False):
    p = _arr(p0).copy()
    m = p.size
    opts = _arr(opts)
    info = np.zeros(10)
    covar = np.zeros((m, m)) if want_covar else None
    jc, cc = REDUCED_JAC_T(jac_cb), REDUCED_COST_T(cost_cb)
    ret = lib().brdfgpu_lm_unc_reduced(C.cast(jc, C.c_void_p), C.cast(cc, C.c_void_p), None, _d(p), m, n, int(itmax),
                                       _d(opts), _d(info), _d(covar))
    return ret, p, info, covar

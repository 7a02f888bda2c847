// gather.cu -- K1: multi-view sample gather.
//
// Replaces CBRDFdata::CalcPixel2SurfaceMapping (brdfdata.cpp:629-681: face centroid -> window
// coordinates -> pixel<->face map, last face wins), GetIntensities_FromPixel (:945-960: 16 radiance
// samples per pixel and colour channel, u8/255.0), GetCosLN / GetCosNH / GetCosRV (:857-899,
// :902-943, :799-855: per (face, LED) cosines), CalcFaceNormals (:314-330) and
// SubtractAmbientLight (:130-147).  The camera is the Tsai .cal calibration.
//
// Bit-exactness contract (oracle/gather_oracle.c is the definition): every arithmetic step is one
// IEEE-754 double operation issued through the round-to-nearest intrinsics below, so no FMA
// contraction can occur whatever flags this file is compiled with; 3-term sums associate the way
// Eigen 3.3 evaluates the reference's expressions (oracle/gather_oracle.c has the derivation): left to
// right everywhere except the final dot products of GetCosLN / GetCosNH, a0*b0 + (a1*b1 + a2*b2)
// (BRDFGPU_GATHER_SEQ_DOT restores left to right there); double -> int truncates.  "Last face wins"
// is atomicMax(face id): the reference walks faces in ascending order, so the survivor is the largest id.
//
// Work decomposition: one thread per (view, face) for the projection, one thread per
// (fit, LED) for the sample rows, so the 16 LEDs of a fit write 16 consecutive doubles.
// Face centroids and normals are per-scene constants (computed once by k_face_geometry); the sample
// kernel writes straight into the arrays of the fit stage (cosphi, model cosine, its log, the channel's
// intensities), so a resident gather needs no device-to-device hand-over and one synchronisation.
#include <cmath>
#include <vector>

#include "brdf_model.cuh"
#include "common.cuh"

// BRDFGPU_TRACE=1: wall-clock milestones of the scene drivers on stderr (host-side diagnosis)
#include <chrono>
#include <cstdlib>
namespace {
struct Trace {
    const char* what;
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(const char* w) : what(w), on(getenv("BRDFGPU_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* step) {
        if (!on) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[brdfgpu trace] %s: %s +%.2f ms\n", what, step, std::chrono::duration<double, std::milli>(t - t0).count());
        t0 = t;
    }
};
}  // namespace

struct brdfgpu_scene {
    int nV = 0, nF = 0, nimg = 0, W = 0, H = 0;
    int gather_flags = 0;          // BRDFGPU_GATHER_* (0 = the reference's behaviour)
    std::vector<double> kappa1;    // per camera of the next gather calls (BRDFGPU_GATHER_KAPPA1)
    double* V = nullptr;           // nV x 3
    int* F = nullptr;              // nF x 3
    double* FN = nullptr;          // nF x 3
    double* FC = nullptr;          // nF x 3 face centroids, brdfdata.cpp:653-660 (the same for every view and LED)
    unsigned char* img = nullptr;  // nimg x H x W x 3 (BGR), ambient already removed
    double* led = nullptr;         // nimg x 3
    // the reference's literal projection (brdfgpu_scene_set_gl_projection); off: the Tsai camera
    bool gl_on = false;
    double gl_mv[16] = {0}, gl_proj[16] = {0};
    int gl_viewport[4] = {0, 0, 0, 0};
};

namespace brdfgpu {

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

struct V3 {
    double x, y, z;
};
constexpr double kPiShade = 3.1415926535897932384626433832795;  // CV_PI
__device__ __forceinline__ double dot3(const V3& a, const V3& b) {
    return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z));
}
// fixed-size vector .cwiseProduct(row of a column-major MatrixXd).sum() in Eigen 3.3 (brdfdata.cpp:893, 937):
// the unrolled scalar reduction splits 3 terms as 1 + 2
__device__ __forceinline__ double dot3_row(const V3& a, const V3& b) {
    return dadd(dmul(a.x, b.x), dadd(dmul(a.y, b.y), dmul(a.z, b.z)));
}
__device__ __forceinline__ V3 normalized(V3 v) {
    const double z = dot3(v, v);
    if (z > 0.0) {
        const double r = __dsqrt_rn(z);
        v.x = ddiv(v.x, r); v.y = ddiv(v.y, r); v.z = ddiv(v.z, r);
    }
    return v;
}
__device__ __forceinline__ V3 load3(const double* p) { return V3{p[0], p[1], p[2]}; }

// brdfdata.cpp:653-660: x = 0; x += v0; x += v1; x += v2; x /= 3.0
__device__ __forceinline__ V3 centroid_of(const double* __restrict__ V, const int* __restrict__ F, int face) {
    const V3 a = load3(V + 3l * F[3l * face + 0]);
    const V3 b = load3(V + 3l * F[3l * face + 1]);
    const V3 c = load3(V + 3l * F[3l * face + 2]);
    V3 s;
    s.x = ddiv(dadd(dadd(dadd(0.0, a.x), b.x), c.x), 3.0);
    s.y = ddiv(dadd(dadd(dadd(0.0, a.y), b.y), c.y), 3.0);
    s.z = ddiv(dadd(dadd(dadd(0.0, a.z), b.z), c.z), 3.0);
    return s;
}

struct Camera {
    double cx, cy, f, sx;
    V3 n, o, a, p;
};
__device__ __forceinline__ Camera load_camera(const double* cam) {
    Camera c;
    c.cx = cam[0]; c.cy = cam[1]; c.f = cam[2]; c.sx = cam[3];
    c.n = load3(cam + 4); c.o = load3(cam + 7); c.a = load3(cam + 10); c.p = load3(cam + 13);
    return c;
}

// Tsai pin-hole projection; returns the pixel index row*W+col or -1 (behind / outside)
__device__ __forceinline__ int project_tsai(const V3& c, const Camera& cam, int W, int H) {
    const V3 d{dsub(c.x, cam.p.x), dsub(c.y, cam.p.y), dsub(c.z, cam.p.z)};
    const double xc = dot3(d, cam.n), yc = dot3(d, cam.o), zc = dot3(d, cam.a);
    if (!(zc > 0.0)) return -1;
    const double u = dadd(cam.cx, ddiv(dmul(dmul(cam.sx, cam.f), xc), zc));
    const double v = dadd(cam.cy, ddiv(dmul(cam.f, yc), zc));
    if (!(u >= 0.0 && v >= 0.0 && u < (double)W && v < (double)H)) return -1;
    return (int)v * W + (int)u;
}

__global__ void k_face_geometry(const double* __restrict__ V, const int* __restrict__ F, int nF, double* __restrict__ FN,
                                double* __restrict__ FC) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nF) return;
    const V3 ctr = centroid_of(V, F, i);
    FC[3l * i] = ctr.x; FC[3l * i + 1] = ctr.y; FC[3l * i + 2] = ctr.z;
    const V3 v0 = load3(V + 3l * F[3l * i]), v1 = load3(V + 3l * F[3l * i + 1]), v2 = load3(V + 3l * F[3l * i + 2]);
    const V3 e1{dsub(v1.x, v0.x), dsub(v1.y, v0.y), dsub(v1.z, v0.z)};
    const V3 e2{dsub(v2.x, v0.x), dsub(v2.y, v0.y), dsub(v2.z, v0.z)};
    V3 n;
    n.x = dsub(dmul(e1.y, e2.z), dmul(e1.z, e2.y));
    n.y = dsub(dmul(e1.z, e2.x), dmul(e1.x, e2.z));
    n.z = dsub(dmul(e1.x, e2.y), dmul(e1.y, e2.x));
    n = normalized(n);
    FN[3l * i] = n.x; FN[3l * i + 1] = n.y; FN[3l * i + 2] = n.z;
}

// img = sat(sat(img - dark) - dark), brdfdata.cpp:140-146; 16 bytes per thread step
__global__ void k_subtract_ambient(unsigned char* __restrict__ img, const unsigned char* __restrict__ dark, long per_image,
                                   int nimg) {
    const long total = per_image * nimg;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int d = dark[i % per_image];
        int v = (int)img[i] - d;
        v = v < 0 ? 0 : v;
        v -= d;
        v = v < 0 ? 0 : v;
        img[i] = (unsigned char)v;
    }
}

// BRDF-shaded preview, one thread per face (glutcallbacks.cpp:346-445): the light sits at the eye,
//   lightDir = normalize(eye - centroid), viewDir = normalize(eye - center), h = normalize(lightDir + viewDir)
//   Blinn-Phong: kd*cosLN + ks*pow(N.h, n)         Phong: kd*cosLN + ks*((n+2)/(2 pi))*pow((float)(viewDir.R), n)
// per colour channel (B, G, R).  LITERAL keeps the reference's cosLN, which indexes the normal with the
// truncated dot product -- face_normals(i, (int)(N.lightDir)) -- instead of using the dot product itself.
// Same single-rounding arithmetic as the gather; only pow() is the device libm's.
template <bool LITERAL>
__global__ void k_shade_faces(const double* __restrict__ FC, const double* __restrict__ FN, int nF,
                              V3 eye, V3 center, int model, int single, const double* __restrict__ brdf,
                              double* __restrict__ bgr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nF) return;
    const V3 c = load3(FC + 3l * i);
    const V3 N = load3(FN + 3l * i);
    const V3 l = normalized(V3{dsub(eye.x, c.x), dsub(eye.y, c.y), dsub(eye.z, c.z)});
    const V3 v = normalized(V3{dsub(eye.x, center.x), dsub(eye.y, center.y), dsub(eye.z, center.z)});
    const double nl = dot3(N, l);
    double cosLN = nl;
    if (LITERAL) {
        int col = (int)nl;  // |N.l| < 1 -> 0: the x component of the normal
        col = col < 0 ? 0 : (col > 2 ? 2 : col);
        cosLN = col == 0 ? N.x : (col == 1 ? N.y : N.z);
    }
    double t;  // the cosine under the power
    if (model == 1) {
        const V3 h = normalized(V3{dadd(l.x, v.x), dadd(l.y, v.y), dadd(l.z, v.z)});
        t = dot3(N, h);
    } else {
        const double sf = -nl;  // P = -scale_factor * N, R = lightDir - 2*P
        const V3 R{dsub(l.x, dmul(2.0, dmul(sf, N.x))), dsub(l.y, dmul(2.0, dmul(sf, N.y))), dsub(l.z, dmul(2.0, dmul(sf, N.z)))};
        t = (double)(float)dot3(v, R);  // `float cosRV`, glutcallbacks.cpp:420
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double* q = brdf + (single ? 3l * ch : 9l * i + 3l * ch);
        const double kd = q[0], ks = q[1], n = q[2];
        const double pw = pow(t, n);
        const double spec = model == 1 ? dmul(ks, pw) : dmul(dmul(ks, ddiv(dadd(n, 2.0), dmul(2.0, kPiShade))), pw);
        bgr[3l * i + ch] = dadd(dmul(kd, cosLN), spec);
    }
}

__global__ void k_fill_int(int* p, long n, int value) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = value;
}

// Projection with the options beyond the reference (the oracle's project_opts): back-face
// culling N.(p - C) > 0, Tsai's radial distortion by five fixed-point steps, depth returned for the z-test.
__device__ __forceinline__ int project_opts(const V3& c, const V3& N, const Camera& cam, double kappa1, int flags, int W, int H,
                                            double* depth) {
    const V3 d{dsub(c.x, cam.p.x), dsub(c.y, cam.p.y), dsub(c.z, cam.p.z)};
    const double xc = dot3(d, cam.n), yc = dot3(d, cam.o), zc = dot3(d, cam.a);
    if (!(zc > 0.0)) return -1;
    if (flags & BRDFGPU_GATHER_CULL_BACKFACES) {
        const V3 toward{-d.x, -d.y, -d.z};
        if (!(dot3(N, toward) > 0.0)) return -1;
    }
    double u, v;
    if (flags & BRDFGPU_GATHER_KAPPA1) {
        const double xu = ddiv(dmul(cam.f, xc), zc), yu = ddiv(dmul(cam.f, yc), zc);
        double xd = xu, yd = yu;
#pragma unroll 1
        for (int it = 0; it < 5; ++it) {
            const double s = dadd(1.0, dmul(kappa1, dadd(dmul(xd, xd), dmul(yd, yd))));
            xd = ddiv(xu, s);
            yd = ddiv(yu, s);
        }
        u = dadd(cam.cx, dmul(cam.sx, xd));
        v = dadd(cam.cy, yd);
    } else {
        u = dadd(cam.cx, ddiv(dmul(dmul(cam.sx, cam.f), xc), zc));
        v = dadd(cam.cy, ddiv(dmul(cam.f, yc), zc));
    }
    if (!(u >= 0.0 && v >= 0.0 && u < (double)W && v < (double)H)) return -1;
    *depth = zc;
    return (int)v * W + (int)u;
}

// one thread per (view, face) with options: pass 0 records the pixel and (depth test) the nearest depth per pixel
// (zc > 0, so the bits of the double order like the value); pass 1 lets the faces at that depth claim the pixel
__global__ void k_project_opts(const double* __restrict__ FC, const double* __restrict__ FN, int nF,
                               const double* __restrict__ cams, const double* __restrict__ kappa1, int flags, int ncam, int W,
                               int H, int pass, int* __restrict__ pix, unsigned long long* __restrict__ zbits,
                               unsigned long long* __restrict__ depth, int* __restrict__ maps) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)ncam * nF) return;
    const int v = (int)(e / nF), face = (int)(e % nF);
    if (pass == 0) {
        const Camera cam = load_camera(cams + 16l * v);
        double z = 0.0;
        const int px = project_opts(load3(FC + 3l * face), load3(FN + 3l * face), cam, kappa1 ? kappa1[v] : 0.0, flags, W, H, &z);
        pix[e] = px;
        if (px < 0) return;
        if (flags & BRDFGPU_GATHER_DEPTH_TEST) {
            const unsigned long long zb = (unsigned long long)__double_as_longlong(z);
            zbits[e] = zb;
            atomicMin(depth + (long)v * W * H + px, zb);
        } else {
            atomicMax(maps + (long)v * W * H + px, face);
        }
    } else {
        const int px = pix[e];
        if (px >= 0 && zbits[e] == depth[(long)v * W * H + px]) atomicMax(maps + (long)v * W * H + px, face);
    }
}

__global__ void k_fill_u64(unsigned long long* p, long n, unsigned long long value) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = value;
}

// one thread per (view, face): pixel of the centroid, and the per-view map by atomicMax
__global__ void k_project(const double* __restrict__ FC, int nF, const double* __restrict__ cams,
                          int ncam, int W, int H, int* __restrict__ pix /*ncam*nF*/, int* __restrict__ maps /*ncam*H*W*/) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)ncam * nF) return;
    const int v = (int)(e / nF), face = (int)(e % nF);
    const Camera cam = load_camera(cams + 16l * v);
    const int px = project_tsai(load3(FC + 3l * face), cam, W, H);
    pix[e] = px;
    if (px >= 0) atomicMax(maps + (long)v * W * H + px, face);
}

// The reference's literal projection (brdfdata.cpp:662-677): gluProject through the GL MODELVIEW / PROJECTION matrices and
// the viewport as the caller read them back from GL.  gluProject is libGLU's (SGI / Mesa GLU 9.0 project.c): two
// column-major 4x4 matrix-vector products summed left to right, perspective divide, * 0.5 + 0.5, viewport scale.  The
// reference writes the map for winX, winY >= 0 with no upper bound; here only pixels inside the map are written.
// Rows are GL rows (bottom-up); the radiance fetch flips them (brdfdata.cpp:955).
struct GlProjection {
    double mv[16], proj[16];
    double vx, vy, vw, vh;  // viewport as doubles (the int -> double conversions of in[0] * viewport[2] + viewport[0])
};
__device__ __forceinline__ void glu_mult(const double* m, const double* in, double* out) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        out[i] = dadd(dadd(dadd(dmul(in[0], m[i]), dmul(in[1], m[4 + i])), dmul(in[2], m[8 + i])), dmul(in[3], m[12 + i]));
}
__global__ void k_project_gl(const double* __restrict__ FC, int nF, GlProjection gl, int W, int H, int* __restrict__ pix,
                             int* __restrict__ maps) {
    const int face = blockIdx.x * blockDim.x + threadIdx.x;
    if (face >= nF) return;
    const double in[4] = {FC[3l * face], FC[3l * face + 1], FC[3l * face + 2], 1.0};
    double eye[4], clip[4];
    glu_mult(gl.mv, in, eye);
    glu_mult(gl.proj, eye, clip);
    int px = -1;
    if (clip[3] != 0.0) {
        const double winx = dadd(dmul(dadd(dmul(ddiv(clip[0], clip[3]), 0.5), 0.5), gl.vw), gl.vx);
        const double winy = dadd(dmul(dadd(dmul(ddiv(clip[1], clip[3]), 0.5), 0.5), gl.vh), gl.vy);
        if (winy >= 0 && winx >= 0 && winx < (double)W && winy < (double)H) px = (int)winy * W + (int)winx;
    }
    pix[face] = px;
    if (px >= 0) atomicMax(maps + px, face);
}

// ---- ordered compaction of the faces that still own their pixel (view-major, ascending face id) ----
constexpr int kScanThreads = 1024;

__device__ __forceinline__ int owner_flag(const int* pix, const int* maps, int nF, int W, int H, long e, long total) {
    if (e >= total) return 0;
    const int px = pix[e];
    if (px < 0) return 0;
    return maps[(e / nF) * (long)W * H + px] == (int)(e % nF);
}

// exclusive scan of one int per thread over a 1024-thread block; returns the prefix, *total the block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += t;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const int base = warp ? warp_sums[warp - 1] : 0;
    *total = warp_sums[31];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) k_owner_count(const int* pix, const int* maps, int nF, int W, int H,
                                                               long total, int* block_sums) {
    const long e = (long)blockIdx.x * kScanThreads + threadIdx.x;
    int tot;
    block_exclusive_scan(owner_flag(pix, maps, nF, W, H, e, total), &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of the block sums in place; grand total to block_sums[nblocks]
__global__ void __launch_bounds__(kScanThreads) k_scan_block_sums(int* block_sums, int nblocks) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? block_sums[i] : 0;
        int tot;
        const int pre = block_exclusive_scan(v, &tot);
        const int c = carry;
        if (i < nblocks) block_sums[i] = c + pre;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_owner_scatter(const int* pix, const int* maps, int nF, int W, int H,
                                                                 long total, const int* block_offsets, int* fit_face,
                                                                 int* fit_pixel, int* fit_cam, int* cam_first /*ncam+1*/,
                                                                 int ncam) {
    const long e = (long)blockIdx.x * kScanThreads + threadIdx.x;
    const int flag = owner_flag(pix, maps, nF, W, H, e, total);
    int tot;
    const int pos = block_offsets[blockIdx.x] + block_exclusive_scan(flag, &tot);
    if (e < total) {
        if (e % nF == 0) cam_first[e / nF] = pos;  // first fit index of this view
        if (flag) {
            fit_face[pos] = (int)(e % nF);
            fit_pixel[pos] = pix[e];
            fit_cam[pos] = (int)(e / nF);
        }
    }
    if (e == total - 1) cam_first[ncam] = pos + flag;
}

// Where one gather writes its samples; every pointer may be null (not wanted).  Besides the plain arrays of
// brdfgpu_gather the kernel fills the arrays of the FIT stage directly -- cosphi, the model's cosine, its
// log (brdf_model.cuh: log_or_flag) and the measurements of a colour channel -- so the resident paths need no
// device-to-device hand-over and no separate log pass.  The number of fits is only known on the device when
// the kernel is launched (no synchronisation in between): offsets that depend on it are formed in the kernel.
struct GatherOut {
    double *phi = nullptr, *thetaDash = nullptr, *theta = nullptr;  // fit-major, nimg per fit
    double* I[3] = {nullptr, nullptr, nullptr};                     // (B, G, R)
    double *fit_c = nullptr, *fit_t = nullptr, *fit_L = nullptr;    // fit stage: cosphi, model cosine, log of it
    int fit_reps = 1;              // the three arrays above are written fit_reps times, block r at offset r * nfit * nimg
    double* fit_x[3] = {nullptr, nullptr, nullptr};  // fit stage: measurements of channel ch
    int fit_x_chan_blocks = 0;     // 1: channel ch is offset by ch * nfit * nimg (the three channel blocks of ONE batch)
    int model = 1;                 // cosine the fit stage reads: 1 = thetaDash (Blinn-Phong), 0 = theta (Phong)
    // second fit-stage destination (brdfgpu_gather_resident asked for a global set AND a batch)
    double *fit2_c = nullptr, *fit2_t = nullptr, *fit2_L = nullptr, *fit2_x = nullptr;
    int fit2_channel = 0;
};

// one thread per (fit, LED): the cosines and the three channel intensities of that sample.
// NH / RV: compute cos(theta') / the literal cos(theta) (each costs a normalisation = 1 sqrt + 3 divisions, and
// bit-exactness forbids anything cheaper); SEQ: left-to-right dots in GetCosLN / GetCosNH (BRDFGPU_GATHER_SEQ_DOT).
#ifndef BG_GATHER_MIN_BLOCKS
#define BG_GATHER_MIN_BLOCKS 6  // 40 registers: 138 vs 148 us per 13-view call at 4 (profiles/r02_summary.md)
#endif
template <bool NH, bool RV, bool SEQ>
__global__ void __launch_bounds__(256, BG_GATHER_MIN_BLOCKS) k_gather_samples(const double* __restrict__ FC, const double* __restrict__ FN,
                                                        const double* __restrict__ led, const unsigned char* __restrict__ img,
                                                        const double* __restrict__ cams, const int* __restrict__ fit_face,
                                                        const int* __restrict__ fit_pixel, const int* __restrict__ fit_cam,
                                                        const int* __restrict__ nfit_dev, int nimg, int W, int H, int flip_rows,
                                                        GatherOut o) {
    const long ns = (long)*nfit_dev * nimg;
    if ((long)blockIdx.x * 256 >= ns) return;  // the grid is sized for the capacity (every face of every view mapped)
    // u8 / 255.0 (brdfdata.cpp:956) for all 256 bytes: one division per thread instead of three
    __shared__ double lut[256];
    lut[threadIdx.x] = ddiv((double)threadIdx.x, 255.0);
    __syncthreads();
    const long s = (long)blockIdx.x * 256 + threadIdx.x;
    if (s >= ns) return;
    const long fit = s / nimg;
    const int k = (int)(s - fit * nimg);
    const int face = fit_face[fit];
    const V3 C = load3(FC + 3l * face);
    const V3 N = load3(FN + 3l * face);
    const V3 Lk = load3(led + 3l * k);

    // GetCosLN, brdfdata.cpp:887-893
    const V3 l = normalized(V3{dsub(Lk.x, C.x), dsub(Lk.y, C.y), dsub(Lk.z, C.z)});
    const double cphi = SEQ ? dot3(l, N) : dot3_row(l, N);
    double cnh = 0.0, crv = 0.0;
    if (NH) {  // GetCosNH, brdfdata.cpp:931-937: H = L - 2C + P
        const V3 P = load3(cams + 16l * fit_cam[fit] + 13);
        const V3 h = normalized(V3{dadd(dsub(Lk.x, dmul(2.0, C.x)), P.x), dadd(dsub(Lk.y, dmul(2.0, C.y)), P.y),
                                   dadd(dsub(Lk.z, dmul(2.0, C.z)), P.z)});
        cnh = SEQ ? dot3(h, N) : dot3_row(h, N);
    }
    if (RV) {  // GetCosRV, brdfdata.cpp:829-851, literal (centroid x in all three components, R.P): SURVEY.md Q8
        const V3 ld = normalized(V3{dsub(C.x, Lk.x), dsub(C.x, Lk.y), dsub(C.x, Lk.z)});
        const double sc = dot3(N, ld);  // Block on the left: Eigen's run-time loop, left to right
        const V3 Pv{dmul(sc, N.x), dmul(sc, N.y), dmul(sc, N.z)};
        const V3 R{dsub(ld.x, dmul(2.0, Pv.x)), dsub(ld.y, dmul(2.0, Pv.y)), dsub(ld.z, dmul(2.0, Pv.z))};
        crv = dot3(R, Pv);              // two plain RowVector3d: packet of two, then the third
    }
    if (o.phi) o.phi[s] = cphi;
    if (NH && o.thetaDash) o.thetaDash[s] = cnh;
    if (RV && o.theta) o.theta[s] = crv;
    const double t = o.model == 1 ? cnh : crv;
    if (o.fit_c) {
        const double Lg = log_or_flag(t);
        for (int r = 0; r < o.fit_reps; ++r) {
            o.fit_c[r * ns + s] = cphi;
            o.fit_t[r * ns + s] = t;
            o.fit_L[r * ns + s] = Lg;
        }
        if (o.fit2_c) {
            o.fit2_c[s] = cphi;
            o.fit2_t[s] = t;
            o.fit2_L[s] = Lg;
        }
    }
    // GetIntensities_FromPixel, brdfdata.cpp:955-956: Tsai rows are top-down (no flip); the literal GL projection
    // yields bottom-up rows and reads image row H-1-y
    int pixel = fit_pixel[fit];
    if (flip_rows) pixel = (H - 1 - pixel / W) * W + pixel % W;
    const unsigned char* px = img + ((long)k * H * W + pixel) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double v = lut[px[ch]];
        if (o.I[ch]) o.I[ch][s] = v;
        if (o.fit_x[ch]) o.fit_x[ch][(o.fit_x_chan_blocks ? ch * ns : 0) + s] = v;
        if (o.fit2_x && ch == o.fit2_channel) o.fit2_x[s] = v;
    }
}

// device-side state of one gather
struct GatherDev {
    long nfit = 0;
    long cap_fits = 0;  // ncam * nF: rows every per-fit array is sized for (the fit count is not known at launch time)
    int ncam = 0;
    // Two stream-ordered allocations (cudaMallocAsync on the context's stream, pool kept by brdfgpu_create): a
    // gather call costs microseconds of allocation instead of the 3-4 ms (and, on a cold box, far more) that
    // a dozen cudaMalloc / cudaFree pairs take.
    void *tmp_block = nullptr, *out_block = nullptr;
    int *pix = nullptr, *maps = nullptr, *fit_face = nullptr, *fit_pixel = nullptr, *fit_cam = nullptr, *cam_first = nullptr;
    int* block_sums = nullptr;
    double *cams = nullptr, *phi = nullptr, *thetaDash = nullptr, *theta = nullptr, *I = nullptr;
    std::vector<int> h_cam_first;
    cudaStream_t stream = nullptr;
    void release() {
        if (tmp_block) cudaFreeAsync(tmp_block, stream);
        if (out_block) cudaFreeAsync(out_block, stream);
        *this = GatherDev();
    }
};

// bump allocation inside one block, 256-byte granules
struct Carver {
    char* base;
    size_t off = 0;
    template <class T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + off);
        off += (count * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
};

enum { kWantPhi = 1, kWantNH = 2, kWantRV = 4, kWantI = 8 };

// Queues the whole gather on the context's stream WITHOUT synchronising: projection, ordered compaction, sample
// kernel.  `plain`: which of the plain arrays (GatherDev::phi ...) to allocate and fill; `out`: fit-stage
// destinations prepared by the caller (may be null).  gather_finish() then waits and reads the counts.
static int gather_device(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cams_host, int ncam, int plain,
                         const GatherOut* fit_out, GatherDev* g) {
    if (ncam < 1) {
        set_error(ctx, "gather: need at least one camera");
        return BRDFGPU_LM_ERROR;
    }
    const long total = (long)ncam * sc->nF, npix = (long)ncam * sc->W * sc->H;
    const int nblocks = (int)((total + kScanThreads - 1) / kScanThreads);
    Trace tr("gather_device");
    g->ncam = ncam;
    g->cap_fits = total;
    g->stream = ctx->stream;
    {
        Carver sz{nullptr};
        sz.take<double>(16 * (size_t)ncam); sz.take<int>(total); sz.take<int>(npix); sz.take<int>(total); sz.take<int>(total);
        sz.take<int>(total); sz.take<int>(ncam + 1); sz.take<int>(nblocks + 1);
        BG_CUDA_OK(ctx, cudaMallocAsync(&g->tmp_block, sz.off, ctx->stream));
        Carver cv{static_cast<char*>(g->tmp_block)};
        g->cams = cv.take<double>(16 * (size_t)ncam); g->pix = cv.take<int>(total); g->maps = cv.take<int>(npix);
        g->fit_face = cv.take<int>(total); g->fit_pixel = cv.take<int>(total); g->fit_cam = cv.take<int>(total);
        g->cam_first = cv.take<int>(ncam + 1); g->block_sums = cv.take<int>(nblocks + 1);
    }
    BG_CUDA_OK(ctx, cudaMemcpyAsync(g->cams, cams_host, sizeof(double) * 16 * ncam, cudaMemcpyHostToDevice, ctx->stream));
    tr.mark("allocation + H2D of the cameras");

    k_fill_int<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(g->maps, npix, -1);
    if (sc->gl_on) {
        if (ncam != 1 || (sc->gather_flags & ~BRDFGPU_GATHER_SEQ_DOT)) {
            set_error(ctx, "gather: the literal GL projection maps one view (one set of GL matrices) and takes no projection options");
            return BRDFGPU_LM_ERROR;
        }
        GlProjection gl;
        for (int i = 0; i < 16; ++i) { gl.mv[i] = sc->gl_mv[i]; gl.proj[i] = sc->gl_proj[i]; }
        gl.vx = (double)sc->gl_viewport[0]; gl.vy = (double)sc->gl_viewport[1];
        gl.vw = (double)sc->gl_viewport[2]; gl.vh = (double)sc->gl_viewport[3];
        k_project_gl<<<(unsigned)((sc->nF + 255) / 256), 256, 0, ctx->stream>>>(sc->FC, sc->nF, gl, sc->W, sc->H, g->pix, g->maps);
    } else if ((sc->gather_flags & ~BRDFGPU_GATHER_SEQ_DOT) == 0) {
        k_project<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(sc->FC, sc->nF, g->cams, ncam, sc->W, sc->H, g->pix,
                                                                            g->maps);
    } else {
        const int flags = sc->gather_flags;
        double* d_kappa = nullptr;
        unsigned long long *d_zbits = nullptr, *d_depth = nullptr;
        if (flags & BRDFGPU_GATHER_KAPPA1) {
            if ((int)sc->kappa1.size() != ncam) {
                set_error(ctx, "gather: BRDFGPU_GATHER_KAPPA1 needs one kappa1 per camera (brdfgpu_scene_set_gather_options)");
                return BRDFGPU_LM_ERROR;
            }
            BG_CUDA_OK(ctx, cudaMalloc(&d_kappa, sizeof(double) * ncam));
            BG_CUDA_OK(ctx, cudaMemcpyAsync(d_kappa, sc->kappa1.data(), sizeof(double) * ncam, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (flags & BRDFGPU_GATHER_DEPTH_TEST) {
            BG_CUDA_OK(ctx, cudaMalloc(&d_zbits, sizeof(unsigned long long) * total));
            BG_CUDA_OK(ctx, cudaMalloc(&d_depth, sizeof(unsigned long long) * npix));
            k_fill_u64<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_depth, npix, ~0ull);
            ++ctx->launches;
        }
        const unsigned blocks = (unsigned)((total + 255) / 256);
        k_project_opts<<<blocks, 256, 0, ctx->stream>>>(sc->FC, sc->FN, sc->nF, g->cams, d_kappa, flags, ncam, sc->W, sc->H, 0,
                                                        g->pix, d_zbits, d_depth, g->maps);
        if (flags & BRDFGPU_GATHER_DEPTH_TEST) {
            k_project_opts<<<blocks, 256, 0, ctx->stream>>>(sc->FC, sc->FN, sc->nF, g->cams, d_kappa, flags, ncam, sc->W, sc->H, 1,
                                                            g->pix, d_zbits, d_depth, g->maps);
            ++ctx->launches;
        }
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_kappa); cudaFree(d_zbits); cudaFree(d_depth);
        BG_CUDA_OK(ctx, e);
    }
    k_owner_count<<<nblocks, kScanThreads, 0, ctx->stream>>>(g->pix, g->maps, sc->nF, sc->W, sc->H, total, g->block_sums);
    k_scan_block_sums<<<1, kScanThreads, 0, ctx->stream>>>(g->block_sums, nblocks);
    k_owner_scatter<<<nblocks, kScanThreads, 0, ctx->stream>>>(g->pix, g->maps, sc->nF, sc->W, sc->H, total, g->block_sums,
                                                               g->fit_face, g->fit_pixel, g->fit_cam, g->cam_first, ncam);
    ctx->launches += 5;
    BG_CUDA_OK(ctx, cudaGetLastError());
    // the per-view fit counts travel back through pinned memory behind everything else: no wait here
    g->h_cam_first.assign(ncam + 1, 0);
    int* staging = (ncam + 1 <= kCountStagingInts) ? ctx->h_counts : g->h_cam_first.data();
    BG_CUDA_OK(ctx, cudaMemcpyAsync(staging, g->cam_first, sizeof(int) * (ncam + 1), cudaMemcpyDeviceToHost, ctx->stream));
    tr.mark("project + compaction kernels queued");
    if (!plain && !fit_out) return 0;

    GatherOut o;
    if (fit_out) o = *fit_out;
    const long cap_ns = total * sc->nimg;
    if (plain) {
        const int narr = ((plain & kWantPhi) ? 1 : 0) + ((plain & kWantNH) ? 1 : 0) + ((plain & kWantRV) ? 1 : 0) +
                         ((plain & kWantI) ? 3 : 0);
        BG_CUDA_OK(ctx, cudaMallocAsync(&g->out_block, sizeof(double) * (size_t)narr * cap_ns, ctx->stream));
        double* q = static_cast<double*>(g->out_block);
        if (plain & kWantPhi) { g->phi = o.phi = q; q += cap_ns; }
        if (plain & kWantNH) { g->thetaDash = o.thetaDash = q; q += cap_ns; }
        if (plain & kWantRV) { g->theta = o.theta = q; q += cap_ns; }
        if (plain & kWantI) {
            g->I = q;
            for (int ch = 0; ch < 3; ++ch) o.I[ch] = q + (size_t)ch * cap_ns;
        }
    }
    const bool nh = (plain & kWantNH) || (o.fit_c && o.model == 1), rv = (plain & kWantRV) || (o.fit_c && o.model == 0);
    const bool seq = (sc->gather_flags & BRDFGPU_GATHER_SEQ_DOT) != 0;
    const unsigned blocks = (unsigned)((cap_ns + 255) / 256);
    const int* nfit_dev = g->cam_first + ncam;
#define BG_GATHER_LAUNCH(NH_, RV_, SEQ_)                                                                                    \
    k_gather_samples<NH_, RV_, SEQ_><<<blocks, 256, 0, ctx->stream>>>(sc->FC, sc->FN, sc->led, sc->img, g->cams, g->fit_face, \
                                                                      g->fit_pixel, g->fit_cam, nfit_dev, sc->nimg, sc->W, sc->H, \
                                                                      sc->gl_on ? 1 : 0, o)
    if (seq) {
        if (nh && rv) BG_GATHER_LAUNCH(true, true, true);
        else if (nh) BG_GATHER_LAUNCH(true, false, true);
        else if (rv) BG_GATHER_LAUNCH(false, true, true);
        else BG_GATHER_LAUNCH(false, false, true);
    } else {
        if (nh && rv) BG_GATHER_LAUNCH(true, true, false);
        else if (nh) BG_GATHER_LAUNCH(true, false, false);
        else if (rv) BG_GATHER_LAUNCH(false, true, false);
        else BG_GATHER_LAUNCH(false, false, false);
    }
#undef BG_GATHER_LAUNCH
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    tr.mark("sample kernel queued");
    return 0;
}

// wait for the queued gather and read the fit counts
static int gather_finish(brdfgpu_ctx* ctx, GatherDev* g) {
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (g->ncam + 1 <= kCountStagingInts)
        for (int v = 0; v <= g->ncam; ++v) g->h_cam_first[v] = ctx->h_counts[v];
    g->nfit = g->h_cam_first[g->ncam];
    return 0;
}

}  // namespace brdfgpu

using namespace brdfgpu;

static brdfgpu_ctx* ctx_or_default(brdfgpu_ctx* ctx) { return ctx ? ctx : default_ctx(); }

// brdfdata.cpp:695-755: 4 x 4 serpentine LED grid, x fixed
static void reference_led_table(double* led) {
    const double x = 303.5, min_y = -157.1, max_y = -2.3, min_z = 555.3, max_z = 645.8;
    const double y_step = (max_y - min_y) / 3, z_step = (max_z - min_z) / 3;
    const double ys[4] = {max_y, max_y - y_step, min_y + y_step, min_y};
    const double zs[4] = {min_z, min_z + z_step, max_z - z_step, max_z};
    for (int i = 0; i < 16; ++i) {
        const int row = i / 4, col = i % 4;
        led[i * 3 + 0] = x;
        led[i * 3 + 1] = (row % 2 == 0) ? ys[col] : ys[3 - col];
        led[i * 3 + 2] = zs[row];
    }
}

extern "C" void brdfgpu_led_table(double* led16x3) { reference_led_table(led16x3); }

extern "C" int brdfgpu_scene_create(brdfgpu_ctx* ctx, const double* V, int nV, const int* F, int nF,
                                    const unsigned char* const* images, int nimg, int W, int H, const unsigned char* dark,
                                    const double* led, brdfgpu_scene** out) {
    ctx = ctx_or_default(ctx);
    if (!ctx) return BRDFGPU_LM_ERROR;
    if (!V || !F || nV < 1 || nF < 1 || nimg < 1 || W < 1 || H < 1 || !images || !out) {
        set_error(ctx, "scene_create: bad arguments");
        return BRDFGPU_LM_ERROR;
    }
    if (!led && nimg != 16) {
        set_error(ctx, "scene_create: the built-in LED table has 16 entries; pass led for other image counts");
        return BRDFGPU_LM_ERROR;
    }
    for (long i = 0; i < 3l * nF; ++i)
        if (F[i] < 0 || F[i] >= nV) {
            set_error(ctx, "scene_create: face index out of range");
            return BRDFGPU_LM_ERROR;
        }
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    brdfgpu_scene* sc = new brdfgpu_scene;
    sc->nV = nV; sc->nF = nF; sc->nimg = nimg; sc->W = W; sc->H = H;
    const long per = (long)W * H * 3;
    double table[48];
    if (!led) reference_led_table(table);
    unsigned char* d_dark = nullptr;
    cudaError_t e = cudaMalloc(&sc->V, sizeof(double) * 3 * nV);
    if (e == cudaSuccess) e = cudaMalloc(&sc->F, sizeof(int) * 3 * nF);
    if (e == cudaSuccess) e = cudaMalloc(&sc->FN, sizeof(double) * 3 * nF);
    if (e == cudaSuccess) e = cudaMalloc(&sc->FC, sizeof(double) * 3 * nF);
    if (e == cudaSuccess) e = cudaMalloc(&sc->img, (size_t)per * nimg);
    if (e == cudaSuccess) e = cudaMalloc(&sc->led, sizeof(double) * 3 * nimg);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sc->V, V, sizeof(double) * 3 * nV, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sc->F, F, sizeof(int) * 3 * nF, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(sc->led, led ? led : table, sizeof(double) * 3 * nimg, cudaMemcpyHostToDevice, ctx->stream);
    for (int k = 0; k < nimg && e == cudaSuccess; ++k)
        e = cudaMemcpyAsync(sc->img + (size_t)per * k, images[k], (size_t)per, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && dark) {
        e = cudaMalloc(&d_dark, (size_t)per);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dark, dark, (size_t)per, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            k_subtract_ambient<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(sc->img, d_dark, per, nimg);
            ++ctx->launches;
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) {
        k_face_geometry<<<(nF + 255) / 256, 256, 0, ctx->stream>>>(sc->V, sc->F, nF, sc->FN, sc->FC);
        ++ctx->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_dark);
    if (e != cudaSuccess) {
        brdfgpu_scene_free(ctx, sc);
        set_error(ctx, std::string("scene_create: ") + cudaGetErrorString(e));
        return BRDFGPU_LM_ERROR;
    }
    *out = sc;
    return 0;
}

extern "C" void brdfgpu_scene_free(brdfgpu_ctx* ctx, brdfgpu_scene* sc) {
    (void)ctx;
    if (!sc) return;
    cudaFree(sc->V); cudaFree(sc->F); cudaFree(sc->FN); cudaFree(sc->FC); cudaFree(sc->img); cudaFree(sc->led);
    delete sc;
}

extern "C" int brdfgpu_shade_faces(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* eye, const double* center, int model,
                                   int single, const double* brdf, int literal_cosln, double* bgr_out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !eye || !center || !brdf || !bgr_out || (model != 0 && model != 1)) return BRDFGPU_LM_ERROR;
    const size_t nb = sizeof(double) * (single ? 9 : 9 * (size_t)sc->nF), no = sizeof(double) * 3 * (size_t)sc->nF;
    double *d_brdf = nullptr, *d_out = nullptr;
    BG_CUDA_OK(ctx, cudaMalloc(&d_brdf, nb));
    cudaError_t e = cudaMalloc(&d_out, no);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_brdf, brdf, nb, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const V3 ey{eye[0], eye[1], eye[2]}, ce{center[0], center[1], center[2]};
        const int blocks = (sc->nF + 255) / 256;
        if (literal_cosln) k_shade_faces<true><<<blocks, 256, 0, ctx->stream>>>(sc->FC, sc->FN, sc->nF, ey, ce, model, single, d_brdf, d_out);
        else k_shade_faces<false><<<blocks, 256, 0, ctx->stream>>>(sc->FC, sc->FN, sc->nF, ey, ce, model, single, d_brdf, d_out);
        ++ctx->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(bgr_out, d_out, no, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_brdf);
    cudaFree(d_out);
    if (e != cudaSuccess) {
        set_error(ctx, std::string("shade_faces: ") + cudaGetErrorString(e));
        return BRDFGPU_LM_ERROR;
    }
    return 0;
}

extern "C" int brdfgpu_scene_set_gather_options(brdfgpu_ctx* ctx, brdfgpu_scene* sc, int flags, const double* kappa1, int ncam) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc) return BRDFGPU_LM_ERROR;
    const int known = BRDFGPU_GATHER_DEPTH_TEST | BRDFGPU_GATHER_CULL_BACKFACES | BRDFGPU_GATHER_KAPPA1 | BRDFGPU_GATHER_SEQ_DOT;
    if (flags & ~known) {
        set_error(ctx, "scene_set_gather_options: unknown flag");
        return BRDFGPU_LM_ERROR;
    }
    if ((flags & BRDFGPU_GATHER_KAPPA1) && (!kappa1 || ncam < 1)) {
        set_error(ctx, "scene_set_gather_options: BRDFGPU_GATHER_KAPPA1 needs kappa1[ncam]");
        return BRDFGPU_LM_ERROR;
    }
    sc->gather_flags = flags;
    sc->kappa1.clear();
    if (flags & BRDFGPU_GATHER_KAPPA1) sc->kappa1.assign(kappa1, kappa1 + ncam);
    return 0;
}

extern "C" int brdfgpu_scene_set_gl_projection(brdfgpu_ctx* ctx, brdfgpu_scene* sc, const double* model_view16,
                                               const double* projection16, const int* viewport4) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc) return BRDFGPU_LM_ERROR;
    if (!model_view16 || !projection16 || !viewport4) {  // back to the Tsai camera
        sc->gl_on = false;
        return 0;
    }
    for (int i = 0; i < 16; ++i) { sc->gl_mv[i] = model_view16[i]; sc->gl_proj[i] = projection16[i]; }
    for (int i = 0; i < 4; ++i) sc->gl_viewport[i] = viewport4[i];
    sc->gl_on = true;
    return 0;
}

// MakeFrustum + glFrustum and gluLookAt(0,0,50, 0,0,0, 0,1,0) as the reference's Display_ sets them up
// (glutcallbacks.cpp:626-642, 672-689), entries rounded to the float32 GL stores them in
extern "C" void brdfgpu_reference_gl_matrices(double cx, double cy, int window_width, int window_height, double* mv,
                                              double* proj) {
    const double DEG2RAD = 3.14159265 / 180, fov = 78, fovV = 49, front = 1.0, back = 1000.0;  // glutcallbacks.cpp:57-62, 628
    const double aspect = fov / fovV;
    const double tangent = tan(fovV / 2 * DEG2RAD), height = front * tangent, width = height * aspect;
    const double offset_y = 2.0 * (window_height / 2.0 - cy) / window_height, offset_x = 2.0 * (window_width / 2.0 - cx) / window_width;
    const double l = -width + offset_x, r = width + offset_x, b = -height - offset_y, t = height - offset_y;
    for (int i = 0; i < 16; ++i) { mv[i] = 0.0; proj[i] = 0.0; }
    proj[0] = 2 * front / (r - l); proj[5] = 2 * front / (t - b);
    proj[8] = (r + l) / (r - l); proj[9] = (t + b) / (t - b); proj[10] = -(back + front) / (back - front); proj[11] = -1.0;
    proj[14] = -2 * back * front / (back - front);
    mv[0] = mv[5] = mv[10] = mv[15] = 1.0;
    mv[14] = -50.0;
    for (int i = 0; i < 16; ++i) { mv[i] = (double)(float)mv[i]; proj[i] = (double)(float)proj[i]; }
}

extern "C" int brdfgpu_scene_dims(const brdfgpu_scene* sc, int* dims5) {
    if (!sc || !dims5) return BRDFGPU_LM_ERROR;
    dims5[0] = sc->nV; dims5[1] = sc->nF; dims5[2] = sc->nimg; dims5[3] = sc->W; dims5[4] = sc->H;
    return 0;
}

extern "C" int brdfgpu_scene_face_normals(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, double* FN) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaMemcpyAsync(FN, sc->FN, sizeof(double) * 3 * sc->nF, cudaMemcpyDeviceToHost, ctx->stream));
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int brdfgpu_scene_image(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, int k, unsigned char* out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || k < 0 || k >= sc->nimg) return BRDFGPU_LM_ERROR;
    const size_t per = (size_t)sc->W * sc->H * 3;
    BG_CUDA_OK(ctx, cudaMemcpyAsync(out, sc->img + per * k, per, cudaMemcpyDeviceToHost, ctx->stream));
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int brdfgpu_calc_pixel2surface(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cam, int* map) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !cam || !map) return BRDFGPU_LM_ERROR;
    GatherDev g;
    int rc = gather_device(ctx, sc, cam, 1, 0, nullptr, &g);
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(map, g.maps, sizeof(int) * (size_t)sc->W * sc->H, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_error(ctx, std::string("calc_pixel2surface: ") + cudaGetErrorString(e));
            rc = BRDFGPU_LM_ERROR;
        }
    }
    g.release();
    return rc;
}

extern "C" long brdfgpu_gather(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cams, int ncam, long capacity,
                               int* maps, long* nfit_cam, int* fit_face, int* fit_pixel, double* phi, double* thetaDash,
                               double* theta, double* I) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !cams) return BRDFGPU_LM_ERROR;
    GatherDev g;
    // only what the caller asked for is computed (each cosine costs a normalisation per sample)
    const int plain = (phi ? kWantPhi : 0) | (thetaDash ? kWantNH : 0) | (theta ? kWantRV : 0) | (I ? kWantI : 0);
    long rc = gather_device(ctx, sc, cams, ncam, plain, nullptr, &g);
    if (rc == 0) rc = gather_finish(ctx, &g);
    if (rc == 0 && g.nfit > capacity && (fit_face || fit_pixel || phi || thetaDash || theta || I)) {
        set_error(ctx, "gather: capacity too small for the number of fits");
        rc = BRDFGPU_LM_ERROR;
    }
    if (rc == 0) {
        const long ns = g.nfit * sc->nimg, cap_ns = g.cap_fits * sc->nimg;
        cudaError_t e = cudaSuccess;
        auto down = [&](void* dst, const void* src, size_t bytes) {
            if (dst && bytes && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        };
        down(maps, g.maps, sizeof(int) * (size_t)ncam * sc->W * sc->H);
        down(fit_face, g.fit_face, sizeof(int) * g.nfit);
        down(fit_pixel, g.fit_pixel, sizeof(int) * g.nfit);
        down(phi, g.phi, sizeof(double) * ns);
        down(thetaDash, g.thetaDash, sizeof(double) * ns);
        down(theta, g.theta, sizeof(double) * ns);
        if (I)
            for (int ch = 0; ch < 3; ++ch) down(I + (size_t)ch * capacity * sc->nimg, g.I + (size_t)ch * cap_ns, sizeof(double) * ns);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_error(ctx, std::string("gather: ") + cudaGetErrorString(e));
            rc = BRDFGPU_LM_ERROR;
        } else {
            if (nfit_cam)
                for (int v = 0; v < ncam; ++v) nfit_cam[v] = g.h_cam_first[v + 1] - g.h_cam_first[v];
            rc = g.nfit;
        }
    }
    g.release();
    return rc;
}

// The sample kernel writes straight into the arrays of the fit stage, which are sized for the capacity of the
// gather (every face of every view mapped) because the fit count is only known after the one synchronisation
// at the end; the handles then get their true sizes.
extern "C" int brdfgpu_gather_resident(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cams, int ncam, int model,
                                       int channel, brdfgpu_samples** global_out, brdfgpu_batch** batch_out,
                                       long* nfit_out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !cams || ncam < 1 || channel < 0 || channel > 2 || (model != 0 && model != 1)) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const long cap_fits = (long)ncam * sc->nF;
    brdfgpu_samples* s = nullptr;
    brdfgpu_batch* b = nullptr;
    if (global_out) *global_out = nullptr;
    if (batch_out) *batch_out = nullptr;
    if (global_out && samples_alloc(ctx, cap_fits * sc->nimg, model, &s) != 0) return BRDFGPU_LM_ERROR;
    if (batch_out && batch_alloc(ctx, cap_fits, sc->nimg, model, &b) != 0) {
        brdfgpu_samples_free(ctx, s);
        return BRDFGPU_LM_ERROR;
    }
    GatherOut o;
    o.model = model;
    if (s) {
        o.fit_c = s->c; o.fit_t = s->traw; o.fit_L = s->L; o.fit_x[channel] = s->x;
        if (b) { o.fit2_c = b->c; o.fit2_t = b->traw; o.fit2_L = b->L; o.fit2_x = b->x; o.fit2_channel = channel; }
    } else if (b) {
        o.fit_c = b->c; o.fit_t = b->traw; o.fit_L = b->L; o.fit_x[channel] = b->x;
    }
    GatherDev g;
    int rc = gather_device(ctx, sc, cams, ncam, 0, (s || b) ? &o : nullptr, &g);
    if (rc == 0) rc = gather_finish(ctx, &g);
    const long nfit = g.nfit;
    g.release();
    if (rc != 0) {
        brdfgpu_samples_free(ctx, s);
        brdfgpu_batch_free(ctx, b);
        return rc;
    }
    if (s) s->n = nfit * sc->nimg;
    if (b) b->nfit = nfit;
    if (nfit_out) *nfit_out = nfit;
    if (global_out) *global_out = s;
    if (batch_out) *batch_out = b;
    return 0;
}

// CBRDFdata::CalcBRDFEquation (brdfdata.cpp:1188-1227): ONE gather, then the fits of all three colour
// channels as one batch of 3 x nfit problems (fit ch * nfit + f = face f, channel ch) in a single launch.  The
// gather kernel writes the three channel blocks of the batch itself.
extern "C" long brdfgpu_calc_brdf_equation(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cam, int model,
                                           double* brdf_surfaces) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !cam || !brdf_surfaces || (model != 0 && model != 1)) return BRDFGPU_LM_ERROR;
    static const double p0[3] = {0.5, 1.0, 1.0}, lb[3] = {0, 0, 0}, ub[3] = {100, 100, 100};
    static const double opts[5] = {1E-03, 1E-15, 1E-15, 1E-20, 1E-06};
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    brdfgpu_batch* b = nullptr;
    if (batch_alloc(ctx, 3l * sc->nF, sc->nimg, model, &b) != 0) return BRDFGPU_LM_ERROR;
    GatherOut o;
    o.model = model;
    o.fit_c = b->c; o.fit_t = b->traw; o.fit_L = b->L; o.fit_reps = 3;
    for (int ch = 0; ch < 3; ++ch) o.fit_x[ch] = b->x;
    o.fit_x_chan_blocks = 1;
    GatherDev g;
    int rc = gather_device(ctx, sc, cam, 1, 0, &o, &g);
    if (rc == 0) rc = gather_finish(ctx, &g);
    const long nfit = g.nfit;
    std::vector<int> faces(nfit > 0 ? nfit : 1);
    std::vector<double> p(9 * (size_t)(nfit > 0 ? nfit : 1));
    if (rc == 0 && nfit > 0) {
        b->nfit = 3 * nfit;
        cudaError_t e = cudaMemcpyAsync(faces.data(), g.fit_face, sizeof(int) * nfit, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) {
            set_error(ctx, std::string("calc_brdf_equation: ") + cudaGetErrorString(e));
            rc = BRDFGPU_LM_ERROR;
        }
        // (levmar-exact: every per-face result equals the reference's, brdfdata.cpp:1119)
        if (rc == 0) rc = brdfgpu_batch_fit(ctx, b, p0, lb, ub, 100, opts, sc->nimg <= 128 ? BRDFGPU_JAC_FD_EXACT : BRDFGPU_JAC_FD);
        if (rc == 0) rc = brdfgpu_batch_results(ctx, b, p.data(), nullptr, nullptr);
    }
    brdfgpu_batch_free(ctx, b);
    g.release();
    if (rc != 0) return BRDFGPU_LM_ERROR;
    for (int ch = 0; ch < 3; ++ch)          // brdfdata.cpp:1205: B, G, R
        for (long f = 0; f < nfit; ++f)     // SaveValuesToSurface, brdfdata.cpp:368-377
            for (int j = 0; j < 3; ++j) brdf_surfaces[((size_t)faces[f] * 3 + ch) * 3 + j] = p[3 * ((size_t)ch * nfit + f) + j];
    return nfit;
}

// CBRDFdata::CalcBRDFEquation_SingleBRDF (brdfdata.cpp:1138-1186): one gather, one global fit per channel.  The
// three channel sets share cosphi / the model cosine / its log and differ in the measurements only.
extern "C" long brdfgpu_calc_brdf_equation_single(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cam, int model,
                                                  double* single_brdf, double* info, int* ret) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !cam || !single_brdf || (model != 0 && model != 1)) return BRDFGPU_LM_ERROR;
    static const double lb[3] = {0, 0, 0}, ub[3] = {100, 100, 100};
    static const double opts[5] = {1E-03, 1E-15, 1E-10, 1E-50, 1.0};
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Trace tr("calc_brdf_equation_single");
    const size_t stride = (((size_t)sc->nF * sc->nimg) + 15) & ~(size_t)15;
    double* block = nullptr;
    BG_CUDA_OK(ctx, cudaMallocAsync(&block, sizeof(double) * 6 * stride, ctx->stream));
    GatherOut o;
    o.model = model;
    o.fit_c = block; o.fit_t = block + stride; o.fit_L = block + 2 * stride;
    for (int ch = 0; ch < 3; ++ch) o.fit_x[ch] = block + (3 + ch) * stride;
    GatherDev g;
    int rc = gather_device(ctx, sc, cam, 1, 0, &o, &g);
    if (rc == 0) rc = gather_finish(ctx, &g);
    const long nfit = g.nfit;
    g.release();
    tr.mark("gather");
    for (int ch = 0; ch < 3 && rc == 0; ++ch) {
        brdfgpu_samples view;  // not an owner: block stays null
        view.n = nfit * sc->nimg; view.capacity = (long)stride; view.model = model; view.stream = ctx->stream;
        view.c = o.fit_c; view.traw = o.fit_t; view.L = o.fit_L; view.x = o.fit_x[ch];
        double p[3] = {0.0, 0.0, 0.0}, inf[10] = {0};
        const int r = brdfgpu_fit_global(ctx, &view, p, 3, lb, ub, nullptr, 2000, opts, inf, nullptr, BRDFGPU_DRIVE_PERSISTENT,
                                         BRDFGPU_JAC_FD);
        tr.mark("global fit");
        for (int j = 0; j < 3; ++j) single_brdf[ch * 3 + j] = p[j];
        if (info)
            for (int j = 0; j < 10; ++j) info[ch * 10 + j] = inf[j];
        if (ret) ret[ch] = r;
    }
    cudaFreeAsync(block, ctx->stream);
    return rc == 0 ? nfit : (long)BRDFGPU_LM_ERROR;
}

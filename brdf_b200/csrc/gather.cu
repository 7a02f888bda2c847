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
// contraction can occur whatever flags this file is compiled with; 3-term sums associate left to
// right; double -> int truncates.  "Last face wins" is atomicMax(face id): the reference walks
// faces in ascending order, so the survivor is the largest id.
//
// Work decomposition: one thread per (view, face) for the projection, one thread per
// (fit, LED) for the sample rows, so the 16 LEDs of a fit write 16 consecutive doubles.
#include <vector>

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
    unsigned char* img = nullptr;  // nimg x H x W x 3 (BGR), ambient already removed
    double* led = nullptr;         // nimg x 3
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

__global__ void k_face_normals(const double* __restrict__ V, const int* __restrict__ F, int nF, double* __restrict__ FN) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nF) return;
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
__global__ void k_shade_faces(const double* __restrict__ V, const int* __restrict__ F, const double* __restrict__ FN, int nF,
                              V3 eye, V3 center, int model, int single, const double* __restrict__ brdf,
                              double* __restrict__ bgr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nF) return;
    const V3 c = centroid_of(V, F, i);
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
__global__ void k_project_opts(const double* __restrict__ V, const int* __restrict__ F, const double* __restrict__ FN, int nF,
                               const double* __restrict__ cams, const double* __restrict__ kappa1, int flags, int ncam, int W,
                               int H, int pass, int* __restrict__ pix, unsigned long long* __restrict__ zbits,
                               unsigned long long* __restrict__ depth, int* __restrict__ maps) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)ncam * nF) return;
    const int v = (int)(e / nF), face = (int)(e % nF);
    if (pass == 0) {
        const Camera cam = load_camera(cams + 16l * v);
        double z = 0.0;
        const int px = project_opts(centroid_of(V, F, face), load3(FN + 3l * face), cam, kappa1 ? kappa1[v] : 0.0, flags, W, H, &z);
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
__global__ void k_project(const double* __restrict__ V, const int* __restrict__ F, int nF, const double* __restrict__ cams,
                          int ncam, int W, int H, int* __restrict__ pix /*ncam*nF*/, int* __restrict__ maps /*ncam*H*W*/) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)ncam * nF) return;
    const int v = (int)(e / nF), face = (int)(e % nF);
    const Camera cam = load_camera(cams + 16l * v);
    const int px = project_tsai(centroid_of(V, F, face), cam, W, H);
    pix[e] = px;
    if (px >= 0) atomicMax(maps + (long)v * W * H + px, face);
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

// one thread per (fit, LED): the three cosines and the three channel intensities of that sample
__global__ void k_gather_samples(const double* __restrict__ V, const int* __restrict__ F, const double* __restrict__ FN,
                                 const double* __restrict__ led, const unsigned char* __restrict__ img,
                                 const double* __restrict__ cams, const int* __restrict__ fit_face,
                                 const int* __restrict__ fit_pixel, const int* __restrict__ fit_cam, long nfit, int nimg,
                                 int W, int H, long chan_stride, double* __restrict__ phi, double* __restrict__ thetaDash,
                                 double* __restrict__ theta, double* __restrict__ I) {
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nfit * nimg) return;
    const long fit = s / nimg;
    const int k = (int)(s % nimg);
    const int face = fit_face[fit];
    const V3 C = centroid_of(V, F, face);
    const V3 N = load3(FN + 3l * face);
    const V3 Lk = load3(led + 3l * k);
    const V3 P = load3(cams + 16l * fit_cam[fit] + 13);

    // GetCosLN, brdfdata.cpp:887-893
    const V3 l = normalized(V3{dsub(Lk.x, C.x), dsub(Lk.y, C.y), dsub(Lk.z, C.z)});
    phi[s] = dot3(l, N);
    // GetCosNH, brdfdata.cpp:931-937: H = L - 2C + P
    const V3 h = normalized(V3{dadd(dsub(Lk.x, dmul(2.0, C.x)), P.x), dadd(dsub(Lk.y, dmul(2.0, C.y)), P.y),
                               dadd(dsub(Lk.z, dmul(2.0, C.z)), P.z)});
    thetaDash[s] = dot3(h, N);
    // GetCosRV, brdfdata.cpp:829-851, literal (centroid x in all three components, R.P): SURVEY.md Q8
    const V3 ld = normalized(V3{dsub(C.x, Lk.x), dsub(C.x, Lk.y), dsub(C.x, Lk.z)});
    const double sc = dot3(N, ld);
    const V3 Pv{dmul(sc, N.x), dmul(sc, N.y), dmul(sc, N.z)};
    const V3 R{dsub(ld.x, dmul(2.0, Pv.x)), dsub(ld.y, dmul(2.0, Pv.y)), dsub(ld.z, dmul(2.0, Pv.z))};
    theta[s] = dot3(R, Pv);
    // GetIntensities_FromPixel, brdfdata.cpp:955-956 (Tsai rows are top-down: no flip)
    const unsigned char* px = img + ((long)k * H * W + fit_pixel[fit]) * 3;
    I[s] = ddiv((double)px[0], 255.0);
    I[chan_stride + s] = ddiv((double)px[1], 255.0);
    I[2 * chan_stride + s] = ddiv((double)px[2], 255.0);
}

__global__ void k_log_flag(const double* __restrict__ t, double* __restrict__ L, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        L[i] = (t[i] >= 0.0) ? log(t[i]) : __longlong_as_double(0x7ff8000000000000LL);
}

// device-side result of one gather
struct GatherDev {
    long nfit = 0;
    int ncam = 0;
    int *pix = nullptr, *maps = nullptr, *fit_face = nullptr, *fit_pixel = nullptr, *fit_cam = nullptr, *cam_first = nullptr;
    int* block_sums = nullptr;
    double *cams = nullptr, *phi = nullptr, *thetaDash = nullptr, *theta = nullptr, *I = nullptr;
    std::vector<int> h_cam_first;
    // Buffers come from the stream-ordered allocator (cudaMallocAsync on the context's stream, pool kept by
    // brdfgpu_create): a gather call costs microseconds of allocation instead of the 3-4 ms (and, on a cold
    // box, far more) that 12 cudaMalloc / cudaFree pairs take.
    cudaStream_t stream = nullptr;
    void release() {
        void* all[] = {pix, maps, fit_face, fit_pixel, fit_cam, cam_first, block_sums, cams, phi, thetaDash, theta, I};
        for (void* q : all)
            if (q) cudaFreeAsync(q, stream);
        *this = GatherDev();
    }
};

static int gather_device(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cams_host, int ncam, bool want_samples,
                         GatherDev* g) {
    if (ncam < 1) {
        set_error(ctx, "gather: need at least one camera");
        return BRDFGPU_LM_ERROR;
    }
    const long total = (long)ncam * sc->nF, npix = (long)ncam * sc->W * sc->H;
    const int nblocks = (int)((total + kScanThreads - 1) / kScanThreads);
    Trace tr("gather_device");
    g->ncam = ncam;
    g->stream = ctx->stream;
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->cams, sizeof(double) * 16 * ncam, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->pix, sizeof(int) * total, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->maps, sizeof(int) * npix, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->fit_face, sizeof(int) * total, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->fit_pixel, sizeof(int) * total, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->fit_cam, sizeof(int) * total, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->cam_first, sizeof(int) * (ncam + 1), ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->block_sums, sizeof(int) * (nblocks + 1), ctx->stream));
    BG_CUDA_OK(ctx, cudaMemcpyAsync(g->cams, cams_host, sizeof(double) * 16 * ncam, cudaMemcpyHostToDevice, ctx->stream));
    tr.mark("8 cudaMalloc + H2D of the cameras");

    k_fill_int<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(g->maps, npix, -1);
    if (sc->gather_flags == 0) {
        k_project<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(sc->V, sc->F, sc->nF, g->cams, ncam, sc->W, sc->H,
                                                                            g->pix, g->maps);
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
        k_project_opts<<<blocks, 256, 0, ctx->stream>>>(sc->V, sc->F, sc->FN, sc->nF, g->cams, d_kappa, flags, ncam, sc->W, sc->H, 0,
                                                        g->pix, d_zbits, d_depth, g->maps);
        if (flags & BRDFGPU_GATHER_DEPTH_TEST) {
            k_project_opts<<<blocks, 256, 0, ctx->stream>>>(sc->V, sc->F, sc->FN, sc->nF, g->cams, d_kappa, flags, ncam, sc->W, sc->H, 1,
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
    g->h_cam_first.resize(ncam + 1);
    BG_CUDA_OK(ctx, cudaMemcpyAsync(g->h_cam_first.data(), g->cam_first, sizeof(int) * (ncam + 1), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    tr.mark("project + compaction kernels, sync");
    g->nfit = g->h_cam_first[ncam];
    if (!want_samples || g->nfit == 0) return 0;

    const long ns = g->nfit * sc->nimg;
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->phi, sizeof(double) * ns, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->thetaDash, sizeof(double) * ns, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->theta, sizeof(double) * ns, ctx->stream));
    BG_CUDA_OK(ctx, cudaMallocAsync(&g->I, sizeof(double) * 3 * ns, ctx->stream));
    k_gather_samples<<<(unsigned)((ns + 255) / 256), 256, 0, ctx->stream>>>(
        sc->V, sc->F, sc->FN, sc->led, sc->img, g->cams, g->fit_face, g->fit_pixel, g->fit_cam, g->nfit, sc->nimg, sc->W,
        sc->H, ns, g->phi, g->thetaDash, g->theta, g->I);
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    tr.mark("4 cudaMalloc + sample kernel launch");
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
        k_face_normals<<<(nF + 255) / 256, 256, 0, ctx->stream>>>(sc->V, sc->F, nF, sc->FN);
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
    cudaFree(sc->V); cudaFree(sc->F); cudaFree(sc->FN); cudaFree(sc->img); cudaFree(sc->led);
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
        if (literal_cosln) k_shade_faces<true><<<blocks, 256, 0, ctx->stream>>>(sc->V, sc->F, sc->FN, sc->nF, ey, ce, model, single, d_brdf, d_out);
        else k_shade_faces<false><<<blocks, 256, 0, ctx->stream>>>(sc->V, sc->F, sc->FN, sc->nF, ey, ce, model, single, d_brdf, d_out);
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
    const int known = BRDFGPU_GATHER_DEPTH_TEST | BRDFGPU_GATHER_CULL_BACKFACES | BRDFGPU_GATHER_KAPPA1;
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
    int rc = gather_device(ctx, sc, cam, 1, false, &g);
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
    long rc = gather_device(ctx, sc, cams, ncam, true, &g);
    if (rc == 0 && g.nfit > capacity && (fit_face || fit_pixel || phi || thetaDash || theta || I)) {
        set_error(ctx, "gather: capacity too small for the number of fits");
        rc = BRDFGPU_LM_ERROR;
    }
    if (rc == 0) {
        const long ns = g.nfit * sc->nimg;
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
            for (int ch = 0; ch < 3; ++ch) down(I + (size_t)ch * capacity * sc->nimg, g.I + (size_t)ch * ns, sizeof(double) * ns);
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

extern "C" int brdfgpu_gather_resident(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cams, int ncam, int model,
                                       int channel, brdfgpu_samples** global_out, brdfgpu_batch** batch_out,
                                       long* nfit_out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !cams || channel < 0 || channel > 2 || (model != 0 && model != 1)) return BRDFGPU_LM_ERROR;
    GatherDev g;
    int rc = gather_device(ctx, sc, cams, ncam, true, &g);
    if (rc != 0) {
        g.release();
        return rc;
    }
    const long ns = g.nfit * sc->nimg;
    if (nfit_out) *nfit_out = g.nfit;
    const double* t = model == 1 ? g.thetaDash : g.theta;
    const size_t nb = sizeof(double) * (size_t)ns;
    cudaError_t e = cudaSuccess;
    if (global_out) {
        *global_out = nullptr;
        brdfgpu_samples* s = nullptr;
        if (samples_alloc(ctx, ns, model, &s) != 0) rc = BRDFGPU_LM_ERROR;
        else if (ns > 0) {
            e = cudaMemcpyAsync(s->c, g.phi, nb, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(s->traw, t, nb, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(s->x, g.I + (size_t)channel * ns, nb, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess && samples_prepare(ctx, s) != 0) rc = BRDFGPU_LM_ERROR;
        }
        if (rc == 0 && e == cudaSuccess) *global_out = s;
        else if (s) brdfgpu_samples_free(ctx, s);
    }
    if (batch_out && rc == 0 && e == cudaSuccess) {
        *batch_out = nullptr;
        brdfgpu_batch* b = nullptr;
        if (batch_alloc(ctx, g.nfit, sc->nimg, model, &b) != 0) rc = BRDFGPU_LM_ERROR;
        else if (ns > 0) {
            e = cudaMemcpyAsync(b->c, g.phi, nb, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(b->traw, t, nb, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(b->x, g.I + (size_t)channel * ns, nb, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess && batch_prepare(ctx, b) != 0) rc = BRDFGPU_LM_ERROR;
        }
        if (rc == 0 && e == cudaSuccess) *batch_out = b;
        else if (b) brdfgpu_batch_free(ctx, b);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    g.release();
    if (e != cudaSuccess) {
        set_error(ctx, std::string("gather_resident: ") + cudaGetErrorString(e));
        return BRDFGPU_LM_ERROR;
    }
    return rc;
}

// resident global-fit sample set of one colour channel from a finished gather
static int samples_from_gather(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const GatherDev& g, int model, int channel,
                               brdfgpu_samples** out) {
    const long ns = g.nfit * sc->nimg;
    const size_t nb = sizeof(double) * (size_t)ns;
    const double* t = model == 1 ? g.thetaDash : g.theta;
    brdfgpu_samples* s = nullptr;
    if (samples_alloc(ctx, ns, model, &s) != 0) return BRDFGPU_LM_ERROR;
    cudaError_t e = cudaSuccess;
    if (ns > 0) {
        e = cudaMemcpyAsync(s->c, g.phi, nb, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->traw, t, nb, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->x, g.I + (size_t)channel * ns, nb, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess && samples_prepare(ctx, s) != 0) e = cudaErrorUnknown;
    }
    if (e != cudaSuccess) {
        brdfgpu_samples_free(ctx, s);
        set_error(ctx, std::string("gather -> samples: ") + cudaGetErrorString(e));
        return BRDFGPU_LM_ERROR;
    }
    *out = s;
    return 0;
}

// CBRDFdata::CalcBRDFEquation (brdfdata.cpp:1188-1227): ONE gather, then the fits of all three colour
// channels as one batch of 3 x nfit problems (fit ch * nfit + f = face f, channel ch: the B, G, R
// intensity blocks of the gather are already laid out that way) in a single launch.
extern "C" long brdfgpu_calc_brdf_equation(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cam, int model,
                                           double* brdf_surfaces) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !cam || !brdf_surfaces) return BRDFGPU_LM_ERROR;
    static const double p0[3] = {0.5, 1.0, 1.0}, lb[3] = {0, 0, 0}, ub[3] = {100, 100, 100};
    static const double opts[5] = {1E-03, 1E-15, 1E-15, 1E-20, 1E-06};
    GatherDev g;
    if (gather_device(ctx, sc, cam, 1, true, &g) != 0) {
        g.release();
        return BRDFGPU_LM_ERROR;
    }
    const long nfit = g.nfit;
    if (nfit == 0) {
        g.release();
        return 0;
    }
    const long ns = nfit * sc->nimg;
    const size_t nb = sizeof(double) * (size_t)ns;
    const double* t = model == 1 ? g.thetaDash : g.theta;
    std::vector<int> faces(nfit);
    std::vector<double> p(9 * (size_t)nfit);
    brdfgpu_batch* b = nullptr;
    int rc = batch_alloc(ctx, 3 * nfit, sc->nimg, model, &b);
    cudaError_t e = cudaSuccess;
    if (rc == 0) {
        e = cudaMemcpyAsync(faces.data(), g.fit_face, sizeof(int) * nfit, cudaMemcpyDeviceToHost, ctx->stream);
        for (int ch = 0; ch < 3 && e == cudaSuccess; ++ch) {
            e = cudaMemcpyAsync(b->c + (size_t)ch * ns, g.phi, nb, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(b->traw + (size_t)ch * ns, t, nb, cudaMemcpyDeviceToDevice, ctx->stream);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(b->x, g.I, 3 * nb, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) rc = BRDFGPU_LM_ERROR;
        if (rc == 0) rc = batch_prepare(ctx, b);
        if (rc == 0) rc = brdfgpu_batch_fit(ctx, b, p0, lb, ub, 100, opts, BRDFGPU_JAC_FD);
        if (rc == 0) rc = brdfgpu_batch_results(ctx, b, p.data(), nullptr, nullptr);
    }
    if (b) brdfgpu_batch_free(ctx, b);
    g.release();
    if (rc != 0) {
        if (e != cudaSuccess) set_error(ctx, std::string("calc_brdf_equation: ") + cudaGetErrorString(e));
        return BRDFGPU_LM_ERROR;
    }
    for (int ch = 0; ch < 3; ++ch)          // brdfdata.cpp:1205: B, G, R
        for (long f = 0; f < nfit; ++f)     // SaveValuesToSurface, brdfdata.cpp:368-377
            for (int j = 0; j < 3; ++j) brdf_surfaces[((size_t)faces[f] * 3 + ch) * 3 + j] = p[3 * ((size_t)ch * nfit + f) + j];
    return nfit;
}

// CBRDFdata::CalcBRDFEquation_SingleBRDF (brdfdata.cpp:1138-1186): one gather, one global fit per channel
extern "C" long brdfgpu_calc_brdf_equation_single(brdfgpu_ctx* ctx, const brdfgpu_scene* sc, const double* cam, int model,
                                                  double* single_brdf, double* info, int* ret) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !sc || !cam || !single_brdf) return BRDFGPU_LM_ERROR;
    static const double lb[3] = {0, 0, 0}, ub[3] = {100, 100, 100};
    static const double opts[5] = {1E-03, 1E-15, 1E-10, 1E-50, 1.0};
    GatherDev g;
    if (gather_device(ctx, sc, cam, 1, true, &g) != 0) {
        g.release();
        return BRDFGPU_LM_ERROR;
    }
    const long nfit = g.nfit;
    Trace tr("calc_brdf_equation_single");
    for (int ch = 0; ch < 3; ++ch) {
        brdfgpu_samples* s = nullptr;
        if (samples_from_gather(ctx, sc, g, model, ch, &s) != 0) {
            g.release();
            return BRDFGPU_LM_ERROR;
        }
        tr.mark("samples from gather (4 cudaMalloc, 3 D2D, log pass)");
        double p[3] = {0.0, 0.0, 0.0}, inf[10] = {0};
        const int r = brdfgpu_fit_global(ctx, s, p, 3, lb, ub, nullptr, 2000, opts, inf, nullptr, BRDFGPU_DRIVE_PERSISTENT,
                                         BRDFGPU_JAC_FD);
        tr.mark("global fit");
        brdfgpu_samples_free(ctx, s);
        tr.mark("samples free");
        for (int j = 0; j < 3; ++j) single_brdf[ch * 3 + j] = p[j];
        if (info)
            for (int j = 0; j < 10; ++j) info[ch * 10 + j] = inf[j];
        if (ret) ret[ch] = r;
    }
    g.release();
    tr.mark("gather release (12 cudaFree)");
    return nfit;
}

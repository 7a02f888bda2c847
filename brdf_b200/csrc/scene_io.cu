// scene_io.cu -- host-side loaders for the reference's own input files: the
// step immediately before the gather.  Pure host code (no kernels): usable without a GPU.
//
//   brdfgpu_read_cal   CBRDFdata::LoadCameraParameters + WriteValue   brdfdata.cpp:149-247
//   brdfgpu_read_obj   CBRDFdata::LoadModel -> igl::readOBJ           brdfdata.cpp:289-312
//   brdfgpu_read_png   cv::imread(path, IMREAD_COLOR) for 8-bit PNGs  brdfdata.cpp:34-61,117-128
//   brdfgpu_scene_load main.cpp:41-58 (LoadModel, LoadImages, SubtractAmbientLight,
//                      LoadCameraParameters, InitLEDs) into a device-resident scene
//
// PNG: the photographs shipped with the reference are 8-bit RGB, non-interlaced; the decoder handles
// 8-bit grey / grey+alpha / RGB / RGBA / palette, non-interlaced, through zlib's inflate and returns
// what IMREAD_COLOR returns for them: H x W x 3, B G R order, alpha dropped.
#include <zlib.h>

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace brdfgpu {

static bool read_whole_file(const char* path, std::vector<unsigned char>* out, brdfgpu_ctx* ctx) {
    FILE* f = path ? fopen(path, "rb") : nullptr;
    if (!f) {
        set_error(ctx, std::string("cannot open ") + (path ? path : "(null)") + ": " + strerror(errno));
        return false;
    }
    out->clear();
    unsigned char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) out->insert(out->end(), buf, buf + got);
    fclose(f);
    return true;
}

// ---- .cal ------------------------------------------------------------------------------------
// The reference scans the file for '<name>value<' groups (brdfdata.cpp:161-189): a '<' opens a name that
// runs to '>', the value runs to the next '<', WriteValue(name, atof(value)) keeps the 16 fields it
// knows, then everything up to the closing '>' of the end tag is skipped.  Scanning stops at a NUL.
static const char* const kCalFields[BRDFGPU_CAM_SZ] = {"cx", "cy", "f",  "sx", "nx", "ny", "nz", "ox",
                                                       "oy", "oz", "ax", "ay", "az", "px", "py", "pz"};

static int parse_cal(const std::vector<unsigned char>& buf, double* cam, double* kappa1 = nullptr, int* has_kappa1 = nullptr) {
    int seen = 0;
    for (int i = 0; i < BRDFGPU_CAM_SZ; ++i) cam[i] = 0.0;
    if (kappa1) *kappa1 = 0.0;
    if (has_kappa1) *has_kappa1 = 0;
    const size_t n = buf.size();
    size_t it = 0;
    while (it < n && buf[it] != 0) {
        if (buf[it] == '<') {
            ++it;
            std::string name, value;
            for (; it < n && buf[it] != '>'; ++it) name += (char)buf[it];
            if (it < n && buf[it] == '>') ++it;
            for (; it < n && buf[it] != '<'; ++it) value += (char)buf[it];
            if (it < n && buf[it] == '<') ++it;
            if (kappa1 && name == "kappa1") {
                *kappa1 = atof(value.c_str());
                if (has_kappa1) *has_kappa1 = 1;
            }
            for (int k = 0; k < BRDFGPU_CAM_SZ; ++k)
                if (name == kCalFields[k]) {
                    cam[k] = atof(value.c_str());  // brdfdata.cpp:197
                    seen |= 1 << k;
                }
            for (; it < n && buf[it] != '>'; ++it) {
            }
            if (it < n && buf[it] == '>') ++it;
        }
        ++it;  // the reference's loop increment applies after a group as well (:159)
    }
    return seen;
}

// ---- .obj ------------------------------------------------------------------------------------
// igl::readOBJ as the reference uses it: `v x y z` rows into V, the VERTEX index of every `f`
// corner (v, v/vt, v//vn, v/vt/vn; 1-based, negative = relative to the vertices read so far) into F.
// The reference only ever reads columns 0..2 of a face (brdfdata.cpp:319-321,652-658), so the first
// three corners are kept; a face with fewer than three is an error.
static bool parse_obj(const std::vector<unsigned char>& buf, std::vector<double>* V, std::vector<int>* F, std::string* why) {
    V->clear();
    F->clear();
    const char* p = reinterpret_cast<const char*>(buf.data());
    const char* end = p + buf.size();
    long line_no = 0;
    while (p < end) {
        const char* eol = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
        if (!eol) eol = end;
        ++line_no;
        std::string line(p, eol);
        p = eol < end ? eol + 1 : end;
        size_t s = line.find_first_not_of(" \t\r");
        if (s == std::string::npos) continue;
        const char* c = line.c_str() + s;
        if (c[0] == 'v' && (c[1] == ' ' || c[1] == '\t')) {
            char* q = nullptr;
            const char* r = c + 1;
            double xyz[3];
            for (int k = 0; k < 3; ++k) {
                xyz[k] = strtod(r, &q);
                if (q == r) {
                    *why = "vertex with fewer than 3 coordinates at line " + std::to_string(line_no);
                    return false;
                }
                r = q;
            }
            V->insert(V->end(), xyz, xyz + 3);
        } else if (c[0] == 'f' && (c[1] == ' ' || c[1] == '\t')) {
            const char* r = c + 1;
            int corners = 0;
            long idx[3];
            while (*r) {
                while (*r == ' ' || *r == '\t' || *r == '\r') ++r;
                if (!*r) break;
                char* q = nullptr;
                const long v = strtol(r, &q, 10);
                if (q == r) {
                    *why = "unreadable face corner at line " + std::to_string(line_no);
                    return false;
                }
                if (corners < 3) idx[corners] = v < 0 ? v + (long)(V->size() / 3) : v - 1;
                ++corners;
                r = q;
                while (*r && *r != ' ' && *r != '\t' && *r != '\r') ++r;  // the /vt/vn part
            }
            if (corners < 3) {
                *why = "face with fewer than 3 corners at line " + std::to_string(line_no);
                return false;
            }
            for (int k = 0; k < 3; ++k) F->push_back((int)idx[k]);
        }
    }
    const long nV = (long)(V->size() / 3);
    for (size_t i = 0; i < F->size(); ++i)
        if ((*F)[i] < 0 || (*F)[i] >= nV) {
            *why = "face " + std::to_string(i / 3) + " refers to vertex " + std::to_string((*F)[i] + 1) + " of " + std::to_string(nV);
            return false;
        }
    return true;
}

// ---- .png ------------------------------------------------------------------------------------
static uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static bool decode_png(const std::vector<unsigned char>& file, int* W, int* H, std::vector<unsigned char>* bgr, bool header_only,
                       std::string* why) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) {
        *why = "not a PNG file";
        return false;
    }
    size_t pos = 8;
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, palette;
    bool have_hdr = false, done = false;
    while (!done && pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const unsigned char* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) {
            *why = "truncated PNG chunk";
            return false;
        }
        const unsigned char* data = &file[pos + 8];
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, 4 + len) != be32(&file[pos + 8 + len])) {  // type + data, PNG spec 5.3
            *why = "PNG chunk CRC mismatch (corrupt file)";
            return false;
        }
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) {
                *why = "bad IHDR";
                return false;
            }
            w = (int)be32(data); h = (int)be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            have_hdr = true;
            if (header_only) break;
        } else if (!memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            done = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_hdr || w <= 0 || h <= 0) {
        *why = "PNG without a valid IHDR";
        return false;
    }
    // a corrupt or crafted IHDR must not turn into a giant allocation (cv::imread returns an empty Mat and the
    // reference fails cleanly): bound each side, and the pixel count against what the IDAT data can inflate to
    // (deflate expands by at most ~1032x)
    if (w > 65535 || h > 65535) {
        *why = "PNG dimensions out of range (each side <= 65535)";
        return false;
    }
    *W = w;
    *H = h;
    if (header_only) return true;
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: *why = "unknown PNG colour type"; return false;
    }
    if (depth != 8 || interlace != 0) {
        *why = "only 8-bit non-interlaced PNGs are supported (the reference's photographs are)";
        return false;
    }
    if (ctype == 3 && palette.size() < 3) {
        *why = "palette PNG without PLTE";
        return false;
    }
    const size_t stride = (size_t)w * channels;
    if ((stride + 1) * (size_t)h > idat.size() * 1032 + 1024) {
        *why = "PNG header announces more pixels than its IDAT data can hold";
        return false;
    }
    std::vector<unsigned char> raw((stride + 1) * (size_t)h);
    uLongf raw_len = (uLongf)raw.size();
    const int zr = uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || raw_len != raw.size()) {
        *why = "PNG pixel data does not inflate to the expected size";
        return false;
    }
    // undo the per-row filters in place (PNG spec 9.2): bpp = bytes per complete pixel
    const int bpp = channels;
    std::vector<unsigned char> zero(stride, 0);
    for (int y = 0; y < h; ++y) {
        unsigned char* row = &raw[(stride + 1) * (size_t)y];
        const int filter = row[0];
        unsigned char* cur = row + 1;
        const unsigned char* up = y ? &raw[(stride + 1) * (size_t)(y - 1) + 1] : zero.data();
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up[i], c = i >= (size_t)bpp ? up[i - bpp] : 0;
            int pred = 0;
            switch (filter) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: *why = "bad PNG filter type"; return false;
            }
            cur[i] = (unsigned char)(cur[i] + pred);
        }
    }
    bgr->resize((size_t)w * h * 3);
    for (int y = 0; y < h; ++y) {
        const unsigned char* cur = &raw[(stride + 1) * (size_t)y + 1];
        unsigned char* dst = &(*bgr)[(size_t)y * w * 3];
        for (int x = 0; x < w; ++x) {
            unsigned char r, g, b;
            const unsigned char* px = cur + (size_t)x * channels;
            if (ctype == 0 || ctype == 4) {
                r = g = b = px[0];
            } else if (ctype == 3) {
                const size_t e = (size_t)px[0] * 3;
                if (e + 3 > palette.size()) {
                    *why = "palette index out of range";
                    return false;
                }
                r = palette[e]; g = palette[e + 1]; b = palette[e + 2];
            } else {
                r = px[0]; g = px[1]; b = px[2];
            }
            dst[3 * x] = b; dst[3 * x + 1] = g; dst[3 * x + 2] = r;
        }
    }
    return true;
}

}  // namespace brdfgpu

using namespace brdfgpu;

static int read_cal_impl(const char* path, double* cam16) {
    brdfgpu_ctx* ctx = nullptr;
    std::vector<unsigned char> buf;
    if (!cam16 || !read_whole_file(path, &buf, ctx)) return BRDFGPU_LM_ERROR;
    return parse_cal(buf, cam16);
}

static int read_cal_kappa1_impl(const char* path, double* kappa1) {
    brdfgpu_ctx* ctx = nullptr;
    std::vector<unsigned char> buf;
    if (!kappa1 || !read_whole_file(path, &buf, ctx)) return BRDFGPU_LM_ERROR;
    double cam[BRDFGPU_CAM_SZ];
    int has = 0;
    parse_cal(buf, cam, kappa1, &has);
    return has;
}

static int read_obj_impl(const char* path, double* V, int* F, int* nV, int* nF) {
    brdfgpu_ctx* ctx = nullptr;
    std::vector<unsigned char> buf;
    if (!read_whole_file(path, &buf, ctx)) return BRDFGPU_LM_ERROR;
    std::vector<double> v;
    std::vector<int> f;
    std::string why;
    if (!parse_obj(buf, &v, &f, &why)) {
        set_error(ctx, std::string(path) + ": " + why);
        return BRDFGPU_LM_ERROR;
    }
    if (V) {  // second call: the caller sized the arrays from the first
        if (!nV || !nF || *nV < (int)(v.size() / 3) || *nF < (int)(f.size() / 3)) {
            set_error(ctx, "brdfgpu_read_obj: arrays too small (pass the counts of the sizing call in *nV / *nF)");
            return BRDFGPU_LM_ERROR;
        }
        memcpy(V, v.data(), v.size() * sizeof(double));
        if (F) memcpy(F, f.data(), f.size() * sizeof(int));
    }
    if (nV) *nV = (int)(v.size() / 3);
    if (nF) *nF = (int)(f.size() / 3);
    return 0;
}

static int read_png_impl(const char* path, unsigned char* bgr, int* W, int* H) {
    brdfgpu_ctx* ctx = nullptr;
    std::vector<unsigned char> file, out;
    std::string why;
    int w = 0, h = 0;
    if (!read_whole_file(path, &file, ctx)) return BRDFGPU_LM_ERROR;
    if (!decode_png(file, &w, &h, &out, bgr == nullptr, &why)) {
        set_error(ctx, std::string(path) + ": " + why);
        return BRDFGPU_LM_ERROR;
    }
    if (bgr) {
        if (!W || !H || *W != w || *H != h) {
            set_error(ctx, "brdfgpu_read_png: pass the size of the sizing call in *W / *H");
            return BRDFGPU_LM_ERROR;
        }
        memcpy(bgr, out.data(), out.size());
    }
    if (W) *W = w;
    if (H) *H = h;
    return 0;
}

static int scene_load_impl(brdfgpu_ctx* ctx, const char* image_folder, const char* obj_path, const char* cal_path,
                           int nimg, brdfgpu_scene** out, double* cam16) {
    if (!ctx) ctx = default_ctx();  // NULL = the process-wide context, as everywhere in this API
    if (!ctx || !image_folder || !obj_path || !out || nimg < 1) return BRDFGPU_LM_ERROR;
    // LoadModel (main.cpp:41)
    std::vector<unsigned char> buf;
    std::vector<double> V;
    std::vector<int> F;
    std::string why;
    if (!read_whole_file(obj_path, &buf, ctx)) return BRDFGPU_LM_ERROR;
    if (!parse_obj(buf, &V, &F, &why)) {
        set_error(ctx, std::string(obj_path) + ": " + why);
        return BRDFGPU_LM_ERROR;
    }
    // LoadImages: <folder>1.png .. <folder>N.png (main.cpp:46, brdfdata.cpp:34-61; the folder string is used
    // as a prefix, exactly as the reference concatenates it)
    std::vector<std::vector<unsigned char>> imgs((size_t)nimg);
    int W = -1, H = -1;
    for (int k = 0; k < nimg; ++k) {
        const std::string path = std::string(image_folder) + std::to_string(k + 1) + ".png";
        int w = 0, h = 0;
        if (!read_whole_file(path.c_str(), &buf, ctx)) return BRDFGPU_LM_ERROR;
        if (!decode_png(buf, &w, &h, &imgs[(size_t)k], false, &why)) {
            set_error(ctx, path + ": " + why);
            return BRDFGPU_LM_ERROR;
        }
        if (W < 0) { W = w; H = h; }  // brdfdata.cpp:52-56: the first image sets the size
        if (w != W || h != H) {
            set_error(ctx, path + ": size differs from the first photograph");
            return BRDFGPU_LM_ERROR;
        }
    }
    // SubtractAmbientLight (main.cpp:49, brdfdata.cpp:130-147): without a dark frame the reference only prints a
    // message and carries on
    std::vector<unsigned char> dark;
    {
        const std::string path = std::string(image_folder) + "dark.png";
        FILE* probe = fopen(path.c_str(), "rb");
        if (probe) {
            fclose(probe);
            int w = 0, h = 0;
            if (!read_whole_file(path.c_str(), &buf, ctx)) return BRDFGPU_LM_ERROR;
            if (!decode_png(buf, &w, &h, &dark, false, &why) || w != W || h != H) {
                set_error(ctx, path + ": " + (why.empty() ? "size differs from the photographs" : why));
                return BRDFGPU_LM_ERROR;
            }
        } else {
            fprintf(stderr, "Could not subtract ambient light\n");
        }
    }
    // LoadCameraParameters (main.cpp:55)
    if (cal_path) {
        if (!cam16 || !read_whole_file(cal_path, &buf, ctx)) return BRDFGPU_LM_ERROR;
        parse_cal(buf, cam16);
    }
    std::vector<const unsigned char*> ptrs((size_t)nimg);
    for (int k = 0; k < nimg; ++k) ptrs[(size_t)k] = imgs[(size_t)k].data();
    // InitLEDs (main.cpp:59): led == NULL selects the reference table when nimg == 16
    return brdfgpu_scene_create(ctx, V.data(), (int)(V.size() / 3), F.data(), (int)(F.size() / 3), ptrs.data(), nimg, W, H,
                                dark.empty() ? nullptr : dark.data(), nullptr, out);
}

// The C boundary: nothing thrown by the parsers above (std::bad_alloc / length_error on absurd sizes in a corrupt
// file) may cross an extern "C" frame -- a C host would see std::terminate.  Failures become BRDFGPU_LM_ERROR with
// brdfgpu_last_error() set, like a cv::imread that returns an empty Mat.
template <class Fn>
static int guarded(const char* what, brdfgpu_ctx* ctx, Fn&& fn) {
    try {
        return fn();
    } catch (const std::exception& e) {
        set_error(ctx, std::string(what) + ": " + e.what());
    } catch (...) {
        set_error(ctx, std::string(what) + ": unknown failure");
    }
    return BRDFGPU_LM_ERROR;
}
extern "C" int brdfgpu_read_cal(const char* path, double* cam16) {
    return guarded("brdfgpu_read_cal", nullptr, [&] { return read_cal_impl(path, cam16); });
}
extern "C" int brdfgpu_read_cal_kappa1(const char* path, double* kappa1) {
    return guarded("brdfgpu_read_cal_kappa1", nullptr, [&] { return read_cal_kappa1_impl(path, kappa1); });
}
extern "C" int brdfgpu_read_obj(const char* path, double* V, int* F, int* nV, int* nF) {
    return guarded("brdfgpu_read_obj", nullptr, [&] { return read_obj_impl(path, V, F, nV, nF); });
}
extern "C" int brdfgpu_read_png(const char* path, unsigned char* bgr, int* W, int* H) {
    return guarded("brdfgpu_read_png", nullptr, [&] { return read_png_impl(path, bgr, W, H); });
}
extern "C" int brdfgpu_scene_load(brdfgpu_ctx* ctx, const char* image_folder, const char* obj_path, const char* cal_path,
                                  int nimg, brdfgpu_scene** out, double* cam16) {
    return guarded("brdfgpu_scene_load", ctx, [&] { return scene_load_impl(ctx, image_folder, obj_path, cal_path, nimg, out, cam16); });
}

/*
 * scene_pipeline_c.c -- the reference's main.cpp (:27-61) plus the 'c' key of its viewer, in plain C against
 * libbrdfgpu: load the model, the 16 photographs, the dark frame and the calibration from the reference's own
 * file formats, fit one BRDF per face (CBRDFdata::CalcBRDFEquation, brdfdata.cpp:1188-1227) and one per colour
 * channel (CalcBRDFEquation_SingleBRDF, :1138-1186), then colour the faces as the viewer's shaded-BRDF mode does
 * (glutcallbacks.cpp:346-445).  Same command line as the reference:
 *
 *     scene_pipeline_c <image folder/> <model.obj> <camera.cal>
 *
 *     gcc -O2 -Iinclude examples/scene_pipeline_c.c -Lbrdf_b200 -lbrdfgpu -Wl,-rpath,$PWD/brdf_b200 -lm -o scene_pipeline_c
 *
 * Prints the number of fitted faces, the three single-BRDF parameter triples and a checksum of the colours
 * (tests/test_gpu_c_example.py compares them with the Python mirror of the same calls).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "brdfgpu.h"

int main(int argc, char **argv) {
    if (argc < 4) { /* main.cpp:29-33 */
        printf("required command line arguments: path to image folder, path to obj file, path to cal file\n");
        return -1;
    }
    const int model = BRDFGPU_MODEL_BLINN_PHONG; /* main.cpp:43 */
    brdfgpu_scene *sc = NULL;
    double cam[BRDFGPU_CAM_SZ];
    if (brdfgpu_scene_load(NULL, argv[1], argv[2], argv[3], 16, &sc, cam) != 0) { /* main.cpp:41-59 */
        fprintf(stderr, "load failed: %s\n", brdfgpu_last_error(NULL));
        return -1;
    }
    int dims[5];
    brdfgpu_scene_dims(sc, dims);
    const int nF = dims[1];
    printf("scene: %d vertices, %d faces, %d photographs %dx%d\n", dims[0], nF, dims[2], dims[3], dims[4]);

    double *surfaces = malloc(sizeof(double) * 9 * (size_t)nF), *bgr = malloc(sizeof(double) * 3 * (size_t)nF);
    if (!surfaces || !bgr) return 2;
    for (long i = 0; i < 9L * nF; ++i) surfaces[i] = 0.0; /* faces no pixel maps to keep {0,0,0} */
    const long nfit = brdfgpu_calc_brdf_equation(NULL, sc, cam, model, surfaces);
    if (nfit < 0) return 3;
    printf("per-face: %ld faces fitted\n", nfit);

    double single[9], info[30];
    int ret[3];
    if (brdfgpu_calc_brdf_equation_single(NULL, sc, cam, model, single, info, ret) < 0) return 4;
    for (int ch = 0; ch < 3; ++ch)
        printf("single[%d]: ret=%d p=%.17g %.17g %.17g reason=%d\n", ch, ret[ch], single[3 * ch], single[3 * ch + 1],
               single[3 * ch + 2], (int)info[10 * ch + 6]);

    /* viewer: eye at the camera, looking along its axis (glutcallbacks.cpp:352-375 uses m_eye / m_center) */
    const double eye[3] = {cam[13], cam[14], cam[15]};
    const double center[3] = {cam[13] + cam[10], cam[14] + cam[11], cam[15] + cam[12]};
    if (brdfgpu_shade_faces(NULL, sc, eye, center, model, 0, surfaces, 1, bgr) != 0) return 5;
    double sum = 0.0;
    long finite = 0;
    for (long i = 0; i < 3L * nF; ++i)
        if (bgr[i] == bgr[i] && bgr[i] - bgr[i] == 0.0) { sum += bgr[i]; ++finite; }
    printf("shaded: %ld finite colour values, sum=%.12g\n", finite, sum);
    brdfgpu_scene_free(NULL, sc);
    free(surfaces);
    free(bgr);
    return 0;
}

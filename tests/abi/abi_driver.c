/* Plain-C driver of the C ABI (include/saro_gs_b200.h): no torch, no C++ — what a cgo / JNI / ctypes binding of the
 * reference's FFI for this path would do.  Reads a scene from a flat binary file, runs sgs_forward + sgs_backward
 * with cudaMalloc-backed resize callbacks and writes the outputs back.  Built and run by tests/test_abi_c_driver.py.
 *
 *   file in : int32 P, D, M, W, H | float tanfovx, tanfovy, scale_modifier | bg[3] view[16] proj[16] campos[3]
 *             means3D[P*3] shs[P*M*3] opacities[P] scales[P*3] rotations[P*4] dL_dpix[3*H*W]
 *   file out: int64 num_rendered | color[3*H*W] depth[H*W] radii[P] (int32) | dmean3D[P*3] dopacity[P] dscale[P*3]
 *             drot[P*4] dsh[P*M*3] dmean2D[P*3]
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "saro_gs_b200.h"

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

typedef struct {
    char* ptr;
    size_t bytes;
} buf_t;

static char* resize_cb(void* user, size_t bytes) { /* the C form of std::function<char*(size_t)> */
    buf_t* b = (buf_t*)user;
    if (b->ptr) cudaFree(b->ptr);
    b->ptr = NULL;
    b->bytes = bytes;
    if (cudaMalloc((void**)&b->ptr, bytes ? bytes : 1) != cudaSuccess) return NULL;
    return b->ptr;
}

static float* upload(FILE* f, size_t n) {
    float* h = (float*)malloc(n * sizeof(float));
    float* d = NULL;
    if (fread(h, sizeof(float), n, f) != n) { free(h); return NULL; }
    if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) { free(h); return NULL; }
    cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice);
    free(h);
    return d;
}

static int download(FILE* f, const void* d, size_t bytes) {
    void* h = malloc(bytes);
    if (cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { free(h); return 1; }
    fwrite(h, 1, bytes, f);
    free(h);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) return 1;
    int32_t hdr[5];
    float fl[3];
    if (fread(hdr, 4, 5, fi) != 5 || fread(fl, 4, 3, fi) != 3) return 1;
    const int P = hdr[0], D = hdr[1], M = hdr[2], W = hdr[3], H = hdr[4];
    if (sgs_abi_version() != SGS_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 3; }
    float *bg = upload(fi, 3), *view = upload(fi, 16), *proj = upload(fi, 16), *campos = upload(fi, 3);
    float* means = upload(fi, (size_t)P * 3);
    float* shs = upload(fi, (size_t)P * M * 3);
    float* opac = upload(fi, (size_t)P);
    float* scales = upload(fi, (size_t)P * 3);
    float* rots = upload(fi, (size_t)P * 4);
    float* dpix = upload(fi, (size_t)3 * H * W);
    fclose(fi);
    if (!bg || !view || !proj || !campos || !means || !shs || !opac || !scales || !rots || !dpix) return 1;

    float *color, *depth, *dm3, *dm2, *dacc, *dop, *dcol, *dcov, *dsh, *dsc, *drot;
    int* radii;
    CK(cudaMalloc((void**)&color, (size_t)3 * H * W * 4));
    CK(cudaMalloc((void**)&depth, (size_t)H * W * 4));
    CK(cudaMalloc((void**)&radii, (size_t)P * 4));
    CK(cudaMalloc((void**)&dm3, (size_t)P * 3 * 4));
    CK(cudaMalloc((void**)&dm2, (size_t)P * 3 * 4));
    CK(cudaMalloc((void**)&dacc, (size_t)P * 12 * 4));
    CK(cudaMalloc((void**)&dop, (size_t)P * 4));
    CK(cudaMalloc((void**)&dcol, (size_t)P * 3 * 4));
    CK(cudaMalloc((void**)&dcov, (size_t)P * 6 * 4));
    CK(cudaMalloc((void**)&dsh, (size_t)P * M * 3 * 4));
    CK(cudaMalloc((void**)&dsc, (size_t)P * 3 * 4));
    CK(cudaMalloc((void**)&drot, (size_t)P * 4 * 4));

    buf_t geom = {0, 0}, binning = {0, 0}, image = {0, 0};
    const int64_t R = sgs_forward(resize_cb, &geom, resize_cb, &binning, resize_cb, &image, P, D, M, bg, W, H, means, shs,
                                  NULL, opac, scales, fl[2], rots, NULL, view, proj, campos, fl[0], fl[1], 0, color, depth,
                                  radii, SGS_FLAG_KEEP_FOR_BACKWARD, NULL);
    if (R < 0) { fprintf(stderr, "sgs_forward: %s\n", sgs_last_error()); return 4; }
    const int rc = sgs_backward(P, D, M, R, bg, W, H, means, shs, NULL, scales, fl[2], rots, NULL, view, proj, campos,
                                fl[0], fl[1], radii, geom.ptr, binning.ptr, image.ptr, dpix, dm2, dacc, dop, dcol, dm3,
                                dcov, dsh, dsc, drot, NULL);
    if (rc < 0) { fprintf(stderr, "sgs_backward: %s\n", sgs_last_error()); return 5; }
    CK(cudaDeviceSynchronize());

    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 1;
    fwrite(&R, 8, 1, fo);
    int bad = download(fo, color, (size_t)3 * H * W * 4) | download(fo, depth, (size_t)H * W * 4) |
              download(fo, radii, (size_t)P * 4) | download(fo, dm3, (size_t)P * 3 * 4) | download(fo, dop, (size_t)P * 4) |
              download(fo, dsc, (size_t)P * 3 * 4) | download(fo, drot, (size_t)P * 4 * 4) |
              download(fo, dsh, (size_t)P * M * 3 * 4) | download(fo, dm2, (size_t)P * 3 * 4);
    fclose(fo);
    /* error path of the ABI: negative sizes are refused with a message */
    if (sgs_forward(resize_cb, &geom, resize_cb, &binning, resize_cb, &image, -1, 0, 0, bg, W, H, means, shs, NULL, opac,
                    scales, 1.f, rots, NULL, view, proj, campos, 1.f, 1.f, 0, color, depth, radii, 0, NULL) !=
        SGS_ERR_INVALID_ARGUMENT)
        return 6;
    printf("ok R=%lld geom=%zu binning=%zu image=%zu bytes\n", (long long)R, geom.bytes, binning.bytes, image.bytes);
    return bad ? 7 : 0;
}

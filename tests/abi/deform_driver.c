/* Plain-C driver of the widened rows of the C ABI (include/saro_gs_b200.h): the deformation hand-off
 * (sgs_deform_pack_mlp x 3, sgs_deform_eval) followed by the densification statistics (sgs_densify_add_view,
 * sgs_densify_commit) — no torch, no C++.  Built and run by tests/test_abi_c_driver.py.
 *
 *   file in : int32 N, feat_dim | float timestamp | xyz[N*3] rotation[N*4] scaling[N*3] opacity[N] features_dc[N*3]
 *             features_rest[N*45] temporal_pos[N] lifespan[N] hexplane_feature[N*feat_dim]
 *             | for each of motion (out 3), rot (7), shs (48): W1[128*in] b1[128] W2[128*128] b2[128] W3[out*128] b3[out]
 *             | dL_dmeans2D[N*3] radii[N] (int32)
 *   file out: int64 selected | means3D[S*3] rotations[S*4] scales[S*3] opacity[S] shs[S*48]
 *             | max_radii2D[N] xyz_gradient_accum[N] denom[N]
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

static void* upload(FILE* f, size_t n) { /* n 4-byte elements */
    void* h = malloc(n * 4 + 4);
    void* d = NULL;
    if (fread(h, 4, n, f) != n) { free(h); return NULL; }
    if (cudaMalloc(&d, n * 4 + 4) != cudaSuccess) { free(h); return NULL; }
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    free(h);
    return d;
}

static int download(FILE* f, const void* d, size_t bytes) {
    void* h = malloc(bytes + 4);
    if (bytes && cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { free(h); return 1; }
    fwrite(h, 1, bytes, f);
    free(h);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) return 1;
    int32_t hdr[2];
    float timestamp;
    if (fread(hdr, 4, 2, fi) != 2 || fread(&timestamp, 4, 1, fi) != 1) return 1;
    const int N = hdr[0], F = hdr[1], in_dim = F + 9;
    const int outs[3] = {3, 7, 48};
    float* xyz = upload(fi, (size_t)N * 3);
    float* rot = upload(fi, (size_t)N * 4);
    float* sca = upload(fi, (size_t)N * 3);
    float* opa = upload(fi, (size_t)N);
    float* dc = upload(fi, (size_t)N * 3);
    float* rest = upload(fi, (size_t)N * 45);
    float* tpos = upload(fi, (size_t)N);
    float* life = upload(fi, (size_t)N);
    float* feat = upload(fi, (size_t)N * F);
    if (!xyz || !rot || !sca || !opa || !dc || !rest || !tpos || !life || !feat) return 1;

    void* packed = NULL;
    CK(cudaMalloc(&packed, sgs_deform_packed_bytes()));
    for (int m = 0; m < 3; ++m) {
        float* W1 = upload(fi, (size_t)128 * in_dim);
        float* b1 = upload(fi, 128);
        float* W2 = upload(fi, (size_t)128 * 128);
        float* b2 = upload(fi, 128);
        float* W3 = upload(fi, (size_t)outs[m] * 128);
        float* b3 = upload(fi, (size_t)outs[m]);
        if (!W1 || !b1 || !W2 || !b2 || !W3 || !b3) return 1;
        if (sgs_deform_pack_mlp(m, in_dim, W1, b1, W2, b2, W3, b3, packed, NULL) != 0) {
            fprintf(stderr, "sgs_deform_pack_mlp(%d): %s\n", m, sgs_last_error());
            return 4;
        }
        CK(cudaDeviceSynchronize());   /* the weights may be freed once packed */
        cudaFree(W1); cudaFree(b1); cudaFree(W2); cudaFree(b2); cudaFree(W3); cudaFree(b3);
    }
    float* dm2 = upload(fi, (size_t)N * 3);
    int* radii = (int*)upload(fi, (size_t)N);
    fclose(fi);
    if (!dm2 || !radii) return 1;

    const size_t ws_bytes = sgs_deform_workspace_bytes(N);
    void* ws = NULL;
    float *o_m3, *o_rot, *o_sca, *o_opa, *o_shs;
    CK(cudaMalloc(&ws, ws_bytes));
    CK(cudaMalloc((void**)&o_m3, (size_t)N * 3 * 4 + 4));
    CK(cudaMalloc((void**)&o_rot, (size_t)N * 4 * 4 + 4));
    CK(cudaMalloc((void**)&o_sca, (size_t)N * 3 * 4 + 4));
    CK(cudaMalloc((void**)&o_opa, (size_t)N * 4 + 4));
    CK(cudaMalloc((void**)&o_shs, (size_t)N * 48 * 4 + 4));
    const int64_t S = sgs_deform_eval(N, F, timestamp, xyz, rot, sca, opa, dc, rest, tpos, life, feat, packed, ws, ws_bytes,
                                      o_m3, o_rot, o_sca, o_opa, o_shs, NULL);
    if (S < 0) { fprintf(stderr, "sgs_deform_eval: %lld %s\n", (long long)S, sgs_last_error()); return 4; }

    /* densification statistics of a one-view "batch" */
    float *grad_sum, *max_r, *acc, *den;
    int *vis, *rmax;
    CK(cudaMalloc((void**)&grad_sum, (size_t)N * 4 + 4)); CK(cudaMemset(grad_sum, 0, (size_t)N * 4));
    CK(cudaMalloc((void**)&vis, (size_t)N * 4 + 4));      CK(cudaMemset(vis, 0, (size_t)N * 4));
    CK(cudaMalloc((void**)&rmax, (size_t)N * 4 + 4));     CK(cudaMemset(rmax, 0, (size_t)N * 4));
    CK(cudaMalloc((void**)&max_r, (size_t)N * 4 + 4));    CK(cudaMemset(max_r, 0, (size_t)N * 4));
    CK(cudaMalloc((void**)&acc, (size_t)N * 4 + 4));      CK(cudaMemset(acc, 0, (size_t)N * 4));
    CK(cudaMalloc((void**)&den, (size_t)N * 4 + 4));      CK(cudaMemset(den, 0, (size_t)N * 4));
    if (sgs_densify_add_view(N, dm2, radii, grad_sum, vis, rmax, NULL) != 0) return 5;
    if (sgs_densify_commit(N, grad_sum, vis, rmax, max_r, acc, den, NULL) != 0) return 5;
    CK(cudaDeviceSynchronize());

    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 1;
    fwrite(&S, 8, 1, fo);
    int bad = download(fo, o_m3, (size_t)S * 3 * 4) | download(fo, o_rot, (size_t)S * 4 * 4) | download(fo, o_sca, (size_t)S * 3 * 4) |
              download(fo, o_opa, (size_t)S * 4) | download(fo, o_shs, (size_t)S * 48 * 4) | download(fo, max_r, (size_t)N * 4) |
              download(fo, acc, (size_t)N * 4) | download(fo, den, (size_t)N * 4);
    fclose(fo);
    return bad ? 6 : 0;
}

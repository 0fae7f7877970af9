/* Plain-C driver of the training-time deformation entry points of the C ABI (include/saro_gs_b200.h) — no torch, no
 * C++: one MLP evaluation (rot_mlp-shaped: in -> 128 -> 128 -> 7) forward, its data-gradient chain and its three
 * weight-gradient GEMMs: sgs_deform_pack_general x 2, sgs_deform_train_forward, sgs_deform_train_backward,
 * sgs_deform_wgrad, with the operand planes and sign bits handed from call to call exactly as a C++ autograd node of
 * the reference would hold them.  Built and run by tests/test_abi_c_driver.py.
 *
 *   file in : int32 N, feat_dim | float timestamp | temporal_pos[N] feature[N*feat_dim]
 *             | W1[128*in] b1[128] W2[128*128] b2[128] W3[7*128] b3[7]   (in = feat_dim + 9) | dL_dout[N*7]
 *   file out: out[N*7] | dL_dfeature[N*feat_dim] | dW1[128*in] db1[128] dW2[128*128] db2[128] dW3[7*128]
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "saro_gs_b200.h"

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
            return 2;                                                                 \
        }                                                                             \
    } while (0)
#define SGS(x)                                                                        \
    do {                                                                              \
        int rc_ = (x);                                                                \
        if (rc_ != 0) {                                                               \
            fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, sgs_last_error());            \
            return 3;                                                                 \
        }                                                                             \
    } while (0)

static void* upload(FILE* f, size_t n) { /* n 4-byte elements */
    void* h = malloc(n * 4 + 4);
    void* d = NULL;
    if (fread(h, 4, n, f) != n) { free(h); return NULL; }
    if (cudaMalloc(&d, n * 4 + 32) != cudaSuccess) { free(h); return NULL; }
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    free(h);
    return d;
}
static void* dalloc(size_t bytes) {
    void* d = NULL;
    return cudaMalloc(&d, bytes + 32) == cudaSuccess ? d : NULL;
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
    const int N = hdr[0], F = hdr[1], in = F + 9, OUT = 7;
    float* tpos = upload(fi, (size_t)N);
    float* feat = upload(fi, (size_t)N * F);
    float* W1 = upload(fi, (size_t)128 * in);
    float* b1 = upload(fi, 128);
    float* W2 = upload(fi, 128 * 128);
    float* b2 = upload(fi, 128);
    float* W3 = upload(fi, (size_t)OUT * 128);
    float* b3 = upload(fi, OUT);
    float* dy = upload(fi, (size_t)N * OUT);
    fclose(fi);
    if (!tpos || !feat || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !dy) return 1;

    /* images: forward, and the data-gradient chain's */
    void* img_f = dalloc(sgs_deform_image_bytes());
    void* img_b = dalloc(sgs_deform_image_bytes());
    SGS(sgs_deform_pack_general(0, in, 128, OUT, F, W1, b1, W2, b2, W3, b3, img_f, NULL));
    SGS(sgs_deform_pack_general(1, in, 128, OUT, F, W1, b1, W2, b2, W3, b3, img_b, NULL));

    /* what an autograd node keeps between forward and backward */
    float* out = dalloc((size_t)N * OUT * 4);
    void* h1 = dalloc(sgs_deform_planes_bytes(N, 16));
    void* h2 = dalloc(sgs_deform_planes_bytes(N, 16));
    void* x = dalloc(sgs_deform_planes_bytes(N, 6));
    void* m1 = dalloc((size_t)N * 16);
    void* m2 = dalloc((size_t)N * 16);
    sgs_mlp_job_t fwd;
    memset(&fwd, 0, sizeof fwd);
    fwd.packed = img_f; fwd.out = out; fwd.save_a = h1; fwd.save_b = h2; fwd.save_in = x; fwd.mask_a = m1; fwd.mask_b = m2;
    fwd.n_io = OUT; fwd.zero_time = 0;
    SGS(sgs_deform_train_forward(N, F, timestamp, tpos, feat, 1, &fwd, NULL));

    float* dfeat = dalloc((size_t)N * F * 4);
    void* dh2 = dalloc(sgs_deform_planes_bytes(N, 16));
    void* dh1 = dalloc(sgs_deform_planes_bytes(N, 16));
    void* dyp = dalloc(sgs_deform_planes_bytes(N, 2));
    sgs_mlp_job_t bwd;
    memset(&bwd, 0, sizeof bwd);
    bwd.packed = img_b; bwd.in = dy; bwd.out = dfeat; bwd.save_a = dh2; bwd.save_b = dh1; bwd.save_in = dyp;
    bwd.mask_a = m2; bwd.mask_b = m1; bwd.n_io = OUT;
    SGS(sgs_deform_train_backward(N, F, 1, &bwd, NULL));

    float* dW1 = dalloc((size_t)128 * in * 4);
    float* db1 = dalloc(128 * 4);
    float* dW2 = dalloc(128 * 128 * 4);
    float* db2 = dalloc(128 * 4);
    float* dW3 = dalloc((size_t)OUT * 128 * 4);
    float* partials = dalloc((size_t)sgs_deform_wgrad_max_ctas() * sgs_deform_wgrad_partial_floats() * 4);
    sgs_wgrad_task_t t[3];
    memset(t, 0, sizeof t);
    t[0].A = dh1; t[0].B = x;   t[0].groups_b = 6;  t[0].dW = dW1; t[0].ldw = in;  t[0].rows = 128; t[0].cols = in;  t[0].db = db1;
    t[1].A = dh2; t[1].B = h1;  t[1].groups_b = 16; t[1].dW = dW2; t[1].ldw = 128; t[1].rows = 128; t[1].cols = 128; t[1].db = db2;
    t[2].A = h2;  t[2].B = dyp; t[2].groups_b = 2;  t[2].dW = dW3; t[2].ldw = 128; t[2].rows = 128; t[2].cols = OUT; t[2].transposed = 1;
    SGS(sgs_deform_wgrad(N, 3, t, partials, NULL));
    CK(cudaDeviceSynchronize());

    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 1;
    int bad = download(fo, out, (size_t)N * OUT * 4) | download(fo, dfeat, (size_t)N * F * 4) |
              download(fo, dW1, (size_t)128 * in * 4) | download(fo, db1, 128 * 4) | download(fo, dW2, 128 * 128 * 4) |
              download(fo, db2, 128 * 4) | download(fo, dW3, (size_t)OUT * 128 * 4);
    fclose(fo);
    return bad ? 4 : 0;
}

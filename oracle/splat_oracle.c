/* splat_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, imported by or called from the
 * product path saro_gs_b200/).
 *
 * CPU restatement of the algorithm of the reference rasterizer
 * (yjb6/SaRO-GS, submodules/gaussian_rasterization_ch3 = $R).  It is written from the maths in
 * conventional row-major matrix form (NOT a transcription of the CUDA code); each block cites
 * the reference lines whose behaviour it restates.  Compiled twice:
 *     -DORACLE_REAL=float   -> liboracle_f32.so   (same precision class as the reference)
 *     -DORACLE_REAL=double  -> liboracle_f64.so   (arbiter for gradients: the reference's
 *                                                  float atomics are order-nondeterministic)
 *
 * PARITY PIN: the reference ships no tests / golden vectors.  This oracle is pinned
 * against outputs of the compiled, unmodified reference (oracle/_ref, built by oracle/build_ref.py)
 * captured on a B200 and committed under tests/golden/ (script: tests/golden/make_golden.py).
 *
 * Reference quirks reproduced on purpose (all documented in SURVEY.md Appendix A):
 *   - cull only on z_view <= 0.2                                   $R/cuda_rasterizer/auxiliary.h:139-164
 *   - quaternion used un-normalised, no normalisation Jacobian     $R/cuda_rasterizer/forward.cu:127, backward.cu:340
 *   - pixel centres at integer coordinates, tile-rect clipping     $R/cuda_rasterizer/forward.cu:272-276,331-380
 *   - alpha capped at 0.99 with gradient passed straight through   $R/cuda_rasterizer/backward.cu:497-538
 *   - clamped t.x,t.y treated as independent of t.z in backward    $R/cuda_rasterizer/backward.cu:175-176,255-264
 *   - 1/(det^2 + 1e-7) regulariser in the conic backward           $R/cuda_rasterizer/backward.cu:203
 *   - median depth (default 15.0), depth not differentiable        $R/cuda_rasterizer/forward.cu:308,368-372
 *   - dL/dmean2D reported in NDC units (x 0.5W, 0.5H)              $R/cuda_rasterizer/backward.cu:460-461,545-546
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORACLE_REAL
#define ORACLE_REAL double
#endif
typedef ORACLE_REAL real;

#define TILE 16

static const double SH_C0 = 0.28209479177387814;
static const double SH_C1 = 0.4886025119029199;
static const double SH_C2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
                                -1.0925484305920792, 0.5462742152960396};
static const double SH_C3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                                -0.4570457994644658, 1.445305721320277, -0.5900435899266435};

typedef struct {
    int P, D, M, W, H, tx, ty;
    int64_t R;
    real tan_fovx, tan_fovy, focal_x, focal_y, scale_modifier;
    real view[16], proj[16], campos[3], bg[3];
    int has_sh, has_scale_rot;
    /* per Gaussian */
    int* radii;
    uint32_t* tiles_touched;
    real* depth;     /* [P] */
    real* mean2D;    /* [2P] */
    real* cov3D;     /* [6P] */
    real* conic_o;   /* [4P] A,B,C,opacity */
    real* rgb;       /* [3P] */
    uint8_t* clamped; /* [3P] */
    /* binning */
    uint32_t* point_list; /* [R] */
    uint32_t* ranges;     /* [2*tiles] */
    /* per pixel */
    real* final_T;
    uint32_t* n_contrib;
    real* out_color; /* [3HW] */
    real* out_depth; /* [HW] */
} oracle_ctx;

typedef struct {
    uint64_t key;
    uint32_t val;
} kv_t;

static int kv_cmp(const void* a, const void* b) {
    const kv_t* x = (const kv_t*)a;
    const kv_t* y = (const kv_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    /* stable radix sort of a stream emitted in ascending Gaussian index: ties keep that order
       ($R/cuda_rasterizer/rasterizer_impl.cu:88-107,304-309) */
    if (x->val != y->val) return x->val < y->val ? -1 : 1;
    return 0;
}

static real clampr(real v, real lo, real hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* tile rectangle of a splat: $R/cuda_rasterizer/auxiliary.h:46-56 (float division, truncation, clamp) */
static void tile_rect(real px, real py, int radius, int tx, int ty, int* x0, int* y0, int* x1, int* y1) {
    int a = (int)((float)(px - radius) / TILE), b = (int)((float)(py - radius) / TILE);
    int c = (int)((float)(px + radius + TILE - 1) / TILE), d = (int)((float)(py + radius + TILE - 1) / TILE);
    *x0 = a < 0 ? 0 : (a > tx ? tx : a);
    *y0 = b < 0 ? 0 : (b > ty ? ty : b);
    *x1 = c < 0 ? 0 : (c > tx ? tx : c);
    *y1 = d < 0 ? 0 : (d > ty ? ty : d);
}

/* quaternion (r,x,y,z) -> rotation matrix, NOT normalised ($R/cuda_rasterizer/forward.cu:127-140) */
static void quat_to_R(const real q[4], real Rm[3][3]) {
    real r = q[0], x = q[1], y = q[2], z = q[3];
    Rm[0][0] = 1 - 2 * (y * y + z * z); Rm[0][1] = 2 * (x * y - r * z); Rm[0][2] = 2 * (x * z + r * y);
    Rm[1][0] = 2 * (x * y + r * z); Rm[1][1] = 1 - 2 * (x * x + z * z); Rm[1][2] = 2 * (y * z - r * x);
    Rm[2][0] = 2 * (x * z - r * y); Rm[2][1] = 2 * (y * z + r * x); Rm[2][2] = 1 - 2 * (x * x + y * y);
}

/* SH basis (degree <= 3) and its partial derivatives w.r.t. the (unit) direction components */
static void sh_basis(int deg, real x, real y, real z, real b[16], real db[16][3]) {
    real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    memset(b, 0, sizeof(real) * 16);
    memset(db, 0, sizeof(real) * 48);
    b[0] = (real)SH_C0;
    if (deg < 1) return;
    b[1] = -(real)SH_C1 * y; db[1][1] = -(real)SH_C1;
    b[2] = (real)SH_C1 * z;  db[2][2] = (real)SH_C1;
    b[3] = -(real)SH_C1 * x; db[3][0] = -(real)SH_C1;
    if (deg < 2) return;
    b[4] = (real)SH_C2[0] * xy; db[4][0] = (real)SH_C2[0] * y; db[4][1] = (real)SH_C2[0] * x;
    b[5] = (real)SH_C2[1] * yz; db[5][1] = (real)SH_C2[1] * z; db[5][2] = (real)SH_C2[1] * y;
    b[6] = (real)SH_C2[2] * (2 * zz - xx - yy);
    db[6][0] = (real)SH_C2[2] * (-2 * x); db[6][1] = (real)SH_C2[2] * (-2 * y); db[6][2] = (real)SH_C2[2] * (4 * z);
    b[7] = (real)SH_C2[3] * xz; db[7][0] = (real)SH_C2[3] * z; db[7][2] = (real)SH_C2[3] * x;
    b[8] = (real)SH_C2[4] * (xx - yy); db[8][0] = (real)SH_C2[4] * 2 * x; db[8][1] = (real)SH_C2[4] * (-2 * y);
    if (deg < 3) return;
    b[9] = (real)SH_C3[0] * y * (3 * xx - yy);
    db[9][0] = (real)SH_C3[0] * 6 * xy; db[9][1] = (real)SH_C3[0] * (3 * xx - 3 * yy);
    b[10] = (real)SH_C3[1] * xy * z;
    db[10][0] = (real)SH_C3[1] * yz; db[10][1] = (real)SH_C3[1] * xz; db[10][2] = (real)SH_C3[1] * xy;
    b[11] = (real)SH_C3[2] * y * (4 * zz - xx - yy);
    db[11][0] = (real)SH_C3[2] * (-2 * xy); db[11][1] = (real)SH_C3[2] * (4 * zz - xx - 3 * yy);
    db[11][2] = (real)SH_C3[2] * 8 * yz;
    b[12] = (real)SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy);
    db[12][0] = (real)SH_C3[3] * (-6 * xz); db[12][1] = (real)SH_C3[3] * (-6 * yz);
    db[12][2] = (real)SH_C3[3] * (6 * zz - 3 * xx - 3 * yy);
    b[13] = (real)SH_C3[4] * x * (4 * zz - xx - yy);
    db[13][0] = (real)SH_C3[4] * (4 * zz - 3 * xx - yy); db[13][1] = (real)SH_C3[4] * (-2 * xy);
    db[13][2] = (real)SH_C3[4] * 8 * xz;
    b[14] = (real)SH_C3[5] * z * (xx - yy);
    db[14][0] = (real)SH_C3[5] * 2 * xz; db[14][1] = (real)SH_C3[5] * (-2 * yz); db[14][2] = (real)SH_C3[5] * (xx - yy);
    b[15] = (real)SH_C3[6] * x * (xx - 3 * yy);
    db[15][0] = (real)SH_C3[6] * (3 * xx - 3 * yy); db[15][1] = (real)SH_C3[6] * (-6 * xy);
}

/* A = Jn * Rcw (2x3), the EWA projection Jacobian applied to world covariance
   ($R/cuda_rasterizer/forward.cu:74-113).  t is the view-space mean with x/z, y/z clamped to
   +-1.3 tan(fov/2).  Returns the clamp masks used by the backward. */
static void ewa_A(const oracle_ctx* c, const real mean[3], real A[2][3], real t[3], int* in_x, int* in_y) {
    const real* V = c->view;
    for (int i = 0; i < 3; i++) t[i] = V[i] * mean[0] + V[4 + i] * mean[1] + V[8 + i] * mean[2] + V[12 + i];
    real limx = (real)1.3f * c->tan_fovx, limy = (real)1.3f * c->tan_fovy;
    real txtz = t[0] / t[2], tytz = t[1] / t[2];
    *in_x = !(txtz < -limx || txtz > limx);
    *in_y = !(tytz < -limy || tytz > limy);
    t[0] = clampr(txtz, -limx, limx) * t[2];
    t[1] = clampr(tytz, -limy, limy) * t[2];
    real Jn[2][3] = {{c->focal_x / t[2], 0, -(c->focal_x * t[0]) / (t[2] * t[2])},
                     {0, c->focal_y / t[2], -(c->focal_y * t[1]) / (t[2] * t[2])}};
    /* Rcw[i][j] = V[4j+i] (row-vector convention => camera rotation is the transpose of the 3x3 block) */
    for (int r = 0; r < 2; r++)
        for (int j = 0; j < 3; j++) {
            real s = 0;
            for (int k = 0; k < 3; k++) s += Jn[r][k] * V[4 * j + k];
            A[r][j] = s;
        }
}

static void sym6_to_mat(const real* s, real Sm[3][3]) {
    Sm[0][0] = s[0]; Sm[0][1] = s[1]; Sm[0][2] = s[2];
    Sm[1][0] = s[1]; Sm[1][1] = s[3]; Sm[1][2] = s[4];
    Sm[2][0] = s[2]; Sm[2][1] = s[4]; Sm[2][2] = s[5];
}

void oracle_free(oracle_ctx* c) {
    if (!c) return;
    free(c->radii); free(c->tiles_touched); free(c->depth); free(c->mean2D); free(c->cov3D); free(c->conic_o);
    free(c->rgb); free(c->clamped); free(c->point_list); free(c->ranges); free(c->final_T); free(c->n_contrib);
    free(c->out_color); free(c->out_depth);
    free(c);
}

oracle_ctx* oracle_forward(int P, int D, int M, const float* bg, int W, int H, const float* means3D,
                           const float* shs, const float* colors_precomp, const float* opacities,
                           const float* scales, float scale_modifier, const float* rotations,
                           const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                           const float* campos, float tan_fovx, float tan_fovy) {
    oracle_ctx* c = (oracle_ctx*)calloc(1, sizeof(oracle_ctx));
    c->P = P; c->D = D; c->M = M; c->W = W; c->H = H;
    c->tx = (W + TILE - 1) / TILE; c->ty = (H + TILE - 1) / TILE;
    c->tan_fovx = tan_fovx; c->tan_fovy = tan_fovy;
    /* focal lengths are derived in float on the host: $R/cuda_rasterizer/rasterizer_impl.cu:222-223 */
    c->focal_y = (real)(H / (2.0f * tan_fovy));
    c->focal_x = (real)(W / (2.0f * tan_fovx));
    c->scale_modifier = scale_modifier;
    for (int i = 0; i < 16; i++) { c->view[i] = viewmatrix[i]; c->proj[i] = projmatrix[i]; }
    for (int i = 0; i < 3; i++) { c->campos[i] = campos[i]; c->bg[i] = bg[i]; }
    c->has_sh = (colors_precomp == NULL);
    c->has_scale_rot = (cov3D_precomp == NULL);
    const size_t N = (size_t)W * H, tiles = (size_t)c->tx * c->ty;
    size_t Pn = P > 0 ? (size_t)P : 1;
    c->radii = (int*)calloc(Pn, sizeof(int));
    c->tiles_touched = (uint32_t*)calloc(Pn, sizeof(uint32_t));
    c->depth = (real*)calloc(Pn, sizeof(real));
    c->mean2D = (real*)calloc(2 * Pn, sizeof(real));
    c->cov3D = (real*)calloc(6 * Pn, sizeof(real));
    c->conic_o = (real*)calloc(4 * Pn, sizeof(real));
    c->rgb = (real*)calloc(3 * Pn, sizeof(real));
    c->clamped = (uint8_t*)calloc(3 * Pn, 1);
    c->ranges = (uint32_t*)calloc(2 * tiles, sizeof(uint32_t));
    c->final_T = (real*)calloc(N, sizeof(real));
    c->n_contrib = (uint32_t*)calloc(N, sizeof(uint32_t));
    c->out_color = (real*)calloc(3 * N, sizeof(real));
    c->out_depth = (real*)calloc(N, sizeof(real));
    if (P == 0) return c; /* zeros everywhere, not background: $R/rasterize_points.cu:67-69,80 */

    /* ---------------- per-Gaussian preprocess: $R/cuda_rasterizer/forward.cu:155-256 ---------------- */
    int64_t R = 0;
#pragma omp parallel for schedule(static) reduction(+ : R)
    for (int i = 0; i < P; i++) {
        const real mean[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
        const real* V = c->view;
        const real* Pm = c->proj;
        real zv = V[2] * mean[0] + V[6] * mean[1] + V[10] * mean[2] + V[14];
        if (zv <= (real)0.2f) continue;
        real hom[4];
        for (int k = 0; k < 4; k++) hom[k] = Pm[k] * mean[0] + Pm[4 + k] * mean[1] + Pm[8 + k] * mean[2] + Pm[12 + k];
        real pw = 1 / (hom[3] + (real)0.0000001f);
        real ndc[2] = {hom[0] * pw, hom[1] * pw};

        real* S6 = c->cov3D + 6 * (size_t)i;
        if (cov3D_precomp) {
            for (int k = 0; k < 6; k++) S6[k] = cov3D_precomp[6 * (size_t)i + k];
        } else {
            real q[4] = {rotations[4 * i], rotations[4 * i + 1], rotations[4 * i + 2], rotations[4 * i + 3]};
            real s[3] = {c->scale_modifier * scales[3 * i], c->scale_modifier * scales[3 * i + 1],
                         c->scale_modifier * scales[3 * i + 2]};
            real Rq[3][3];
            quat_to_R(q, Rq);
            /* Sigma = Rq diag(s^2) Rq^T  ($R/cuda_rasterizer/forward.cu:118-152) */
            real Sg[3][3];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) {
                    real acc = 0;
                    for (int k = 0; k < 3; k++) acc += (Rq[a][k] * s[k]) * (Rq[b][k] * s[k]);
                    Sg[a][b] = acc;
                }
            S6[0] = Sg[0][0]; S6[1] = Sg[0][1]; S6[2] = Sg[0][2]; S6[3] = Sg[1][1]; S6[4] = Sg[1][2]; S6[5] = Sg[2][2];
        }
        real A[2][3], t[3];
        int inx, iny;
        ewa_A(c, mean, A, t, &inx, &iny);
        real Sm[3][3];
        sym6_to_mat(S6, Sm);
        real AS[2][3];
        for (int r = 0; r < 2; r++)
            for (int j = 0; j < 3; j++) AS[r][j] = A[r][0] * Sm[0][j] + A[r][1] * Sm[1][j] + A[r][2] * Sm[2][j];
        real a = AS[0][0] * A[0][0] + AS[0][1] * A[0][1] + AS[0][2] * A[0][2] + (real)0.3f;
        real b = AS[0][0] * A[1][0] + AS[0][1] * A[1][1] + AS[0][2] * A[1][2];
        real cc = AS[1][0] * A[1][0] + AS[1][1] * A[1][1] + AS[1][2] * A[1][2] + (real)0.3f;
        real det = a * cc - b * b;
        if (det == 0) continue;
        real det_inv = 1 / det;
        real mid = (real)0.5 * (a + cc);
        real disc = mid * mid - det;
        if (disc < (real)0.1f) disc = (real)0.1f;
        real l1 = mid + sqrt(disc), l2 = mid - sqrt(disc);
        real lm = l1 > l2 ? l1 : l2;
        real my_radius = ceil(3 * sqrt(lm));
        /* ndc -> pixel in double: $R/cuda_rasterizer/auxiliary.h:41-44 */
        real px = (real)(((double)ndc[0] + 1.0) * W - 1.0) * 0.5;
        real py = (real)(((double)ndc[1] + 1.0) * H - 1.0) * 0.5;
        int x0, y0, x1, y1;
        tile_rect(px, py, (int)my_radius, c->tx, c->ty, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;

        real col[3];
        if (c->has_sh) {
            real dir[3] = {mean[0] - c->campos[0], mean[1] - c->campos[1], mean[2] - c->campos[2]};
            real len = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
            real basis[16], dbasis[16][3];
            sh_basis(D, dir[0] / len, dir[1] / len, dir[2] / len, basis, dbasis);
            int nco = (D + 1) * (D + 1);
            for (int ch = 0; ch < 3; ch++) {
                real acc = 0;
                for (int k = 0; k < nco; k++) acc += basis[k] * shs[((size_t)i * M + k) * 3 + ch];
                acc += (real)0.5;
                c->clamped[3 * (size_t)i + ch] = acc < 0;
                col[ch] = acc < 0 ? 0 : acc;
            }
        } else {
            for (int ch = 0; ch < 3; ch++) col[ch] = colors_precomp[3 * (size_t)i + ch];
        }
        c->depth[i] = zv;
        c->radii[i] = (int)my_radius;
        c->mean2D[2 * (size_t)i] = px;
        c->mean2D[2 * (size_t)i + 1] = py;
        c->conic_o[4 * (size_t)i] = cc * det_inv;
        c->conic_o[4 * (size_t)i + 1] = -b * det_inv;
        c->conic_o[4 * (size_t)i + 2] = a * det_inv;
        c->conic_o[4 * (size_t)i + 3] = opacities[i];
        for (int ch = 0; ch < 3; ch++) c->rgb[3 * (size_t)i + ch] = col[ch];
        c->tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
        R += (x1 - x0) * (y1 - y0);
    }
    c->R = R;

    /* ---------------- binning: $R/cuda_rasterizer/rasterizer_impl.cu:70-138,299-319 ---------------- */
    kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(R > 0 ? R : 1));
    {
        size_t off = 0;
        for (int i = 0; i < P; i++) {
            if (c->radii[i] <= 0) continue;
            int x0, y0, x1, y1;
            tile_rect(c->mean2D[2 * (size_t)i], c->mean2D[2 * (size_t)i + 1], c->radii[i], c->tx, c->ty, &x0, &y0, &x1, &y1);
            float df = (float)c->depth[i];
            uint32_t dbits;
            memcpy(&dbits, &df, 4);
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) {
                    kv[off].key = ((uint64_t)(y * c->tx + x) << 32) | dbits;
                    kv[off].val = (uint32_t)i;
                    off++;
                }
        }
    }
    qsort(kv, (size_t)R, sizeof(kv_t), kv_cmp);
    c->point_list = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(R > 0 ? R : 1));
    for (int64_t k = 0; k < R; k++) {
        c->point_list[k] = kv[k].val;
        uint32_t tile = (uint32_t)(kv[k].key >> 32);
        if (k == 0 || tile != (uint32_t)(kv[k - 1].key >> 32)) c->ranges[2 * tile] = (uint32_t)k;
        c->ranges[2 * tile + 1] = (uint32_t)(k + 1);
    }
    free(kv);

    /* ---------------- compositing: $R/cuda_rasterizer/forward.cu:261-393 ---------------- */
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < (int)tiles; tile++) {
        const int tx0 = (tile % c->tx) * TILE, ty0 = (tile / c->tx) * TILE;
        const uint32_t lo = c->ranges[2 * tile], hi = c->ranges[2 * tile + 1];
        for (int yy = ty0; yy < ty0 + TILE && yy < H; yy++)
            for (int xx = tx0; xx < tx0 + TILE && xx < W; xx++) {
                real T = 1, C[3] = {0, 0, 0}, Dm = 15;
                uint32_t last = 0, contributor = 0;
                for (uint32_t k = lo; k < hi; k++) {
                    contributor++;
                    const uint32_t g = c->point_list[k];
                    const real dx = c->mean2D[2 * (size_t)g] - (real)xx, dy = c->mean2D[2 * (size_t)g + 1] - (real)yy;
                    const real* co = c->conic_o + 4 * (size_t)g;
                    const real power = -(real)0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0) continue;
                    real alpha = co[3] * exp(power);
                    if (alpha > (real)0.99f) alpha = (real)0.99f;
                    if (alpha < (real)(1.0f / 255.0f)) continue;
                    const real test_T = T * (1 - alpha);
                    if (test_T < (real)0.0001f) break;
                    for (int ch = 0; ch < 3; ch++) C[ch] += c->rgb[3 * (size_t)g + ch] * alpha * T;
                    if (T > (real)0.5 && test_T < (real)0.5) Dm = c->depth[g];
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)yy * W + xx;
                c->final_T[pix] = T;
                c->n_contrib[pix] = last;
                for (int ch = 0; ch < 3; ch++) c->out_color[ch * N + pix] = C[ch] + T * c->bg[ch];
                c->out_depth[pix] = Dm;
            }
    }
    return c;
}

/* getters (so the Python wrapper needs no struct layout knowledge) */
int64_t oracle_num_rendered(const oracle_ctx* c) { return c->R; }
const int* oracle_radii(const oracle_ctx* c) { return c->radii; }
const uint32_t* oracle_tiles_touched(const oracle_ctx* c) { return c->tiles_touched; }
const uint32_t* oracle_point_list(const oracle_ctx* c) { return c->point_list; }
const uint32_t* oracle_ranges(const oracle_ctx* c) { return c->ranges; }
const uint32_t* oracle_n_contrib(const oracle_ctx* c) { return c->n_contrib; }
const real* oracle_final_T(const oracle_ctx* c) { return c->final_T; }
const real* oracle_color(const oracle_ctx* c) { return c->out_color; }
const real* oracle_depth_img(const oracle_ctx* c) { return c->out_depth; }
const real* oracle_means2D(const oracle_ctx* c) { return c->mean2D; }
const real* oracle_conic_opacity(const oracle_ctx* c) { return c->conic_o; }
const real* oracle_rgb(const oracle_ctx* c) { return c->rgb; }
const real* oracle_cov3D(const oracle_ctx* c) { return c->cov3D; }
int oracle_real_bytes(void) { return (int)sizeof(real); }

/* Backward.  dL_dpix [3][H][W] (float).  Outputs are `real` arrays sized as in the reference:
 * dmean2D [P][3], dcolor [P][3], dopacity [P], dmean3D [P][3], dcov3D [P][6], dsh [P][M][3],
 * dscale [P][3], drot [P][4]; all fully written. */
void oracle_backward(const oracle_ctx* c, const float* dL_dpix, const float* means3D, const float* shs,
                     const float* scales, const float* rotations, real* dmean2D, real* dcolor, real* dopacity,
                     real* dmean3D, real* dcov3D, real* dsh, real* dscale, real* drot) {
    const int P = c->P, W = c->W, H = c->H, M = c->M, D = c->D;
    const size_t N = (size_t)W * H, tiles = (size_t)c->tx * c->ty;
    memset(dmean2D, 0, sizeof(real) * 3 * (size_t)P);
    memset(dcolor, 0, sizeof(real) * 3 * (size_t)P);
    memset(dopacity, 0, sizeof(real) * (size_t)P);
    memset(dmean3D, 0, sizeof(real) * 3 * (size_t)P);
    memset(dcov3D, 0, sizeof(real) * 6 * (size_t)P);
    if (M > 0) memset(dsh, 0, sizeof(real) * 3 * (size_t)M * P);
    memset(dscale, 0, sizeof(real) * 3 * (size_t)P);
    memset(drot, 0, sizeof(real) * 4 * (size_t)P);
    if (P == 0) return;
    /* true dL/d(A,B,C) of the conic (the reference stores HALF of dL/dB) */
    real* dconic = (real*)calloc(3 * (size_t)P, sizeof(real));

    /* ------------- compositing backward: $R/cuda_rasterizer/backward.cu:399-557 -------------
       Closed form instead of the reference's running recursion:
         C = sum_i c_i a_i T_i + T_N bg,  T_i = prod_{j<i} (1 - a_j)
         dC/da_i = c_i T_i - (sum_{j>i} c_j a_j T_j + T_N bg) / (1 - a_i)                       */
#pragma omp parallel
    {
        uint32_t cap = 1024;
        uint32_t* idx = (uint32_t*)malloc(sizeof(uint32_t) * cap);
        real* al = (real*)malloc(sizeof(real) * cap);
        real* Ts = (real*)malloc(sizeof(real) * cap);
        real* Gs = (real*)malloc(sizeof(real) * cap);
#pragma omp for schedule(dynamic, 4)
        for (int tile = 0; tile < (int)tiles; tile++) {
            const int tx0 = (tile % c->tx) * TILE, ty0 = (tile / c->tx) * TILE;
            const uint32_t lo = c->ranges[2 * tile];
            for (int yy = ty0; yy < ty0 + TILE && yy < H; yy++)
                for (int xx = tx0; xx < tx0 + TILE && xx < W; xx++) {
                    const size_t pix = (size_t)yy * W + xx;
                    const uint32_t last = c->n_contrib[pix];
                    if (last == 0) continue;
                    if (last > cap) {
                        cap = last * 2;
                        idx = (uint32_t*)realloc(idx, sizeof(uint32_t) * cap);
                        al = (real*)realloc(al, sizeof(real) * cap);
                        Ts = (real*)realloc(Ts, sizeof(real) * cap);
                        Gs = (real*)realloc(Gs, sizeof(real) * cap);
                    }
                    const real dpix[3] = {dL_dpix[pix], dL_dpix[N + pix], dL_dpix[2 * N + pix]};
                    /* replay the forward for this pixel over list entries [0, last) */
                    uint32_t n = 0;
                    real T = 1;
                    for (uint32_t k = 0; k < last; k++) {
                        const uint32_t g = c->point_list[lo + k];
                        const real dx = c->mean2D[2 * (size_t)g] - (real)xx, dy = c->mean2D[2 * (size_t)g + 1] - (real)yy;
                        const real* co = c->conic_o + 4 * (size_t)g;
                        const real power = -(real)0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (power > 0) continue;
                        const real G = exp(power);
                        real alpha = co[3] * G;
                        if (alpha > (real)0.99f) alpha = (real)0.99f;
                        if (alpha < (real)(1.0f / 255.0f)) continue;
                        idx[n] = g; al[n] = alpha; Ts[n] = T; Gs[n] = G;
                        n++;
                        T *= (1 - alpha);
                    }
                    /* T now equals final_T of the forward (entries after `last` never blended) */
                    real suffix[3] = {T * c->bg[0], T * c->bg[1], T * c->bg[2]};
                    for (int k = (int)n - 1; k >= 0; k--) {
                        const uint32_t g = idx[k];
                        const real alpha = al[k], Tk = Ts[k], G = Gs[k];
                        const real* col = c->rgb + 3 * (size_t)g;
                        real dL_dalpha = 0;
                        for (int ch = 0; ch < 3; ch++) {
                            dL_dalpha += (col[ch] * Tk - suffix[ch] / (1 - alpha)) * dpix[ch];
#pragma omp atomic
                            dcolor[3 * (size_t)g + ch] += alpha * Tk * dpix[ch];
                            suffix[ch] += col[ch] * alpha * Tk;
                        }
                        const real* co = c->conic_o + 4 * (size_t)g;
                        const real dx = c->mean2D[2 * (size_t)g] - (real)xx, dy = c->mean2D[2 * (size_t)g + 1] - (real)yy;
                        const real dL_dG = co[3] * dL_dalpha; /* 0.99 cap passes the gradient through */
                        /* G = exp(-0.5(A dx^2 + C dy^2) - B dx dy) */
                        const real dG_ddx = -G * (co[0] * dx + co[1] * dy);
                        const real dG_ddy = -G * (co[2] * dy + co[1] * dx);
#pragma omp atomic
                        dmean2D[3 * (size_t)g] += dL_dG * dG_ddx * (real)(0.5 * W);
#pragma omp atomic
                        dmean2D[3 * (size_t)g + 1] += dL_dG * dG_ddy * (real)(0.5 * H);
#pragma omp atomic
                        dconic[3 * (size_t)g] += dL_dG * (-(real)0.5 * G * dx * dx);
#pragma omp atomic
                        dconic[3 * (size_t)g + 1] += dL_dG * (-G * dx * dy);
#pragma omp atomic
                        dconic[3 * (size_t)g + 2] += dL_dG * (-(real)0.5 * G * dy * dy);
#pragma omp atomic
                        dopacity[g] += G * dL_dalpha;
                    }
                }
        }
        free(idx); free(al); free(Ts); free(Gs);
    }

    /* ------------- per-Gaussian backward: $R/cuda_rasterizer/backward.cu:144-396 ------------- */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (!(c->radii[i] > 0)) continue;
        const real mean[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
        const real* V = c->view;
        real A[2][3], t[3];
        int inx, iny;
        ewa_A(c, mean, A, t, &inx, &iny);
        real Sm[3][3];
        sym6_to_mat(c->cov3D + 6 * (size_t)i, Sm);
        real AS[2][3];
        for (int r = 0; r < 2; r++)
            for (int j = 0; j < 3; j++) AS[r][j] = A[r][0] * Sm[0][j] + A[r][1] * Sm[1][j] + A[r][2] * Sm[2][j];
        const real a = AS[0][0] * A[0][0] + AS[0][1] * A[0][1] + AS[0][2] * A[0][2] + (real)0.3f;
        const real b = AS[0][0] * A[1][0] + AS[0][1] * A[1][1] + AS[0][2] * A[1][2];
        const real cc = AS[1][0] * A[1][0] + AS[1][1] * A[1][1] + AS[1][2] * A[1][2] + (real)0.3f;
        const real det = a * cc - b * b;
        const real d2i = 1 / (det * det + (real)0.0000001f);
        const real gA = dconic[3 * (size_t)i], gB = dconic[3 * (size_t)i + 1], gC = dconic[3 * (size_t)i + 2];
        real ga = 0, gb = 0, gc = 0;
        real dSig[6] = {0, 0, 0, 0, 0, 0};
        if (d2i != 0) {
            /* conic = (c, -b, a)/det */
            ga = d2i * (-cc * cc * gA + b * cc * gB + (det - a * cc) * gC);
            gc = d2i * (-a * a * gC + a * b * gB + (det - a * cc) * gA);
            gb = d2i * (2 * b * cc * gA - (det + 2 * b * b) * gB + 2 * a * b * gC);
            /* dL/dSigma = A^T G2 A with G2 = [[ga, gb/2],[gb/2, gc]]; off-diagonals appear twice */
            real G2[2][2] = {{ga, gb / 2}, {gb / 2, gc}};
            real full[3][3];
            for (int p = 0; p < 3; p++)
                for (int q = 0; q < 3; q++) {
                    real s = 0;
                    for (int r = 0; r < 2; r++)
                        for (int u = 0; u < 2; u++) s += A[r][p] * G2[r][u] * A[u][q];
                    full[p][q] = s;
                }
            dSig[0] = full[0][0]; dSig[3] = full[1][1]; dSig[5] = full[2][2];
            dSig[1] = 2 * full[0][1]; dSig[2] = 2 * full[0][2]; dSig[4] = 2 * full[1][2];
        }
        for (int k = 0; k < 6; k++) dcov3D[6 * (size_t)i + k] = dSig[k];
        /* dL/dA = 2 G2 A Sigma */
        real dA[2][3];
        {
            real G2[2][2] = {{ga, gb / 2}, {gb / 2, gc}};
            for (int r = 0; r < 2; r++)
                for (int j = 0; j < 3; j++) dA[r][j] = 2 * (G2[r][0] * AS[0][j] + G2[r][1] * AS[1][j]);
        }
        /* A = Jn Rcw  =>  dL/dJn = dL/dA Rcw^T ;  Rcw[k][j] = V[4j+k] */
        real dJ[2][3];
        for (int r = 0; r < 2; r++)
            for (int k = 0; k < 3; k++) dJ[r][k] = dA[r][0] * V[k] + dA[r][1] * V[4 + k] + dA[r][2] * V[8 + k];
        const real tz = 1 / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const real fx = c->focal_x, fy = c->focal_y;
        real dt[3];
        dt[0] = (inx ? 1 : 0) * (-fx * tz2) * dJ[0][2];
        dt[1] = (iny ? 1 : 0) * (-fy * tz2) * dJ[1][2];
        dt[2] = -fx * tz2 * dJ[0][0] - fy * tz2 * dJ[1][1] + 2 * fx * t[0] * tz3 * dJ[0][2] + 2 * fy * t[1] * tz3 * dJ[1][2];
        real dm[3];
        for (int k = 0; k < 3; k++) dm[k] = V[4 * k] * dt[0] + V[4 * k + 1] * dt[1] + V[4 * k + 2] * dt[2];

        /* projection: ndc = hom.xy / (hom.w + 1e-7)   ($R/cuda_rasterizer/backward.cu:365-387) */
        {
            const real* Pm = c->proj;
            real hom[4];
            for (int k = 0; k < 4; k++) hom[k] = Pm[k] * mean[0] + Pm[4 + k] * mean[1] + Pm[8 + k] * mean[2] + Pm[12 + k];
            const real mw = 1 / (hom[3] + (real)0.0000001f);
            const real gx = dmean2D[3 * (size_t)i], gy = dmean2D[3 * (size_t)i + 1];
            for (int k = 0; k < 3; k++) {
                const real dndcx = Pm[4 * k] * mw - Pm[4 * k + 3] * hom[0] * mw * mw;
                const real dndcy = Pm[4 * k + 1] * mw - Pm[4 * k + 3] * hom[1] * mw * mw;
                dm[k] += dndcx * gx + dndcy * gy;
            }
        }
        /* SH: $R/cuda_rasterizer/backward.cu:20-139 */
        if (c->has_sh) {
            real dirv[3] = {mean[0] - c->campos[0], mean[1] - c->campos[1], mean[2] - c->campos[2]};
            real len = sqrt(dirv[0] * dirv[0] + dirv[1] * dirv[1] + dirv[2] * dirv[2]);
            real u[3] = {dirv[0] / len, dirv[1] / len, dirv[2] / len};
            real basis[16], dbasis[16][3];
            sh_basis(D, u[0], u[1], u[2], basis, dbasis);
            real dRGB[3];
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = c->clamped[3 * (size_t)i + ch] ? 0 : dcolor[3 * (size_t)i + ch];
            int nco = (D + 1) * (D + 1);
            real ddir[3] = {0, 0, 0};
            for (int k = 0; k < nco; k++) {
                real dotc = 0;
                for (int ch = 0; ch < 3; ch++) {
                    dsh[((size_t)i * M + k) * 3 + ch] = basis[k] * dRGB[ch];
                    dotc += shs[((size_t)i * M + k) * 3 + ch] * dRGB[ch];
                }
                for (int ax = 0; ax < 3; ax++) ddir[ax] += dbasis[k][ax] * dotc;
            }
            /* through normalisation: (I - u u^T)/len */
            real udot = u[0] * ddir[0] + u[1] * ddir[1] + u[2] * ddir[2];
            for (int k = 0; k < 3; k++) dm[k] += (ddir[k] - u[k] * udot) / len;
        }
        for (int k = 0; k < 3; k++) dmean3D[3 * (size_t)i + k] = dm[k];

        /* Sigma = Rq diag(s^2) Rq^T: $R/cuda_rasterizer/backward.cu:278-341 */
        if (c->has_scale_rot) {
            real q[4] = {rotations[4 * i], rotations[4 * i + 1], rotations[4 * i + 2], rotations[4 * i + 3]};
            real s[3] = {c->scale_modifier * scales[3 * i], c->scale_modifier * scales[3 * i + 1],
                         c->scale_modifier * scales[3 * i + 2]};
            real Rq[3][3];
            quat_to_R(q, Rq);
            real Gs[3][3] = {{dSig[0], dSig[1] / 2, dSig[2] / 2}, {dSig[1] / 2, dSig[3], dSig[4] / 2},
                             {dSig[2] / 2, dSig[4] / 2, dSig[5]}};
            /* dL/dRq = 2 Gs Rq S^2 ;  dL/ds_k = 2 s_k (Rq^T Gs Rq)_kk */
            real GR[3][3];
            for (int p = 0; p < 3; p++)
                for (int k = 0; k < 3; k++) GR[p][k] = Gs[p][0] * Rq[0][k] + Gs[p][1] * Rq[1][k] + Gs[p][2] * Rq[2][k];
            for (int k = 0; k < 3; k++) {
                real rgr = Rq[0][k] * GR[0][k] + Rq[1][k] * GR[1][k] + Rq[2][k] * GR[2][k];
                /* gradient w.r.t. the *input* scale (before scale_modifier)?  the reference returns
                   d/d(modified scale): dL_dscale = dot(Rt[k], dL_dMt[k]) with no modifier factor */
                dscale[3 * (size_t)i + k] = 2 * s[k] * rgr;
            }
            real dR[3][3];
            for (int p = 0; p < 3; p++)
                for (int k = 0; k < 3; k++) dR[p][k] = 2 * GR[p][k] * s[k] * s[k];
            const real r = q[0], x = q[1], y = q[2], z = q[3];
            const real dRdr[3][3] = {{0, -2 * z, 2 * y}, {2 * z, 0, -2 * x}, {-2 * y, 2 * x, 0}};
            const real dRdx[3][3] = {{0, 2 * y, 2 * z}, {2 * y, -4 * x, -2 * r}, {2 * z, 2 * r, -4 * x}};
            const real dRdy[3][3] = {{-4 * y, 2 * x, 2 * r}, {2 * x, 0, 2 * z}, {-2 * r, 2 * z, -4 * y}};
            const real dRdz[3][3] = {{-4 * z, -2 * r, 2 * x}, {2 * r, -4 * z, 2 * y}, {2 * x, 2 * y, 0}};
            real gq[4] = {0, 0, 0, 0};
            for (int p = 0; p < 3; p++)
                for (int k = 0; k < 3; k++) {
                    gq[0] += dR[p][k] * dRdr[p][k];
                    gq[1] += dR[p][k] * dRdx[p][k];
                    gq[2] += dR[p][k] * dRdy[p][k];
                    gq[3] += dR[p][k] * dRdz[p][k];
                }
            for (int k = 0; k < 4; k++) drot[4 * (size_t)i + k] = gq[k];
        }
    }
    free(dconic);
}

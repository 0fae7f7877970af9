// Forward per-Gaussian preprocess + markVisible.
//
// The per-Gaussian arithmetic (eval_sh, cov3d_from_scale_rot, cov2d_from_cov3d) deliberately keeps the reference's
// expression trees — same association, same FMA contraction points — because radii / tile counts must be bit-identical
// (SURVEY.md section 7 "hard parts"); the kernel around it (up-front loads, coalesced SH rows through shared memory,
// alpha-box rect clip, fused housekeeping) is new.  Rules followed:
//   near cull  z_view <= 0.2                          $R/cuda_rasterizer/auxiliary.h:139-164
//   projection p_hom, p_w = 1/(w + 1e-7)              $R/cuda_rasterizer/forward.cu:196-200
//   cov3D = (S R)^T (S R)                             $R/cuda_rasterizer/forward.cu:118-152
//   EWA cov2D, +0.3 low-pass                          $R/cuda_rasterizer/forward.cu:74-113
//   conic, 3-sigma radius, tile rect                  $R/cuda_rasterizer/forward.cu:216-237
//   SH(deg<=3) -> RGB, +0.5, clamp mask               $R/cuda_rasterizer/forward.cu:20-71
//
// B200 design differences: per-view constants come from the kernel-parameter constant
// bank (no per-thread pointer chasing); SH rows are fetched with 128-bit loads; all
// outputs are SoA and coalesced; colour+depth are packed into one float4 so the render
// kernel stages an instance with three 16/8-byte gathers; the depth-sort key of the
// two-level binning (see sgs_binning.cu) is emitted here.
#include "sgs_common.cuh"
#include <cstdio>

namespace sgs {

struct V3 {
    float x, y, z;
};
__forceinline__ __device__ V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__forceinline__ __device__ V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__forceinline__ __device__ V3 operator*(float s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
__forceinline__ __device__ V3 operator+(const V3& a, float s) { return {a.x + s, a.y + s, a.z + s}; }

template <bool VEC_SH>
struct ShRow {
    const float* base;  // row of this Gaussian: [M][3]
    __forceinline__ __device__ V3 get(int k) const { return {base[3 * k], base[3 * k + 1], base[3 * k + 2]}; }
};

// SH evaluation on a row already held in registers (sh[k] as V3).
__forceinline__ __device__ V3 eval_sh(int deg, const V3* sh, const V3& pos, const float* campos,
                                      uint8_t& clamp_mask) {
    V3 dir = {pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]};
    float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir = {dir.x / len, dir.y / len, dir.z / len};

    V3 result = SGS_SH_C0 * sh[0];
    if (deg > 0) {
        float x = dir.x, y = dir.y, z = dir.z;
        result = result - SGS_SH_C1 * y * sh[1] + SGS_SH_C1 * z * sh[2] - SGS_SH_C1 * x * sh[3];
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;
            result = result + SGS_SH_C2_0 * xy * sh[4] + SGS_SH_C2_1 * yz * sh[5] +
                     SGS_SH_C2_2 * (2.0f * zz - xx - yy) * sh[6] + SGS_SH_C2_3 * xz * sh[7] +
                     SGS_SH_C2_4 * (xx - yy) * sh[8];
            if (deg > 2) {
                result = result + SGS_SH_C3_0 * y * (3.0f * xx - yy) * sh[9] + SGS_SH_C3_1 * xy * z * sh[10] +
                         SGS_SH_C3_2 * y * (4.0f * zz - xx - yy) * sh[11] +
                         SGS_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12] +
                         SGS_SH_C3_4 * x * (4.0f * zz - xx - yy) * sh[13] + SGS_SH_C3_5 * z * (xx - yy) * sh[14] +
                         SGS_SH_C3_6 * x * (xx - 3.0f * yy) * sh[15];
            }
        }
    }
    result = result + 0.5f;
    clamp_mask = (uint8_t)((result.x < 0 ? 1 : 0) | (result.y < 0 ? 2 : 0) | (result.z < 0 ? 4 : 0));
    return {fmaxf(result.x, 0.0f), fmaxf(result.y, 0.0f), fmaxf(result.z, 0.0f)};
}

__forceinline__ __device__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 rot, float* cov3D) {
    Mat3 S = mat3_cols(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
    S.c[0][0] = mod * scale.x;
    S.c[1][1] = mod * scale.y;
    S.c[2][2] = mod * scale.z;
    // quaternion used as given (callers normalise): (r, x, y, z)
    float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
    Mat3 R = mat3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                       2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                       2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    Mat3 M = mat3_mul(S, R);
    Mat3 Sigma = mat3_mul(mat3_T(M), M);
    cov3D[0] = Sigma.c[0][0];
    cov3D[1] = Sigma.c[0][1];
    cov3D[2] = Sigma.c[0][2];
    cov3D[3] = Sigma.c[1][1];
    cov3D[4] = Sigma.c[1][2];
    cov3D[5] = Sigma.c[2][2];
}

__forceinline__ __device__ float3 cov2d_from_cov3d(const float3& mean, float focal_x, float focal_y, float tan_fovx,
                                                   float tan_fovy, const float* cov3D, const float* view) {
    float3 t = xform_point_4x3(mean, view);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = min(limx, max(-limx, txtz)) * t.z;
    t.y = min(limy, max(-limy, tytz)) * t.z;

    Mat3 J = mat3_cols(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
                       0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
                       0, 0, 0);
    Mat3 Wm = mat3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    Mat3 T = mat3_mul(Wm, J);
    Mat3 Vrk = mat3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    Mat3 cov = mat3_mul(mat3_mul(mat3_T(T), mat3_T(Vrk)), T);
    cov.c[0][0] += 0.3f;
    cov.c[1][1] += 0.3f;
    return {cov.c[0][0], cov.c[0][1], cov.c[1][1]};
}

// Clip the reference's tile rect [rmin, rmax) to the tiles that contain at least one pixel able to reach
// alpha >= 1/255 (the reference `continue`s on every other pixel, $R/cuda_rasterizer/forward.cu:342-352).
//   alpha = o exp(power) >= 1/255  <=>  q(d) = -power <= tau = ln(255 o)
//   max |dx| on the ellipse q(d) <= tau is sqrt(2 tau / (A - B^2/C)), max |dy| = sqrt(2 tau / (C - B^2/A)).
// Everything is conservative: tau is inflated by a bound on the float error of `power` anywhere inside the
// rect, the Schur complements are deflated by their own rounding error, the extents get a relative and an
// absolute margin, and anything numerically unusual keeps the full rect.
__forceinline__ __device__ void clip_rect_to_alpha_box(float2 p, float3 conic, float o, float radius, uint2& rmin,
                                                       uint2& rmax) {
    const float A = conic.x, B = conic.y, C = conic.z;
    if (!(A > 0.f && C > 0.f) || !(fabsf(p.x) < 1e6f && fabsf(p.y) < 1e6f) || !(fabsf(B) < 1e15f)) return;
    if (!(o > 0.f)) {
        if (o == o) rmax = rmin;          // alpha <= 0 < 1/255 everywhere (NaN opacity: keep)
        return;
    }
    const float ext = radius + 32.f;      // |dx|, |dy| of any pixel of any tile of the rect
    const float tau = logf(255.f * o) + 2e-5f * (A + C) * ext * ext + 1e-3f;
    if (!(tau < 1e30f)) return;           // NaN / inf
    if (tau < 0.f) {                      // o exp(power <= 0) < 1/255: no pixel can blend
        rmax = rmin;
        return;
    }
    const float Sx = A - B * (B / C) - 1e-6f * A;
    const float Sy = C - B * (B / A) - 1e-6f * C;
    if (Sx > 0.f) {
        const float ex = sqrtf(2.f * tau / Sx) * 1.00001f + 1e-2f;
        if (ex < 1e6f) {
            const float lo = floorf((p.x - ex) * (1.f / SGS_TILE_X)), hi = floorf((p.x + ex) * (1.f / SGS_TILE_X)) + 1.f;
            rmin.x = max(rmin.x, (uint32_t)fminf(fmaxf(lo, 0.f), 65535.f));
            rmax.x = min(rmax.x, (uint32_t)fminf(fmaxf(hi, 0.f), 65535.f));
        }
    }
    if (Sy > 0.f) {
        const float ey = sqrtf(2.f * tau / Sy) * 1.00001f + 1e-2f;
        if (ey < 1e6f) {
            const float lo = floorf((p.y - ey) * (1.f / SGS_TILE_Y)), hi = floorf((p.y + ey) * (1.f / SGS_TILE_Y)) + 1.f;
            rmin.y = max(rmin.y, (uint32_t)fminf(fmaxf(lo, 0.f), 65535.f));
            rmax.y = min(rmax.y, (uint32_t)fminf(fmaxf(hi, 0.f), 65535.f));
        }
    }
    if (rmax.x <= rmin.x || rmax.y <= rmin.y) rmax = rmin;
}

#define SGS_PRE_THREADS 128
#define SGS_SH_ROW4 12          // float4 per 16-coefficient SH row (192 B)
#define SGS_SH_PAD4 13          // padded row stride in shared memory (conflict-free 128-bit accesses)

// Warp-cooperative, fully coalesced load of the 32 SH rows of a warp (6 KB contiguous) into shared memory.
__forceinline__ __device__ void load_sh_rows(float4* s_rows, const float* __restrict__ shs, int first_row, int nrows) {
    const int lane = threadIdx.x & 31;
    const float4* src = reinterpret_cast<const float4*>(shs + (size_t)first_row * 48);
    const int total = nrows * SGS_SH_ROW4;
#pragma unroll
    for (int it = 0; it < SGS_SH_ROW4; it++) {
        const int e = it * 32 + lane;
        if (e < total) {
            const int row = e / SGS_SH_ROW4, c = e - row * SGS_SH_ROW4;
            s_rows[row * SGS_SH_PAD4 + c] = __ldg(src + e);
        }
    }
    __syncwarp();
}

template <bool VEC_SH>
__global__ void __launch_bounds__(SGS_PRE_THREADS)
preprocess_fwd_kernel(int P, const __grid_constant__ ViewParams vp, const float* __restrict__ means3D,
                      const float* __restrict__ scales, const float* __restrict__ rotations,
                      const float* __restrict__ opacities, const float* __restrict__ shs,
                      const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp,
                      int* __restrict__ radii, GeomState g, uint32_t* __restrict__ zero_words, uint32_t n_zero,
                      int cull, int rot_vec) {
    __shared__ ViewSmem cam;
    __shared__ float4 s_sh[VEC_SH ? (SGS_PRE_THREADS / 32) * 32 * SGS_SH_PAD4 : 1];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < P;

    // Housekeeping that used to be two memsets: this kernel precedes every consumer in stream order, so it zeroes
    // the binning control block (barrier counters, key range, totals) and the per-tile ranges / counts.
    if (blockIdx.x == 0 && threadIdx.x < sizeof(BinCtl) / 4) reinterpret_cast<uint32_t*>(g.ctl)[threadIdx.x] = 0u;
    for (uint32_t i = (uint32_t)idx; i < n_zero; i += gridDim.x * blockDim.x) zero_words[i] = 0u;

    // The kernel is latency-bound (ncu: 44 % of the samples wait on global loads at 38 % occupancy), so every load a
    // Gaussian may need is issued up front — its own parameters and the warp's 32 SH rows (coalesced, through shared
    // memory) — before the camera staging barrier and before the first dependent instruction.
    float3 pre_p = {0.f, 0.f, 0.f}, pre_sc = {0.f, 0.f, 0.f};
    float4 pre_q = {1.f, 0.f, 0.f, 0.f};
    float pre_o = 0.f;
    if (valid) {
        pre_p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
        pre_o = opacities[idx];
        if (cov3D_precomp == nullptr) {
            pre_sc = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
            if (rot_vec) pre_q = reinterpret_cast<const float4*>(rotations)[idx];
            else pre_q = make_float4(rotations[4 * idx], rotations[4 * idx + 1], rotations[4 * idx + 2], rotations[4 * idx + 3]);
        }
    }
    if (VEC_SH && colors_precomp == nullptr) {
        const int warp = threadIdx.x >> 5;
        const int first_row = blockIdx.x * blockDim.x + warp * 32;
        if (first_row < P) load_sh_rows(s_sh + warp * 32 * SGS_SH_PAD4, shs, first_row, min(32, P - first_row));
    }
    stage_view(cam, vp);

    // defaults for a Gaussian that takes no further part
    int out_radius = 0;
    uint32_t out_tiles = 0;
    uint32_t out_key = 0xFFFFFFFFu;
    ushort4 out_rect = {0, 0, 0, 0};
    bool visible = false;
    float3 p_orig = {0.f, 0.f, 0.f};
    float depth = 0.f;

    if (valid) {
        p_orig = pre_p;
        const float3 p_view = xform_point_4x3(p_orig, cam.view);
        depth = p_view.z;

        bool alive = !(p_view.z <= 0.2f);
        if (!alive && vp.prefiltered) {
            printf("Point is filtered although prefiltered is set. This shouldn't happen!");
            __trap();
        }

        if (alive) {
            float4 p_hom = xform_point_4x4(p_orig, cam.proj);
            float p_w = 1.0f / (p_hom.w + 0.0000001f);
            float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

            float cov3D[6];
            if (cov3D_precomp != nullptr) {
#pragma unroll
                for (int i = 0; i < 6; i++) cov3D[i] = cov3D_precomp[6 * idx + i];
            } else {
                cov3d_from_scale_rot(pre_sc, vp.scale_modifier, pre_q, cov3D);
#pragma unroll
                for (int i = 0; i < 6; i++) g.cov3D[6 * idx + i] = cov3D[i];
            }

            float3 cov = cov2d_from_cov3d(p_orig, vp.focal_x, vp.focal_y, vp.tan_fovx, vp.tan_fovy, cov3D, cam.view);

            float det = (cov.x * cov.z - cov.y * cov.y);
            if (det != 0.0f) {
                float det_inv = 1.f / det;
                float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

                float mid = 0.5f * (cov.x + cov.z);
                float lambda1 = mid + sqrtf(max(0.1f, mid * mid - det));
                float lambda2 = mid - sqrtf(max(0.1f, mid * mid - det));
                float my_radius = ceilf(3.f * sqrtf(max(lambda1, lambda2)));
                float2 point_image = {ndc2pix(p_proj.x, vp.W), ndc2pix(p_proj.y, vp.H)};
                uint2 rmin, rmax;
                get_rect(point_image, (int)my_radius, rmin, rmax, vp.tiles_x, vp.tiles_y);
                uint32_t ntiles = (rmax.x - rmin.x) * (rmax.y - rmin.y);
                if (ntiles != 0) {
                    visible = true;
                    const float o = pre_o;
                    g.depths[idx] = p_view.z;
                    g.means2D[idx] = point_image;
                    g.conic_opacity[idx] = make_float4(conic.x, conic.y, conic.z, o);
                    out_radius = (int)my_radius;
                    out_tiles = ntiles;
                    if (cull) clip_rect_to_alpha_box(point_image, conic, o, my_radius, rmin, rmax);
                    out_rect = make_ushort4((unsigned short)rmin.x, (unsigned short)rmax.x, (unsigned short)rmin.y,
                                            (unsigned short)rmax.y);
                    out_key = __float_as_uint(p_view.z);
                }
            }
        }
    }

    // ---- colour: SH rows are fetched by the whole warp (coalesced) when any of its Gaussians is visible
    if (colors_precomp == nullptr) {
        V3 sh[16];
        if (VEC_SH) {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            float4* rows = s_sh + warp * 32 * SGS_SH_PAD4;
            {
                if (visible) {
                    float f[48];
#pragma unroll
                    for (int k = 0; k < SGS_SH_ROW4; k++) {
                        const float4 v = rows[lane * SGS_SH_PAD4 + k];
                        f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
                    }
#pragma unroll
                    for (int k = 0; k < 16; k++) sh[k] = {f[3 * k], f[3 * k + 1], f[3 * k + 2]};
                }
            }
        } else if (visible) {
            const int M = vp.sh_coeffs;
            const float* row = shs + (size_t)idx * M * 3;
            const int ncoef = (vp.sh_degree + 1) * (vp.sh_degree + 1);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k < ncoef) sh[k] = {row[3 * k], row[3 * k + 1], row[3 * k + 2]};
                else sh[k] = {0.f, 0.f, 0.f};
            }
        }
        if (visible) {
            uint8_t cmask = 0;
            V3 c = eval_sh(vp.sh_degree, sh, V3{p_orig.x, p_orig.y, p_orig.z}, cam.campos, cmask);
            g.rgbd[idx] = make_float4(c.x, c.y, c.z, depth);
            g.clamped[idx] = cmask;
        }
    } else if (visible) {
        g.rgbd[idx] = make_float4(colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2], depth);
        g.clamped[idx] = 0;
    }

    // per-block range of the visible depth keys: the depth-sort kernel normalises its keys to the frame's
    // [min, max] (fewer radix passes) and gets that range from these partials without a pass over the keys
    {
        __shared__ uint32_t s_kmax[SGS_PRE_THREADS / 32], s_knmin[SGS_PRE_THREADS / 32];
        __shared__ uint32_t s_kept[SGS_PRE_THREADS / 32], s_touched[SGS_PRE_THREADS / 32], s_vis[SGS_PRE_THREADS / 32];
        const uint32_t kx = __reduce_max_sync(0xFFFFFFFFu, visible ? out_key : 0u);
        const uint32_t kn = __reduce_max_sync(0xFFFFFFFFu, visible ? ~out_key : 0u);
        // the instance totals do not depend on the depth order: block partials here, reduced by the binning kernel
        const uint32_t area = (uint32_t)(out_rect.y - out_rect.x) * (uint32_t)(out_rect.w - out_rect.z);
        const uint32_t sk = __reduce_add_sync(0xFFFFFFFFu, area);
        const uint32_t st = __reduce_add_sync(0xFFFFFFFFu, out_tiles);
        const uint32_t sv = __reduce_add_sync(0xFFFFFFFFu, visible ? 1u : 0u);
        if ((threadIdx.x & 31) == 0) {
            s_kmax[threadIdx.x >> 5] = kx;
            s_knmin[threadIdx.x >> 5] = kn;
            s_kept[threadIdx.x >> 5] = sk;
            s_touched[threadIdx.x >> 5] = st;
            s_vis[threadIdx.x >> 5] = sv;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t a = 0u, b = 0u, c = 0u, d = 0u, e = 0u;
#pragma unroll
            for (int w = 0; w < SGS_PRE_THREADS / 32; w++) {
                a = max(a, s_kmax[w]);
                b = max(b, s_knmin[w]);
                c += s_kept[w];
                d += s_touched[w];
                e += s_vis[w];
            }
            g.blk_range[blockIdx.x] = make_uint2(a, b);
            g.blk_sums[blockIdx.x] = make_uint4(c, d, e, 0u);
        }
    }
    if (!valid) return;
    radii[idx] = out_radius;
    g.tiles_touched[idx] = out_tiles;
    g.rect_kept[idx] = out_rect;
    g.depth_raw[idx] = out_key;
}

int preprocess_blocks(int P) { return (P + SGS_PRE_THREADS - 1) / SGS_PRE_THREADS; }

void launch_preprocess_fwd(int P, const ViewParams& vp, const float* means3D, const float* scales,
                           const float* rotations, const float* opacities, const float* shs,
                           const float* cov3D_precomp, const float* colors_precomp, int* radii, GeomState g,
                           uint32_t* zero_words, size_t n_zero, int cull, cudaStream_t s) {
    if (P <= 0) return;
    const int block = SGS_PRE_THREADS;
    const int grid = preprocess_blocks(P);
    const bool vec = (shs != nullptr) && vp.sh_coeffs == 16 && ((reinterpret_cast<size_t>(shs) & 15) == 0);
    // 128-bit quaternion loads only when the caller's pointer allows them (a torch view with a storage offset, or a
    // plain C caller, may hand over a 4-byte-aligned array: include/saro_gs_b200.h promises to accept that)
    const int rot_vec = (reinterpret_cast<size_t>(rotations) & 15) == 0 ? 1 : 0;
    if (vec)
        preprocess_fwd_kernel<true><<<grid, block, 0, s>>>(P, vp, means3D, scales, rotations, opacities, shs,
                                                          cov3D_precomp, colors_precomp, radii, g, zero_words,
                                                          (uint32_t)n_zero, cull, rot_vec);
    else
        preprocess_fwd_kernel<false><<<grid, block, 0, s>>>(P, vp, means3D, scales, rotations, opacities, shs,
                                                           cov3D_precomp, colors_precomp, radii, g, zero_words,
                                                           (uint32_t)n_zero, cull, rot_vec);
}

// markVisible: bool per point = (z_view > 0.2).  $R/cuda_rasterizer/rasterizer_impl.cu:54-66,141-153
__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    const float3 pv = xform_point_4x3(p, view);
    present[idx] = (pv.z <= 0.2f) ? 0 : 1;
}

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s) {
    if (P <= 0) return;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
}

}  // namespace sgs

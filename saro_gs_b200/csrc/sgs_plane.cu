// Scale-aware plane sampler: the residual-field lookup of the reference,
//   ScaleAwareResField.forward / get_density / get_level        scene/hexplane.py:231-286
//   interpolate_ms_features (6 coordinate planes, summed)       scene/hexplane.py:91-137
//   grid_sample_wrapper -> nvdiffrast.torch.texture(mip_level_bias = min level of the two axes,
//                          boundary_mode = "clamp", max_mip_level = 7 (space planes) | 0 (time planes))   :26-60
// nvdiffrast is an un-vendored third-party package; the op is rebuilt here from its published algorithm
// (linear-mipmap-linear with an explicit level, texel centres at half-integers, clamp-to-edge, 2x2 box mip stack) —
// see oracle/plane_oracle.py for the statement this kernel is tested against.
//
// B200 design:
//   * the planes are nn.Parameters in NCHW; sampling them there costs C scattered sectors per tap.  A CHANNELS-LAST
//     MIP PYRAMID (one 128-byte line per texel at C = 32) is built once per parameter update (plane_build_*), not
//     once per call as the reference's permute(0,2,3,1).contiguous() + mip construction does (hexplane.py:35,49);
//   * one kernel samples all six planes of a resolution level: a group of min(C, 32) lanes owns one point, a lane owns
//     a channel; all eight taps of a plane (two mip levels x four texels) are issued before the first use; the sum
//     over planes stays in registers and is written once, in the [N, C_total] layout the MLPs consume;
//   * backward (the planes are the only inputs that carry a gradient: the reference samples at detached positions and
//     scales, scene/saro_gaussian.py:765,780,865): the same traversal scatters w * dL/dout with one RED per lane into a
//     channels-last gradient pyramid, and one fold kernel applies the adjoint of the box filters and transposes back
//     to NCHW through shared memory.
#include "../../include/saro_gs_b200.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>

namespace sgs_plane {

constexpr int kMaxLevels = 8;   // max_mip_level 7 => 8 levels

struct PlaneDev {
    float* lv[kMaxLevels];   // channels-last levels [h][w][C]
    int H, W;                // level-0 extents
    int du, dv;              // coordinate index (0 x, 1 y, 2 z, 3 t) driving the W axis / the H axis
    int levels;              // levels present (1 .. 8)
};

struct FieldParams {
    PlaneDev planes[6];
    int n_planes;
    const float* pts;         // [N][3]
    const float* timestamps;  // [N]
    const float* scales;      // [N][3]
    const float* aabb;        // [2][3] device: row 0 = xyz_max, row 1 = xyz_min (as set_aabb stores them)
    const float* base_scale;  // [3] device
    float time_scale;         // duration / (duration - 1)
    int reso0[3];             // resolution of the coarsest grid (get_level's clamp)
    int N, C;
    int out_stride, out_offset;
};

__host__ __device__ __forceinline__ int level_extent(int e, int l) {
    const int v = e >> l;
    return v > 0 ? v : 1;
}

struct Taps {
    int idx[4];     // texel index (iv * w + iu) of the four taps
    float w[4];
};

__device__ __forceinline__ Taps make_taps(float ux, float uy, int h, int w, float scale) {
    float u = fminf(fmaxf(ux * (float)w - 0.5f, 0.f), (float)w - 1.f);
    float v = fminf(fmaxf(uy * (float)h - 0.5f, 0.f), (float)h - 1.f);
    const int iu0 = (int)floorf(u), iv0 = (int)floorf(v);
    const float fu = u - (float)iu0, fv = v - (float)iv0;
    const int iu1 = min(iu0 + 1, w - 1), iv1 = min(iv0 + 1, h - 1);
    Taps t;
    t.idx[0] = iv0 * w + iu0;
    t.idx[1] = iv0 * w + iu1;
    t.idx[2] = iv1 * w + iu0;
    t.idx[3] = iv1 * w + iu1;
    t.w[0] = (1.f - fu) * (1.f - fv) * scale;
    t.w[1] = fu * (1.f - fv) * scale;
    t.w[2] = (1.f - fu) * fv * scale;
    t.w[3] = fu * fv * scale;
    return t;
}

// normalised 4-D coordinate and per-axis mip level of one point (scene/hexplane.py:19-23, 231-242)
__device__ __forceinline__ bool point_setup(const FieldParams& p, int n, float p4[4], float level[4]) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float mx = p.aabb[i], mn = p.aabb[3 + i];
        p4[i] = (p.pts[3 * (size_t)n + i] - mx) / (mn - mx);
        const float bs = p.base_scale[i];
        const float min_scale = bs / 2.f;
        const float max_scale = min_scale * (float)p.reso0[i];
        const float s = fminf(fmaxf(p.scales[3 * (size_t)n + i], min_scale), max_scale);
        level[i] = log2f(2.f * s / bs);
    }
    p4[3] = p.timestamps[n] * p.time_scale;
    level[3] = 0.f;
    return true;
}

// BACKWARD == false: out[n][c] = sum over planes of the mip-blended bilinear sample
// BACKWARD == true : scatter  weight * dout[n][c]  into the (gradient) pyramids
template <bool BACKWARD>
__global__ void __launch_bounds__(256) plane_field_kernel(const __grid_constant__ FieldParams p, float* __restrict__ out,
                                                          const float* __restrict__ dout) {
    const int C = p.C;
    const int lpp = C < 32 ? C : 32;                 // lanes per point
    const int gthread = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = gthread / lpp;
    const int c0 = gthread % lpp;
    if (n >= p.N) return;
    float p4[4], level[4];
    point_setup(p, n, p4, level);

    for (int cb = c0; cb < C; cb += 32) {             // one iteration unless C > 32
        float acc = 0.f;
        float g = 0.f;
        if (BACKWARD) g = dout[(size_t)n * p.out_stride + p.out_offset + cb];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (k >= p.n_planes) break;
            const PlaneDev& pl = p.planes[k];
            const float ux = p4[pl.du], uy = p4[pl.dv];
            const float bias = fminf(level[pl.du], level[pl.dv]);
            const float top = (float)(pl.levels - 1);
            const float lvl = fminf(fmaxf(bias, 0.f), top);
            const int l0 = (int)floorf(lvl);
            const float f = lvl - (float)l0;
            const int l1 = min(l0 + 1, pl.levels - 1);
            const Taps t0 = make_taps(ux, uy, level_extent(pl.H, l0), level_extent(pl.W, l0), 1.f - f);
            const Taps t1 = make_taps(ux, uy, level_extent(pl.H, l1), level_extent(pl.W, l1), f);
            float* b0 = pl.lv[l0];
            float* b1 = pl.lv[l1];
            if (!BACKWARD) {
                float v0[4], v1[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v0[j] = __ldg(b0 + (size_t)t0.idx[j] * C + cb);
                if (f > 0.f) {
#pragma unroll
                    for (int j = 0; j < 4; j++) v1[j] = __ldg(b1 + (size_t)t1.idx[j] * C + cb);
                }
                float s = t0.w[0] * v0[0] + t0.w[1] * v0[1] + t0.w[2] * v0[2] + t0.w[3] * v0[3];
                if (f > 0.f) s += t1.w[0] * v1[0] + t1.w[1] * v1[1] + t1.w[2] * v1[2] + t1.w[3] * v1[3];
                acc += s;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (t0.w[j] != 0.f) atomicAdd(b0 + (size_t)t0.idx[j] * C + cb, t0.w[j] * g);
                if (f > 0.f) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (t1.w[j] != 0.f) atomicAdd(b1 + (size_t)t1.idx[j] * C + cb, t1.w[j] * g);
                }
            }
        }
        if (!BACKWARD) out[(size_t)n * p.out_stride + p.out_offset + cb] = acc;
    }
}

// Four channels per lane (C % 4 == 0, 16-byte aligned rows): a group of C / 4 lanes (8 at C = 32) owns one point, so
// a warp serves four points per instruction instead of one — a quarter of the tap-address arithmetic, LDG.128 instead
// of four LDG.32 per tap, and in backward ONE red.global.add.v4.f32 per lane and tap instead of four scalar REDs.
// Same arithmetic per channel as plane_field_kernel.
template <bool BACKWARD>
__global__ void __launch_bounds__(256) plane_field_kernel_v4(const __grid_constant__ FieldParams p, float* __restrict__ out,
                                                             const float* __restrict__ dout) {
    const int C = p.C;
    const int lpp = C / 4 < 8 ? C / 4 : 8;           // lanes per point (a power of two: C is 8, 16, 24 -> 6?, 32)
    const int gthread = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = gthread / lpp;
    const int q0 = gthread % lpp;
    if (n >= p.N) return;
    float p4[4], level[4];
    point_setup(p, n, p4, level);

    for (int cb = q0 * 4; cb < C; cb += lpp * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 g = acc;
        if (BACKWARD) g = *reinterpret_cast<const float4*>(dout + (size_t)n * p.out_stride + p.out_offset + cb);
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (k >= p.n_planes) break;
            const PlaneDev& pl = p.planes[k];
            const float ux = p4[pl.du], uy = p4[pl.dv];
            const float bias = fminf(level[pl.du], level[pl.dv]);
            const float top = (float)(pl.levels - 1);
            const float lvl = fminf(fmaxf(bias, 0.f), top);
            const int l0 = (int)floorf(lvl);
            const float f = lvl - (float)l0;
            const int l1 = min(l0 + 1, pl.levels - 1);
            const Taps t0 = make_taps(ux, uy, level_extent(pl.H, l0), level_extent(pl.W, l0), 1.f - f);
            const Taps t1 = make_taps(ux, uy, level_extent(pl.H, l1), level_extent(pl.W, l1), f);
            float* b0 = pl.lv[l0];
            float* b1 = pl.lv[l1];
            if (!BACKWARD) {
                float4 v0[4], v1[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v0[j] = __ldg(reinterpret_cast<const float4*>(b0 + (size_t)t0.idx[j] * C + cb));
                if (f > 0.f) {
#pragma unroll
                    for (int j = 0; j < 4; j++) v1[j] = __ldg(reinterpret_cast<const float4*>(b1 + (size_t)t1.idx[j] * C + cb));
                }
                // per channel exactly the scalar kernel's expression: ((w0 v0 + w1 v1) + w2 v2) + w3 v3, then + level 1
#define SGS_PLANE_MIX(T, V, m) (T.w[0] * V[0].m + T.w[1] * V[1].m + T.w[2] * V[2].m + T.w[3] * V[3].m)
                float4 s = make_float4(SGS_PLANE_MIX(t0, v0, x), SGS_PLANE_MIX(t0, v0, y), SGS_PLANE_MIX(t0, v0, z),
                                       SGS_PLANE_MIX(t0, v0, w));
                if (f > 0.f) {
                    s.x += SGS_PLANE_MIX(t1, v1, x);
                    s.y += SGS_PLANE_MIX(t1, v1, y);
                    s.z += SGS_PLANE_MIX(t1, v1, z);
                    s.w += SGS_PLANE_MIX(t1, v1, w);
                }
#undef SGS_PLANE_MIX
                acc.x += s.x;
                acc.y += s.y;
                acc.z += s.z;
                acc.w += s.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (t0.w[j] != 0.f)
                        atomicAdd(reinterpret_cast<float4*>(b0 + (size_t)t0.idx[j] * C + cb),
                                  make_float4(t0.w[j] * g.x, t0.w[j] * g.y, t0.w[j] * g.z, t0.w[j] * g.w));
                if (f > 0.f) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (t1.w[j] != 0.f)
                            atomicAdd(reinterpret_cast<float4*>(b1 + (size_t)t1.idx[j] * C + cb),
                                      make_float4(t1.w[j] * g.x, t1.w[j] * g.y, t1.w[j] * g.z, t1.w[j] * g.w));
                }
            }
        }
        if (!BACKWARD) *reinterpret_cast<float4*>(out + (size_t)n * p.out_stride + p.out_offset + cb) = acc;
    }
}

// level 0 of the pyramid: NCHW parameter -> channels-last, 32 x-positions x all channels per block through smem
__global__ void __launch_bounds__(256) plane_to_channels_last_kernel(int C, int H, int W, const float* __restrict__ src,
                                                                     float* __restrict__ dst) {
    extern __shared__ float s_tile[];   // [C][33]
    const int x0 = blockIdx.x * 32, y = blockIdx.y;
    for (int e = threadIdx.x; e < C * 32; e += blockDim.x) {
        const int c = e >> 5, xx = e & 31;
        if (x0 + xx < W) s_tile[c * 33 + xx] = src[((size_t)c * H + y) * W + x0 + xx];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < C * 32; e += blockDim.x) {
        const int xx = e / C, c = e - xx * C;
        if (x0 + xx < W) dst[((size_t)y * W + x0 + xx) * C + c] = s_tile[c * 33 + xx];
    }
}

// level l+1 from level l (channels-last): 2x2 box (2x1 / 1x2 when an extent is already 1)
__global__ void __launch_bounds__(256) plane_downsample_kernel(int C, int h, int w, const float* __restrict__ src,
                                                               float* __restrict__ dst) {
    const int h2 = h > 1 ? h >> 1 : 1, w2 = w > 1 ? w >> 1 : 1;
    const size_t total = (size_t)h2 * w2 * C;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % C);
    const size_t t = e / C;
    const int x = (int)(t % w2), y = (int)(t / w2);
    const int sy = h > 1 ? 2 : 1, sx = w > 1 ? 2 : 1;
    float s = 0.f;
    for (int dy = 0; dy < sy; dy++)
        for (int dx = 0; dx < sx; dx++) s += src[((size_t)(y * sy + dy) * w + (x * sx + dx)) * C + c];
    dst[e] = s * (1.f / (float)(sy * sx));
}

struct FoldParams {
    const float* lv[kMaxLevels];
    int C, H, W, levels;
};

// dL/dplane (NCHW) = sum over levels of the gradient pyramid pushed down through the adjoint of the box filters
__global__ void __launch_bounds__(256) plane_fold_kernel(const __grid_constant__ FoldParams p, float* __restrict__ dplane) {
    extern __shared__ float s_tile[];   // [C][33]
    const int C = p.C, H = p.H, W = p.W;
    const int x0 = blockIdx.x * 32, y = blockIdx.y;
    for (int e = threadIdx.x; e < C * 32; e += blockDim.x) {
        const int xx = e / C, c = e - xx * C;
        const int x = x0 + xx;
        if (x < W) {
            float s = 0.f, fac = 1.f;
            int h = H, w = W;
            for (int l = 0; l < p.levels; l++) {
                const int yy = (H >> l) > 0 ? (y >> l) : 0, xl = (W >> l) > 0 ? (x >> l) : 0;
                const int wl = level_extent(W, l);
                s += fac * p.lv[l][((size_t)yy * wl + xl) * C + c];
                fac *= (h > 1 ? 0.5f : 1.f) * (w > 1 ? 0.5f : 1.f);
                h = h > 1 ? h >> 1 : 1;
                w = w > 1 ? w >> 1 : 1;
            }
            s_tile[c * 33 + xx] = s;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < C * 32; e += blockDim.x) {
        const int c = e >> 5, xx = e & 31;
        if (x0 + xx < W) dplane[((size_t)c * H + y) * W + x0 + xx] = s_tile[c * 33 + xx];
    }
}

// Four channels per thread (C % 4 == 0): the level walk's index arithmetic is paid once per float4 — the scalar kernel
// spends ~120 instructions per output element on it and runs at a fifth of the memory roofline.
__global__ void __launch_bounds__(256) plane_fold_kernel_v4(const __grid_constant__ FoldParams p, float* __restrict__ dplane) {
    extern __shared__ float s_tile[];   // [C][33]
    const int C = p.C, H = p.H, W = p.W, cpt = C >> 2;
    const int x0 = blockIdx.x * 32, y = blockIdx.y;
    for (int e = threadIdx.x; e < cpt * 32; e += blockDim.x) {
        const int xx = e / cpt, c4 = e - xx * cpt;
        const int x = x0 + xx;
        if (x < W) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            float fac = 1.f;
            int h = H, w = W;
            for (int l = 0; l < p.levels; l++) {
                const int yy = (H >> l) > 0 ? (y >> l) : 0, xl = (W >> l) > 0 ? (x >> l) : 0;
                const int wl = level_extent(W, l);
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.lv[l] + ((size_t)yy * wl + xl) * C + c4 * 4));
                s.x += fac * v.x; s.y += fac * v.y; s.z += fac * v.z; s.w += fac * v.w;     // same order as the scalar kernel
                fac *= (h > 1 ? 0.5f : 1.f) * (w > 1 ? 0.5f : 1.f);
                h = h > 1 ? h >> 1 : 1;
                w = w > 1 ? w >> 1 : 1;
            }
            float* t = s_tile + (c4 * 4) * 33 + xx;
            t[0] = s.x; t[33] = s.y; t[66] = s.z; t[99] = s.w;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < C * 32; e += blockDim.x) {
        const int c = e >> 5, xx = e & 31;
        if (x0 + xx < W) dplane[((size_t)c * H + y) * W + x0 + xx] = s_tile[c * 33 + xx];
    }
}

int levels_for(int H, int W, int max_mip_level) {
    int l = 1, h = H, w = W;
    while ((h > 1 || w > 1) && l - 1 < max_mip_level) {
        if ((h > 1 && (h & 1)) || (w > 1 && (w & 1))) return -1;   // extents must be even at every level that is built
        h = h > 1 ? h >> 1 : 1;
        w = w > 1 ? w >> 1 : 1;
        l++;
    }
    return l;
}

size_t level_offset(int C, int H, int W, int l) {   // floats before level l
    size_t off = 0;
    for (int k = 0; k < l; k++) off += (size_t)level_extent(H, k) * level_extent(W, k) * C;
    return off;
}

bool channels_ok(int C) { return C > 0 && (C >= 32 ? (C % 32 == 0) : ((C & (C - 1)) == 0)); }

}  // namespace sgs_plane

extern "C" {

int sgs_plane_levels(int H, int W, int max_mip_level) {
    if (H <= 0 || W <= 0 || max_mip_level < 0) return SGS_ERR_INVALID_ARGUMENT;
    const int l = sgs_plane::levels_for(H, W, max_mip_level > 7 ? 7 : max_mip_level);
    return l < 0 ? SGS_ERR_INVALID_ARGUMENT : l;
}

size_t sgs_plane_pyramid_floats(int C, int H, int W, int max_mip_level) {
    const int l = sgs_plane_levels(H, W, max_mip_level);
    if (l < 0 || C <= 0) return 0;
    return sgs_plane::level_offset(C, H, W, l);
}

int sgs_plane_build(int C, int H, int W, int max_mip_level, const float* plane_nchw, float* pyramid, void* stream) {
    using namespace sgs_plane;
    cudaStream_t s = (cudaStream_t)stream;
    const int L = sgs_plane_levels(H, W, max_mip_level);
    if (L < 0 || !channels_ok(C) || !plane_nchw || !pyramid) return SGS_ERR_INVALID_ARGUMENT;
    const size_t smem = (size_t)C * 33 * sizeof(float);
    if (smem > 48 * 1024) return SGS_ERR_INVALID_ARGUMENT;
    plane_to_channels_last_kernel<<<dim3((W + 31) / 32, H), 256, smem, s>>>(C, H, W, plane_nchw, pyramid);
    for (int l = 0; l + 1 < L; l++) {
        const int h = level_extent(H, l), w = level_extent(W, l);
        const size_t total = (size_t)level_extent(H, l + 1) * level_extent(W, l + 1) * C;
        plane_downsample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(C, h, w, pyramid + level_offset(C, H, W, l),
                                                                               pyramid + level_offset(C, H, W, l + 1));
    }
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

namespace sgs_plane {
// the four-channels-per-lane kernels need 16-byte aligned rows everywhere: channel count, output row layout, the
// caller's output / gradient pointer and every pyramid level (level offsets are multiples of C floats)
static bool vec4_ok(const FieldParams& fp, const float* io) {
    if (getenv("SGS_PLANE_SCALAR")) return false;     // developer switch for A/B measurements
    if (fp.C % 4 != 0 || fp.out_stride % 4 != 0 || fp.out_offset % 4 != 0) return false;
    if (reinterpret_cast<size_t>(io) & 15) return false;
    for (int k = 0; k < fp.n_planes; k++)
        if (reinterpret_cast<size_t>(fp.planes[k].lv[0]) & 15) return false;
    return true;
}
}  // namespace sgs_plane

static int fill_field(sgs_plane::FieldParams& fp, int N, int C, const float* pts, const float* timestamps,
                      const float* scales, const float* aabb, const float* base_scale, float time_scale, const int* reso0,
                      int n_planes, const sgs_plane_t* planes, int out_stride, int out_offset) {
    using namespace sgs_plane;
    if (N < 0 || n_planes < 1 || n_planes > 6 || !channels_ok(C) || !planes || !reso0) return SGS_ERR_INVALID_ARGUMENT;
    if (N > 0 && (!pts || !timestamps || !scales || !aabb || !base_scale)) return SGS_ERR_INVALID_ARGUMENT;
    if (out_offset < 0 || out_offset + C > out_stride) return SGS_ERR_INVALID_ARGUMENT;
    fp.n_planes = n_planes;
    for (int k = 0; k < n_planes; k++) {
        const sgs_plane_t& pl = planes[k];
        const int L = sgs_plane_levels(pl.H, pl.W, pl.max_mip_level);
        if (L < 0 || !pl.pyramid || pl.dim_u < 0 || pl.dim_u > 3 || pl.dim_v < 0 || pl.dim_v > 3) return SGS_ERR_INVALID_ARGUMENT;
        PlaneDev& d = fp.planes[k];
        d.H = pl.H;
        d.W = pl.W;
        d.du = pl.dim_u;
        d.dv = pl.dim_v;
        d.levels = L;
        for (int l = 0; l < kMaxLevels; l++) d.lv[l] = l < L ? pl.pyramid + level_offset(C, pl.H, pl.W, l) : nullptr;
    }
    fp.pts = pts;
    fp.timestamps = timestamps;
    fp.scales = scales;
    fp.aabb = aabb;
    fp.base_scale = base_scale;
    fp.time_scale = time_scale;
    for (int i = 0; i < 3; i++) fp.reso0[i] = reso0[i];
    fp.N = N;
    fp.C = C;
    fp.out_stride = out_stride;
    fp.out_offset = out_offset;
    return 0;
}

int sgs_plane_sample_forward(int N, int C, const float* pts, const float* timestamps, const float* scales,
                             const float* aabb, const float* base_scale, float time_scale, const int* reso0,
                             int n_planes, const sgs_plane_t* planes, int out_stride, int out_offset, float* out,
                             void* stream) {
    sgs_plane::FieldParams fp;
    if (int rc = fill_field(fp, N, C, pts, timestamps, scales, aabb, base_scale, time_scale, reso0, n_planes, planes,
                            out_stride, out_offset))
        return rc;
    if (N == 0) return 0;
    if (!out) return SGS_ERR_INVALID_ARGUMENT;
    if (sgs_plane::vec4_ok(fp, out)) {
        const size_t threads = (size_t)N * (C / 4 < 8 ? C / 4 : 8);
        sgs_plane::plane_field_kernel_v4<false><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fp, out, nullptr);
        return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
    }
    const int lpp = C < 32 ? C : 32;
    const size_t threads = (size_t)N * lpp;
    sgs_plane::plane_field_kernel<false><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fp, out, nullptr);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int sgs_plane_sample_backward(int N, int C, const float* pts, const float* timestamps, const float* scales,
                              const float* aabb, const float* base_scale, float time_scale, const int* reso0,
                              int n_planes, const sgs_plane_t* grad_planes, int out_stride, int out_offset,
                              const float* dout, void* stream) {
    sgs_plane::FieldParams fp;
    if (int rc = fill_field(fp, N, C, pts, timestamps, scales, aabb, base_scale, time_scale, reso0, n_planes, grad_planes,
                            out_stride, out_offset))
        return rc;
    if (N == 0) return 0;
    if (!dout) return SGS_ERR_INVALID_ARGUMENT;
    if (sgs_plane::vec4_ok(fp, dout)) {
        const size_t threads = (size_t)N * (C / 4 < 8 ? C / 4 : 8);
        sgs_plane::plane_field_kernel_v4<true><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fp, nullptr, dout);
        return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
    }
    const int lpp = C < 32 ? C : 32;
    const size_t threads = (size_t)N * lpp;
    sgs_plane::plane_field_kernel<true><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fp, nullptr, dout);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int sgs_plane_fold(int C, int H, int W, int max_mip_level, const float* grad_pyramid, float* dplane_nchw, void* stream) {
    using namespace sgs_plane;
    const int L = sgs_plane_levels(H, W, max_mip_level);
    if (L < 0 || !channels_ok(C) || !grad_pyramid || !dplane_nchw) return SGS_ERR_INVALID_ARGUMENT;
    const size_t smem = (size_t)C * 33 * sizeof(float);
    if (smem > 48 * 1024) return SGS_ERR_INVALID_ARGUMENT;
    FoldParams fp;
    fp.C = C;
    fp.H = H;
    fp.W = W;
    fp.levels = L;
    for (int l = 0; l < kMaxLevels; l++) fp.lv[l] = l < L ? grad_pyramid + level_offset(C, H, W, l) : nullptr;
    if (C % 4 == 0 && !(reinterpret_cast<size_t>(grad_pyramid) & 15) && !getenv("SGS_PLANE_SCALAR"))
        plane_fold_kernel_v4<<<dim3((W + 31) / 32, H), 256, smem, (cudaStream_t)stream>>>(fp, dplane_nchw);
    else
        plane_fold_kernel<<<dim3((W + 31) / 32, H), 256, smem, (cudaStream_t)stream>>>(fp, dplane_nchw);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

}  // extern "C"

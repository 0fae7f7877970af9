// Forward tile compositing: front-to-back alpha blending of colour + median depth.
//
// Semantics follow $R/cuda_rasterizer/forward.cu:261-393 exactly (same thresholds, same
// float expression trees, same n_contrib bookkeeping):
//   power = -0.5 (A dx^2 + C dy^2) - B dx dy ; skip if power > 0
//   alpha = min(0.99, o * exp(power))        ; skip if alpha < 1/255
//   test_T = T (1 - alpha)                   ; if < 1e-4 the pixel is done (instance NOT blended)
//   C += rgb * alpha * T ; median depth when T crosses 0.5 ; T = test_T
//   out = C + T * bg ; depth default 15.0 ; n_contrib = 1-based list position of last blended
//
// B200 design (differs from the reference's one-gather-per-thread + 256 evaluations):
//   * Exact tile-level culling while staging: an instance whose alpha is provably < 1/255 on
//     every pixel of this tile (conservative ellipse/rectangle bound with an explicit
//     rounding margin) is dropped before it reaches shared memory.  The reference would
//     `continue` on all 256 pixels for such an instance, so results are bit-identical.
//   * The survivors are compacted with warp ballots and staged as 3x float4 (48 B) records —
//     colour and depth come from shared memory instead of per-pixel global gathers.
//   * A per-instance threshold on `power` short-circuits the exp() for pixels that the
//     reference would skip after computing alpha (same decision, proven conservative).
//   * Optionally the compacted records are streamed to HBM (`packed`) so the backward
//     kernel reads a dense, tile-ordered list with bulk copies and never gathers.
//   * Warps own 8x4 pixel patches (not 16x2 rows) so whole-warp skips are more frequent.
#include "sgs_common.cuh"

namespace sgs {

// Returns false only if NO pixel of the tile [tx0,tx0+15]x[ty0,ty0+15] can pass the
// reference's  alpha >= 1/255  test for this instance.  `thr` is the power threshold
// (pixels with power < thr are skipped by the reference after exp()).
__forceinline__ __device__ bool tile_may_contribute(float mx, float my, float A, float B, float C, float thr,
                                                    float tx0, float ty0) {
    const float big = fmaxf(fmaxf(fabsf(A), fabsf(B)), fabsf(C));
    const bool safe = (big < 1e15f) && (fabsf(mx) < 1e6f) && (fabsf(my) < 1e6f);
    if (!safe || thr != thr) return true;  // NaN/inf/huge: let the exact per-pixel path decide
    if (thr > 0.f) return false;      // opacity < 1/255: o*exp(power<=0) < 1/255 everywhere
    if (!(A > 0.f && C > 0.f && A * C - B * B > 0.f)) return true;  // not positive definite
    // d = mean - pixel ranges over [x0,x1] x [y0,y1]
    const float x1 = mx - tx0, x0 = x1 - (SGS_TILE_X - 1);
    const float y1 = my - ty0, y0 = y1 - (SGS_TILE_Y - 1);
    if (x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f) return true;  // centre inside the tile
    // q(dx,dy) = 0.5 (A dx^2 + C dy^2) + B dx dy = -power ; minimise over the four edges
    float qmin = 3.0e38f;
    {
        const float inv = 1.f / C;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float cx = e ? x1 : x0;
            const float dy = fminf(y1, fmaxf(y0, -B * cx * inv));
            qmin = fminf(qmin, 0.5f * (A * cx * cx + C * dy * dy) + B * cx * dy);
        }
    }
    {
        const float inv = 1.f / A;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float cy = e ? y1 : y0;
            const float dx = fminf(x1, fmaxf(x0, -B * cy * inv));
            qmin = fminf(qmin, 0.5f * (A * dx * dx + C * cy * cy) + B * dx * cy);
        }
    }
    const float Dx = fmaxf(fabsf(x0), fabsf(x1)), Dy = fmaxf(fabsf(y0), fabsf(y1));
    const float margin = 1e-5f * (A * Dx * Dx + C * Dy * Dy + 2.f * fabsf(B) * Dx * Dy) + 1e-3f;
    return (qmin - margin) <= -thr;
}

// power threshold below which  o * exp(power) < 1/255  is certain (margin 1e-3 in log space)
__forceinline__ __device__ float power_threshold(float o) {
    if (!(o > 0.f)) return (o == o) ? __int_as_float(0x7f800000) : __int_as_float(0x7fc00000);
    return -logf(255.f * o) - 1e-3f;
}

__forceinline__ __device__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__forceinline__ __device__ float pin_reg(float v) {
    asm volatile("" : "+f"(v));
    return v;
}

template <bool WRITE_PACKED, bool TILE_CULL>
__global__ void __launch_bounds__(SGS_TILE_PIX)
render_fwd_kernel(const __grid_constant__ ViewParams vp, const uint2* __restrict__ ranges,
                  const uint32_t* __restrict__ point_list, const float2* __restrict__ means2D,
                  const float4* __restrict__ conic_opacity, const float4* __restrict__ rgbd,
                  float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                  uint32_t* __restrict__ tile_count, PackedInst* __restrict__ packed,
                  float* __restrict__ out_color, float* __restrict__ out_depth) {
    __shared__ float4 s_stage[3 * SGS_TILE_PIX];
    float4* const s_a = s_stage;                     // x, y, A, B
    float4* const s_b = s_stage + SGS_TILE_PIX;      // C, opacity, thr, list_pos(bits)
    float4* const s_c = s_stage + 2 * SGS_TILE_PIX;  // r, g, b, depth
    __shared__ uint32_t s_wcount[SGS_TILE_PIX / 32];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int W = vp.W, H = vp.H;
    const uint32_t tile = blockIdx.y * vp.tiles_x + blockIdx.x;
    const uint32_t tx0 = blockIdx.x * SGS_TILE_X, ty0 = blockIdx.y * SGS_TILE_Y;
    // warp -> 8x4 pixel patch
    const uint32_t px = tx0 + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = ty0 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    const float2 pixf = {pin_reg((float)px), pin_reg((float)py)};
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(s_stage);
    asm volatile("" : "+r"(sa));   // keep the shared-window base in a register (no re-derivation in the loop)
    constexpr uint32_t kB = SGS_TILE_PIX * 16u, kC = 2u * SGS_TILE_PIX * 16u;

    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C[SGS_CH] = {0.f, 0.f, 0.f};
    float D = 15.0f;  // reference default for "median depth never crossed" ($R/.../forward.cu:308)
    uint32_t packed_count = 0;

    for (int base = 0; base < n; base += SGS_TILE_PIX) {
        // also the barrier that protects the staging buffers of the previous round
        if (__syncthreads_count(done) == SGS_TILE_PIX) break;

        // ---- stage: gather, cull, compact ------------------------------------------------
        const int i = base + tid;
        bool keep = false;
        float4 ra, rb, rc;
        uint32_t gid = 0;
        if (i < n) {
            gid = point_list[range.x + i];
            const float2 xy = means2D[gid];
            const float4 co = conic_opacity[gid];
            rc = rgbd[gid];
            const float thr = power_threshold(co.w);
            ra = make_float4(xy.x, xy.y, co.x, co.y);
            rb = make_float4(co.z, co.w, thr, __uint_as_float((uint32_t)i));
            keep = TILE_CULL ? tile_may_contribute(xy.x, xy.y, co.x, co.y, co.z, thr, (float)tx0, (float)ty0) : true;
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SGS_TILE_PIX / 32; w++) {
            const uint32_t c = s_wcount[w];
            if (w < warp) woff += c;
            total += c;
        }
        if (keep) {
            const uint32_t pos = woff + __popc(bal & ((1u << lane) - 1u));
            s_a[pos] = ra;
            s_b[pos] = rb;
            s_c[pos] = rc;
            if (WRITE_PACKED) {
                float4* dst = reinterpret_cast<float4*>(packed + (size_t)range.x + packed_count + pos);
                dst[0] = ra;
                dst[1] = rb;
                dst[2] = make_float4(rc.x, rc.y, rc.z, __uint_as_float(gid));
            }
        }
        packed_count += total;
        __syncthreads();

        // ---- composite -------------------------------------------------------------------
        for (uint32_t j = 0, off = 0; !done && j < total; j++, off += 16u) {
            const float4 a = lds128(sa + off);
            const float4 b = lds128(sa + off + kB);
            const float2 d = {a.x - pixf.x, a.y - pixf.y};
            const float power = -0.5f * (a.z * d.x * d.x + b.x * d.y * d.y) - a.w * d.x * d.y;
            if (power > 0.0f) continue;
            if (power < b.z) continue;  // provably alpha < 1/255 (see power_threshold)
            const float alpha = min(0.99f, b.y * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) {
                done = true;
                continue;
            }
            const float4 c = lds128(sa + off + kC);
            C[0] += c.x * alpha * T;
            C[1] += c.y * alpha * T;
            C[2] += c.z * alpha * T;
            if (T > 0.5f && test_T < 0.5) D = c.w;
            T = test_T;
            last_contributor = __float_as_uint(b.w) + 1u;
        }
    }

    if (WRITE_PACKED && tid == 0) tile_count[tile] = packed_count;
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        const size_t HW = (size_t)H * W;
#pragma unroll
        for (int ch = 0; ch < SGS_CH; ch++) out_color[ch * HW + pix_id] = C[ch] + T * vp.bg[ch];
        out_depth[pix_id] = D;
    }
}

void launch_render_fwd(const ViewParams& vp, GeomState g, BinningState b, ImageState img,
                       const uint32_t* point_list, int write_packed, int tile_cull, float* out_color,
                       float* out_depth, cudaStream_t s) {
    dim3 grid(vp.tiles_x, vp.tiles_y, 1);
#define SGS_LAUNCH_RF(WP, TC)                                                                              \
    render_fwd_kernel<WP, TC><<<grid, SGS_TILE_PIX, 0, s>>>(vp, img.ranges, point_list, g.means2D,          \
                                                            g.conic_opacity, g.rgbd, img.final_T,          \
                                                            img.n_contrib, img.tile_count, b.packed,       \
                                                            out_color, out_depth)
    if (write_packed) {
        if (tile_cull) SGS_LAUNCH_RF(true, true); else SGS_LAUNCH_RF(true, false);
    } else {
        if (tile_cull) SGS_LAUNCH_RF(false, true); else SGS_LAUNCH_RF(false, false);
    }
#undef SGS_LAUNCH_RF
}

}  // namespace sgs

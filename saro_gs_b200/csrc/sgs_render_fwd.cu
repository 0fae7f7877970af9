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
// B200 design (differs from the reference's one-gather-per-thread + 256 evaluations per instance):
//   * 128 threads per tile, 2 vertically adjacent pixels per thread, warp = 8x8 quadrant; the quadratic
//     form of both pixels is evaluated with packed f32x2 instructions (sgs_render_common.cuh).
//   * Exact quadrant-level culling while staging: for every instance a 4-bit mask says which 8x8
//     quadrants can reach alpha >= 1/255 at all (conservative ellipse/rectangle bound with explicit
//     rounding margins).  Instances with an empty mask never reach shared memory; a warp walks only
//     the instances whose mask has its quadrant bit (ballot bitmaps), so a warp-visit almost always
//     blends something.  The reference would `continue` on every culled pixel: bit-identical results.
//   * Survivors are compacted with warp ballots and staged as 3x float4 records — colour and depth
//     come from shared memory instead of per-pixel global gathers.
//   * A per-instance threshold on `power` short-circuits exp() for pixels the reference would skip
//     after computing alpha (same decision, proven conservative).
//   * Optionally the compacted records are streamed to HBM (`packed`) so the backward kernel reads a
//     dense, tile-ordered list and never gathers.
#include "sgs_render_common.cuh"

namespace sgs {

// State of the two pixels of a thread as f32x2 pairs: the whole blend step runs on packed FMUL2 / FFMA2 / FADD2
// instructions and selects instead of two divergent scalar paths.
struct FwdPix2 {
    float2 T, C0, C1, C2, D;
    uint32_t last0, last1;
    bool done0, done1;
};

// two pixels x one instance, exactly the reference's sequence of tests and roundings per pixel
// ($R/cuda_rasterizer/forward.cu:342-379).  A pixel that does not blend (inactive, alpha < 1/255, or saturating)
// runs through the same packed code with alpha = 0: then test_T = T * 1 = T and C += (c * 0) * T = C exactly
// (C is never -0: it starts at +0 and a round-to-nearest sum that cancels gives +0).
// `median_open` (warp-uniform): some pixel of the warp still has T > 0.5, i.e. the median-depth test can still fire;
// once every pixel is past it the six instructions of the test are skipped (T only decreases).
__forceinline__ __device__ void blend_pixels2(FwdPix2& s, float2 pw, bool act0, bool act1, float o, const float4 c,
                                              uint32_t pos, bool median_open) {
    const float2 e = expf2_exact(pw);
    float2 alpha = __fmul2_rn(make_float2(o, o), e);
    alpha.x = min(0.99f, alpha.x);
    alpha.y = min(0.99f, alpha.y);
    bool go0 = act0 && !(alpha.x < 1.0f / 255.0f);
    bool go1 = act1 && !(alpha.y < 1.0f / 255.0f);
    const float2 test_T = __fmul2_rn(s.T, __fadd2_rn(make_float2(1.f, 1.f), make_float2(-alpha.x, -alpha.y)));
    if (go0 && test_T.x < 0.0001f) { s.done0 = true; go0 = false; }
    if (go1 && test_T.y < 0.0001f) { s.done1 = true; go1 = false; }
    if (!go0) alpha.x = 0.f;
    if (!go1) alpha.y = 0.f;
    s.C0 = __ffma2_rn(__fmul2_rn(make_float2(c.x, c.x), alpha), s.T, s.C0);
    s.C1 = __ffma2_rn(__fmul2_rn(make_float2(c.y, c.y), alpha), s.T, s.C1);
    s.C2 = __ffma2_rn(__fmul2_rn(make_float2(c.z, c.z), alpha), s.T, s.C2);
    if (median_open) {
        if (go0 && s.T.x > 0.5f && test_T.x < 0.5) s.D.x = c.w;
        if (go1 && s.T.y > 0.5f && test_T.y < 0.5) s.D.y = c.w;
    }
    if (go0) {
        s.T.x = test_T.x;
        s.last0 = pos + 1u;
    }
    if (go1) {
        s.T.y = test_T.y;
        s.last1 = pos + 1u;
    }
}

template <bool WRITE_PACKED, bool TILE_CULL>
__global__ void __launch_bounds__(SGS_R_THREADS, 10)
render_fwd_kernel(const __grid_constant__ ViewParams vp, const uint2* __restrict__ ranges,
                  const uint32_t* __restrict__ point_list, const float2* __restrict__ means2D,
                  const float4* __restrict__ conic_opacity, const float4* __restrict__ rgbd,
                  float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                  uint32_t* __restrict__ tile_count, PackedInst* __restrict__ packed,
                  float* __restrict__ out_color, float* __restrict__ out_depth, float4* __restrict__ acc4,
                  uint32_t acc_n4) {
    __shared__ float4 s_g0[SGS_R_BATCH];   // x, y, A, -B
    __shared__ float4 s_g1[SGS_R_BATCH];   // C, thr, list_pos(bits), opacity
    __shared__ float4 s_g2[SGS_R_BATCH];   // r, g, b, depth
    __shared__ uint32_t s_mask[SGS_R_BATCH];
    __shared__ uint32_t s_wcount[SGS_R_THREADS / 32];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int W = vp.W, H = vp.H;
    const uint32_t tile = blockIdx.y * vp.tiles_x + blockIdx.x;
    const uint32_t tx0 = blockIdx.x * SGS_TILE_X, ty0 = blockIdx.y * SGS_TILE_Y;
    const uint32_t px = tx0 + (warp & 1) * SGS_Q + (lane & 7);
    const uint32_t py0 = ty0 + (warp >> 1) * SGS_Q + 2 * (lane >> 3);
    const uint32_t py1 = py0 + 1;
    const bool in0 = px < (uint32_t)W && py0 < (uint32_t)H;
    const bool in1 = px < (uint32_t)W && py1 < (uint32_t)H;
    const float pxf = pin_reg((float)px);
    const float2 npy = {pin_reg(-(float)py0), pin_reg(-(float)py1)};
    uint32_t a0 = (uint32_t)__cvta_generic_to_shared(s_g0);
    uint32_t a1 = (uint32_t)__cvta_generic_to_shared(s_g1);
    uint32_t a2 = (uint32_t)__cvta_generic_to_shared(s_g2);
    asm volatile("" : "+r"(a0), "+r"(a1), "+r"(a2));

    if (WRITE_PACKED) {
        // zero this CTA's share of the backward pass's moment accumulator ([P][12] floats): the kernel is bound by
        // instruction issue and its memory system is idle, so the 14 MB fill costs nothing here and the backward pass
        // needs no memset launch.  Independent of the predecessor kernel: done before the dependency wait.
        const uint32_t ctas = gridDim.x * gridDim.y;
        const uint32_t per = (acc_n4 + ctas - 1) / ctas;
        const uint32_t b0 = tile * per, b1 = min(acc_n4, b0 + per);
        for (uint32_t i = b0 + tid; i < b1; i += SGS_R_THREADS) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pdl_wait();     // launched programmatically dependent on the expansion kernel that writes ranges / point_list
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);

    // reference defaults: T = 1, depth 15.0 when the median is never crossed ($R/.../forward.cu:303-308)
    FwdPix2 px2;
    px2.T = make_float2(1.0f, 1.0f);
    px2.C0 = px2.C1 = px2.C2 = make_float2(0.f, 0.f);
    px2.D = make_float2(15.0f, 15.0f);
    px2.last0 = px2.last1 = 0u;
    px2.done0 = !in0;
    px2.done1 = !in1;
    uint32_t packed_count = 0;
    bool both_done = px2.done0 && px2.done1;

    for (int base = 0; base < n; base += SGS_R_BATCH) {
        // also the barrier that protects the staging buffers of the previous round
        if (__syncthreads_count(both_done) == SGS_R_THREADS) break;

        // ---- stage: gather, cull, compact ------------------------------------------------
        const int i = base + tid;
        uint32_t mask = 0;
        float4 r0, r1, r2;
        uint32_t gid = 0;
        float Bc = 0.f;
        if (i < n) {
            gid = point_list[range.x + i];
            const float2 xy = means2D[gid];
            const float4 co = conic_opacity[gid];
            const float4 cd = rgbd[gid];
            const float thr = power_threshold(co.w);
            r0 = make_float4(xy.x, xy.y, co.x, -co.y);
            r1 = make_float4(co.z, thr, __uint_as_float((uint32_t)i), co.w);
            r2 = cd;
            Bc = co.y;
            mask = TILE_CULL ? quadrant_mask(xy.x, xy.y, co.x, co.y, co.z, thr, (float)tx0, (float)ty0) : 0xFu;
        }
        const bool keep = mask != 0u;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SGS_R_THREADS / 32; w++) {
            const uint32_t c = s_wcount[w];
            if (w < warp) woff += c;
            total += c;
        }
        if (keep) {
            const uint32_t pos = woff + __popc(bal & ((1u << lane) - 1u));
            s_g0[pos] = r0;
            s_g1[pos] = r1;
            s_g2[pos] = r2;
            s_mask[pos] = mask;
            if (WRITE_PACKED) {
                float4* dst = reinterpret_cast<float4*>(packed + (size_t)range.x + packed_count + pos);
                dst[0] = make_float4(r0.x, r0.y, r0.z, Bc);                      // x, y, A, B
                dst[1] = make_float4(r1.x, r1.w, r1.y, r1.z);                    // C, opacity, thr, list_pos
                dst[2] = make_float4(r2.x, r2.y, r2.z, __uint_as_float(gid | (mask << 28)));   // r, g, b, gid | mask
            }
        }
        packed_count += total;
        __syncthreads();

        // ---- per-warp visit bitmaps: bit b of word k <=> staged slot 32k+b can reach this quadrant
        uint32_t mywords = 0;
#pragma unroll
        for (int k = 0; k < SGS_R_BATCH / 32; k++) {
            const uint32_t slot = k * 32 + lane;
            const uint32_t m = slot < total ? s_mask[slot] : 0u;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, (m >> warp) & 1u);
            if (lane == k) mywords = b;
        }
        if (__all_sync(0xFFFFFFFFu, both_done)) continue;   // whole quadrant saturated

        // ---- composite -------------------------------------------------------------------
        for (int k = 0; k < SGS_R_BATCH / 32; k++) {
            // saturation of the whole quadrant is tested once per 32 staged slots, not once per visit
            both_done = px2.done0 && px2.done1;
            if (__all_sync(0xFFFFFFFFu, both_done)) break;
            // can the median-depth test still fire for some pixel of the warp?  (T never increases; a finished or
            // out-of-image pixel cannot blend any more)
            const bool median_open = __any_sync(0xFFFFFFFFu, (!px2.done0 && px2.T.x > 0.5f) || (!px2.done1 && px2.T.y > 0.5f));
            uint32_t word = __shfl_sync(0xFFFFFFFFu, mywords, k);
            while (word) {
                const uint32_t j = k * 32 + (__ffs(word) - 1);
                word &= word - 1;
                const float4 g0 = lds128(a0 + j * 16u);
                const float4 g1 = lds128(a1 + j * 16u);
                float dx;
                float2 dy;
                const float2 pw = power2(g0, g1.x, pxf, npy, dx, dy);
                // same tests as the reference: power > 0 -> skip; power < thr -> provably alpha < 1/255
                const bool act0 = !px2.done0 && !(pw.x > 0.0f) && !(pw.x < g1.y);
                const bool act1 = !px2.done1 && !(pw.y > 0.0f) && !(pw.y < g1.y);
                if (!__any_sync(0xFFFFFFFFu, act0 || act1)) continue;
                const float4 c = lds128(a2 + j * 16u);
                const uint32_t pos = __float_as_uint(g1.z);
                blend_pixels2(px2, pw, act0, act1, g1.w, c, pos, median_open);
            }
        }
        both_done = px2.done0 && px2.done1;
    }

    if (WRITE_PACKED && tid == 0) tile_count[tile] = packed_count;
    const size_t HW = (size_t)H * W;
    if (in0) {
        const uint32_t pix_id = (uint32_t)W * py0 + px;
        final_T[pix_id] = px2.T.x;
        n_contrib[pix_id] = px2.last0;
        out_color[pix_id] = px2.C0.x + px2.T.x * vp.bg[0];
        out_color[HW + pix_id] = px2.C1.x + px2.T.x * vp.bg[1];
        out_color[2 * HW + pix_id] = px2.C2.x + px2.T.x * vp.bg[2];
        out_depth[pix_id] = px2.D.x;
    }
    if (in1) {
        const uint32_t pix_id = (uint32_t)W * py1 + px;
        final_T[pix_id] = px2.T.y;
        n_contrib[pix_id] = px2.last1;
        out_color[pix_id] = px2.C0.y + px2.T.y * vp.bg[0];
        out_color[HW + pix_id] = px2.C1.y + px2.T.y * vp.bg[1];
        out_color[2 * HW + pix_id] = px2.C2.y + px2.T.y * vp.bg[2];
        out_depth[pix_id] = px2.D.y;
    }
}

void launch_render_fwd(int P, const ViewParams& vp, GeomState g, BinningState b, ImageState img,
                       const uint32_t* point_list, int write_packed, int tile_cull, float* out_color,
                       float* out_depth, cudaStream_t s) {
    dim3 grid(vp.tiles_x, vp.tiles_y, 1);
#define SGS_LAUNCH_RF(WP, TC)                                                                              \
    launch_pdl(render_fwd_kernel<WP, TC>, grid, dim3(SGS_R_THREADS), 0, s, vp, img.ranges, point_list,      \
               g.means2D, g.conic_opacity, g.rgbd, img.final_T, img.n_contrib, img.tile_count, b.packed,    \
               out_color, out_depth, reinterpret_cast<float4*>(g.acc), (uint32_t)P * 3u)
    if (write_packed) {
        if (tile_cull) SGS_LAUNCH_RF(true, true); else SGS_LAUNCH_RF(true, false);
    } else {
        if (tile_cull) SGS_LAUNCH_RF(false, true); else SGS_LAUNCH_RF(false, false);
    }
#undef SGS_LAUNCH_RF
}

}  // namespace sgs

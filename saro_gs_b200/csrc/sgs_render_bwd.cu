// Backward tile compositing: per-pixel back-to-front gradient walk.
//
// Semantics follow $R/cuda_rasterizer/backward.cu:399-557:
//   walk the tile list from the back; skip entries with list position >= n_contrib(pixel);
//   recompute alpha with the same tests as forward; recover T by division T /= (1-alpha);
//   dL/dcolor = alpha*T*dL/dpix ; dL/dalpha via the accum_rec recursion + background term;
//   dL/dmean2D (scaled by 0.5W, 0.5H), dL/dconic (A,B,C), dL/dopacity.
//
// B200 design (the reference issues 9 global float atomics per contributing pixel x instance,
// SURVEY.md §2.1):
//   * reads the dense tile-ordered `PackedInst` list written by the forward kernel with
//     coalesced 16-byte loads (no index gathers, tile-culled instances never appear);
//   * each warp owns an 8x4 pixel patch and walks the staged batch on its own: it starts at the
//     last record that any of its pixels blended (binary search on the sorted list positions),
//     so records behind every pixel's last contributor cost nothing;
//   * the 9 partial derivatives of an instance are reduced across the 32 pixels of a warp
//     with a 12-shuffle "transposing" butterfly (values are split between lane halves at
//     every step instead of reducing each value with 5 shuffles), only when at least one
//     lane contributes;
//   * one RED per value per (warp, instance) lands in a [P][12] accumulator (48-byte rows)
//     => 32x fewer L2 atomics than the reference.
//   The kernel is FP32-issue bound (ncu: issue active 90%), so the inner loop is written to keep
//   the not-contributing path at ~a dozen instructions.
#include "sgs_common.cuh"

namespace sgs {

#define SGS_BWD_BATCH 256

// send `hi` to the partner if this lane keeps `lo`, and vice versa; returns kept + received
__forceinline__ __device__ float xsplit(float lo, float hi, bool upper, int xorm) {
    const float send = upper ? lo : hi;
    const float keep = upper ? hi : lo;
    return keep + __shfl_xor_sync(0xFFFFFFFFu, send, xorm);
}

// 128-bit / 32-bit shared-memory loads from a 32-bit shared-window address.  Using explicit
// shared addresses keeps nvcc from re-deriving the generic->shared base (S2R SR_CgaCtaId + LEA)
// inside the hot loop.
__forceinline__ __device__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// keep a loop-invariant value in a register (stops nvcc from rematerialising it inside the loop)
__forceinline__ __device__ float pin_reg(float v) {
    asm volatile("" : "+f"(v));
    return v;
}

__global__ void __launch_bounds__(SGS_TILE_PIX, 3)
render_bwd_kernel(const __grid_constant__ ViewParams vp, const uint2* __restrict__ ranges,
                  const uint32_t* __restrict__ tile_count, const PackedInst* __restrict__ packed,
                  const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ dL_dpix, float* __restrict__ acc) {
    __shared__ float4 s_rec[SGS_BWD_BATCH * 3];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int W = vp.W, H = vp.H;
    const uint32_t tile = blockIdx.y * vp.tiles_x + blockIdx.x;
    const uint32_t tx0 = blockIdx.x * SGS_TILE_X, ty0 = blockIdx.y * SGS_TILE_Y;
    const uint32_t px = tx0 + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = ty0 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    const float pixx = pin_reg((float)px), pixy = pin_reg((float)py);

    const uint32_t start = ranges[tile].x;
    const int count = (int)tile_count[tile];
    if (count == 0) return;

    const float T_final = inside ? final_T[pix_id] : 0.f;
    float T = T_final;
    const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0u;
    // warp-uniform bound: records at list positions >= this were blended by no pixel of the warp
    const uint32_t warp_last = __reduce_max_sync(0xFFFFFFFFu, last_contributor);

    float accum_rec[SGS_CH] = {0.f, 0.f, 0.f};
    float dL_dpixel[SGS_CH] = {0.f, 0.f, 0.f};
    if (inside) {
        const size_t HW = (size_t)H * W;
#pragma unroll
        for (int ch = 0; ch < SGS_CH; ch++) dL_dpixel[ch] = dL_dpix[ch * HW + pix_id];
    }
    float last_alpha = 0.f;
    float last_color[SGS_CH] = {0.f, 0.f, 0.f};
    // $R/cuda_rasterizer/backward.cu:460-461 (double product rounded to float once)
    const float ddelx_dx = pin_reg((float)(0.5 * W));
    const float ddely_dy = pin_reg((float)(0.5 * H));
    float bg_dot_dpixel = 0.f;
#pragma unroll
    for (int ch = 0; ch < SGS_CH; ch++) bg_dot_dpixel += vp.bg[ch] * dL_dpixel[ch];
    const float neg_Tfinal_bg = pin_reg(-T_final * bg_dot_dpixel);

    // which accumulator slot this lane owns after the butterfly (see reduce below)
    //   bit1 set -> value 4 ; else value = (bit4 ? 5 : 0) + (bit2 ? 2 : 0) + (bit3 ? 1 : 0)
    const int my_slot = (lane & 2) ? 4 : (((lane & 16) ? 5 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 8) ? 1 : 0));
    const bool writer = (lane & 1) == 0 && ((lane & 2) == 0 || lane == 2);
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
    float* const acc_lane = acc + my_slot;

    const float4* src = reinterpret_cast<const float4*>(packed + start);
    uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_rec);
    asm volatile("" : "+r"(s_base));

    for (int hi = count; hi > 0; hi -= SGS_BWD_BATCH) {
        const int lo = max(0, hi - SGS_BWD_BATCH);
        const int nrec = hi - lo;
        __syncthreads();
        for (int k = tid; k < nrec * 3; k += SGS_TILE_PIX) s_rec[k] = src[(size_t)lo * 3 + k];
        __syncthreads();

        // first record (from the back) that some pixel of this warp blended: list positions are
        // strictly increasing within the batch -> binary search for the count of records < warp_last
        int nvalid;
        {
            int a0 = 0, a1 = nrec;
            while (a0 < a1) {
                const int mid = (a0 + a1) >> 1;
                if (__float_as_uint(s_rec[3 * mid + 1].w) < warp_last) a0 = mid + 1;
                else a1 = mid;
            }
            nvalid = a0;
        }

        uint32_t rec = s_base + (uint32_t)nvalid * 48u;   // one past the first record to visit
        for (int j = nvalid; j > 0; j--) {
            rec -= 48u;
            const float4 a = lds128(rec);         // x, y, A, B
            const float4 b = lds128(rec + 16u);   // C, opacity, thr, list_pos
            const float dx = a.x - pixx, dy = a.y - pixy;
            const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
            // same tests as forward: power > 0 -> skip; power < thr -> provably alpha < 1/255
            bool active = (__float_as_uint(b.w) < last_contributor) && !(power > 0.0f) && !(power < b.z);
            float G = 0.f, alpha = 0.f;
            if (active) {
                G = expf(power);
                alpha = min(0.99f, b.y * G);
                active = !(alpha < 1.0f / 255.0f);
            }
            if (!__any_sync(0xFFFFFFFFu, active)) continue;

            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f;
            const float4 c4 = lds128(rec + 32u);  // r, g, b, gid
            if (active) {
                // 1/(1-alpha) once (IEEE reciprocal), shared by the T recovery and the background term;
                // the reference divides twice ($R/cuda_rasterizer/backward.cu:503,534): <= 1 ulp apart
                const float inv = __frcp_rn(1.f - alpha);
                T = T * inv;
                const float dchannel_dcolor = alpha * T;
                const float om = 1.f - last_alpha;
                accum_rec[0] = last_alpha * last_color[0] + om * accum_rec[0];
                accum_rec[1] = last_alpha * last_color[1] + om * accum_rec[1];
                accum_rec[2] = last_alpha * last_color[2] + om * accum_rec[2];
                last_color[0] = c4.x;
                last_color[1] = c4.y;
                last_color[2] = c4.z;
                float dL_dalpha = (c4.x - accum_rec[0]) * dL_dpixel[0];
                dL_dalpha += (c4.y - accum_rec[1]) * dL_dpixel[1];
                dL_dalpha += (c4.z - accum_rec[2]) * dL_dpixel[2];
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += neg_Tfinal_bg * inv;

                const float dL_dG = b.y * dL_dalpha;
                const float gdx = G * dx;
                const float gdy = G * dy;
                const float dG_ddelx = -gdx * a.z - gdy * a.w;
                const float dG_ddely = -gdy * b.x - gdx * a.w;
                v0 = dL_dG * dG_ddelx * ddelx_dx;
                v1 = dL_dG * dG_ddely * ddely_dy;
                const float h = -0.5f * dL_dG;
                v2 = h * gdx * dx;
                v3 = h * gdx * dy;
                v4 = h * gdy * dy;
                v5 = G * dL_dalpha;
                v6 = dchannel_dcolor * dL_dpixel[0];
                v7 = dchannel_dcolor * dL_dpixel[1];
                v8 = dchannel_dcolor * dL_dpixel[2];
            }

            // transposing butterfly: 9 values x 32 lanes -> one value per writer lane
            // step 1 (xor 16): pairs (v0,v5) (v1,v6) (v2,v7) (v3,v8); v4 reduced plainly
            float w0 = xsplit(v0, v5, b16, 16);
            float w1 = xsplit(v1, v6, b16, 16);
            float w2 = xsplit(v2, v7, b16, 16);
            float w3 = xsplit(v3, v8, b16, 16);
            v4 += __shfl_xor_sync(0xFFFFFFFFu, v4, 16);
            // step 2 (xor 8): pairs (w0,w1) (w2,w3)
            float u0 = xsplit(w0, w1, b8, 8);
            float u1 = xsplit(w2, w3, b8, 8);
            v4 += __shfl_xor_sync(0xFFFFFFFFu, v4, 8);
            // step 3 (xor 4): pair (u0,u1)
            float t0 = xsplit(u0, u1, b4, 4);
            v4 += __shfl_xor_sync(0xFFFFFFFFu, v4, 4);
            // step 4 (xor 2): pair (t0, v4)
            float r = xsplit(t0, v4, b2, 2);
            // step 5 (xor 1)
            r += __shfl_xor_sync(0xFFFFFFFFu, r, 1);

            if (writer) atomicAdd(acc_lane + (size_t)__float_as_uint(c4.w) * 12, r);
        }
    }
}

void launch_render_bwd(const ViewParams& vp, BinningState b, ImageState img, const float* dL_dpix, float* acc,
                       cudaStream_t s) {
    dim3 grid(vp.tiles_x, vp.tiles_y, 1);
    render_bwd_kernel<<<grid, SGS_TILE_PIX, 0, s>>>(vp, img.ranges, img.tile_count, b.packed, img.final_T,
                                                   img.n_contrib, dL_dpix, acc);
}

}  // namespace sgs

// Backward tile compositing: per-pixel back-to-front gradient walk.
//
// Semantics follow $R/cuda_rasterizer/backward.cu:399-557:
//   walk the tile list from the back; skip entries with list position >= n_contrib(pixel);
//   recompute alpha with the same tests as forward; recover T by division T /= (1-alpha);
//   dL/dcolor = alpha*T*dL/dpix ; dL/dalpha via the accum_rec recursion + background term;
//   dL/dmean2D (scaled by 0.5W, 0.5H), dL/dconic (A,B,C), dL/dopacity.
//
// B200 design (the reference issues 9 global float atomics per contributing pixel x instance,
// SURVEY.md §2.1; the kernel is FP32-issue bound, so everything below is about instructions per
// blended pair):
//   * streams the dense tile-ordered `PackedInst` list written by the forward kernel through a 4-stage
//     shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier): no index gathers, no staging
//     instructions, and NO block-wide barrier in the main loop — a warp whose quadrant is cheap runs ahead of
//     its neighbours by up to 3 batches, and the last warp to leave a stage refills it;
//   * 128 threads per tile, 2 vertically adjacent pixels per thread, warp = 8x8 quadrant; packed f32x2
//     evaluation of the quadratic form (sgs_render_common.cuh);
//   * per-warp visit bitmaps: a warp only visits instances whose quadrant mask (computed once by the forward
//     kernel, stored in the top 4 bits of the record's Gaussian-id word) has its bit AND whose list position is
//     below the last contributor of some pixel of the warp;
//   * per pair only what depends on the pixel is computed:  g = G dL/dalpha  and  w = alpha T ;
//     the warp accumulates the MOMENTS  sum g, sum g dx, sum g dy, sum g dx^2, sum g dx dy, sum g dy^2
//     and  sum w dL/dpix[c]  — the multiplications by opacity, conic and 0.5W/0.5H happen once per
//     Gaussian in the fused backward-preprocess kernel;
//   * the accum_rec recursion is carried as ONE scalar  a = sum_c accum_rec[c] dL/dpix[c]  (the
//     reference carries 3 colours and re-dots them with dL/dpix for every pair);
//   * the 9 sums are reduced across the warp's 64 pixels with a 12-shuffle "transposing" butterfly,
//     then one RED.ADD.F32 per value per (warp, instance) into a [P][12] accumulator
//     => 64x fewer L2 atomics than the reference.
#include "sgs_render_common.cuh"

namespace sgs {

// send `hi` to the partner if this lane keeps `lo`, and vice versa; returns kept + received
__forceinline__ __device__ float xsplit(float lo, float hi, bool upper, int xorm) {
    const float send = upper ? lo : hi;
    const float keep = upper ? hi : lo;
    return keep + __shfl_xor_sync(0xFFFFFFFFu, send, xorm);
}

// State of the two pixels of a thread, kept as f32x2 pairs so that the per-pair maths after exp() runs on the
// packed FMUL2 / FFMA2 / FADD2 instructions (one issue slot for both pixels).
struct BwdPix2 {
    float2 T;           // transmittance in front of the instance being visited
    float2 a_rec;       // sum_c accum_rec[c] * dL/dpix[c]
    float2 last_alpha;  // alpha of the previously visited (= next deeper) blended instance
    float2 last_cd;     // its colour . dL/dpix
    float2 last_om;     // 1 - last_alpha
    float2 d0, d1, d2;  // dL/dpix
    float2 bgT;         // -T_final * (bg . dL/dpix)
    uint32_t lc0, lc1;  // n_contrib of the two pixels
};

__forceinline__ __device__ float2 bcast2(float v) { return make_float2(v, v); }
__forceinline__ __device__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }

// two pixels x one instance: g = G * dL/dalpha and w = alpha * T per pixel (0 when the pair did not blend).
// A pixel that does not blend runs through the same packed code with alpha = G = 0: then 1 - alpha = 1, its
// reciprocal is exactly 1, T is unchanged, g = w = 0, and the pending accum_rec fold  a <- la*lcd + lom*a  is
// merely applied one step early (afterwards la = 0, lom = 1, so the next fold returns a bit for bit).
__forceinline__ __device__ void grad_pixels2(BwdPix2& s, float2 pw, bool act0, bool act1, float o, const float4 c,
                                             float2& g, float2& w) {
    // accurate expf() and a < 1 ulp reciprocal, like forward / the reference: alpha must equal forward's bit for
    // bit and T is recovered by a long product of 1/(1-alpha) factors — a bare ex2.approx / rcp.approx (~1e-7
    // each, but biased) drifts T by ~n * 1e-7 over n blended instances (measured: 2.5x the error on dL/dmeans3D).
    float2 G = expf2_exact(pw);        // both pixels, branch-free, bit-identical to expf (sgs_render_common.cuh)
    if (!act0) G.x = 0.f;
    if (!act1) G.y = 0.f;
    float2 alpha = __fmul2_rn(bcast2(o), G);
    alpha.x = min(0.99f, alpha.x);
    alpha.y = min(0.99f, alpha.y);
    if (alpha.x < 1.0f / 255.0f) { alpha.x = 0.f; G.x = 0.f; }
    if (alpha.y < 1.0f / 255.0f) { alpha.y = 0.f; G.y = 0.f; }
    const float2 om = __fadd2_rn(bcast2(1.f), neg2(alpha));   // in [0.01, 1]
    float2 inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(om.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(om.y));
    inv = __ffma2_rn(inv, __ffma2_rn(neg2(om), inv, bcast2(1.f)), inv);   // one Newton step: < 1 ulp
    s.T = __fmul2_rn(s.T, inv);                                            // $R/.../backward.cu:503
    s.a_rec = __ffma2_rn(s.last_alpha, s.last_cd, __fmul2_rn(s.last_om, s.a_rec));   // :515-519, dotted with dL/dpix
    const float2 cd = __ffma2_rn(bcast2(c.z), s.d2, __ffma2_rn(bcast2(c.y), s.d1, __fmul2_rn(bcast2(c.x), s.d0)));
    const float2 dL_dalpha = __ffma2_rn(__fadd2_rn(cd, neg2(s.a_rec)), s.T, __fmul2_rn(s.bgT, inv));   // :519-534
    g = __fmul2_rn(G, dL_dalpha);
    w = __fmul2_rn(alpha, s.T);
    s.last_alpha = alpha;
    s.last_cd = cd;
    s.last_om = om;
}

// ---- TMA / mbarrier helpers (sm_90+ PTX; SASS: UBLKCP + SYNCS) ---------------------------------------
__forceinline__ __device__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__forceinline__ __device__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__forceinline__ __device__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes: multiple of 16, 16-B aligned)
__forceinline__ __device__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__forceinline__ __device__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

#define SGS_B_STAGES 3
#define SGS_B_STAGE_BYTES (SGS_R_BATCH * 48)

__global__ void __launch_bounds__(SGS_R_THREADS)
render_bwd_kernel(const __grid_constant__ ViewParams vp, const uint2* __restrict__ ranges,
                  const uint32_t* __restrict__ tile_count, const PackedInst* __restrict__ packed,
                  const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ dL_dpix, float* __restrict__ acc) {
    // ring of record batches filled by TMA bulk copies; no block-wide barrier in the main loop
    __shared__ __align__(128) unsigned char s_rec[SGS_B_STAGES * SGS_B_STAGE_BYTES];
    __shared__ __align__(8) uint64_t s_full[SGS_B_STAGES];   // mbarriers: "batch has landed"
    __shared__ uint32_t s_done[SGS_B_STAGES];                // warps that finished the batch in this stage

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

    const uint32_t start = ranges[tile].x;
    const int count = (int)tile_count[tile];
    if (count == 0) return;
    const int nb = (count + SGS_R_BATCH - 1) / SGS_R_BATCH;      // batches, walked from the back of the list
    const unsigned char* src = reinterpret_cast<const unsigned char*>(packed + start);
    uint32_t rec_base = (uint32_t)__cvta_generic_to_shared(s_rec);
    const uint32_t full_base = (uint32_t)__cvta_generic_to_shared(s_full);
    asm volatile("" : "+r"(rec_base));

    // batch j covers records [lo, hi) with hi = count - 128 j
    auto issue = [&](int j) {
        const int hi = count - j * SGS_R_BATCH, lo = max(0, hi - SGS_R_BATCH);
        const uint32_t bytes = (uint32_t)(hi - lo) * 48u;
        const int stage = j % SGS_B_STAGES;
        mbar_expect_tx(full_base + stage * 8, bytes);
        tma_load_1d(rec_base + stage * SGS_B_STAGE_BYTES, src + (size_t)lo * 48, bytes, full_base + stage * 8);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SGS_B_STAGES; s++) {
            mbar_init(full_base + s * 8, 1);
            s_done[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int j = 0; j < min(nb, SGS_B_STAGES); j++) issue(j);
    }

    BwdPix2 st;
    {
        const size_t HW = (size_t)H * W;
        const uint32_t id0 = (uint32_t)W * py0 + px, id1 = (uint32_t)W * py1 + px;
        const float Tf0 = in0 ? final_T[id0] : 0.f, Tf1 = in1 ? final_T[id1] : 0.f;
        st.T = make_float2(Tf0, Tf1);
        st.lc0 = in0 ? n_contrib[id0] : 0u;
        st.lc1 = in1 ? n_contrib[id1] : 0u;
        st.d0 = make_float2(in0 ? dL_dpix[id0] : 0.f, in1 ? dL_dpix[id1] : 0.f);
        st.d1 = make_float2(in0 ? dL_dpix[HW + id0] : 0.f, in1 ? dL_dpix[HW + id1] : 0.f);
        st.d2 = make_float2(in0 ? dL_dpix[2 * HW + id0] : 0.f, in1 ? dL_dpix[2 * HW + id1] : 0.f);
        const float b0 = vp.bg[0], b1 = vp.bg[1], b2 = vp.bg[2];
        st.bgT = make_float2(-Tf0 * (b0 * st.d0.x + b1 * st.d1.x + b2 * st.d2.x),     // :531-534
                             -Tf1 * (b0 * st.d0.y + b1 * st.d1.y + b2 * st.d2.y));
        st.a_rec = st.last_alpha = st.last_cd = make_float2(0.f, 0.f);
        st.last_om = make_float2(1.f, 1.f);
    }
    // warp-uniform bound: records at list positions >= this were blended by no pixel of the warp
    const uint32_t warp_last = __reduce_max_sync(0xFFFFFFFFu, max(st.lc0, st.lc1));

    // which accumulator slot this lane owns after the butterfly (see the reduction below)
    //   bit1 set -> value 4 ; else value = (bit4 ? 5 : 0) + (bit2 ? 2 : 0) + (bit3 ? 1 : 0)
    const int my_slot = (lane & 2) ? 4 : (((lane & 16) ? 5 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 8) ? 1 : 0));
    const bool writer = (lane & 1) == 0 && ((lane & 2) == 0 || lane == 2);
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
    float* const acc_lane = acc + my_slot;

    __syncthreads();   // mbarrier initialisation visible to every warp (the only block-wide barrier)

    for (int j = 0; j < nb; j++) {
        const int stage = j % SGS_B_STAGES;
        const int hi = count - j * SGS_R_BATCH, lo = max(0, hi - SGS_R_BATCH);
        const int nrec = hi - lo;
        const uint32_t sbase = rec_base + stage * SGS_B_STAGE_BYTES;
        mbar_wait(full_base + stage * 8, (uint32_t)((j / SGS_B_STAGES) & 1));

        // visit bitmap of this warp: quadrant bit (top 4 bits of the gid word) set and list position below the
        // warp's last contributor
        uint32_t mywords = 0;
#pragma unroll
        for (int k = 0; k < SGS_R_BATCH / 32; k++) {
            const int slot = k * 32 + lane;
            bool v = false;
            if (slot < nrec) {
                const uint32_t gw = lds32(sbase + slot * 48u + 44u);
                const uint32_t pos = lds32(sbase + slot * 48u + 28u);
                v = ((gw >> (28 + warp)) & 1u) && (pos < warp_last);
            }
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, v);
            if (lane == k) mywords = b;
        }

        for (int k = SGS_R_BATCH / 32 - 1; k >= 0; k--) {
            uint32_t word = __shfl_sync(0xFFFFFFFFu, mywords, k);
            while (word) {
                const uint32_t bit = 31u - __clz(word);
                word ^= 1u << bit;
                const uint32_t rec = sbase + (k * 32 + bit) * 48u;
                const float4 ra = lds128(rec);          // x, y, A, B
                const float4 rb = lds128(rec + 16u);    // C, opacity, thr, list_pos
                float dx;
                float2 dy;
                const float2 pw = power2(make_float4(ra.x, ra.y, ra.z, -ra.w), rb.x, pxf, npy, dx, dy);
                const uint32_t pos = __float_as_uint(rb.w);
                // same tests as forward: power > 0 -> skip; power < thr -> provably alpha < 1/255
                // (a separate cheaper warp-level pre-test was measured slower: with the visit bitmaps
                // nearly every visit blends something, so the exact tests are needed anyway)
                const bool act0 = (pos < st.lc0) && !(pw.x > 0.0f) && !(pw.x < rb.z);
                const bool act1 = (pos < st.lc1) && !(pw.y > 0.0f) && !(pw.y < rb.z);
                if (!__any_sync(0xFFFFFFFFu, act0 || act1)) continue;

                const float4 c = lds128(rec + 32u);   // r, g, b, gid | quadrant mask << 28
                float2 g, w;
                grad_pixels2(st, pw, act0, act1, rb.y, c, g, w);

                // moments of g over this thread's two pixels (dx shared)
                const float2 gy = __fmul2_rn(g, dy);
                float v5 = g.x + g.y;                       // sum g            -> dL/dopacity
                float v0 = v5 * dx;                         // sum g dx
                float v1 = gy.x + gy.y;                     // sum g dy
                float v2 = v0 * dx;                         // sum g dx^2
                float v3 = v1 * dx;                         // sum g dx dy
                float v4 = fmaf(gy.y, dy.y, gy.x * dy.x);   // sum g dy^2
                float v6 = fmaf(w.y, st.d0.y, w.x * st.d0.x);   // sum alpha T dL/dpix[c]  -> dL/dcolour
                float v7 = fmaf(w.y, st.d1.y, w.x * st.d1.x);
                float v8 = fmaf(w.y, st.d2.y, w.x * st.d2.x);

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

                if (writer) atomicAdd(acc_lane + (size_t)(__float_as_uint(c.w) & 0x0FFFFFFFu) * 12, r);
            }
        }

        // this warp is done with the stage; the LAST warp to finish refills it with batch j + STAGES
        __syncwarp();
        if (lane == 0 && j + SGS_B_STAGES < nb) {
            const uint32_t old = atomicAdd(&s_done[stage], 1u);
            if (old == SGS_R_THREADS / 32 - 1) {
                s_done[stage] = 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(j + SGS_B_STAGES);
            }
        }
    }
}

void launch_render_bwd(const ViewParams& vp, BinningState b, ImageState img, const float* dL_dpix, float* acc,
                       cudaStream_t s) {
    dim3 grid(vp.tiles_x, vp.tiles_y, 1);
    render_bwd_kernel<<<grid, SGS_R_THREADS, 0, s>>>(vp, img.ranges, img.tile_count, b.packed, img.final_T,
                                                    img.n_contrib, dL_dpix, acc);
}

}  // namespace sgs

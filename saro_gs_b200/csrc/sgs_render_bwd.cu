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
//   * reduction across the warp's 64 pixels (G = SGS_B_STASH visits at a time): every lane parks its SIX
//     dx-independent partial sums (sum g, sum g dy, sum g dy^2, sum w dL/dpix[c]) of a visit in a per-warp
//     shared-memory stash; after G visits lane l owns (visit l / 6, value l % 6): it reads the 32 partials as
//     8 x LDS.128, adds the four lane rows, and only then forms the dx moments from the 8 COLUMN sums
//     (all lanes of a column share dx) — sum g dx, sum g dx^2, sum g dx dy cost nothing per visit.  One
//     RED.ADD.F32 per value per (warp, instance) into a [P][12] accumulator => 64x fewer L2 atomics than the
//     reference, and ~25 instead of ~55 instructions per visit for the reduction compared with the 12-shuffle
//     transposing butterfly of round 1 (kept as the G = 0 variant for A/B measurements).
#include <cstdlib>

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
    float2 a_rec;       // sum_c accum_rec[c] * dL/dpix[c] over the instances behind the one being visited
    float2 d0, d1, d2;  // dL/dpix
    float2 bgT;         // -T_final * (bg . dL/dpix)
    uint32_t lc0, lc1;  // n_contrib of the two pixels
};

__forceinline__ __device__ float2 bcast2(float v) { return make_float2(v, v); }
__forceinline__ __device__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }

// two pixels x one instance: g = G * dL/dalpha and w = alpha * T per pixel (0 when the pair did not blend).
// A pixel that does not blend runs through the same packed code with alpha = G = 0: then 1 - alpha = 1, its
// reciprocal is exactly 1, T is unchanged, g = w = 0, and the accum_rec fold  a <- alpha*cd + (1 - alpha)*a  returns a
// bit for bit.  The reference folds the PREVIOUS blended instance at the start of the next one
// ($R/cuda_rasterizer/backward.cu:515-519); folding the current one at the end of its own visit is the same
// sequence of operations on the same values without three carried register pairs.
__forceinline__ __device__ void grad_pixels2(BwdPix2& s, float2 pw, bool act0, bool act1, float o, const float4 c,
                                             float2& g, float2& w) {
    // accurate expf() and a < 1 ulp reciprocal, like forward / the reference: alpha must equal forward's bit for
    // bit and T is recovered by a long product of 1/(1-alpha) factors — a bare ex2.approx / rcp.approx (~1e-7
    // each, but biased) drifts T by ~n * 1e-7 over n blended instances (measured: 2.5x the error on dL/dmeans3D).
    float2 G = expf2_exact(pw);        // both pixels, branch-free, bit-identical to expf (sgs_render_common.cuh)
    float2 alpha = __fmul2_rn(bcast2(o), G);
    // one predicate per pixel: passed the power tests AND alpha >= 1/255 (an inactive lane's G may be anything,
    // NaN included: the comparison is then false or irrelevant, the selects below drop the value)
    const bool go0 = act0 && !(alpha.x < 1.0f / 255.0f);
    const bool go1 = act1 && !(alpha.y < 1.0f / 255.0f);
    alpha.x = go0 ? min(0.99f, alpha.x) : 0.f;
    alpha.y = go1 ? min(0.99f, alpha.y) : 0.f;
    G.x = go0 ? G.x : 0.f;
    G.y = go1 ? G.y : 0.f;
    const float2 om = __fadd2_rn(bcast2(1.f), neg2(alpha));   // in [0.01, 1]
    float2 inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(om.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(om.y));
    inv = __ffma2_rn(inv, __ffma2_rn(neg2(om), inv, bcast2(1.f)), inv);   // one Newton step: < 1 ulp
    s.T = __fmul2_rn(s.T, inv);                                            // $R/.../backward.cu:503
    const float2 cd = __ffma2_rn(bcast2(c.z), s.d2, __ffma2_rn(bcast2(c.y), s.d1, __fmul2_rn(bcast2(c.x), s.d0)));
    const float2 dL_dalpha = __ffma2_rn(__fadd2_rn(cd, neg2(s.a_rec)), s.T, __fmul2_rn(s.bgT, inv));   // :519-534
    g = __fmul2_rn(G, dL_dalpha);
    w = __fmul2_rn(alpha, s.T);
    s.a_rec = __ffma2_rn(alpha, cd, __fmul2_rn(om, s.a_rec));              // :515-519, dotted with dL/dpix
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

template <int OFF>
__forceinline__ __device__ void sts32_off(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(addr), "f"(v), "n"(OFF) : "memory");
}
template <int OFF>
__forceinline__ __device__ float4 lds128_off(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr), "n"(OFF));
    return v;
}
__forceinline__ __device__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__forceinline__ __device__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__forceinline__ __device__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }

#define SGS_B_STAGES 3
#define SGS_B_PAIR_STRIDE 144   // bytes per stashed (visit, value) row: 32 lanes x 4 B + 16 B pad (conflict-free LDS.128)

// BATCH: records per TMA batch.  G: visits per stash flush (0 = round-1 butterfly reduction).
// (Measured and dropped: one single-warp CTA per 8x8 quadrant with its own record ring — no warp waits for a slower
// quadrant of its tile, but 4x the L2 reads and 24 instead of 32 resident warps: 348 us against 326 us.)
template <int BATCH, int G, int MINB>
__global__ void __launch_bounds__(SGS_R_THREADS, MINB)
render_bwd_kernel(const __grid_constant__ ViewParams vp, const uint2* __restrict__ ranges,
                  const uint32_t* __restrict__ tile_count, const PackedInst* __restrict__ packed,
                  const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ dL_dpix, float* __restrict__ acc) {
    // ring of record batches filled by TMA bulk copies; no block-wide barrier in the main loop
    constexpr int SGS_B_STAGE_BYTES = BATCH * 48;
    constexpr int STASH_WARP_BYTES = G > 0 ? G * 8 * SGS_B_PAIR_STRIDE : 16;
    __shared__ __align__(128) unsigned char s_rec[SGS_B_STAGES * SGS_B_STAGE_BYTES];
    __shared__ __align__(16) unsigned char s_stash[(SGS_R_THREADS / 32) * STASH_WARP_BYTES];
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

    pdl_wait();     // launched programmatically dependent on whatever precedes it in the stream (forward render, loss)
    const uint32_t start = ranges[tile].x;
    const int count = (int)tile_count[tile];
    if (count == 0) return;
    const int nb = (count + BATCH - 1) / BATCH;      // batches, walked from the back of the list
    const unsigned char* src = reinterpret_cast<const unsigned char*>(packed + start);
    uint32_t rec_base = (uint32_t)__cvta_generic_to_shared(s_rec);
    const uint32_t full_base = (uint32_t)__cvta_generic_to_shared(s_full);
    asm volatile("" : "+r"(rec_base));

    // batch j covers records [lo, hi) with hi = count - BATCH j
    auto issue = [&](int j) {
        const int hi = count - j * BATCH, lo = max(0, hi - BATCH);
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
        st.a_rec = make_float2(0.f, 0.f);
    }
    // warp-uniform bound: records at list positions >= this were blended by no pixel of the warp
    const uint32_t warp_last = __reduce_max_sync(0xFFFFFFFFu, max(st.lc0, st.lc1));

    // G == 0: which accumulator slot this lane owns after the butterfly (see the reduction below)
    //   bit1 set -> value 4 ; else value = (bit4 ? 5 : 0) + (bit2 ? 2 : 0) + (bit3 ? 1 : 0)
    const int my_slot = (lane & 2) ? 4 : (((lane & 16) ? 5 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 8) ? 1 : 0));
    const bool writer = (lane & 1) == 0 && ((lane & 2) == 0 || lane == 2);
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
    float* const acc_lane = acc + my_slot;

    // G > 0: the warp's stash = G visits x 8 rows of 32 lane values (+ pad): rows 0..5 the partial sums
    // (0 sum g, 1 sum g dy, 2 sum g dy^2, 3..5 sum w dL/dpix[c]), row 6 the lanes' dx, row 7 the Gaussian id (32
    // copies: unconditional full-warp stores need neither a predicate nor an address of their own).
    // At a flush lane l reduces row (visit l / 6, value l % 6).  Accumulator slots (consumed by
    // preprocess_bwd_kernel): 0 g dx, 1 g dy, 2 g dx^2, 3 g dx dy, 4 g dy^2, 5 g, 6..8 colour.
    const uint32_t stash_base = (uint32_t)__cvta_generic_to_shared(s_stash) + warp * STASH_WARP_BYTES;
    const int fl_v = lane / 6, fl_k = lane - 6 * fl_v;
    const int fl_slot = fl_k == 0 ? 5 : (fl_k == 1 ? 1 : (fl_k == 2 ? 4 : fl_k + 3));
    uint32_t stash_wr = stash_base + lane * 4u;
    int nst = 0;                                            // visits in the stash (warp-uniform)
    auto flush = [&](int nv) {
        __syncwarp();
        if (lane < nv * 6) {
            const uint32_t vis = stash_base + fl_v * (8 * SGS_B_PAIR_STRIDE);
            const uint32_t row = vis + fl_k * SGS_B_PAIR_STRIDE;
            // partial i came from lane i = 8 r + c (r: pixel-row pair, c: pixel column): add the four r
            const float4 q0 = lds128_off<0>(row), q1 = lds128_off<16>(row), q2 = lds128_off<32>(row),
                         q3 = lds128_off<48>(row), q4 = lds128_off<64>(row), q5 = lds128_off<80>(row),
                         q6 = lds128_off<96>(row), q7 = lds128_off<112>(row);
            // dx of the 8 pixel columns (as the visit computed them) and the Gaussian id
            const float4 dA = lds128_off<6 * SGS_B_PAIR_STRIDE>(vis), dB = lds128_off<6 * SGS_B_PAIR_STRIDE + 16>(vis);
            const uint32_t gid = lds32(vis + 7 * SGS_B_PAIR_STRIDE);
            const float2 c01 = add2(add2(lo2(q0), lo2(q2)), add2(lo2(q4), lo2(q6)));
            const float2 c23 = add2(add2(hi2(q0), hi2(q2)), add2(hi2(q4), hi2(q6)));
            const float2 c45 = add2(add2(lo2(q1), lo2(q3)), add2(lo2(q5), lo2(q7)));
            const float2 c67 = add2(add2(hi2(q1), hi2(q3)), add2(hi2(q5), hi2(q7)));
            const float2 d01 = lo2(dA), d23 = hi2(dA), d45 = lo2(dB), d67 = hi2(dB);
            const float2 t01 = __fmul2_rn(c01, d01), t23 = __fmul2_rn(c23, d23), t45 = __fmul2_rn(c45, d45),
                         t67 = __fmul2_rn(c67, d67);
            const float2 s0 = add2(add2(c01, c23), add2(c45, c67));
            const float2 s1 = add2(add2(t01, t23), add2(t45, t67));
            const float2 s2 = __ffma2_rn(t01, d01, __ffma2_rn(t23, d23, __ffma2_rn(t45, d45, __fmul2_rn(t67, d67))));
            float* const dst = acc + (size_t)(gid & 0x0FFFFFFFu) * 12;
            atomicAdd(dst + fl_slot, s0.x + s0.y);
            if (fl_k < 2) atomicAdd(dst + (fl_k == 0 ? 0 : 3), s1.x + s1.y);
            if (fl_k == 0) atomicAdd(dst + 2, s2.x + s2.y);
        }
        __syncwarp();
    };

    __syncthreads();   // mbarrier initialisation visible to every warp (the only block-wide barrier)

    for (int j = 0; j < nb; j++) {
        const int stage = j % SGS_B_STAGES;
        const int hi = count - j * BATCH, lo = max(0, hi - BATCH);
        const int nrec = hi - lo;
        const uint32_t sbase = rec_base + stage * SGS_B_STAGE_BYTES;
        mbar_wait(full_base + stage * 8, (uint32_t)((j / SGS_B_STAGES) & 1));

        // visit bitmap of this warp: quadrant bit (top 4 bits of the gid word) set and list position below the
        // warp's last contributor
        uint32_t mywords = 0;
#pragma unroll
        for (int k = 0; k < BATCH / 32; k++) {
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

        for (int k = BATCH / 32 - 1; k >= 0; k--) {
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
                float v1 = gy.x + gy.y;                     // sum g dy
                float v4 = fmaf(gy.y, dy.y, gy.x * dy.x);   // sum g dy^2
                float v6 = fmaf(w.y, st.d0.y, w.x * st.d0.x);   // sum alpha T dL/dpix[c]  -> dL/dcolour
                float v7 = fmaf(w.y, st.d1.y, w.x * st.d1.x);
                float v8 = fmaf(w.y, st.d2.y, w.x * st.d2.x);

                if (G > 0) {
                    // park the six dx-independent partials; the dx moments are formed per column at the flush
                    sts32_off<0 * SGS_B_PAIR_STRIDE>(stash_wr, v5);
                    sts32_off<1 * SGS_B_PAIR_STRIDE>(stash_wr, v1);
                    sts32_off<2 * SGS_B_PAIR_STRIDE>(stash_wr, v4);
                    sts32_off<3 * SGS_B_PAIR_STRIDE>(stash_wr, v6);
                    sts32_off<4 * SGS_B_PAIR_STRIDE>(stash_wr, v7);
                    sts32_off<5 * SGS_B_PAIR_STRIDE>(stash_wr, v8);
                    sts32_off<6 * SGS_B_PAIR_STRIDE>(stash_wr, dx);
                    sts32_off<7 * SGS_B_PAIR_STRIDE>(stash_wr, c.w);
                    stash_wr += 8 * SGS_B_PAIR_STRIDE;
                    if (++nst == G) {
                        flush(G);
                        nst = 0;
                        stash_wr = stash_base + lane * 4u;
                    }
                    continue;
                }
                float v0 = v5 * dx;                         // sum g dx
                float v2 = v0 * dx;                         // sum g dx^2
                float v3 = v1 * dx;                         // sum g dx dy

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
    if (G > 0 && nst) flush(nst);
}

void launch_render_bwd(const ViewParams& vp, BinningState b, ImageState img, const float* dL_dpix, float* acc,
                       cudaStream_t s) {
    dim3 grid(vp.tiles_x, vp.tiles_y, 1);
    // developer switch for A/B measurements (tools/gpu_ab_bwd.sh): SGS_BWD_VARIANT = 0 round-1 butterfly reduction
    // (369 us at configs[1]); stash reduction: 1 (default) 32-record batches, 5 visits per flush, 8 CTAs per SM
    // (326 us); 2 64-record batches (334 us); 3 64-record batches, 4 visits per flush (330 us); 4 128-record batches,
    // 5 CTAs per SM (352 us)
    static const int variant = [] {
        const char* e = getenv("SGS_BWD_VARIANT");
        return e ? atoi(e) : 1;
    }();
#define SGS_LAUNCH_RB(BATCH, G, MINB)                                                                        \
    launch_pdl(render_bwd_kernel<BATCH, G, MINB>, grid, dim3(SGS_R_THREADS), 0, s, vp, (const uint2*)img.ranges,   \
               (const uint32_t*)img.tile_count, (const PackedInst*)b.packed, (const float*)img.final_T,          \
               (const uint32_t*)img.n_contrib, dL_dpix, acc)
    if (variant == 0) SGS_LAUNCH_RB(128, 0, 10);
    else if (variant == 2) SGS_LAUNCH_RB(64, 5, 7);
    else if (variant == 3) SGS_LAUNCH_RB(64, 4, 8);
    else if (variant == 4) SGS_LAUNCH_RB(128, 5, 6);
    else SGS_LAUNCH_RB(32, 5, 8);
#undef SGS_LAUNCH_RB
}

}  // namespace sgs

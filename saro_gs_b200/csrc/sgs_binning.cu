// Tile binning: produces, for every 16x16 tile, the list of Gaussian instances that touch
// it, ordered by (depth bits, Gaussian index) — the exact order the reference obtains
// from one stable 64-bit radix sort of  key = tile<<32 | float_bits(depth)
//   $R/cuda_rasterizer/rasterizer_impl.cu:70-111 (duplicateWithKeys)
//   $R/cuda_rasterizer/rasterizer_impl.cu:299-309 (SortPairs on bits [0, 32+bit))
//   $R/cuda_rasterizer/rasterizer_impl.cu:116-138 (identifyTileRanges)
//
// B200 design: the R = 3.9 M (tile, depth) instances of the reference are NEVER sorted.
//   0. (in the preprocess kernel) the reference's 3-sigma tile rect of every Gaussian is clipped to
//      the exact axis-aligned bounding box of its  alpha >= 1/255  ellipse (`rect_kept`).  The
//      reference's own count (`tiles_touched`) is kept for the API-visible num_rendered; a dropped
//      instance would `continue` on all 256 pixels of its tile in the reference, so image, depth and
//      gradients are unchanged bit for bit.  (SGS_FLAG_NO_TILE_CULL keeps the full rect: then ranges /
//      point_list / n_contrib equal the reference's bit for bit.)
//   1. depth_sort_kernel (persistent, cooperative): stable LSD radix sort of the P Gaussians by their depth bits,
//      normalised to the frame's [min, max] key range (26 bits = 3 passes of 9 at config 2), then the scan, in depth
//      order, of the number of SUPERTILES (4x4 tiles = 64x64 pixels) each Gaussian's kept rect overlaps.  The
//      totals (kept instances, the reference's num_rendered, visible Gaussians) go to the host through a pinned slot.
//   2. coarse_sort_kernel (persistent, cooperative): every block generates its slice of the depth-ordered
//      (supertile, Gaussian) stream straight into shared memory and ONE stable radix pass (9 bits: 352 supertiles at
//      1352x1014) buckets it: ~0.75 M coarse instances instead of 2.2 M tile instances, 4 bytes scattered each.
//   3. tile_count_kernel / tile_fill_kernel (one block per supertile): stream the supertile's depth-ordered list and
//      expand it into its 16 per-tile lists with ballot-based stable compaction — the only per-instance cost of the
//      whole binning is one coalesced 4-byte store.  Per-tile ranges come from the counts (scan by the last block).
//   A stable bucketing of a depth-ordered stream is ordered by (bucket, depth, index), and the expansion keeps the
//   order inside every tile: the reference order restricted to the kept instances.
//
// One radix pass = every block ranks its slice in shared memory (warp-private histograms, ballot-based peer
// masks), publishes its digit histogram, grid-wide barrier, the blocks share the column scans of the histogram
// matrix, grid-wide barrier, every block scatters.  All counts are read on the device: the host never waits between
// the stages (see sgs_api.cu).  Round 1 used CUB here: 10 launches and 34-CTA decoupled look-back chains, 205 us.
#include "sgs_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace sgs {

#define SGS_SORT_THREADS 1024
#define SGS_SORT_WARPS 32
#define SGS_SORT_ND 512        // max digits per pass (9 bits)
#define SGS_SORT_CHUNK 16384   // items a block keeps resident in shared memory
#define SGS_DUP_SMALL 4
#define SGS_ST 4               // supertile edge in tiles
#define SGS_EXP_THREADS 512    // expansion kernels: threads per supertile block

// ------------------------------------------------------------------------------------------------
// grid-wide barrier (all blocks co-resident: cooperative launch, one block per SM)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_barrier(uint32_t* ctr, uint32_t& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire(ctr) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// Optional phase timestamps (developer aid, sgs_debug_binning_profile): block 0 writes %globaltimer into a pinned host
// buffer at the phase boundaries.  `prof` is NULL in normal operation.
__device__ __forceinline__ void prof_mark(unsigned long long* prof, int& slot) {
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        prof[slot] = t;
    }
    slot++;
}

// ------------------------------------------------------------------------------------------------
// One radix pass over a slice held in shared memory
// ------------------------------------------------------------------------------------------------
struct SortSmem {
    uint16_t whist[SGS_SORT_WARPS][SGS_SORT_ND];   // per-warp digit counts -> exclusive prefix over the warps
    uint32_t cnt[SGS_SORT_ND];                      // digit counts of the slice
    uint32_t base[SGS_SORT_ND];                     // global offset of the slice's first element of every digit
    uint32_t wsum[32];
    uint32_t key[SGS_SORT_CHUNK];
    uint32_t val[SGS_SORT_CHUNK];
    uint16_t rank[SGS_SORT_CHUNK];
};

__device__ __forceinline__ uint32_t keys_per_warp(uint32_t n) {
    return (((n + SGS_SORT_WARPS - 1) / SGS_SORT_WARPS) + 31u) & ~31u;
}

// Stable ranks of sm.key[0..n) on digit (key >> shift) & (nd - 1), nd = 1 << dbits.  Warp w owns the contiguous keys
// [w * per, (w + 1) * per); on return  position within the slice's digit group = whist[w][d] + rank[i], and
// sm.cnt[d] holds the slice's digit histogram.  Peer masks are built from one ballot per digit bit (match.any costs
// ~200 cycles per step on sm_100 when the 32 digits differ, which is the common case here).
__device__ __forceinline__ void rank_slice(SortSmem& sm, uint32_t n, uint32_t shift, uint32_t dbits) {
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nd = 1u << dbits;
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&sm.whist[0][0]);
        constexpr uint32_t words = SGS_SORT_WARPS * SGS_SORT_ND / 2;
        for (uint32_t i = tid; i < words; i += SGS_SORT_THREADS) z[i] = 0u;
    }
    __syncthreads();
    const uint32_t per = keys_per_warp(n);
    const uint32_t beg = warp * per, end = min(n, beg + per);
    uint16_t* wh = sm.whist[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t i0 = beg; i0 < end; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < end;
        const uint32_t d = valid ? ((sm.key[i] >> shift) & (nd - 1u)) : 0u;
        uint32_t peers = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
        for (uint32_t b = 0; b < 9; b++) {
            if (b < dbits) {
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? m : ~m;
            }
        }
        const uint32_t before = peers & lt_mask;
        uint32_t old = 0;
        if (valid) old = wh[d];
        __syncwarp();
        if (valid && before == 0u) wh[d] = (uint16_t)(old + __popc(peers));
        __syncwarp();
        if (valid) sm.rank[i] = (uint16_t)(old + __popc(before));
    }
    __syncthreads();
    // exclusive prefix over the 32 warps, per digit: (digit, half) pairs so that all 1024 threads work
    {
        const uint32_t d = tid & (SGS_SORT_ND - 1u), half = tid >> 9;   // warps [16 half, 16 half + 16)
        uint32_t run = 0;
        if (d < nd) {
#pragma unroll 8
            for (int w = 0; w < 16; w++) {
                const uint32_t c = sm.whist[half * 16 + w][d];
                sm.whist[half * 16 + w][d] = (uint16_t)run;
                run += c;
            }
            if (half == 0) sm.cnt[d] = run;
        }
        __syncthreads();
        if (d < nd && half == 1) {
            const uint32_t first = sm.cnt[d];
            sm.cnt[d] = first + run;
#pragma unroll 8
            for (int w = 16; w < 32; w++) sm.whist[w][d] = (uint16_t)(sm.whist[w][d] + first);
        }
    }
    __syncthreads();
}

// hist layout: [vblocks + 1][nd]; row `vblocks` receives the digit totals
__device__ __forceinline__ void scatter_slice_pairs(SortSmem& sm, uint32_t n, uint32_t shift, uint32_t nd,
                                                    unsigned long long* __restrict__ out) {
    const uint32_t per = keys_per_warp(n);
    for (uint32_t i = threadIdx.x; i < n; i += SGS_SORT_THREADS) {
        const uint32_t k = sm.key[i];
        const uint32_t d = (k >> shift) & (nd - 1u);
        const uint32_t dst = sm.base[d] + sm.whist[i / per][d] + sm.rank[i];
        out[dst] = (unsigned long long)sm.val[i] | ((unsigned long long)k << 32);
    }
}

__device__ __forceinline__ void publish_hist(const uint32_t* cnt, uint32_t* __restrict__ hist, uint32_t v, uint32_t nd) {
    for (uint32_t d = threadIdx.x; d < nd; d += SGS_SORT_THREADS) __stcg(hist + (size_t)v * nd + d, cnt[d]);
}

// Column scans of the histogram matrix, shared by the blocks: block b owns the digits [b dpb, (b + 1) dpb), one warp
// per digit; hist[v][d] becomes the number of keys with digit d in the slices before v, row `vblocks` the totals.
__device__ __forceinline__ void column_scans(uint32_t* __restrict__ hist, uint32_t vblocks, uint32_t nd) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t dpb = (nd + gridDim.x - 1) / gridDim.x;
    for (uint32_t dd = warp; dd < dpb; dd += SGS_SORT_WARPS) {
        const uint32_t d = blockIdx.x * dpb + dd;
        if (d >= nd) break;
        uint32_t carry = 0;
        for (uint32_t v0 = 0; v0 < vblocks; v0 += 128) {
            uint32_t x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t v = v0 + u * 32 + lane;
                x[u] = v < vblocks ? __ldcg(hist + (size_t)v * nd + d) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t v = v0 + u * 32 + lane;
                uint32_t inc = x[u];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if (lane >= (uint32_t)o) inc += y;
                }
                if (v < vblocks) __stcg(hist + (size_t)v * nd + d, carry + inc - x[u]);
                carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
        }
        if (lane == 0) __stcg(hist + (size_t)vblocks * nd + d, carry);
    }
}

// sm.base[d] = (number of keys with a smaller digit anywhere) + (keys with digit d in the slices before `v`);
// also leaves the exclusive scan of the digit totals in sm.cnt (the bucket starts of a single-pass sort).
__device__ __forceinline__ void slice_bases(SortSmem& sm, const uint32_t* __restrict__ hist, uint32_t v, uint32_t vblocks,
                                            uint32_t nd) {
    const uint32_t tid = threadIdx.x;
    if (tid < 512) {
        const uint32_t T = tid < nd ? __ldcg(hist + (size_t)vblocks * nd + tid) : 0u;
        const uint32_t pre = tid < nd ? __ldcg(hist + (size_t)v * nd + tid) : 0u;
        const uint32_t lane = tid & 31, warp = tid >> 5;
        uint32_t inc = T;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        if (lane == 31) sm.wsum[warp] = inc;
        asm volatile("bar.sync 1, 512;");
        uint32_t woff = 0;
        for (uint32_t w = 0; w < warp; w++) woff += sm.wsum[w];
        if (tid < nd) {
            sm.cnt[tid] = woff + inc - T;
            sm.base[tid] = woff + inc - T + pre;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void scatter_slice(SortSmem& sm, uint32_t n, uint32_t shift, uint32_t nd,
                                              uint32_t* __restrict__ out_key, uint32_t* __restrict__ out_val) {
    const uint32_t per = keys_per_warp(n);
    for (uint32_t i = threadIdx.x; i < n; i += SGS_SORT_THREADS) {
        const uint32_t k = sm.key[i];
        const uint32_t d = (k >> shift) & (nd - 1u);
        const uint32_t dst = sm.base[d] + sm.whist[i / per][d] + sm.rank[i];
        if (out_key) out_key[dst] = k;
        out_val[dst] = sm.val[i];
    }
}

// bit (4 ly + lx) set  <=>  the kept rect covers tile (tx0 + lx, ty0 + ly)
__device__ __forceinline__ uint32_t tile_mask16(const ushort4 r, uint32_t tx0, uint32_t ty0) {
    uint32_t xm = 0, ym = 0;
#pragma unroll
    for (uint32_t l = 0; l < SGS_ST; l++) {
        xm |= ((tx0 + l >= r.x) && (tx0 + l < r.y)) ? (1u << l) : 0u;
        ym |= ((ty0 + l >= r.z) && (ty0 + l < r.w)) ? (1u << l) : 0u;
    }
    uint32_t m = 0;
#pragma unroll
    for (uint32_t l = 0; l < SGS_ST; l++) m |= ((ym >> l) & 1u) ? (xm << (4 * l)) : 0u;
    return m;
}

__device__ __forceinline__ ushort4 as_rect(unsigned long long v) {
    return make_ushort4((unsigned short)(v & 0xFFFF), (unsigned short)((v >> 16) & 0xFFFF), (unsigned short)((v >> 32) & 0xFFFF),
                        (unsigned short)(v >> 48));
}

// supertiles a kept tile rect overlaps: [x0, x1) x [y0, y1) in supertile units
__device__ __forceinline__ uint4 super_rect(const ushort4 r) {
    if (r.y <= r.x || r.w <= r.z) return make_uint4(0u, 0u, 0u, 0u);
    return make_uint4((uint32_t)r.x / SGS_ST, ((uint32_t)r.y + SGS_ST - 1) / SGS_ST, (uint32_t)r.z / SGS_ST,
                      ((uint32_t)r.w + SGS_ST - 1) / SGS_ST);
}

// ------------------------------------------------------------------------------------------------
// Kernel 1: depth sort of the Gaussians + scan of their supertile counts in depth order
// ------------------------------------------------------------------------------------------------
struct DepthArgs {
    int P;
    int sort;         // 1: depth-sort the Gaussians first (round-2a design); 0: index order, the depth sort happens inside
                      //    every supertile (tile_fill_sorted_kernel)
    int vblocks;      // virtual blocks (slices); multiple of gridDim.x; == gridDim.x  <=>  slices stay resident
    int slice;        // keys per slice
    const uint32_t* raw;
    const uint2* blk_range;   // per-preprocess-block (max key, max ~key) of the visible Gaussians
    const uint4* blk_sums;    // per-preprocess-block (kept, touched, visible) sums
    int n_blk_range;
    ushort4* rect_sorted;     // out: rect_kept in depth order
    uint32_t* keys[2];
    uint32_t* vals[2];
    const ushort4* rect_kept;
    const uint32_t* tiles_touched;
    uint32_t* coffs;
    uint32_t* hist;
    uint32_t* blocksum;
    BinCtl* ctl;
    HostSlot* slot;
    unsigned long long ticket;
    unsigned long long* prof;
};

// block-wide sum of three 64-bit counters (every thread gets the result)
struct Tri {
    unsigned long long kept, touched, vis;
};
__device__ __forceinline__ Tri block_sum3(Tri q, Tri* s_w) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        q.kept += __shfl_xor_sync(0xFFFFFFFFu, q.kept, o);
        q.touched += __shfl_xor_sync(0xFFFFFFFFu, q.touched, o);
        q.vis += __shfl_xor_sync(0xFFFFFFFFu, q.vis, o);
    }
    __syncthreads();
    if (lane == 0) s_w[warp] = q;
    __syncthreads();
    Tri t = s_w[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t.kept += __shfl_xor_sync(0xFFFFFFFFu, t.kept, o);
        t.touched += __shfl_xor_sync(0xFFFFFFFFu, t.touched, o);
        t.vis += __shfl_xor_sync(0xFFFFFFFFu, t.vis, o);
    }
    return t;
}

__device__ __forceinline__ void depth_sort_phases(const DepthArgs& a, SortSmem& sm, Tri* s_w, uint32_t* bar,
                                                  uint32_t& bar_target) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t P = (uint32_t)a.P, VB = (uint32_t)a.vblocks, SL = (uint32_t)a.slice;
    const bool resident = VB == gridDim.x;
    int pslot = 0;
    prof_mark(a.prof, pslot);

    // ---- phase 0: key range of the visible Gaussians from the preprocess kernel's per-block partials; every block
    // reduces all of them itself (a few KB from L2), so no grid-wide barrier is needed for the range
    uint32_t key_max, key_nmin;
    {
        uint32_t kmax = 0u, knmin = 0u;
        Tri tot = {0ull, 0ull, 0ull};
        for (int i = tid; i < a.n_blk_range; i += SGS_SORT_THREADS) {
            const uint2 r = __ldcg(a.blk_range + i);
            kmax = max(kmax, r.x);
            knmin = max(knmin, r.y);
            if (blockIdx.x == 0) {
                const uint4 q = __ldcg(a.blk_sums + i);
                tot.kept += q.x;
                tot.touched += q.y;
                tot.vis += q.z;
            }
        }
        if (blockIdx.x == 0) {
            // the instance totals do not depend on the order: block 0 reports them to the host right away, at the
            // START of the binning, and leaves them in the control block for the later stages
            tot = block_sum3(tot, s_w);
            if (tid == 0) {
                a.ctl->kept = tot.kept;
                a.ctl->touched = tot.touched;
                a.ctl->visible = (uint32_t)tot.vis;
                if (a.slot) {
                    volatile HostSlot* hs = a.slot;
                    hs->kept = tot.kept;
                    hs->touched = tot.touched;
                    hs->visible = (uint32_t)tot.vis;
                    __threadfence_system();
                    hs->ticket = a.ticket;
                }
            }
        }
        kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
        knmin = __reduce_max_sync(0xFFFFFFFFu, knmin);
        if (lane == 0) {
            sm.base[warp] = kmax;
            sm.cnt[warp] = knmin;
        }
        __syncthreads();
        key_max = __reduce_max_sync(0xFFFFFFFFu, sm.base[lane]);
        key_nmin = __reduce_max_sync(0xFFFFFFFFu, sm.cnt[lane]);
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) {   // the frame's key range, for the per-supertile depth sort
            a.ctl->key_max = key_max;
            a.ctl->key_nmin = key_nmin;
        }
        prof_mark(a.prof, pslot);
    }
    const uint32_t key_min = ~key_nmin;
    // normalised key: visible -> raw - min in [0, span) ; culled -> span (sorted behind everything, stable)
    const uint32_t span = (key_nmin != 0u && key_max >= key_min) ? key_max - key_min + 1u : 0u;
    const uint32_t nbits = span ? 32u - (uint32_t)__clz(span) : 0u;
    const uint32_t npass = a.sort ? (nbits + 8u) / 9u : 0u;
    const uint32_t dbits = npass ? (nbits + npass - 1u) / npass : 0u;
    const uint32_t nd = 1u << dbits;

    if (npass == 0u && a.sort) {   // nothing visible: identity order
        for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
            const uint32_t lo = v * SL, hi = min(P, lo + SL);
            for (uint32_t i = lo + tid; i < hi; i += SGS_SORT_THREADS) a.vals[0][i] = i;
        }
        grid_barrier(bar, bar_target);
    }

    for (uint32_t p = 1; p <= npass; p++) {
        const uint32_t shift = (p - 1u) * dbits;
        const uint32_t out = (npass - p) & 1u;            // the last pass lands in side 0
        const uint32_t* in_key = a.keys[out ^ 1u];
        const uint32_t* in_val = a.vals[out ^ 1u];
        uint32_t* out_key = (p == npass) ? nullptr : a.keys[out];   // the keys are dead after the last pass
        uint32_t* out_val = a.vals[out];

        auto load = [&](uint32_t lo, uint32_t n) {
            if (p == 1u) {
                for (uint32_t i = tid; i < n; i += SGS_SORT_THREADS) {
                    const uint32_t k = a.raw[lo + i];
                    sm.key[i] = (k == 0xFFFFFFFFu) ? span : k - key_min;
                    sm.val[i] = lo + i;
                }
            } else {
                for (uint32_t i = tid; i < n; i += SGS_SORT_THREADS) {
                    sm.key[i] = __ldcg(in_key + lo + i);
                    sm.val[i] = __ldcg(in_val + lo + i);
                }
            }
            __syncthreads();
        };

        for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
            const uint32_t lo = v * SL, hi = min(P, lo + SL), n = hi > lo ? hi - lo : 0u;
            load(lo, n);
            rank_slice(sm, n, shift, dbits);
            publish_hist(sm.cnt, a.hist, v, nd);
        }
        prof_mark(a.prof, pslot);
        grid_barrier(bar, bar_target);
        column_scans(a.hist, VB, nd);
        grid_barrier(bar, bar_target);
        prof_mark(a.prof, pslot);
        for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
            const uint32_t lo = v * SL, hi = min(P, lo + SL), n = hi > lo ? hi - lo : 0u;
            if (!resident) {
                load(lo, n);
                rank_slice(sm, n, shift, dbits);
            }
            slice_bases(sm, a.hist, v, VB, nd);
            scatter_slice(sm, n, shift, nd, out_key, out_val);
            __syncthreads();
        }
        prof_mark(a.prof, pslot);
        grid_barrier(bar, bar_target);
        prof_mark(a.prof, pslot);
    }

    // ---- scan, in depth order (sort = 0: in index order), of the supertile counts (+ the kept rects in depth order
    // for the next stage)
    const uint32_t* order = a.vals[0];
    const bool sorted = a.sort != 0;
    // one slice: supertile count of every item -> sm.key, its rect -> rect_sorted; two items per thread in flight
    auto gather = [&](uint32_t lo, uint32_t n) -> uint32_t {
        uint32_t sum = 0;
        for (uint32_t i = tid; i < n; i += 2 * SGS_SORT_THREADS) {
            const uint32_t i1 = i + SGS_SORT_THREADS;
            const uint32_t g0 = sorted ? __ldcg(order + lo + i) : lo + i;
            const uint32_t g1 = i1 < n ? (sorted ? __ldcg(order + lo + i1) : lo + i1) : g0;
            const ushort4 r0 = a.rect_kept[g0];
            const ushort4 r1 = a.rect_kept[g1];
            const uint4 s0 = super_rect(r0), s1 = super_rect(r1);
            const uint32_t c0 = (s0.y - s0.x) * (s0.w - s0.z), c1 = (s1.y - s1.x) * (s1.w - s1.z);
            sm.key[i] = c0;
            if (sorted) a.rect_sorted[lo + i] = r0;
            sum += c0;
            if (i1 < n) {
                sm.key[i1] = c1;
                if (sorted) a.rect_sorted[lo + i1] = r1;
                sum += c1;
            }
        }
        return sum;
    };
    auto block_total = [&](uint32_t v) -> uint32_t {
        v = __reduce_add_sync(0xFFFFFFFFu, v);
        __syncthreads();
        if (lane == 0) sm.wsum[warp] = v;
        __syncthreads();
        return __reduce_add_sync(0xFFFFFFFFu, sm.wsum[lane]);
    };
    for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
        const uint32_t lo = v * SL, hi = min(P, lo + SL), n = hi > lo ? hi - lo : 0u;
        const uint32_t tot = block_total(gather(lo, n));
        if (tid == 0) __stcg(a.blocksum + v, tot);
    }
    prof_mark(a.prof, pslot);
    grid_barrier(bar, bar_target);
    prof_mark(a.prof, pslot);
    for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
        const uint32_t lo = v * SL, hi = min(P, lo + SL), n = hi > lo ? hi - lo : 0u;
        uint32_t pre = 0;     // slices before this one
        for (uint32_t vv = tid; vv < v; vv += SGS_SORT_THREADS) pre += __ldcg(a.blocksum + vv);
        pre = block_total(pre);
        if (!resident) (void)gather(lo, n);
        __syncthreads();
        // block-wide inclusive scan of sm.key[0..n): contiguous segment per thread
        const uint32_t ipt = (n + SGS_SORT_THREADS - 1) / SGS_SORT_THREADS;
        const uint32_t b = min(n, tid * ipt), e = min(n, b + ipt);
        uint32_t mine = 0;
        for (uint32_t i = b; i < e; i++) mine += sm.key[i];
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        __syncthreads();
        if (lane == 31) sm.wsum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (uint32_t w = 0; w < warp; w++) woff += sm.wsum[w];
        uint32_t run = pre + woff + inc - mine;
        for (uint32_t i = b; i < e; i++) {
            run += sm.key[i];
            sm.val[i] = run;
        }
        __syncthreads();
        for (uint32_t i = tid; i < n; i += SGS_SORT_THREADS) a.coffs[lo + i] = sm.val[i];
        if (v == VB - 1u && tid == 0) a.ctl->coarse = pre + __ldcg(a.blocksum + v);
        __syncthreads();
    }
    prof_mark(a.prof, pslot);
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: (supertile, Gaussian) instance generation + stable bucketing by supertile
// ------------------------------------------------------------------------------------------------
struct CoarseArgs {
    int P;
    int super_x;      // supertiles per row
    int n_super;
    int super_bits;
    int keep;
    unsigned long long cap;
    const uint32_t* order;      // Gaussians in depth order, or NULL: index order (the supertiles sort themselves)
    const uint32_t* coffs;      // inclusive scan of the supertile counts, in that order
    const ushort4* rect_sorted; // kept rects in that order
    uint32_t* keys[2];          // key = supertile id | (mask of the supertile's 16 tiles the rect covers) << 16
    uint32_t* vals[2];
    unsigned long long* pairs;  // final list: Gaussian index | key << 32, supertile-major, depth order inside
    uint32_t* hist;
    uint2* cranges;             // [n_super] bucket of every supertile in the coarse list
    uint32_t* header;
    BinCtl* ctl;
    unsigned long long* prof;
};

// Fill sm.key / sm.val with the instances [s, e) of the depth-ordered coarse stream: Gaussian k (depth order) owns
// the instances [coffs[k] - cc_k, coffs[k]), row-major over the supertiles its kept rect overlaps.
__device__ __forceinline__ void generate_slice(SortSmem& sm, const CoarseArgs& a, uint32_t s, uint32_t e) {
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t P = (uint32_t)a.P;
    // first Gaussian whose inclusive offset exceeds s: 1024-ary search on the monotone coffs[]
    uint32_t lo = 0, len = P;
    while (len > 1u) {
        const uint32_t step = (len + SGS_SORT_THREADS - 1) / SGS_SORT_THREADS;
        const uint64_t idx = (uint64_t)lo + (uint64_t)(tid + 1u) * step - 1u;
        const bool probe = idx < (uint64_t)lo + len;
        const int le = (probe && __ldcg(a.coffs + idx) <= s) ? 1 : 0;
        const uint32_t c = (uint32_t)__syncthreads_count(le);
        const uint32_t nlo = lo + c * step;
        const uint32_t end = lo + len;
        lo = nlo;
        len = (nlo >= end) ? 0u : min(step, end - nlo);
        if (len == 0u) break;
    }
    for (uint32_t kb = lo;; kb += SGS_SORT_THREADS) {
        const uint32_t k = kb + tid;
        uint32_t n = 0, start = 0, gid = 0;
        unsigned long long rk = 0ull;
        uint4 sr = make_uint4(0u, 0u, 0u, 0u);
        bool beyond = true;    // this Gaussian's instances end at or after e (nothing more to do past it)
        if (k < P) {
            const uint32_t incl = __ldcg(a.coffs + k);
            gid = a.order ? __ldcg(a.order + k) : k;
            rk = __ldcg(reinterpret_cast<const unsigned long long*>(a.rect_sorted) + k);
            sr = super_rect(as_rect(rk));
            n = (sr.y - sr.x) * (sr.w - sr.z);
            start = incl - n;
            beyond = incl >= e;
            if (start >= e) n = 0;
        }
        const uint32_t w = sr.y - sr.x;
        if (n > 0 && n <= SGS_DUP_SMALL) {
            uint32_t j = start;
            for (uint32_t y = sr.z; y < sr.w; y++)
                for (uint32_t x = sr.x; x < sr.y; x++, j++)
                    if (j >= s && j < e) {
                        sm.key[j - s] = (y * (uint32_t)a.super_x + x) | (tile_mask16(as_rect(rk), x * SGS_ST, y * SGS_ST) << 16);
                        sm.val[j - s] = gid;
                    }
        }
        unsigned big = __ballot_sync(0xFFFFFFFFu, n > SGS_DUP_SMALL);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const uint32_t sn = __shfl_sync(0xFFFFFFFFu, n, src);
            const uint32_t sstart = __shfl_sync(0xFFFFFFFFu, start, src);
            const uint32_t sgid = __shfl_sync(0xFFFFFFFFu, gid, src);
            const uint32_t sx0 = __shfl_sync(0xFFFFFFFFu, sr.x, src);
            const uint32_t sy0 = __shfl_sync(0xFFFFFFFFu, sr.z, src);
            const uint32_t sw = __shfl_sync(0xFFFFFFFFu, w, src);
            const unsigned long long srk = __shfl_sync(0xFFFFFFFFu, rk, src);
            const uint32_t j0 = max(sstart, s) - sstart, j1 = min(sstart + sn, e) - sstart;
            for (uint32_t i = j0 + lane; i < j1; i += 32) {
                const uint32_t yy = i / sw, xx = i - yy * sw;
                const uint32_t sx = sx0 + xx, sy = sy0 + yy;
                sm.key[sstart + i - s] = (sy * (uint32_t)a.super_x + sx) | (tile_mask16(as_rect(srk), sx * SGS_ST, sy * SGS_ST) << 16);
                sm.val[sstart + i - s] = sgid;
            }
        }
        if (__syncthreads_or((tid == SGS_SORT_THREADS - 1 && !beyond) ? 1 : 0) == 0) break;
    }
    __syncthreads();
}

__device__ __forceinline__ void coarse_sort_phases(const CoarseArgs& a, SortSmem& sm, uint32_t* bar, uint32_t& bar_target) {
    const uint32_t tid = threadIdx.x;
    int pslot = 64;
    prof_mark(a.prof, pslot);
    const unsigned long long kept64 = __ldcg(&a.ctl->kept);
    const uint32_t Rc = __ldcg(&a.ctl->coarse);
    const uint32_t npass = (uint32_t)(a.super_bits + 8) / 9u;
    const uint32_t side_final = (npass & 1u) ^ 1u;     // pass p writes side (p - 1) & 1
    if (kept64 > a.cap || kept64 == 0ull) {
        // over capacity: the host re-launches with a larger buffer; nothing may be written beyond the header
        // (the supertile buckets stay empty, so the expansion kernels and the render kernel see empty tiles)
        if (blockIdx.x == 0 && tid == 0)
            *reinterpret_cast<uint4*>(a.header) = make_uint4(side_final, 0u, (uint32_t)a.keep, (uint32_t)a.cap);
        return;
    }
    if (blockIdx.x == 0 && tid == 0)
        *reinterpret_cast<uint4*>(a.header) = make_uint4(side_final, (uint32_t)kept64, (uint32_t)a.keep, (uint32_t)a.cap);

    const uint32_t G = gridDim.x;
    const uint32_t per_block = (Rc + G - 1) / G;
    const uint32_t slices_per_block = (per_block + SGS_SORT_CHUNK - 1) / SGS_SORT_CHUNK;
    const uint32_t VB = G * slices_per_block;
    const uint32_t SL = (Rc + VB - 1) / VB;
    const bool resident = slices_per_block == 1u;
    const uint32_t dbits = ((uint32_t)a.super_bits + npass - 1u) / npass;
    const uint32_t nd = 1u << dbits;

    for (uint32_t p = 1; p <= npass; p++) {
        const uint32_t shift = (p - 1u) * dbits;
        const uint32_t out = (p - 1u) & 1u;
        const uint32_t* in_key = a.keys[out ^ 1u];
        const uint32_t* in_val = a.vals[out ^ 1u];
        const bool last = p == npass;       // the last pass writes (Gaussian, key) pairs, 8 bytes per store
        auto load = [&](uint32_t lo, uint32_t n) {
            if (p == 1u) {
                generate_slice(sm, a, lo, lo + n);
            } else {
                for (uint32_t i = tid; i < n; i += SGS_SORT_THREADS) {
                    sm.key[i] = __ldcg(in_key + lo + i);
                    sm.val[i] = __ldcg(in_val + lo + i);
                }
                __syncthreads();
            }
        };
        for (uint32_t v = blockIdx.x; v < VB; v += G) {
            const uint32_t lo = min(Rc, v * SL), hi = min(Rc, lo + SL), n = hi - lo;
            if (n) load(lo, n);
            prof_mark(a.prof, pslot);
            rank_slice(sm, n, shift, dbits);
            publish_hist(sm.cnt, a.hist, v, nd);
        }
        prof_mark(a.prof, pslot);
        grid_barrier(bar, bar_target);
        column_scans(a.hist, VB, nd);
        grid_barrier(bar, bar_target);
        prof_mark(a.prof, pslot);
        for (uint32_t v = blockIdx.x; v < VB; v += G) {
            const uint32_t lo = min(Rc, v * SL), hi = min(Rc, lo + SL), n = hi - lo;
            if (!resident) {
                if (n) load(lo, n);
                rank_slice(sm, n, shift, dbits);
            }
            slice_bases(sm, a.hist, v, VB, nd);
            if (npass == 1u && v == 0u) {
                for (uint32_t d = tid; d < (uint32_t)a.n_super; d += SGS_SORT_THREADS) {
                    const uint32_t st = sm.cnt[d];
                    const uint32_t tot = __ldcg(a.hist + (size_t)VB * nd + d);
                    a.cranges[d] = tot ? make_uint2(st, st + tot) : make_uint2(0u, 0u);
                }
            }
            if (last) scatter_slice_pairs(sm, n, shift, nd, a.pairs);
            else scatter_slice(sm, n, shift, nd, a.keys[out], a.vals[out]);
            __syncthreads();
        }
        prof_mark(a.prof, pslot);
        if (p < npass) grid_barrier(bar, bar_target);
    }
    prof_mark(a.prof, pslot);
}

__global__ void __launch_bounds__(SGS_SORT_THREADS, 1) depth_sort_kernel(const DepthArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem& sm = *reinterpret_cast<SortSmem*>(smem_raw);
    __shared__ Tri s_w[32];
    uint32_t bar_target = 0;
    depth_sort_phases(a, sm, s_w, &a.ctl->bar_depth, bar_target);
}

__global__ void __launch_bounds__(SGS_SORT_THREADS, 1) coarse_sort_kernel(const CoarseArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem& sm = *reinterpret_cast<SortSmem*>(smem_raw);
    uint32_t bar_target = 0;
    coarse_sort_phases(a, sm, &a.ctl->bar_tile, bar_target);
}

// Both persistent stages in ONE cooperative launch (the normal path: a launch boundary between two cooperative
// kernels costs ~7 us on B200).  The host ticket is written at the end of the depth stage, i.e. in the middle of this
// kernel; the stand-alone coarse kernel remains for the capacity re-launch and for per-stage profiling.
__global__ void __launch_bounds__(SGS_SORT_THREADS, 1) binning_fused_kernel(const DepthArgs d, const CoarseArgs c) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem& sm = *reinterpret_cast<SortSmem*>(smem_raw);
    __shared__ Tri s_w[32];
    uint32_t bar_target = 0;
    depth_sort_phases(d, sm, s_w, &d.ctl->bar_depth, bar_target);
    grid_barrier(&d.ctl->bar_depth, bar_target);     // totals + coffs of every block visible to every block
    coarse_sort_phases(c, sm, &d.ctl->bar_depth, bar_target);
}

// multi-pass sorts only (more than 512 supertiles): supertile buckets from the sorted pairs, one thread per instance
__global__ void __launch_bounds__(256) coarse_ranges_kernel(const unsigned long long* __restrict__ pairs, const BinCtl* ctl,
                                                            unsigned long long cap, uint2* __restrict__ cranges) {
    if (__ldcg(&ctl->kept) > cap) return;
    const uint32_t Rc = __ldcg(&ctl->coarse);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < Rc; i += gridDim.x * blockDim.x) {
        const uint32_t cur = (uint32_t)(pairs[i] >> 32) & 0xFFFFu;
        const uint32_t prev = i ? ((uint32_t)(pairs[i - 1] >> 32) & 0xFFFFu) : 0xFFFFFFFFu;
        if (cur != prev) {
            cranges[cur].x = i;
            if (i) cranges[prev].y = i;
        }
        if (i == Rc - 1u) cranges[cur].y = Rc;
    }
}

// ------------------------------------------------------------------------------------------------
// Kernels 3a / 3b: expansion of every supertile's depth-ordered list into its 16 per-tile lists
// ------------------------------------------------------------------------------------------------
struct ExpandArgs {
    int tiles_x, tiles_y, super_x, n_tiles;
    const unsigned long long* pairs;   // supertile-major coarse list: Gaussian index | (supertile | tile mask << 16) << 32
    const uint2* cranges;
    uint2* ranges;          // [n_tiles]: the count kernel writes .y = count, its last block turns that into [start, end)
    uint32_t* point_list;
    uint32_t* done;         // last-block ticket (zero before and after the kernel)
    uint32_t* counts;       // [n_tiles] per-tile instance counts (tile_count_kernel<false>; the forward render kernel
                            // later overwrites the array with its packed-record counts)
    int local_scan;         // tile_fill_sorted_kernel derives its tiles' list starts from `counts` itself
    // tile_fill_sorted_kernel only
    const uint32_t* depth_raw;   // [P] float bits of the view-space depth
    const BinCtl* ctl;           // key range of the frame, kept count
    unsigned long long cap;
    uint32_t* scratch_key[2];    // [cap] per-bucket scratch (the coarse sort's ping-pong buffers, free by now): only
    uint32_t* scratch_val[2];    //       buckets longer than SGS_FS_CAP use them
    unsigned long long* prof;    // developer aid (sgs_debug_binning_profile): phase timestamps of blocks 0 and 200
};

// SCAN = true: counts go to ranges[t].y and the last block to finish turns them into [start, end) (any grid size).
// SCAN = false (grids of at most SGS_LOCALSCAN_TILES tiles): counts go to `counts` and every block of
// tile_fill_sorted_kernel derives the starts of its own 16 tiles from them — no serial tail behind a last block
// (14 of the kernel's 18 us at configs[1]).
template <bool SCAN>
__global__ void __launch_bounds__(SGS_EXP_THREADS) tile_count_kernel(const ExpandArgs a) {
    __shared__ uint32_t s_cnt[16];
    __shared__ uint32_t s_last;
    __shared__ uint32_t s_scan[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t s = blockIdx.x;
    const uint32_t tx0 = (s % a.super_x) * SGS_ST, ty0 = (s / a.super_x) * SGS_ST;
    pdl_wait();
    const uint2 cr = a.cranges[s];
    if (tid < 16) s_cnt[tid] = 0;
    __syncthreads();
    uint32_t mine = 0;   // lane t < 16 of every warp accumulates tile t
    (void)tx0;
    unsigned long long nxt = (cr.x + tid < cr.y) ? __ldcg(a.pairs + cr.x + tid) : 0ull;
    for (uint32_t i0 = cr.x + (tid & ~31u); i0 < cr.y; i0 += SGS_EXP_THREADS) {
        const uint32_t i = i0 + lane;
        const uint32_t m = (i < cr.y) ? (uint32_t)(nxt >> 48) : 0u;
        nxt = (i + SGS_EXP_THREADS < cr.y) ? __ldcg(a.pairs + i + SGS_EXP_THREADS) : 0ull;   // next chunk in flight
#pragma unroll
        for (uint32_t t = 0; t < 16; t++) {
            const uint32_t c = __popc(__ballot_sync(0xFFFFFFFFu, (m >> t) & 1u));
            if (lane == t) mine += c;
        }
    }
    if (lane < 16 && mine) atomicAdd(&s_cnt[lane], mine);
    __syncthreads();
    if (tid < 16) {
        const uint32_t tx = tx0 + (tid & 3u), ty = ty0 + (tid >> 2);
        if (tx < (uint32_t)a.tiles_x && ty < (uint32_t)a.tiles_y) {
            if (SCAN) a.ranges[ty * a.tiles_x + tx].y = s_cnt[tid];
            else a.counts[ty * a.tiles_x + tx] = s_cnt[tid];
        }
    }
    if (!SCAN) return;
    // the last block to finish turns the counts into [start, end) ranges (exclusive scan over the tile ids)
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.done, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const uint32_t n = (uint32_t)a.n_tiles;
    const uint32_t ipt = (n + SGS_EXP_THREADS - 1) / SGS_EXP_THREADS;
    const uint32_t b = min(n, tid * ipt), e = min(n, b + ipt);
    uint32_t sum = 0;
    for (uint32_t t = b; t < e; t++) sum += __ldcg(&a.ranges[t].y);
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (uint32_t)o) inc += y;
    }
    if (lane == 31) s_scan[tid >> 5] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t w = 0; w < (tid >> 5); w++) woff += s_scan[w];
    uint32_t run = woff + inc - sum;
    for (uint32_t t = b; t < e; t++) {
        const uint32_t c = __ldcg(&a.ranges[t].y);
        a.ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
        run += c;
    }
    if (tid == 0) *a.done = 0u;
}

__global__ void __launch_bounds__(SGS_EXP_THREADS) tile_fill_kernel(const ExpandArgs a) {
    constexpr int NW = SGS_EXP_THREADS / 32;
    __shared__ uint32_t s_wc[16][NW];     // per chunk: instances of tile t found by warp w
    __shared__ uint32_t s_run[16];        // next free position of every tile list
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s = blockIdx.x;
    const uint32_t tx0 = (s % a.super_x) * SGS_ST, ty0 = (s / a.super_x) * SGS_ST;
    pdl_wait();
    const uint2 cr = a.cranges[s];
    if (cr.y == cr.x) return;
    if (tid < 16) {
        const uint32_t tx = tx0 + (tid & 3u), ty = ty0 + (tid >> 2);
        s_run[tid] = (tx < (uint32_t)a.tiles_x && ty < (uint32_t)a.tiles_y) ? a.ranges[ty * a.tiles_x + tx].x : 0u;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned long long nxt = (cr.x + tid < cr.y) ? __ldcg(a.pairs + cr.x + tid) : 0ull;
    for (uint32_t i0 = cr.x; i0 < cr.y; i0 += SGS_EXP_THREADS) {
        const uint32_t i = i0 + tid;
        const uint32_t gid = (uint32_t)nxt;
        const uint32_t m = (i < cr.y) ? (uint32_t)(nxt >> 48) : 0u;
        nxt = (i + SGS_EXP_THREADS < cr.y) ? __ldcg(a.pairs + i + SGS_EXP_THREADS) : 0ull;       // next chunk in flight
        uint32_t before[16];
#pragma unroll
        for (uint32_t t = 0; t < 16; t++) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (m >> t) & 1u);
            before[t] = __popc(bal & lt_mask);
            if (lane == t) s_wc[t][warp] = __popc(bal);
        }
        __syncthreads();   // also: s_run of the previous chunk is final
        // lane t < 16: this warp's offset in the list of tile t
        uint32_t myoff = 0, total = 0;
        if (lane < 16) {
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const uint32_t c = s_wc[lane][w];
                if ((uint32_t)w < warp) myoff += c;
                total += c;
            }
            myoff += s_run[lane];
        }
#pragma unroll
        for (uint32_t t = 0; t < 16; t++) {
            const uint32_t off = __shfl_sync(0xFFFFFFFFu, myoff, t);
            if ((m >> t) & 1u) a.point_list[off + before[t]] = gid;
        }
        __syncthreads();   // every warp has read s_wc / s_run of this chunk
        if (warp == 0 && lane < 16) s_run[lane] += total;
    }
}


// ------------------------------------------------------------------------------------------------
// Kernel 3b': depth sort INSIDE the supertile + expansion (round 2b).  The coarse list arrives in Gaussian-index
// order inside every supertile bucket (stage A no longer sorts the P Gaussians: three radix passes with six grid-wide
// barriers, 58 us at configs[1]); one block per supertile sorts its bucket — a few thousand entries — by the depth
// bits in shared memory with the same stable LSD passes (9-bit digits over the frame's key range), no grid barrier,
// and then expands it into the 16 tile lists exactly like tile_fill_kernel.  Stable passes over an index-ordered
// bucket give the order (depth bits, Gaussian index): the reference's order restricted to the supertile.
// Buckets longer than SGS_FS_CAP run the same passes chunk by chunk through global scratch (slow but exact).
// ------------------------------------------------------------------------------------------------
#define SGS_FS_CAP 4096
#define SGS_FS_FIX_MAX 96          // longest top-digit bucket the single-pass sort places by counting
#define SGS_LOCALSCAN_TILES 8192   // up to here every fill block reads all tile counts itself (32 KB from L2)
template <int NT>
struct FillSmem {
    uint16_t whist[NT / 32][SGS_SORT_ND];   // per-warp digit counts -> exclusive prefix over the warps
    uint32_t cnt[SGS_SORT_ND];                            // exclusive scan of the digit totals (bucket starts)
    uint32_t key[2][SGS_FS_CAP];
    uint16_t idx[2][SGS_FS_CAP];                          // position in the index-ordered bucket; the side being
                                                          // written doubles as the rank array of the pass
    uint32_t wsum[NT / 32];
    uint32_t wc[16][NT / 32];
    uint32_t run[16];
};

// Stable ranks of key[0..n) on digit (key >> shift) & (nd - 1): rank[i] = position among the keys of the same digit
// within the owning warp's contiguous range; whist[w][d] = keys of digit d in the warps before w.  Returns, in the
// thread d < nd, the total of digit d.
template <int NT>
__device__ __forceinline__ uint32_t fs_rank(FillSmem<NT>& sm, const uint32_t* key, uint16_t* rank, uint32_t n, uint32_t shift,
                                            uint32_t dbits, uint32_t per) {
    constexpr uint32_t NW = NT / 32;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nd = 1u << dbits;
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&sm.whist[0][0]);
        for (uint32_t i = tid; i < NW * SGS_SORT_ND / 2; i += NT) z[i] = 0u;
    }
    __syncthreads();
    const uint32_t beg = warp * per, end = min(n, beg + per);
    uint16_t* wh = sm.whist[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t i0 = beg; i0 < end; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < end;
        const uint32_t d = valid ? ((key[i] >> shift) & (nd - 1u)) : 0u;
        // peers = lanes holding the same digit.  One match.any instead of the nine ballots of the persistent sort
        // kernels: there a warp is alone on its scheduler and the instruction's latency (~200 cycles when the 32
        // digits differ) is exposed; here 48 warps per SM hide it and the kernel is bound by instruction issue.
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? d : 0xFFFFFFFFu) & __ballot_sync(0xFFFFFFFFu, valid);
        const uint32_t before = peers & lt_mask;
        uint32_t old = 0;
        if (valid) old = wh[d];
        __syncwarp();
        if (valid && before == 0u) wh[d] = (uint16_t)(old + __popc(peers));
        __syncwarp();
        if (valid) rank[i] = (uint16_t)(old + __popc(before));
    }
    __syncthreads();
    uint32_t total = 0;
    if (tid < nd) {
#pragma unroll 4
        for (uint32_t w = 0; w < NW; w++) {
            const uint32_t c = sm.whist[w][tid];
            sm.whist[w][tid] = (uint16_t)total;
            total += c;
        }
    }
    return total;
}

// block-wide exclusive scan over the digits: thread d < nd holds `v`, sm.cnt[d] receives the sum of the smaller digits
template <int NT>
__device__ __forceinline__ void fs_digit_scan(FillSmem<NT>& sm, uint32_t v, uint32_t nd) {
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (uint32_t)o) inc += y;
    }
    if (lane == 31) sm.wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t w = 0; w < warp; w++) woff += sm.wsum[w];
    if (tid < nd) sm.cnt[tid] = woff + inc - v;
    __syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT, NT >= 512 ? 3 : 4) tile_fill_sorted_kernel(const ExpandArgs a) {
    constexpr int SGS_FS_IPT = SGS_FS_CAP / NT;
    constexpr uint32_t NW = NT / 32;
    extern __shared__ __align__(16) unsigned char fs_raw[];
    FillSmem<NT>& sm = *reinterpret_cast<FillSmem<NT>*>(fs_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s = blockIdx.x;
    const uint32_t tx0 = (s % a.super_x) * SGS_ST, ty0 = (s / a.super_x) * SGS_ST;
    pdl_wait();
    const uint2 cr = a.cranges[s];
    if (cr.y == cr.x) return;
    if (__ldcg(&a.ctl->kept) > a.cap) return;      // over capacity: the host re-launches
    const uint32_t n = cr.y - cr.x;
    int pslot = blockIdx.x == 0 ? 96 : 112;
    auto mark = [&]() {
        if (a.prof && (blockIdx.x == 0 || blockIdx.x == 200) && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.prof[pslot] = t;
        }
        pslot++;
    };
    mark();
    const unsigned long long* bucket = a.pairs + cr.x;

    // the frame's key range -> passes (every block derives the same numbers)
    const uint32_t key_max = __ldcg(&a.ctl->key_max), key_nmin = __ldcg(&a.ctl->key_nmin);
    const uint32_t key_min = ~key_nmin;
    const uint32_t span = (key_nmin != 0u && key_max >= key_min) ? key_max - key_min + 1u : 0u;
    const uint32_t nbits = span > 1u ? 32u - (uint32_t)__clz(span - 1u) : 0u;   // keys are in [0, span)
    constexpr uint32_t MAXB = NT >= 512 ? 9u : 8u;      // one thread per digit in the prefix phases
    const uint32_t npass = (nbits + MAXB - 1u) / MAXB;
    const uint32_t dbits = npass ? (nbits + npass - 1u) / npass : 0u;
    const uint32_t nd = 1u << dbits;

    if (a.local_scan) {
        // list start of tile t = number of instances in the tiles before t (tile-id order, like the reference's sorted
        // list): one pass over all counts accumulates, per tile row r of this supertile, the sum before the row's
        // first tile here
        const uint32_t nt = (uint32_t)a.n_tiles;
        uint32_t rowfirst[SGS_ST], acc[SGS_ST];
#pragma unroll
        for (uint32_t r = 0; r < SGS_ST; r++) {
            rowfirst[r] = (ty0 + r < (uint32_t)a.tiles_y) ? (ty0 + r) * (uint32_t)a.tiles_x + tx0 : nt;
            acc[r] = 0u;
        }
        for (uint32_t t = tid; t < nt; t += NT) {
            const uint32_t c = __ldcg(a.counts + t);
#pragma unroll
            for (uint32_t r = 0; r < SGS_ST; r++) acc[r] += (t < rowfirst[r]) ? c : 0u;
        }
#pragma unroll
        for (uint32_t r = 0; r < SGS_ST; r++) {
            const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, acc[r]);
            if (lane == 0) sm.wc[r][warp] = v;
        }
        __syncthreads();
        if (tid < 16) {
            const uint32_t r = tid >> 2, x = tid & 3u;
            const uint32_t tx = tx0 + x, ty = ty0 + r;
            uint32_t start = 0;
            if (tx < (uint32_t)a.tiles_x && ty < (uint32_t)a.tiles_y) {
                uint32_t cx[SGS_ST];      // the row's counts up to this tile, all loads in flight together
#pragma unroll
                for (uint32_t xx = 0; xx < SGS_ST; xx++) cx[xx] = xx <= x ? __ldcg(a.counts + rowfirst[r] + xx) : 0u;
                for (uint32_t w = 0; w < NW; w++) start += sm.wc[r][w];
#pragma unroll
                for (uint32_t xx = 0; xx < SGS_ST; xx++) start += xx < x ? cx[xx] : 0u;
                const uint32_t c = cx[x];
                a.ranges[ty * a.tiles_x + tx] = c ? make_uint2(start, start + c) : make_uint2(0u, 0u);
            }
            sm.run[tid] = start;
        }
        __syncthreads();
    } else if (tid < 16) {
        const uint32_t tx = tx0 + (tid & 3u), ty = ty0 + (tid >> 2);
        sm.run[tid] = (tx < (uint32_t)a.tiles_x && ty < (uint32_t)a.tiles_y) ? a.ranges[ty * a.tiles_x + tx].x : 0u;
    }

    mark();
    uint32_t cur = 0;                    // side of key / idx holding the current order (resident path)
    const uint32_t* gorder = nullptr;    // long buckets: the order lives in global scratch
    if (n <= SGS_FS_CAP) {
        for (uint32_t i = tid; i < n; i += NT) {
            const uint32_t gid = (uint32_t)__ldcg(bucket + i);
            sm.key[0][i] = __ldcg(a.depth_raw + gid) - key_min;
            sm.idx[0][i] = (uint16_t)i;
        }
        __syncthreads();
        mark();
        const uint32_t per = (((n + NW - 1) / NW) + 31u) & ~31u;
        // i / per for i < 4096 as a multiplication: m = ceil(2^24 / per), error m per - 2^24 < per <= 256, i * 256 < 2^24
        const uint32_t per_inv = ((1u << 24) + per - 1u) / per;
        const uint32_t rounds = (n + NT - 1) / NT;
        // Fast path for keys that spread over the frame's depth range (the normal case): ONE stable pass on the TOP
        // digit, then every entry finds its exact place inside its digit bucket — a handful of entries — by counting
        // the bucket's entries that precede it in (key, position) order.  8 barrier-separated phases instead of 18.
        // Buckets longer than SGS_FS_FIX_MAX (depths clustered in a sliver of the range) fall through to the LSD
        // passes below; sides 0 of key / idx are untouched until then.
        bool sorted_msd = false;
        if (npass >= 2u) {
            const uint32_t mbits = min(9u, nbits), mshift = nbits - mbits, mnd = 1u << mbits;
            uint16_t* rank = sm.idx[1];
            const uint32_t total = fs_rank(sm, sm.key[0], rank, n, mshift, mbits, per);
            fs_digit_scan(sm, total, mnd);
            if (__syncthreads_or(total > SGS_FS_FIX_MAX ? 1 : 0) == 0) {
                uint32_t dv[SGS_FS_IPT];
#pragma unroll
                for (int u = 0; u < SGS_FS_IPT; u++) {
                    if ((uint32_t)u >= rounds) break;     // block-uniform
                    const uint32_t i = tid + u * NT;
                    if (i < n) {
                        const uint32_t k = sm.key[0][i];
                        const uint32_t d = k >> mshift;
                        const uint32_t dst = sm.cnt[d] + sm.whist[(i * per_inv) >> 24][d] + rank[i];
                        sm.key[1][dst] = k;
                        dv[u] = dst | ((uint32_t)sm.idx[0][i] << 16);
                    }
                }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < SGS_FS_IPT; u++) {
                    if ((uint32_t)u >= rounds) break;
                    const uint32_t i = tid + u * NT;
                    if (i < n) sm.idx[1][dv[u] & 0xFFFFu] = (uint16_t)(dv[u] >> 16);
                }
                __syncthreads();
                // exact placement inside the digit bucket [b0, b1)
#pragma unroll
                for (int u = 0; u < SGS_FS_IPT; u++) {
                    if ((uint32_t)u >= rounds) break;
                    const uint32_t j = tid + u * NT;
                    if (j < n) {
                        const uint32_t k = sm.key[1][j];
                        const uint32_t d = k >> mshift;
                        const uint32_t b0 = sm.cnt[d], b1 = (d + 1u < mnd) ? sm.cnt[d + 1u] : n;
                        uint32_t c = 0;
                        for (uint32_t q = b0; q < b1; q++) {
                            const uint32_t kq = sm.key[1][q];
                            c += (kq < k || (kq == k && q < j)) ? 1u : 0u;
                        }
                        sm.idx[0][b0 + c] = sm.idx[1][j];
                    }
                }
                __syncthreads();
                cur = 0u;
                sorted_msd = true;
                mark();
            }
        }
        for (uint32_t p = 0; p < (sorted_msd ? 0u : npass); p++) {
            const uint32_t shift = p * dbits;
            uint16_t* rank = sm.idx[cur ^ 1u];
            const uint32_t total = fs_rank(sm, sm.key[cur], rank, n, shift, dbits, per);
            fs_digit_scan(sm, total, nd);
            // scatter: the keys move at once; the rank array aliases the destination index array, so the indices
            // wait in registers (destination and index packed, 12 bits each) until every rank has been read
            uint32_t dv[SGS_FS_IPT];
#pragma unroll
            for (int u = 0; u < SGS_FS_IPT; u++) {
                if ((uint32_t)u >= rounds) break;     // block-uniform
                const uint32_t i = tid + u * NT;
                if (i < n) {
                    const uint32_t k = sm.key[cur][i];
                    const uint32_t d = (k >> shift) & (nd - 1u);
                    const uint32_t dst = sm.cnt[d] + sm.whist[(i * per_inv) >> 24][d] + rank[i];
                    sm.key[cur ^ 1u][dst] = k;
                    dv[u] = dst | ((uint32_t)sm.idx[cur][i] << 16);
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < SGS_FS_IPT; u++) {
                if ((uint32_t)u >= rounds) break;
                const uint32_t i = tid + u * NT;
                if (i < n) sm.idx[cur ^ 1u][dv[u] & 0xFFFFu] = (uint16_t)(dv[u] >> 16);
            }
            __syncthreads();
            cur ^= 1u;
            mark();
        }
    } else {
        uint32_t* K[2] = {a.scratch_key[0] + cr.x, a.scratch_key[1] + cr.x};
        uint32_t* V[2] = {a.scratch_val[0] + cr.x, a.scratch_val[1] + cr.x};
        for (uint32_t i = tid; i < n; i += NT) {
            const uint32_t gid = (uint32_t)__ldcg(bucket + i);
            K[0][i] = __ldcg(a.depth_raw + gid) - key_min;
            V[0][i] = i;
        }
        __syncthreads();
        uint32_t side = 0;
        for (uint32_t p = 0; p < npass; p++) {
            const uint32_t shift = p * dbits;
            // digit totals of the whole bucket -> bucket starts
            for (uint32_t d = tid; d < SGS_SORT_ND; d += NT) sm.cnt[d] = 0u;
            __syncthreads();
            for (uint32_t i = tid; i < n; i += NT) atomicAdd(&sm.cnt[(K[side][i] >> shift) & (nd - 1u)], 1u);
            __syncthreads();
            const uint32_t tot_all = tid < nd ? sm.cnt[tid] : 0u;
            __syncthreads();
            fs_digit_scan(sm, tot_all, nd);
            for (uint32_t c0 = 0; c0 < n; c0 += SGS_FS_CAP) {
                const uint32_t m = min((uint32_t)SGS_FS_CAP, n - c0);
                for (uint32_t i = tid; i < m; i += NT) sm.key[0][i] = K[side][c0 + i];
                __syncthreads();
                const uint32_t per = (((m + NW - 1) / NW) + 31u) & ~31u;
                const uint32_t total = fs_rank(sm, sm.key[0], sm.idx[0], m, shift, dbits, per);
                __syncthreads();
                for (uint32_t i = tid; i < m; i += NT) {
                    const uint32_t kk = sm.key[0][i];
                    const uint32_t d = (kk >> shift) & (nd - 1u);
                    const uint32_t dst = sm.cnt[d] + sm.whist[i / per][d] + sm.idx[0][i];
                    K[side ^ 1u][dst] = kk;
                    V[side ^ 1u][dst] = V[side][c0 + i];
                }
                __syncthreads();
                if (tid < nd) sm.cnt[tid] += total;     // the next chunk continues behind this one
                __syncthreads();
            }
            side ^= 1u;
        }
        gorder = V[side];
    }

    // ---- expansion of the depth-ordered bucket into the 16 tile lists (as tile_fill_kernel)
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t i0 = 0; i0 < n; i0 += NT) {
        const uint32_t i = i0 + tid;
        unsigned long long pr = 0ull;
        if (i < n) pr = __ldcg(bucket + (gorder ? gorder[i] : (uint32_t)sm.idx[cur][i]));
        const uint32_t gid = (uint32_t)pr;
        const uint32_t m = (uint32_t)(pr >> 48);
        uint32_t before[16];
#pragma unroll
        for (uint32_t t = 0; t < 16; t++) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (m >> t) & 1u);
            before[t] = __popc(bal & lt_mask);
            if (lane == t) sm.wc[t][warp] = __popc(bal);
        }
        __syncthreads();   // also: run[] of the previous chunk is final
        uint32_t myoff = 0, total = 0;
        if (lane < 16) {
#pragma unroll
            for (uint32_t w = 0; w < NW; w++) {
                const uint32_t c = sm.wc[lane][w];
                if (w < warp) myoff += c;
                total += c;
            }
            myoff += sm.run[lane];
        }
#pragma unroll
        for (uint32_t t = 0; t < 16; t++) {
            const uint32_t off = __shfl_sync(0xFFFFFFFFu, myoff, t);
            if ((m >> t) & 1u) a.point_list[off + before[t]] = gid;
        }
        __syncthreads();   // every warp has read wc / run of this chunk
        if (warp == 0 && lane < 16) sm.run[lane] += total;
    }
    mark();
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static int bits_for(int n) {
    int b = 1;
    while ((1 << b) < n) b++;
    return b;
}
int binning_supertiles(int tiles_x, int tiles_y, int* super_x) {
    const int sx = (tiles_x + SGS_ST - 1) / SGS_ST, sy = (tiles_y + SGS_ST - 1) / SGS_ST;
    if (super_x) *super_x = sx;
    return sx * sy;
}
static int coarse_passes(int n_super) { return (bits_for(n_super) + 8) / 9; }
int binning_coarse_list_side(int n_super) { return (coarse_passes(n_super) & 1) ^ 1; }

static unsigned long long* g_prof_host = nullptr;   // 128 timestamps, pinned + mapped; allocated by binning_profile()
unsigned long long* binning_profile(bool enable) {
    if (enable && !g_prof_host) {
        void* h = nullptr;
        if (cudaHostAlloc(&h, 128 * sizeof(unsigned long long), cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
            memset(h, 0, 128 * sizeof(unsigned long long));
            g_prof_host = reinterpret_cast<unsigned long long*>(h);
        }
    }
    return g_prof_host;
}
static bool g_prof_on = false;
void binning_profile_enable(bool on) {
    g_prof_on = on;
    if (on) binning_profile(true);
}

int binning_grid_blocks() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(depth_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
        cudaFuncSetAttribute(coarse_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
        cudaFuncSetAttribute(binning_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
        cudaFuncSetAttribute(tile_fill_sorted_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FillSmem<512>));
        cudaFuncSetAttribute(tile_fill_sorted_kernel<512>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);   // 3 blocks per SM
        cached = sms;
        cached_dev = dev;
    }
    return cached;
}

static int vblocks_for(size_t n) {
    const size_t G = (size_t)binning_grid_blocks();
    if (G == 0) return 0;
    const size_t per_block = (n + G - 1) / G;
    size_t slices = (per_block + SGS_SORT_CHUNK - 1) / SGS_SORT_CHUNK;
    if (slices == 0) slices = 1;
    return (int)(G * slices);
}
int binning_depth_vblocks(int P) { return vblocks_for((size_t)P); }
size_t binning_hist_words(size_t n) { return ((size_t)vblocks_for(n) + 1) * SGS_SORT_ND; }

// Where the depth sort happens.  1: inside every supertile (tile_fill_sorted_kernel; fastest while a bucket fits a
// block's shared memory); 0: one global sort of the P Gaussians first (round-2a design; scales to dense scenes).
// Chosen per forward call by sgs_api.cu (binning_set_mode) from the previous frame's counts; both are exact.
static thread_local int g_binning_mode = 1;
void binning_set_mode(int mode) { g_binning_mode = mode ? 1 : 0; }
static int binning_mode() { return g_binning_mode; }
int binning_bucket_capacity() { return SGS_FS_CAP; }

static DepthArgs make_depth_args(int P, const GeomState& g, HostSlot* slot, unsigned long long ticket) {
    DepthArgs a;
    a.P = P;
    a.sort = binning_mode() == 0 ? 1 : 0;
    a.vblocks = g.depth_vblocks;
    a.slice = (P + a.vblocks - 1) / a.vblocks;
    a.raw = g.depth_raw;
    a.blk_range = g.blk_range;
    a.blk_sums = g.blk_sums;
    a.n_blk_range = g.n_blk_range;
    a.rect_sorted = g.rect_sorted;
    a.keys[0] = g.depth_keys[0];
    a.keys[1] = g.depth_keys[1];
    a.vals[0] = g.depth_vals[0];
    a.vals[1] = g.depth_vals[1];
    a.rect_kept = g.rect_kept;
    a.tiles_touched = g.tiles_touched;
    a.coffs = g.coffs;
    a.hist = g.hist;
    a.blocksum = g.blocksum;
    a.ctl = g.ctl;
    a.slot = slot;
    a.ticket = ticket;
    a.prof = g_prof_on ? g_prof_host : nullptr;
    return a;
}

static CoarseArgs make_coarse_args(int P, const ViewParams& vp, const GeomState& g, const BinningState& b,
                                   const ImageState& img, int keep) {
    CoarseArgs a;
    a.P = P;
    a.n_super = binning_supertiles(vp.tiles_x, vp.tiles_y, &a.super_x);
    a.super_bits = bits_for(a.n_super);
    a.keep = keep;
    a.cap = (unsigned long long)b.cap;
    a.order = binning_mode() == 0 ? g.depth_vals[0] : nullptr;
    a.coffs = g.coffs;
    a.rect_sorted = binning_mode() == 0 ? g.rect_sorted : g.rect_kept;
    a.pairs = b.coarse_pairs;
    a.keys[0] = b.coarse_keys[0];
    a.keys[1] = b.coarse_keys[1];
    a.vals[0] = b.coarse_vals[0];
    a.vals[1] = b.coarse_vals[1];
    a.hist = b.hist;
    a.cranges = img.cranges;
    a.header = b.header;
    a.ctl = g.ctl;
    a.prof = g_prof_on ? g_prof_host : nullptr;
    return a;
}

// supertile buckets (multi-pass sorts) + expansion into the per-tile lists and ranges
static cudaError_t launch_expand(const CoarseArgs& a, const ViewParams& vp, const GeomState& g, const BinningState& b,
                                 const ImageState& img, cudaStream_t s) {
    const int G = binning_grid_blocks();
    if (coarse_passes(a.n_super) > 1)
        coarse_ranges_kernel<<<4 * G, 256, 0, s>>>(b.coarse_pairs, g.ctl, a.cap, img.cranges);
    ExpandArgs x;
    x.tiles_x = vp.tiles_x;
    x.tiles_y = vp.tiles_y;
    x.super_x = a.super_x;
    x.n_tiles = vp.tiles_x * vp.tiles_y;
    x.pairs = b.coarse_pairs;
    x.cranges = img.cranges;
    x.ranges = img.ranges;
    x.point_list = b.point_list;
    x.done = &g.ctl->expand_done;
    x.depth_raw = g.depth_raw;
    x.ctl = g.ctl;
    x.cap = a.cap;
    x.scratch_key[0] = b.coarse_keys[0];
    x.scratch_key[1] = b.coarse_keys[1];
    x.scratch_val[0] = b.coarse_vals[0];
    x.scratch_val[1] = b.coarse_vals[1];
    x.counts = img.tile_count;
    x.prof = g_prof_on ? g_prof_host : nullptr;
    x.local_scan = (binning_mode() != 0 && x.n_tiles <= SGS_LOCALSCAN_TILES) ? 1 : 0;
    const dim3 grid(a.n_super);
    cudaError_t e;
    if (x.local_scan) e = launch_pdl(tile_count_kernel<false>, grid, dim3(SGS_EXP_THREADS), 0, s, x);
    else e = launch_pdl(tile_count_kernel<true>, grid, dim3(SGS_EXP_THREADS), 0, s, x);
    if (e != cudaSuccess) return e;
    if (binning_mode() == 0) e = launch_pdl(tile_fill_kernel, grid, dim3(SGS_EXP_THREADS), 0, s, x);
    else e = launch_pdl(tile_fill_sorted_kernel<512>, grid, dim3(512), sizeof(FillSmem<512>), s, x);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_depth_sort(int P, GeomState g, HostSlot* slot, unsigned long long ticket, cudaStream_t s) {
    const int G = binning_grid_blocks();
    if (G <= 0) return cudaErrorInvalidDevice;
    DepthArgs a = make_depth_args(P, g, slot, ticket);
    void* args[] = {&a};
    return cudaLaunchCooperativeKernel((const void*)depth_sort_kernel, dim3(G), dim3(SGS_SORT_THREADS), args,
                                       sizeof(SortSmem), s);
}

cudaError_t launch_tile_binning(int P, const ViewParams& vp, GeomState g, BinningState b, ImageState img, int keep,
                                cudaStream_t s) {
    const int G = binning_grid_blocks();
    if (G <= 0) return cudaErrorInvalidDevice;
    CoarseArgs a = make_coarse_args(P, vp, g, b, img, keep);
    void* args[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)coarse_sort_kernel, dim3(G), dim3(SGS_SORT_THREADS), args,
                                                sizeof(SortSmem), s);
    if (e != cudaSuccess) return e;
    return launch_expand(a, vp, g, b, img, s);
}

// depth sort + supertile bucketing in one cooperative launch, then the expansion kernels
cudaError_t launch_binning_fused(int P, const ViewParams& vp, GeomState g, BinningState b, ImageState img, int keep,
                                 HostSlot* slot, unsigned long long ticket, cudaStream_t s) {
    const int G = binning_grid_blocks();
    if (G <= 0) return cudaErrorInvalidDevice;
    DepthArgs d = make_depth_args(P, g, slot, ticket);
    CoarseArgs c = make_coarse_args(P, vp, g, b, img, keep);
    void* args[] = {&d, &c};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)binning_fused_kernel, dim3(G), dim3(SGS_SORT_THREADS), args,
                                                sizeof(SortSmem), s);
    if (e != cudaSuccess) return e;
    return launch_expand(c, vp, g, b, img, s);
}

}  // namespace sgs

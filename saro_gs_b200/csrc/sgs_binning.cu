// Tile binning: produces, for every 16x16 tile, the list of Gaussian instances that touch
// it, ordered by (depth bits, Gaussian index) — the exact order the reference obtains
// from one stable 64-bit radix sort of  key = tile<<32 | float_bits(depth)
//   $R/cuda_rasterizer/rasterizer_impl.cu:70-111 (duplicateWithKeys)
//   $R/cuda_rasterizer/rasterizer_impl.cu:299-309 (SortPairs on bits [0, 32+bit))
//   $R/cuda_rasterizer/rasterizer_impl.cu:116-138 (identifyTileRanges)
//
// B200 design: a TWO-LEVEL sort over a CULLED instance set, done by TWO persistent kernels.
//   0. (in the preprocess kernel) the reference's 3-sigma tile rect of every Gaussian is clipped to
//      the exact axis-aligned bounding box of its  alpha >= 1/255  ellipse (`rect_kept`).  The
//      reference's own count (`tiles_touched`) is kept for the API-visible num_rendered; a dropped
//      instance would `continue` on all 256 pixels of its tile in the reference, so image, depth and
//      gradients are unchanged bit for bit.  (SGS_FLAG_NO_TILE_CULL keeps the full rect: then ranges /
//      point_list / n_contrib equal the reference's bit for bit.)
//   1. depth_sort_kernel: stable LSD radix sort of the P Gaussians (not the R >> P instances) by their depth
//      bits, normalised to the frame's [min, max] key range (26 bits = 3 passes of 9 at config 2 instead of 4 x 8),
//      then the scan of area(rect_kept) in depth order.  The totals go to the host through a pinned slot.
//   2. tile_sort_kernel: every block GENERATES its slice of the depth-ordered (tile, Gaussian) instance stream
//      straight into shared memory (the unsorted stream never exists in HBM), then a stable LSD radix sort on
//      ceil(log2 #tiles) bits (13 bits = 7 + 6 at 1352x1014), then the per-tile ranges.
//   A stable sort by tile of a depth-ordered stream is ordered by (tile, depth, index): the reference order
//   restricted to the kept instances.
//
// Why not CUB (round 1): both sorts are tiny (2.4 MB and 17 MB per pass) and CUB's onesweep spends its time in
// launch gaps and in 34-CTA decoupled look-back chains (10 launches, 170 us at config 2).  Here one pass is:
// every block ranks its slice in shared memory (warp-private histograms + match.any), publishes its digit
// histogram, ONE grid-wide barrier, every block derives its global offsets from the histogram matrix and scatters.
// The kernels are launched cooperatively with one 1024-thread block per SM; all counts (P-dependent pass count,
// number of instances) are read on the device, so the host never waits between the stages.
#include "sgs_common.cuh"
#include <cstring>

namespace sgs {

#define SGS_SORT_THREADS 1024
#define SGS_SORT_WARPS 32
#define SGS_DEPTH_ND 512       // max digits per pass, depth sort (9 bits)
#define SGS_TILE_ND 256        // max digits per pass, tile sort (8 bits)
#define SGS_DEPTH_CHUNK 16384  // keys a block keeps resident in shared memory (depth sort)
#define SGS_TILE_CHUNK 20480   // instances a block keeps resident in shared memory (tile sort)
#define SGS_DUP_SMALL 8

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
template <int NDMAX, int CHUNK>
struct SortSmem {
    uint16_t whist[SGS_SORT_WARPS][NDMAX + 2];   // per-warp digit counts -> exclusive prefix over warps (+ dump bin)
    uint32_t cnt[NDMAX];                          // digit counts of the slice
    uint32_t base[NDMAX];                         // global offset of the slice's first element of every digit
    uint32_t red[2][SGS_SORT_THREADS];            // reduction scratch
    uint32_t wsum[32];
    uint32_t key[CHUNK];
    uint32_t val[CHUNK];
    uint16_t rank[CHUNK];
};

__device__ __forceinline__ uint32_t keys_per_warp(uint32_t n) {
    return (((n + SGS_SORT_WARPS - 1) / SGS_SORT_WARPS) + 31u) & ~31u;
}

// Stable ranks of sm.key[0..n) on digit (key >> shift) & (nd - 1).  Warp w owns the contiguous keys
// [w * per, (w + 1) * per); on return  position within the slice's digit group = whist[w][d] + rank[i], and
// sm.cnt[d] holds the slice's digit histogram.
template <int NDMAX, int CHUNK>
__device__ __forceinline__ void rank_slice(SortSmem<NDMAX, CHUNK>& sm, uint32_t n, uint32_t shift, uint32_t nd) {
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&sm.whist[0][0]);
        constexpr uint32_t words = SGS_SORT_WARPS * (NDMAX + 2) / 2;
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
        const uint32_t d = valid ? ((sm.key[i] >> shift) & (nd - 1u)) : nd;   // nd = dump bin of the idle lanes
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t before = peers & lt_mask;
        const uint32_t old = wh[d];
        __syncwarp();
        if (before == 0u) wh[d] = (uint16_t)(old + __popc(peers));
        __syncwarp();
        if (valid) sm.rank[i] = (uint16_t)(old + __popc(before));
    }
    __syncthreads();
    for (uint32_t d = tid; d < nd; d += SGS_SORT_THREADS) {
        uint32_t run = 0;
#pragma unroll 8
        for (int w = 0; w < SGS_SORT_WARPS; w++) {
            const uint32_t c = sm.whist[w][d];
            sm.whist[w][d] = (uint16_t)run;
            run += c;
        }
        sm.cnt[d] = run;
    }
    __syncthreads();
}

// sm.base[d] = (number of keys with a smaller digit anywhere) + (keys with digit d in slices before `v`),
// from the published histogram matrix hist[vblocks][nd].
template <int NDMAX, int CHUNK>
__device__ __forceinline__ void slice_bases(SortSmem<NDMAX, CHUNK>& sm, const uint32_t* __restrict__ hist, uint32_t v,
                                            uint32_t vblocks, uint32_t nd) {
    const uint32_t tid = threadIdx.x;
    const uint32_t parts = SGS_SORT_THREADS / nd;    // nd is a power of two <= 512
    const uint32_t d = tid & (nd - 1u), part = tid / nd;
    uint32_t tot = 0, below = 0;
#pragma unroll 4
    for (uint32_t vv = part; vv < vblocks; vv += parts) {
        const uint32_t x = __ldcg(hist + (size_t)vv * nd + d);
        tot += x;
        if (vv < v) below += x;
    }
    sm.red[0][tid] = tot;
    sm.red[1][tid] = below;
    __syncthreads();
    // digit totals + exclusive scan over the digits (threads 0..511 take part in the shuffles, idle ones add 0)
    uint32_t T = 0, Bl = 0;
    if (tid < nd) {
        for (uint32_t p = 0; p < parts; p++) {
            T += sm.red[0][p * nd + tid];
            Bl += sm.red[1][p * nd + tid];
        }
    }
    if (tid < 512) {
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
        if (tid < nd) sm.base[tid] = woff + inc - T + Bl;
    }
    __syncthreads();
}

template <int NDMAX, int CHUNK>
__device__ __forceinline__ void scatter_slice(SortSmem<NDMAX, CHUNK>& sm, uint32_t n, uint32_t shift, uint32_t nd,
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

__device__ __forceinline__ void publish_hist(const uint32_t* cnt, uint32_t* __restrict__ hist, uint32_t v, uint32_t nd) {
    for (uint32_t d = threadIdx.x; d < nd; d += SGS_SORT_THREADS) __stcg(hist + (size_t)v * nd + d, cnt[d]);
}

// ------------------------------------------------------------------------------------------------
// Kernel 1: depth sort of the Gaussians + scan of the kept tile counts in depth order
// ------------------------------------------------------------------------------------------------
struct DepthArgs {
    int P;
    int vblocks;      // virtual blocks (slices); multiple of gridDim.x; == gridDim.x  <=>  slices stay resident
    int slice;        // keys per slice
    const uint32_t* raw;
    uint32_t* keys[2];
    uint32_t* vals[2];
    const ushort4* rect_kept;
    const uint32_t* tiles_touched;
    uint32_t* offs;
    uint32_t* hist;
    unsigned long long* blocksum;
    BinCtl* ctl;
    HostSlot* slot;
    unsigned long long ticket;
    unsigned long long* prof;
};

using DepthSmem = SortSmem<SGS_DEPTH_ND, SGS_DEPTH_CHUNK>;

struct Tri {
    unsigned long long kept, touched;
    uint32_t vis;
};
__device__ __forceinline__ Tri tri_add(const Tri& a, const Tri& b) { return {a.kept + b.kept, a.touched + b.touched, a.vis + b.vis}; }
__device__ __forceinline__ Tri tri_shfl_up(const Tri& a, int o) {
    Tri r;
    r.kept = __shfl_up_sync(0xFFFFFFFFu, a.kept, o);
    r.touched = __shfl_up_sync(0xFFFFFFFFu, a.touched, o);
    r.vis = __shfl_up_sync(0xFFFFFFFFu, a.vis, o);
    return r;
}

__global__ void __launch_bounds__(SGS_SORT_THREADS, 1) depth_sort_kernel(const DepthArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DepthSmem& sm = *reinterpret_cast<DepthSmem*>(smem_raw);
    __shared__ Tri s_w[32], s_p[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t P = (uint32_t)a.P, VB = (uint32_t)a.vblocks, SL = (uint32_t)a.slice;
    const bool resident = VB == gridDim.x;
    uint32_t bar_target = 0;
    int pslot = 0;
    prof_mark(a.prof, pslot);

    // ---- phase 0: key range of the visible Gaussians (resident mode: the raw keys stay in shared memory)
    {
        uint32_t kmax = 0u, knmin = 0u;
        for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
            const uint32_t lo = v * SL, hi = min(P, lo + SL);
            for (uint32_t i = lo + tid; i < hi; i += SGS_SORT_THREADS) {
                const uint32_t k = a.raw[i];
                if (resident) sm.key[i - lo] = k;
                if (k != 0xFFFFFFFFu) {
                    kmax = max(kmax, k);
                    knmin = max(knmin, ~k);
                }
            }
        }
        kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
        knmin = __reduce_max_sync(0xFFFFFFFFu, knmin);
        if (lane == 0) {
            sm.red[0][warp] = kmax;
            sm.red[1][warp] = knmin;
        }
        __syncthreads();
        if (warp == 0) {
            kmax = __reduce_max_sync(0xFFFFFFFFu, sm.red[0][lane]);
            knmin = __reduce_max_sync(0xFFFFFFFFu, sm.red[1][lane]);
            if (lane == 0) {
                if (kmax) atomicMax(&a.ctl->key_max, kmax);
                if (knmin) atomicMax(&a.ctl->key_nmin, knmin);
            }
        }
        { prof_mark(a.prof, pslot); grid_barrier(&a.ctl->bar_depth, bar_target); prof_mark(a.prof, pslot); }
    }
    const uint32_t key_max = __ldcg(&a.ctl->key_max), key_nmin = __ldcg(&a.ctl->key_nmin);
    const uint32_t key_min = ~key_nmin;
    // normalised key: visible -> raw - min in [0, span) ; culled -> span (sorted behind everything, stable)
    const uint32_t span = (key_nmin != 0u && key_max >= key_min) ? key_max - key_min + 1u : 0u;
    const uint32_t nbits = span ? 32u - (uint32_t)__clz(span) : 0u;
    const uint32_t npass = (nbits + 8u) / 9u;
    const uint32_t dbits = npass ? (nbits + npass - 1u) / npass : 0u;
    const uint32_t nd = 1u << dbits;

    if (npass == 0u) {   // nothing visible: identity order
        for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
            const uint32_t lo = v * SL, hi = min(P, lo + SL);
            for (uint32_t i = lo + tid; i < hi; i += SGS_SORT_THREADS) a.vals[0][i] = i;
        }
    }

    for (uint32_t p = 1; p <= npass; p++) {
        const uint32_t shift = (p - 1u) * dbits;
        const uint32_t out = (npass - p) & 1u;            // the last pass lands in side 0
        const uint32_t* in_key = a.keys[out ^ 1u];
        const uint32_t* in_val = a.vals[out ^ 1u];
        uint32_t* out_key = (p == npass) ? nullptr : a.keys[out];   // the keys are dead after the last pass
        uint32_t* out_val = a.vals[out];

        auto load = [&](uint32_t v, uint32_t lo, uint32_t n) {
            (void)v;
            if (p == 1u) {
                for (uint32_t i = tid; i < n; i += SGS_SORT_THREADS) {
                    const uint32_t k = resident ? sm.key[i] : a.raw[lo + i];
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
            load(v, lo, n);
            prof_mark(a.prof, pslot);
            rank_slice(sm, n, shift, nd);
            publish_hist(sm.cnt, a.hist, v, nd);
        }
        { prof_mark(a.prof, pslot); grid_barrier(&a.ctl->bar_depth, bar_target); prof_mark(a.prof, pslot); }
        for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
            const uint32_t lo = v * SL, hi = min(P, lo + SL), n = hi > lo ? hi - lo : 0u;
            if (!resident) {
                load(v, lo, n);
                rank_slice(sm, n, shift, nd);
            }
            slice_bases(sm, a.hist, v, VB, nd);
            prof_mark(a.prof, pslot);
            scatter_slice(sm, n, shift, nd, out_key, out_val);
            __syncthreads();
        }
        { prof_mark(a.prof, pslot); grid_barrier(&a.ctl->bar_depth, bar_target); prof_mark(a.prof, pslot); }
    }
    if (npass == 0u) { prof_mark(a.prof, pslot); grid_barrier(&a.ctl->bar_depth, bar_target); prof_mark(a.prof, pslot); }

    // ---- scan of (area(rect_kept), tiles_touched, visible) in depth order
    const uint32_t* order = a.vals[0];
    auto slice_scan = [&](uint32_t lo, uint32_t n, Tri& mine_excl, Tri& block_total, uint32_t& ipt) {
        ipt = (n + SGS_SORT_THREADS - 1) / SGS_SORT_THREADS;
        Tri t = {0ull, 0ull, 0u};
        const uint32_t b = lo + tid * ipt, e = min(lo + n, b + ipt);
        for (uint32_t k = b; k < e; k++) {
            const uint32_t gid = __ldcg(order + k);
            const ushort4 r = a.rect_kept[gid];
            const uint32_t tt = a.tiles_touched[gid];
            t.kept += (uint32_t)(r.y - r.x) * (uint32_t)(r.w - r.z);
            t.touched += tt;
            t.vis += tt ? 1u : 0u;
        }
        Tri inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const Tri y = tri_shfl_up(inc, o);
            if (lane >= (uint32_t)o) inc = tri_add(inc, y);
        }
        __syncthreads();
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        Tri woff = {0ull, 0ull, 0u}, tot = {0ull, 0ull, 0u};
        for (uint32_t w = 0; w < 32; w++) {
            if (w < warp) woff = tri_add(woff, s_w[w]);
            tot = tri_add(tot, s_w[w]);
        }
        mine_excl = tri_add(woff, inc);
        mine_excl.kept -= t.kept;
        mine_excl.touched -= t.touched;
        mine_excl.vis -= t.vis;
        block_total = tot;
    };
    for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
        const uint32_t lo = v * SL, hi = min(P, lo + SL), n = hi > lo ? hi - lo : 0u;
        Tri ex, tot;
        uint32_t ipt;
        slice_scan(lo, n, ex, tot, ipt);
        if (tid == 0) {
            __stcg(a.blocksum + 3 * (size_t)v, tot.kept);
            __stcg(a.blocksum + 3 * (size_t)v + 1, tot.touched);
            __stcg(a.blocksum + 3 * (size_t)v + 2, (unsigned long long)tot.vis);
        }
    }
    { prof_mark(a.prof, pslot); grid_barrier(&a.ctl->bar_depth, bar_target); prof_mark(a.prof, pslot); }
    for (uint32_t v = blockIdx.x; v < VB; v += gridDim.x) {
        const uint32_t lo = v * SL, hi = min(P, lo + SL), n = hi > lo ? hi - lo : 0u;
        // slices before this one
        Tri pre = {0ull, 0ull, 0u};
        for (uint32_t vv = tid; vv < v; vv += SGS_SORT_THREADS) {
            pre.kept += __ldcg(a.blocksum + 3 * (size_t)vv);
            pre.touched += __ldcg(a.blocksum + 3 * (size_t)vv + 1);
            pre.vis += (uint32_t)__ldcg(a.blocksum + 3 * (size_t)vv + 2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pre.kept += __shfl_xor_sync(0xFFFFFFFFu, pre.kept, o);
            pre.touched += __shfl_xor_sync(0xFFFFFFFFu, pre.touched, o);
            pre.vis += __shfl_xor_sync(0xFFFFFFFFu, pre.vis, o);
        }
        __syncthreads();
        if (lane == 0) s_p[warp] = pre;
        __syncthreads();
        Tri base = {0ull, 0ull, 0u};
        for (uint32_t w = 0; w < 32; w++) base = tri_add(base, s_p[w]);

        Tri ex, tot;
        uint32_t ipt;
        slice_scan(lo, n, ex, tot, ipt);
        unsigned long long run = base.kept + ex.kept;
        const uint32_t b = lo + tid * ipt, e = min(lo + n, b + ipt);
        for (uint32_t k = b; k < e; k++) {
            const uint32_t gid = __ldcg(order + k);
            const ushort4 r = a.rect_kept[gid];
            run += (uint32_t)(r.y - r.x) * (uint32_t)(r.w - r.z);
            a.offs[k] = (uint32_t)run;
        }
        if (v == VB - 1u && tid == 0) {
            const Tri all = tri_add(base, tot);
            a.ctl->kept = all.kept;
            a.ctl->touched = all.touched;
            a.ctl->visible = all.vis;
            if (a.slot) {
                volatile HostSlot* hs = a.slot;
                hs->kept = all.kept;
                hs->touched = all.touched;
                hs->visible = all.vis;
                __threadfence_system();
                hs->ticket = a.ticket;
            }
        }
        __syncthreads();
    }
    prof_mark(a.prof, pslot);
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: instance generation + stable sort by tile + tile ranges
// ------------------------------------------------------------------------------------------------
struct TileArgs {
    int P;
    int tiles_x;
    int n_tiles;
    int tile_bits;
    int keep;
    unsigned long long cap;
    const uint32_t* order;      // Gaussians in depth order
    const uint32_t* offs;       // inclusive scan of the kept tile counts, in depth order
    const ushort4* rect_kept;
    uint32_t* keys[2];
    uint32_t* vals[2];
    uint32_t* hist;
    uint2* ranges;
    uint32_t* header;
    BinCtl* ctl;
    unsigned long long* prof;
};

using TileSmem = SortSmem<SGS_TILE_ND, SGS_TILE_CHUNK>;

// Fill sm.key / sm.val with the instances [s, e) of the depth-ordered stream: Gaussian k (depth order) owns the
// instances [offs[k] - area_k, offs[k]), row-major over its kept tile rect.
__device__ __forceinline__ void generate_slice(TileSmem& sm, const TileArgs& a, uint32_t s, uint32_t e) {
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t P = (uint32_t)a.P;
    // first Gaussian whose inclusive offset exceeds s: 1024-ary search on the monotone offs[]
    uint32_t lo = 0, len = P;
    while (len > 1u) {
        const uint32_t step = (len + SGS_SORT_THREADS - 1) / SGS_SORT_THREADS;
        const uint64_t idx = (uint64_t)lo + (uint64_t)(tid + 1u) * step - 1u;
        const bool probe = idx < (uint64_t)lo + len;
        const int le = (probe && __ldcg(a.offs + idx) <= s) ? 1 : 0;
        const uint32_t c = (uint32_t)__syncthreads_count(le);
        const uint32_t nlo = lo + c * step;
        const uint32_t end = lo + len;
        lo = nlo;
        len = (nlo >= end) ? 0u : min(step, end - nlo);
        if (len == 0u) break;
    }
    // lo = first Gaussian (in depth order) with offs > s
    for (uint32_t kb = lo;; kb += SGS_SORT_THREADS) {
        const uint32_t k = kb + tid;
        uint32_t n = 0, start = 0, gid = 0;
        ushort4 r = {0, 0, 0, 0};
        bool beyond = true;    // this Gaussian's instances end at or after e (nothing more to do past it)
        if (k < P) {
            const uint32_t incl = __ldcg(a.offs + k);
            gid = __ldcg(a.order + k);
            r = a.rect_kept[gid];
            n = (uint32_t)(r.y - r.x) * (uint32_t)(r.w - r.z);
            start = incl - n;
            beyond = incl >= e;
            if (start >= e) n = 0;
        }
        const uint32_t w = (uint32_t)(r.y - r.x);
        if (n > 0 && n <= SGS_DUP_SMALL) {
            uint32_t j = start;
            for (uint32_t y = r.z; y < r.w; y++)
                for (uint32_t x = r.x; x < r.y; x++, j++)
                    if (j >= s && j < e) {
                        sm.key[j - s] = y * (uint32_t)a.tiles_x + x;
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
            const uint32_t sx0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)r.x, src);
            const uint32_t sy0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)r.z, src);
            const uint32_t sw = __shfl_sync(0xFFFFFFFFu, w, src);
            // instances of this Gaussian that fall inside [s, e)
            const uint32_t j0 = max(sstart, s) - sstart, j1 = min(sstart + sn, e) - sstart;
            for (uint32_t i = j0 + lane; i < j1; i += 32) {
                const uint32_t yy = i / sw, xx = i - yy * sw;
                sm.key[sstart + i - s] = (sy0 + yy) * (uint32_t)a.tiles_x + (sx0 + xx);
                sm.val[sstart + i - s] = sgid;
            }
        }
        if (__syncthreads_or((tid == SGS_SORT_THREADS - 1 && !beyond) ? 1 : 0) == 0) break;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SGS_SORT_THREADS, 1) tile_sort_kernel(const TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& sm = *reinterpret_cast<TileSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    int pslot = 64;
    prof_mark(a.prof, pslot);
    const unsigned long long kept64 = __ldcg(&a.ctl->kept);
    const uint32_t npass = (uint32_t)(a.tile_bits + 7) / 8u;
    const uint32_t side_final = (npass & 1u) ^ 1u;     // pass p writes side (p - 1) & 1
    if (kept64 > a.cap || kept64 == 0ull) {
        // over capacity: the host re-launches with a larger buffer; nothing may be written beyond the header
        if (blockIdx.x == 0 && tid == 0)
            *reinterpret_cast<uint4*>(a.header) = make_uint4(side_final, 0u, (uint32_t)a.keep, (uint32_t)a.cap);
        return;
    }
    const uint32_t Rk = (uint32_t)kept64;
    if (blockIdx.x == 0 && tid == 0)
        *reinterpret_cast<uint4*>(a.header) = make_uint4(side_final, Rk, (uint32_t)a.keep, (uint32_t)a.cap);

    const uint32_t G = gridDim.x;
    const uint32_t per_block = (Rk + G - 1) / G;
    const uint32_t slices_per_block = (per_block + SGS_TILE_CHUNK - 1) / SGS_TILE_CHUNK;
    const uint32_t VB = G * slices_per_block;
    const uint32_t SL = (Rk + VB - 1) / VB;
    const bool resident = slices_per_block == 1u;
    const uint32_t dbits = ((uint32_t)a.tile_bits + npass - 1u) / npass;
    const uint32_t nd = 1u << dbits;
    uint32_t bar_target = 0;

    for (uint32_t p = 1; p <= npass; p++) {
        const uint32_t shift = (p - 1u) * dbits;
        const uint32_t out = (p - 1u) & 1u;
        const uint32_t* in_key = a.keys[out ^ 1u];
        const uint32_t* in_val = a.vals[out ^ 1u];
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
            const uint32_t lo = min(Rk, v * SL), hi = min(Rk, lo + SL), n = hi - lo;
            if (n) load(lo, n);
            prof_mark(a.prof, pslot);
            rank_slice(sm, n, shift, nd);
            publish_hist(sm.cnt, a.hist, v, nd);
        }
        { prof_mark(a.prof, pslot); grid_barrier(&a.ctl->bar_tile, bar_target); prof_mark(a.prof, pslot); }
        for (uint32_t v = blockIdx.x; v < VB; v += G) {
            const uint32_t lo = min(Rk, v * SL), hi = min(Rk, lo + SL), n = hi - lo;
            if (!resident) {
                if (n) load(lo, n);
                rank_slice(sm, n, shift, nd);
            }
            slice_bases(sm, a.hist, v, VB, nd);
            prof_mark(a.prof, pslot);
            scatter_slice(sm, n, shift, nd, a.keys[out], a.vals[out]);
            __syncthreads();
        }
        { prof_mark(a.prof, pslot); grid_barrier(&a.ctl->bar_tile, bar_target); prof_mark(a.prof, pslot); }
    }

    // ---- per-tile [start, end) from the sorted tile ids
    const uint32_t* sorted = a.keys[side_final];
    for (uint32_t v = blockIdx.x; v < VB; v += G) {
        const uint32_t lo = min(Rk, v * SL), hi = min(Rk, lo + SL);
        for (uint32_t i = lo + tid; i < hi; i += SGS_SORT_THREADS) {
            const uint32_t cur = __ldcg(sorted + i);
            const uint32_t prev = i ? __ldcg(sorted + i - 1) : 0xFFFFFFFFu;
            if (cur != prev) {
                a.ranges[cur].x = i;
                if (i) a.ranges[prev].y = i;
            }
            if (i == Rk - 1u) a.ranges[cur].y = Rk;
        }
    }
    prof_mark(a.prof, pslot);
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static int bits_for_tiles(int n_tiles) {
    int b = 1;
    while ((1 << b) < n_tiles) b++;
    return b;
}
int binning_tile_bits(int n_tiles) { return bits_for_tiles(n_tiles); }
int binning_point_list_side(int n_tiles) { return (((bits_for_tiles(n_tiles) + 7) / 8) & 1) ^ 1; }

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
void binning_profile_enable(bool on) { g_prof_on = on; if (on) binning_profile(true); }

int binning_grid_blocks() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(depth_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DepthSmem));
        cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
        cached = sms;
        cached_dev = dev;
    }
    return cached;
}

int binning_depth_vblocks(int P) {
    const int G = binning_grid_blocks();
    if (G <= 0) return 0;
    const long per_block = ((long)P + G - 1) / G;
    const long slices = per_block > 0 ? (per_block + SGS_DEPTH_CHUNK - 1) / SGS_DEPTH_CHUNK : 1;
    return (int)(G * (slices > 0 ? slices : 1));
}

size_t binning_tile_hist_words(size_t cap) {
    const size_t G = (size_t)binning_grid_blocks();
    const size_t per_block = (cap + G - 1) / (G ? G : 1);
    size_t slices = (per_block + SGS_TILE_CHUNK - 1) / SGS_TILE_CHUNK;
    if (slices == 0) slices = 1;
    return G * slices * SGS_TILE_ND;
}

cudaError_t launch_depth_sort(int P, GeomState g, HostSlot* slot, unsigned long long ticket, cudaStream_t s) {
    const int G = binning_grid_blocks();
    if (G <= 0) return cudaErrorInvalidDevice;
    DepthArgs a;
    a.P = P;
    a.vblocks = g.depth_vblocks;
    a.slice = (P + a.vblocks - 1) / a.vblocks;
    a.raw = g.depth_raw;
    a.keys[0] = g.depth_keys[0];
    a.keys[1] = g.depth_keys[1];
    a.vals[0] = g.depth_vals[0];
    a.vals[1] = g.depth_vals[1];
    a.rect_kept = g.rect_kept;
    a.tiles_touched = g.tiles_touched;
    a.offs = g.offs;
    a.hist = g.hist;
    a.blocksum = g.blocksum;
    a.ctl = g.ctl;
    a.slot = slot;
    a.ticket = ticket;
    a.prof = g_prof_on ? g_prof_host : nullptr;
    void* args[] = {&a};
    return cudaLaunchCooperativeKernel((const void*)depth_sort_kernel, dim3(G), dim3(SGS_SORT_THREADS), args,
                                       sizeof(DepthSmem), s);
}

cudaError_t launch_tile_sort(int P, const ViewParams& vp, GeomState g, BinningState b, ImageState img, int keep,
                             cudaStream_t s) {
    const int G = binning_grid_blocks();
    if (G <= 0) return cudaErrorInvalidDevice;
    TileArgs a;
    a.P = P;
    a.tiles_x = vp.tiles_x;
    a.n_tiles = vp.tiles_x * vp.tiles_y;
    a.tile_bits = bits_for_tiles(a.n_tiles);
    a.keep = keep;
    a.cap = (unsigned long long)b.cap;
    a.order = g.depth_vals[0];
    a.offs = g.offs;
    a.rect_kept = g.rect_kept;
    a.keys[0] = b.tile_keys[0];
    a.keys[1] = b.tile_keys[1];
    a.vals[0] = b.gauss_vals[0];
    a.vals[1] = b.gauss_vals[1];
    a.hist = b.hist;
    a.ranges = img.ranges;
    a.header = b.header;
    a.ctl = g.ctl;
    a.prof = g_prof_on ? g_prof_host : nullptr;
    void* args[] = {&a};
    return cudaLaunchCooperativeKernel((const void*)tile_sort_kernel, dim3(G), dim3(SGS_SORT_THREADS), args,
                                       sizeof(TileSmem), s);
}

}  // namespace sgs

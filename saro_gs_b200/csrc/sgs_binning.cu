// Tile binning: produces, for every 16x16 tile, the list of Gaussian instances that touch
// it, ordered by (depth bits, Gaussian index) — the exact order the reference obtains
// from one stable 64-bit radix sort of  key = tile<<32 | float_bits(depth)
//   $R/cuda_rasterizer/rasterizer_impl.cu:70-111 (duplicateWithKeys)
//   $R/cuda_rasterizer/rasterizer_impl.cu:299-309 (SortPairs on bits [0, 32+bit))
//   $R/cuda_rasterizer/rasterizer_impl.cu:116-138 (identifyTileRanges)
//
// B200 design: a TWO-LEVEL sort over a CULLED instance set — far fewer bytes through HBM.
//   0. (in the preprocess kernel) the reference's 3-sigma tile rect of every Gaussian is clipped to
//      the exact axis-aligned bounding box of its  alpha >= 1/255  ellipse (`rect_kept`, with
//      explicit rounding margins).  The reference's own count (`tiles_touched`, rect area) is
//      kept for the API-visible num_rendered and the bit-exact tile-count checks; at config 2
//      the clipped rects sum to ~55 % of it.  A dropped instance would `continue` on all 256
//      pixels of its tile in the reference, so the image, depth and gradients are unchanged bit
//      for bit.  (SGS_FLAG_NO_TILE_CULL keeps the full rect: then ranges / point_list /
//      n_contrib equal the reference's bit for bit.)
//   1. sort the P Gaussians (not the R >> P instances) by their 32-bit depth bits
//      (stable => ties keep ascending Gaussian index);
//   2. scan area(rect_kept) in that depth order and emit the instances in depth order
//      (key = tile id only, value = Gaussian index);
//   3. one stable radix sort of the instances on ceil(log2(#tiles)) bits (13 bits at
//      1352x1014 => 2 CUB onesweep passes over 8-byte pairs instead of 6 passes over 12-byte pairs).
// A stable sort by tile of a depth-ordered stream is ordered by (tile, depth, index): the
// reference order restricted to the kept instances.
#include "sgs_common.cuh"
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

namespace sgs {

// One scan yields both counts: low word = instances kept (area of rect_kept), high word = the reference's
// tiles_touched.  Both totals are < 2^32 (checked on the host), so the words never interfere.
struct KeptInDepthOrder {
    const ushort4* rect_kept;
    const uint32_t* tiles_touched;
    __host__ __device__ __forceinline__ uint64_t operator()(const uint32_t& gid) const {
        const ushort4 r = rect_kept[gid];
        return (uint64_t)((uint32_t)(r.y - r.x) * (uint32_t)(r.w - r.z)) | ((uint64_t)tiles_touched[gid] << 32);
    }
};

void binning_geom_temp_bytes(int P, size_t* bytes) {
    size_t a = 0, b = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, a, k, v, P, 0, 32);
    KeptInDepthOrder op{nullptr, nullptr};
    auto it = thrust::make_transform_iterator((const uint32_t*)nullptr, op);
    cub::DeviceScan::InclusiveSum(nullptr, b, it, (uint64_t*)nullptr, P);
    *bytes = (a > b ? a : b) + 256;
}

// Tile ids are sorted as 16-bit keys whenever the grid has at most 65536 tiles (16.7 Mpixel): 6 instead of 8 bytes
// per instance through the emit kernel, both onesweep passes and the range detection.
void binning_inst_temp_bytes(size_t R, int tile_bits, size_t* bytes) {
    size_t a = 0, b = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, a, k, v, (int64_t)R, 0, tile_bits);
    cub::DoubleBuffer<uint16_t> k16(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b, k16, v, (int64_t)R, 0, tile_bits < 16 ? tile_bits : 16);
    *bytes = (a > b ? a : b) + 256;
}

#define SGS_DUP_SMALL 8

// Step 1+2a: depth sort of Gaussians, then inclusive scan of tiles_kept in depth order.
// On return (stream order) g.depth_vals[0] holds the depth-ordered Gaussian indices and
// g.sorted_offsets the scan; the number of kept instances = sorted_offsets[P-1].
cudaError_t launch_depth_sort_scan(int P, GeomState g, cudaStream_t s) {
    cub::DoubleBuffer<uint32_t> keys(g.depth_keys[0], g.depth_keys[1]);
    cub::DoubleBuffer<uint32_t> vals(g.depth_vals[0], g.depth_vals[1]);
    size_t tb = g.temp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(g.temp, tb, keys, vals, P, 0, 32, s);
    if (e != cudaSuccess) return e;
    // 4 passes of 8 bits: the result lands back in buffer 0; keep the code robust anyway.
    if (vals.Current() != g.depth_vals[0])
        cudaMemcpyAsync(g.depth_vals[0], vals.Current(), sizeof(uint32_t) * (size_t)P, cudaMemcpyDeviceToDevice, s);
    KeptInDepthOrder op{g.rect_kept, g.tiles_touched};
    auto it = thrust::make_transform_iterator((const uint32_t*)g.depth_vals[0], op);
    tb = g.temp_bytes;
    return cub::DeviceScan::InclusiveSum(g.temp, tb, it, g.sorted_offsets, P, s);
}

// Step 2b: emit (tile id, Gaussian index) for every tile of rect_kept of every visible Gaussian,
// visiting Gaussians in depth order.  Small rects are written by the owning thread; large rects are
// written by the whole warp (coalesced), which removes the long divergent per-thread loops
// of the reference's duplicateWithKeys.
template <typename KeyT>
__global__ void __launch_bounds__(256)
duplicate_kernel(int P, int tiles_x, const uint32_t* __restrict__ order, const uint64_t* __restrict__ sorted_offsets,
                 const ushort4* __restrict__ rect_kept, KeyT* __restrict__ tile_keys,
                 uint32_t* __restrict__ gauss_vals) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    uint32_t n = 0, off = 0, gid = 0;
    ushort4 r = {0, 0, 0, 0};
    if (k < P) {
        gid = order[k];
        r = rect_kept[gid];
        n = (uint32_t)(r.y - r.x) * (uint32_t)(r.w - r.z);
        if (n > 0) off = (k == 0) ? 0u : (uint32_t)sorted_offsets[k - 1];
    }
    if (n > 0 && n <= SGS_DUP_SMALL) {
        for (uint32_t y = r.z; y < r.w; y++)
            for (uint32_t x = r.x; x < r.y; x++) {
                tile_keys[off] = (KeyT)(y * tiles_x + x);
                gauss_vals[off] = gid;
                off++;
            }
    }
    unsigned big = __ballot_sync(0xFFFFFFFFu, n > SGS_DUP_SMALL);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t sn = __shfl_sync(0xFFFFFFFFu, n, src);
        const uint32_t soff = __shfl_sync(0xFFFFFFFFu, off, src);
        const uint32_t sgid = __shfl_sync(0xFFFFFFFFu, gid, src);
        const uint32_t sx0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)r.x, src);
        const uint32_t sy0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)r.z, src);
        const uint32_t sw = __shfl_sync(0xFFFFFFFFu, (uint32_t)r.y, src) - sx0;
        for (uint32_t i = lane; i < sn; i += 32) {
            const uint32_t yy = i / sw, xx = i - yy * sw;
            tile_keys[soff + i] = (KeyT)((sy0 + yy) * tiles_x + (sx0 + xx));
            gauss_vals[soff + i] = sgid;
        }
    }
}

// Step 4: per-tile [start,end) in the sorted instance list.
// Each thread scans 16 bytes of sorted keys (8 x u16 or 4 x u32, one vector load) plus the key before them.
template <typename KeyT>
__global__ void __launch_bounds__(256)
tile_ranges_kernel(uint32_t R, const KeyT* __restrict__ sorted_tiles, uint2* __restrict__ ranges,
                   uint32_t* __restrict__ header, uint4 header_words) {
    constexpr uint32_t kPer = 16 / sizeof(KeyT);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *reinterpret_cast<uint4*>(header) = header_words;   // binning-buffer header rides along
    const uint32_t i0 = t * kPer;
    if (i0 >= R) return;
    KeyT k[kPer];
    if (i0 + kPer <= R) {
        *reinterpret_cast<uint4*>(k) = *reinterpret_cast<const uint4*>(sorted_tiles + i0);
    } else {
#pragma unroll
        for (uint32_t j = 0; j < kPer; j++) k[j] = (i0 + j < R) ? sorted_tiles[i0 + j] : (KeyT)0;
    }
    uint32_t prev = (i0 == 0) ? 0xFFFFFFFFu : (uint32_t)sorted_tiles[i0 - 1];
#pragma unroll
    for (uint32_t j = 0; j < kPer; j++) {
        const uint32_t i = i0 + j;
        if (i < R) {
            const uint32_t cur = (uint32_t)k[j];
            if (cur != prev) {
                if (i != 0) ranges[prev].y = i;
                ranges[cur].x = i;
            }
            if (i == R - 1) ranges[cur].y = R;
            prev = cur;
        }
    }
}

static int bits_for_tiles(int n_tiles) {
    int b = 1;
    while ((1 << b) < n_tiles) b++;
    return b;
}

// Step 2b launcher: emit the depth-ordered (tile, Gaussian) stream.
static bool keys16(int n_tiles) { return n_tiles <= 65536; }

cudaError_t launch_duplicate(int P, const ViewParams& vp, GeomState g, BinningState b, cudaStream_t s) {
    if (keys16(vp.tiles_x * vp.tiles_y))
        duplicate_kernel<uint16_t><<<(P + 255) / 256, 256, 0, s>>>(P, vp.tiles_x, g.depth_vals[0], g.sorted_offsets,
                                                                  g.rect_kept, reinterpret_cast<uint16_t*>(b.tile_keys[0]),
                                                                  b.gauss_vals[0]);
    else
        duplicate_kernel<uint32_t><<<(P + 255) / 256, 256, 0, s>>>(P, vp.tiles_x, g.depth_vals[0], g.sorted_offsets,
                                                                  g.rect_kept, b.tile_keys[0], b.gauss_vals[0]);
    return cudaGetLastError();
}

// Step 3: stable sort by tile id.  Returns (through *point_list / *sorted_tiles) the buffers
// holding the sorted Gaussian indices and tile ids.
cudaError_t launch_tile_sort(size_t R, int n_tiles, BinningState b, const uint32_t** point_list,
                             const uint32_t** sorted_tiles, cudaStream_t s) {
    cub::DoubleBuffer<uint32_t> vals(b.gauss_vals[0], b.gauss_vals[1]);
    size_t tb = b.temp_bytes;
    cudaError_t e;
    if (keys16(n_tiles)) {
        cub::DoubleBuffer<uint16_t> keys(reinterpret_cast<uint16_t*>(b.tile_keys[0]),
                                         reinterpret_cast<uint16_t*>(b.tile_keys[1]));
        e = cub::DeviceRadixSort::SortPairs(b.temp, tb, keys, vals, (int64_t)R, 0, bits_for_tiles(n_tiles), s);
        *sorted_tiles = reinterpret_cast<const uint32_t*>(keys.Current());
    } else {
        cub::DoubleBuffer<uint32_t> keys(b.tile_keys[0], b.tile_keys[1]);
        e = cub::DeviceRadixSort::SortPairs(b.temp, tb, keys, vals, (int64_t)R, 0, bits_for_tiles(n_tiles), s);
        *sorted_tiles = keys.Current();
    }
    *point_list = vals.Current();
    return e;
}

// Step 4 launcher.
cudaError_t launch_tile_ranges(size_t R, int n_tiles, const uint32_t* sorted_tiles, ImageState img, uint32_t* header,
                               const uint32_t header_words[4], cudaStream_t s) {
    const uint4 hw = make_uint4(header_words[0], header_words[1], header_words[2], header_words[3]);
    const size_t per = (n_tiles <= 65536) ? 8 : 4;
    const unsigned grid = (unsigned)(((R + per - 1) / per + 255) / 256);
    if (n_tiles <= 65536)
        tile_ranges_kernel<uint16_t><<<grid, 256, 0, s>>>((uint32_t)R, reinterpret_cast<const uint16_t*>(sorted_tiles),
                                                         img.ranges, header, hw);
    else
        tile_ranges_kernel<uint32_t><<<grid, 256, 0, s>>>((uint32_t)R, sorted_tiles, img.ranges, header, hw);
    return cudaGetLastError();
}

int binning_tile_bits(int n_tiles) { return bits_for_tiles(n_tiles); }

}  // namespace sgs

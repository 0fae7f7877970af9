// saro_gs_b200 — B200-native differentiable 3D-Gaussian tile rasterizer (sm_100a).
// Shared device helpers, state layouts and launch declarations.
//
// Behavioural contract followed here (reference = yjb6/SaRO-GS,
// submodules/gaussian_rasterization_ch3, cited as $R/...):
//   tile size 16x16, 3 colour channels            $R/cuda_rasterizer/config.h:15-17
//   SH basis constants                            $R/cuda_rasterizer/auxiliary.h:22-39
//   ndc->pixel, tile rect, point transforms       $R/cuda_rasterizer/auxiliary.h:41-95
// The arithmetic below is written so that nvcc forms the same floating-point
// expression trees (same association, same shared sub-expressions) as the reference's
// glm-based code: radii / tile counts must be bit-identical (SURVEY.md §7 "hard parts").
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cstdlib>

#define SGS_TILE_X 16
#define SGS_TILE_Y 16
#define SGS_TILE_PIX (SGS_TILE_X * SGS_TILE_Y)
#define SGS_CH 3

namespace sgs {

// ---------------------------------------------------------------------------------------
// SH constants ($R/cuda_rasterizer/auxiliary.h:22-39)
// ---------------------------------------------------------------------------------------
#define SGS_SH_C0 0.28209479177387814f
#define SGS_SH_C1 0.4886025119029199f
#define SGS_SH_C2_0 1.0925484305920792f
#define SGS_SH_C2_1 -1.0925484305920792f
#define SGS_SH_C2_2 0.31539156525252005f
#define SGS_SH_C2_3 -1.0925484305920792f
#define SGS_SH_C2_4 0.5462742152960396f
#define SGS_SH_C3_0 -0.5900435899266435f
#define SGS_SH_C3_1 2.890611442640554f
#define SGS_SH_C3_2 -0.4570457994644658f
#define SGS_SH_C3_3 0.3731763325901154f
#define SGS_SH_C3_4 -0.4570457994644658f
#define SGS_SH_C3_5 1.445305721320277f
#define SGS_SH_C3_6 -0.5900435899266435f

// Per-view constants, passed by value to kernels (kernel-parameter constant bank).  The
// camera matrices / campos / background stay DEVICE pointers exactly as in the reference
// API (the caller's tensors live on the GPU; reading them on the host would force a sync);
// per-Gaussian kernels stage them once per block into shared memory.
struct ViewParams {
    const float* view;    // [16] world->view, row-vector convention => indexed column-major
    const float* proj;    // [16] full projection, same convention
    const float* campos;  // [3]
    const float* bg;      // [3]
    int   W, H;
    int   tiles_x, tiles_y;
    float tan_fovx, tan_fovy;
    float focal_x, focal_y;
    float scale_modifier;
    int   sh_degree;      // active degree D
    int   sh_coeffs;      // M = coefficients per Gaussian in the sh tensor
    int   prefiltered;
};

// Block-wide staging of the camera constants (call before any early return).
struct ViewSmem {
    float view[16];
    float proj[16];
    float campos[3];
};
__forceinline__ __device__ void stage_view(ViewSmem& sm, const ViewParams& vp) {
    const int t = threadIdx.x;
    if (t < 16) sm.view[t] = vp.view[t];
    else if (t < 32) sm.proj[t - 16] = vp.proj[t - 16];
    else if (t < 35) sm.campos[t - 32] = vp.campos[t - 32];
    __syncthreads();
}

// Device-resident control block of one forward call (lives in the geometry buffer; zeroed by the preprocess
// kernel, filled by the binning kernels).  The host never has to read it: the counts it needs travel through a
// pinned host slot written by the depth-sort kernel (sgs_api.cu).
struct BinCtl {
    uint32_t bar_depth;   // grid-barrier arrival counter of depth_sort_kernel
    uint32_t bar_tile;    // grid-barrier arrival counter of tile_sort_kernel
    uint32_t key_max;     // max depth key over the visible Gaussians
    uint32_t key_nmin;    // max of ~key  (min key = ~key_nmin)
    unsigned long long kept;      // instances to bin: sum of area(rect_kept)
    unsigned long long touched;   // the reference's num_rendered: sum of tiles_touched
    uint32_t visible;     // Gaussians with radii > 0
    uint32_t coarse;      // (supertile, Gaussian) instances: sum over Gaussians of the supertiles their kept rect overlaps
    uint32_t expand_done; // last-block ticket of tile_count_kernel (self-resetting)
    uint32_t pad[21];
};
static_assert(sizeof(BinCtl) == 128, "BinCtl must be 128 bytes");

// What the depth-sort kernel reports to the host through pinned, device-mapped memory (one 64-byte slot per
// forward call in flight): payload first, then a system-scope fence, then the ticket the host is spinning on.
struct HostSlot {
    unsigned long long kept;
    unsigned long long touched;
    uint32_t visible;
    uint32_t pad0;
    unsigned long long pad1[4];
    unsigned long long ticket;
};
static_assert(sizeof(HostSlot) == 64, "HostSlot must be 64 bytes");

// Opaque "geometry" state: one entry per input Gaussian. SoA, every array 128-B aligned.
// Replaces GeometryState of $R/cuda_rasterizer/rasterizer_impl.h:33-48 (layout is free:
// the buffer is opaque to Python).
struct GeomState {
    float*    depths;         // [P]   view-space z
    float2*   means2D;        // [P]   pixel-space centre
    float4*   conic_opacity;  // [P]   (A, B, C, opacity)
    float4*   rgbd;           // [P]   (r, g, b, depth): colour after SH eval / precomputed copy
    float*    cov3D;          // [6P]  upper triangle of world covariance
    uint8_t*  clamped;        // [P]   bit c set  <=>  channel c was clamped at 0
    uint32_t* tiles_touched;  // [P]   tiles of the 3-sigma rect (the reference's count; sums to num_rendered)
    ushort4*  rect_kept;      // [P]   tile rect [x0,x1) x [y0,y1) actually binned: the reference's 3-sigma rect
                              //       clipped to the exact bounding box of the alpha >= 1/255 ellipse
    uint32_t* depth_raw;      // [P]   float bits of depth (0xFFFFFFFF when culled), written by preprocess
    uint2*    blk_range;      // [preprocess blocks] (max key, max ~key) over the block's visible Gaussians
    uint4*    blk_sums;       // [preprocess blocks] (kept tiles, touched tiles, visible Gaussians, -) of the block
    ushort4*  rect_sorted;    // [P]   rect_kept in depth order (written by the depth-sort kernel's scan)
    int       n_blk_range;
    uint32_t* depth_keys[2];  // [P]   radix-sort ping-pong: normalised keys
    uint32_t* depth_vals[2];  // [P]   radix-sort ping-pong: Gaussian index; [0] ends up holding the depth order
    uint32_t* coffs;          // [P]   inclusive scan, in depth order, of the supertiles each kept rect overlaps
    BinCtl*   ctl;            // control block (see above)
    uint32_t* hist;           // [depth_vblocks + 1][512] per-block digit histograms of the current radix pass
    uint32_t* blocksum;       // [depth_vblocks] per-slice supertile count of the scan
    int       depth_vblocks;  // virtual blocks of the depth sort (multiple of the grid size)
    float*    acc;            // [P][12] moment accumulators of the backward render kernel; present only when the state
                              // is kept for backward: zeroed by the forward render kernel (its memory system is idle),
                              // consumed AND re-zeroed by the backward-preprocess kernel (a second backward pass over
                              // the same state starts clean) — no memset launch in backward
};

// Opaque "image" state: per pixel + per tile.  Replaces ImageState ($R/.../rasterizer_impl.h:50-57).
struct ImageState {
    float*    final_T;     // [W*H]
    uint32_t* n_contrib;   // [W*H]  1-based list position of last blended instance (original list index)
    uint2*    ranges;      // [tiles] [start, end) into the sorted instance list
    uint32_t* tile_count;  // [tiles] number of packed (tile-culled) records written by forward render
    uint2*    cranges;     // [supertiles] [start, end) of every 4x4-tile supertile in the coarse list
};

// One packed, tile-ordered instance record (48 B, 16-B aligned) written by the forward
// render kernel and streamed back-to-front by the backward render kernel.
struct __align__(16) PackedInst {
    float x, y, A, B;          // centre (pixels), conic A, B
    float C, opacity, thr;     // conic C, opacity, skip threshold on `power`
    uint32_t list_pos;         // 0-based position in the tile's original sorted list
    float r, g, b;             // colour
    uint32_t gid;              // Gaussian index (low 28 bits) | quadrant visit mask (top 4 bits)
};
static_assert(sizeof(PackedInst) == 48, "PackedInst must be 48 bytes");

// Opaque "binning" state: per tile instance.  Replaces BinningState ($R/.../rasterizer_impl.h:59-69).
// Sized for `cap` instances (a prediction from the previous frame); the kernels read the true count from BinCtl.
struct BinningState {
    uint32_t*   header;         // [32] word 1: kept instances; 2: packed records present; 3: cap
    PackedInst* packed;         // [cap] tile-ordered packed records (tile t uses [ranges[t].x, +tile_count[t]))
    uint32_t*   point_list;     // [cap] Gaussian index of every kept instance, tile-major, depth order inside a tile
    uint32_t*   coarse_keys[2]; // [cap] supertile id | tile mask << 16; radix-sort ping-pong (only when > 512 supertiles)
    uint32_t*   coarse_vals[2]; // [cap] Gaussian index; ping-pong (only when > 512 supertiles)
    unsigned long long* coarse_pairs;  // [cap] the supertile-major coarse list: Gaussian index | key << 32
    uint32_t*   hist;           // per-block digit histograms of the coarse radix pass
    size_t      cap;
};

template <typename T>
__host__ __device__ inline void carve(char*& chunk, T*& ptr, size_t count, size_t alignment = 128) {
    size_t off = (reinterpret_cast<size_t>(chunk) + alignment - 1) & ~(alignment - 1);
    ptr = reinterpret_cast<T*>(off);
    chunk = reinterpret_cast<char*>(ptr + count);
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl may be scheduled while its predecessor in
// the stream is still draining (its launch latency, ~3 us per kernel boundary on B200, disappears); it must execute
// pdl_wait() before it touches anything the predecessor wrote.  pdl_wait() is a no-op for a normal launch.
// SGS_NO_PDL=1 turns the attribute off (A/B measurements).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("SGS_NO_PDL");
        return !(e && e[0] == '1');
    }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------
// Small device helpers
// ---------------------------------------------------------------------------------------

// $R/cuda_rasterizer/auxiliary.h:41-44 — evaluated in double, exactly as the reference's
// un-suffixed literals force it to be.
__forceinline__ __device__ float ndc2pix(float v, int S) {
    return ((v + 1.0) * S - 1.0) * 0.5;
}

// $R/cuda_rasterizer/auxiliary.h:46-56 — integer radius, float division, int truncation.
__forceinline__ __device__ void get_rect(const float2 p, int max_radius, uint2& rmin, uint2& rmax,
                                         int gx, int gy) {
    rmin.x = (unsigned)min(gx, max(0, (int)((p.x - max_radius) / SGS_TILE_X)));
    rmin.y = (unsigned)min(gy, max(0, (int)((p.y - max_radius) / SGS_TILE_Y)));
    rmax.x = (unsigned)min(gx, max(0, (int)((p.x + max_radius + SGS_TILE_X - 1) / SGS_TILE_X)));
    rmax.y = (unsigned)min(gy, max(0, (int)((p.y + max_radius + SGS_TILE_Y - 1) / SGS_TILE_Y)));
}

// $R/cuda_rasterizer/auxiliary.h:58-76
__forceinline__ __device__ float3 xform_point_4x3(const float3& p, const float* m) {
    float3 t = {
        m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
        m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
        m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
    };
    return t;
}
__forceinline__ __device__ float4 xform_point_4x4(const float3& p, const float* m) {
    float4 t = {
        m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
        m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
        m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
        m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]
    };
    return t;
}

// Minimal column-major 3x3 with the same product association as the reference's matrix
// library: R[c][r] = (A[0][r]*B[c][0] + A[1][r]*B[c][1]) + A[2][r]*B[c][2].
struct Mat3 {
    float c[3][3];  // c[col][row]
};
__forceinline__ __device__ Mat3 mat3_cols(float a0, float a1, float a2, float b0, float b1, float b2,
                                          float c0, float c1, float c2) {
    Mat3 m;
    m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
    m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
    m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
    return m;
}
__forceinline__ __device__ Mat3 mat3_mul(const Mat3& a, const Mat3& b) {
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.c[i][j] = a.c[0][j] * b.c[i][0] + a.c[1][j] * b.c[i][1] + a.c[2][j] * b.c[i][2];
    return r;
}
__forceinline__ __device__ Mat3 mat3_T(const Mat3& a) {
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.c[i][j] = a.c[j][i];
    return r;
}

}  // namespace sgs

// ---------------------------------------------------------------------------------------
// Host-side launchers implemented in the individual .cu files
// ---------------------------------------------------------------------------------------
namespace sgs {

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                         cudaStream_t s);

int  preprocess_blocks(int P);
void launch_preprocess_fwd(int P, const ViewParams& vp, const float* means3D, const float* scales,
                           const float* rotations, const float* opacities, const float* shs,
                           const float* cov3D_precomp, const float* colors_precomp, int* radii,
                           GeomState g, uint32_t* zero_words, size_t n_zero, int cull, cudaStream_t s);

// Binning (sgs_binning.cu): no host-visible sizes between the stages.
//   depth_sort  : persistent kernel; stable radix sort of the P depth keys + scan of the supertile counts in depth
//                 order; reports (kept, touched, visible) to `slot` (pinned host memory) under `ticket`.
//   tile_binning: persistent kernel bucketing the (supertile, Gaussian) stream + two expansion kernels that write
//                 the per-tile lists and ranges.  Writes nothing when kept > b.cap (the host then re-launches it
//                 with a larger buffer).
unsigned long long* binning_profile(bool enable);   // developer aid: pinned buffer of 128 phase timestamps (ns)
void   binning_profile_enable(bool on);
int    binning_grid_blocks();                       // co-resident blocks of the persistent kernels (= #SMs)
int    binning_depth_vblocks(int P);
size_t binning_hist_words(size_t n);
int    binning_supertiles(int tiles_x, int tiles_y, int* super_x);
int    binning_coarse_list_side(int n_super);
void   binning_set_mode(int mode);                  // this thread's next launches: 1 sort inside the supertiles, 0 globally
int    binning_bucket_capacity();                   // entries a supertile block sorts in shared memory
cudaError_t launch_depth_sort(int P, GeomState g, HostSlot* slot, unsigned long long ticket, cudaStream_t s);
cudaError_t launch_tile_binning(int P, const ViewParams& vp, GeomState g, BinningState b, ImageState img, int keep,
                                cudaStream_t s);
// the normal path: depth_sort + the persistent part of tile_binning in ONE cooperative launch, then the expansion
cudaError_t launch_binning_fused(int P, const ViewParams& vp, GeomState g, BinningState b, ImageState img, int keep,
                                 HostSlot* slot, unsigned long long ticket, cudaStream_t s);

void launch_render_fwd(int P, const ViewParams& vp, GeomState g, BinningState b, ImageState img,
                       const uint32_t* point_list, int write_packed, int tile_cull, float* out_color,
                       float* out_depth, cudaStream_t s);

void launch_render_bwd(const ViewParams& vp, BinningState b, ImageState img, const float* dL_dpix,
                       float* acc /*[P][12]*/, cudaStream_t s);

// running buffers of the densification statistics (sgs_densify_attach): all NULL = not attached
struct DensifySink {
    float* grad_sum;
    int* vis_count;
    int* radii_max;
};

void launch_preprocess_bwd(int P, const ViewParams& vp, const float* means3D, const int* radii, const float* shs,
                           const float* scales, const float* rotations, const float* cov3D, GeomState g,
                           float* acc, float* dL_dmean2D, float* dL_dopacity, float* dL_dcolor,
                           float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                           DensifySink sink, cudaStream_t s);

}  // namespace sgs

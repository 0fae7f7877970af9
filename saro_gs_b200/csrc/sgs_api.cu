// C-ABI entry points (include/saro_gs_b200.h) and host orchestration.
//
// Replaces the host side of $R/cuda_rasterizer/rasterizer_impl.cu:198-436
// (Rasterizer::forward / backward / markVisible): buffer carving, kernel sequencing, the
// single device->host read of num_rendered.  No torch types cross this boundary.
#include "../../include/saro_gs_b200.h"
#include "sgs_common.cuh"

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx when a tool is attached

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <chrono>
#include <cstdlib>
#include <atomic>

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
    char buf[512];
    if (ce != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(ce));
    else snprintf(buf, sizeof(buf), "%s", what);
    g_last_error = buf;
    return code;
}

#define SGS_CUDA_OK(expr)                                                        \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) return fail(SGS_ERR_CUDA, #expr, _e);             \
    } while (0)

sgs::GeomState carve_geom(char*& chunk, size_t P) {
    sgs::GeomState g;
    sgs::carve(chunk, g.depths, P);
    sgs::carve(chunk, g.means2D, P);
    sgs::carve(chunk, g.conic_opacity, P);
    sgs::carve(chunk, g.rgbd, P);
    sgs::carve(chunk, g.cov3D, P * 6);
    sgs::carve(chunk, g.clamped, P);
    sgs::carve(chunk, g.tiles_touched, P);
    sgs::carve(chunk, g.rect_kept, P);
    sgs::carve(chunk, g.depth_raw, P);
    g.n_blk_range = sgs::preprocess_blocks((int)P);
    sgs::carve(chunk, g.blk_range, (size_t)g.n_blk_range);
    sgs::carve(chunk, g.blk_sums, (size_t)g.n_blk_range);
    sgs::carve(chunk, g.rect_sorted, P);
    sgs::carve(chunk, g.depth_keys[0], P);
    sgs::carve(chunk, g.depth_keys[1], P);
    sgs::carve(chunk, g.depth_vals[0], P);
    sgs::carve(chunk, g.depth_vals[1], P);
    sgs::carve(chunk, g.coffs, P);
    sgs::carve(chunk, g.ctl, 1);
    g.depth_vblocks = sgs::binning_depth_vblocks((int)P);
    sgs::carve(chunk, g.hist, ((size_t)g.depth_vblocks + 1) * 512);
    sgs::carve(chunk, g.blocksum, (size_t)g.depth_vblocks);
    // always carved (an inference-mode forward merely never touches it): a backward call on such a state must not
    // write outside the buffer
    sgs::carve(chunk, g.acc, P * 12);
    return g;
}

sgs::ImageState carve_image(char*& chunk, size_t N, size_t tiles, size_t supers) {
    sgs::ImageState im;
    sgs::carve(chunk, im.final_T, N);
    sgs::carve(chunk, im.n_contrib, N);
    sgs::carve(chunk, im.ranges, tiles);
    sgs::carve(chunk, im.tile_count, tiles);
    sgs::carve(chunk, im.cranges, supers);
    return im;
}

// Layout: header | packed records (only if backward will run) | sort ping-pong buffers | histogram matrix, all sized
// for `cap` instances.  `packed` comes right after the header so that sgs_backward — which knows neither the
// capacity nor the number of kept instances — finds it without a device->host read.
sgs::BinningState carve_binning(char*& chunk, size_t cap, bool with_packed, bool header_and_packed_only = false) {
    sgs::BinningState b;
    sgs::carve(chunk, b.header, 32);
    if (with_packed) sgs::carve(chunk, b.packed, cap);
    else b.packed = nullptr;
    b.point_list = nullptr;
    b.coarse_keys[0] = b.coarse_keys[1] = b.coarse_vals[0] = b.coarse_vals[1] = nullptr;
    b.hist = nullptr;
    b.coarse_pairs = nullptr;
    b.cap = cap;
    if (header_and_packed_only) return b;
    sgs::carve(chunk, b.point_list, cap);
    // coarse instances <= kept instances <= cap
    sgs::carve(chunk, b.coarse_keys[0], cap);
    sgs::carve(chunk, b.coarse_vals[0], cap);
    sgs::carve(chunk, b.coarse_keys[1], cap);
    sgs::carve(chunk, b.coarse_vals[1], cap);
    sgs::carve(chunk, b.coarse_pairs, cap);
    sgs::carve(chunk, b.hist, sgs::binning_hist_words(cap));
    return b;
}

template <typename F>
size_t required_bytes(F f) {
    char* p = nullptr;
    f(p);
    return reinterpret_cast<size_t>(p) + 128;
}

sgs::ViewParams make_view(int W, int H, const float* view, const float* proj, const float* campos,
                          const float* bg, float tan_fovx, float tan_fovy, float scale_modifier, int D, int M,
                          int prefiltered) {
    sgs::ViewParams vp;
    vp.view = view;
    vp.proj = proj;
    vp.campos = campos;
    vp.bg = bg;
    vp.W = W;
    vp.H = H;
    vp.tiles_x = (W + SGS_TILE_X - 1) / SGS_TILE_X;
    vp.tiles_y = (H + SGS_TILE_Y - 1) / SGS_TILE_Y;
    vp.tan_fovx = tan_fovx;
    vp.tan_fovy = tan_fovy;
    // $R/cuda_rasterizer/rasterizer_impl.cu:222-223 (float arithmetic on the host)
    vp.focal_y = H / (2.0f * tan_fovy);
    vp.focal_x = W / (2.0f * tan_fovx);
    vp.scale_modifier = scale_modifier;
    vp.sh_degree = D;
    vp.sh_coeffs = M;
    vp.prefiltered = prefiltered;
    return vp;
}

// SGS_TRACE=1: print host-side timestamps (us) of sgs_forward phases to stderr (debug aid)
bool trace_on() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SGS_TRACE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Pinned, device-mapped ring of report slots: the depth-sort kernel writes (kept, touched, visible) + the call's
// ticket straight into host memory, and the host spins on the ticket — no stream synchronisation, no copy command
// in the stream, and (because every kernel of the forward pass is already queued when the host starts waiting) no
// GPU idle time around the one device->host dependency of the path ($R/cuda_rasterizer/rasterizer_impl.cu:281-286).
constexpr int kSlots = 64;
struct SlotRing {
    sgs::HostSlot* host = nullptr;
    sgs::HostSlot* dev = nullptr;
    unsigned long long next = 1;
};
SlotRing* slot_ring() {
    thread_local SlotRing ring;
    if (!ring.host) {
        void* h = nullptr;
        if (cudaHostAlloc(&h, sizeof(sgs::HostSlot) * kSlots, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
            return nullptr;
        memset(h, 0, sizeof(sgs::HostSlot) * kSlots);
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
            cudaFreeHost(h);
            return nullptr;
        }
        ring.host = reinterpret_cast<sgs::HostSlot*>(h);
        ring.dev = reinterpret_cast<sgs::HostSlot*>(d);
    }
    return &ring;
}

// spin until the kernel has published `ticket`; polls the stream now and then so that a failed launch or a sticky
// device error ends the wait instead of hanging the caller
int wait_for_ticket(const volatile sgs::HostSlot* hs, unsigned long long ticket, cudaStream_t s) {
    unsigned spins = 0;
    while (hs->ticket != ticket) {
        if ((++spins & 0x1FFFu) == 0u) {
            const cudaError_t q = cudaStreamQuery(s);
            if (q == cudaSuccess) {
                if (hs->ticket == ticket) break;
                return fail(SGS_ERR_CUDA, "sgs_forward: the binning kernel finished without reporting its counts");
            }
            if (q != cudaErrorNotReady) return fail(SGS_ERR_CUDA, "sgs_forward: device error while waiting for the instance count", q);
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return 0;
}

// Capacity of the binning buffer is predicted from the previous call of this thread (frames of a sequence, views of
// a training batch): 25 % head-room, scaled when the number of Gaussians grew.  A wrong guess is not an error — the
// binning kernel leaves the buffer untouched and the host re-launches it with the exact size.
struct CapPredictor {
    size_t last_kept = 0;
    int last_P = 0;
    // counts of the most recent forward call of this thread (sgs_last_forward_counts)
    unsigned long long kept = 0, touched = 0, visible = 0, cap = 0, relaunched = 0;
};
CapPredictor& cap_predictor() {
    thread_local CapPredictor p;
    return p;
}
std::atomic<long long> g_forced_cap{-1};   // sgs_debug_set_capacity: tests force the over-capacity path
std::atomic<int> g_forced_bin_mode{-1};    // sgs_debug_set_binning_mode: -1 automatic, 0 global depth sort, 1 per supertile

// Depth sort inside the supertiles (mode 1) while the expected bucket — (supertile, Gaussian) pairs per supertile,
// about two supertiles per visible Gaussian — stays well inside what a block sorts in shared memory; the global sort
// (mode 0) for denser frames.  Estimated from the previous call of this thread; a wrong guess only costs time.
int choose_binning_mode(int P, size_t supers) {
    const int forced = g_forced_bin_mode.load();
    if (forced >= 0) return forced ? 1 : 0;
    static const int env = [] {
        const char* e = getenv("SGS_BIN_MODE");
        return e ? atoi(e) : -1;
    }();
    if (env >= 0) return env ? 1 : 0;
    const CapPredictor& pr = cap_predictor();
    const unsigned long long vis = pr.last_P > 0 ? (unsigned long long)((double)pr.visible * (P > pr.last_P ? (double)P / pr.last_P : 1.0))
                                                 : (unsigned long long)P;
    const unsigned long long est = 2ull * vis / (supers ? supers : 1);
    // measured (tools/sweep.py, profiles/r2n_workload_sweep.json): buckets of ~4 700 entries (P = 1 M at 1352x1014, chunked
    // path for most supertiles) still favour the per-supertile sort by 3 %, ~6 400 (P = 2 M at 1920x1080) the global one
    return est * 10ull > 13ull * (unsigned long long)sgs::binning_bucket_capacity() ? 0 : 1;
}
size_t predict_capacity(int P) {
    const long long forced = g_forced_cap.load();
    size_t cap;
    const CapPredictor& pr = cap_predictor();
    if (forced >= 0) cap = (size_t)forced;
    else if (pr.last_P <= 0) cap = (size_t)P * 8 + 16384;
    else {
        const double grow = P > pr.last_P ? (double)P / pr.last_P : 1.0;
        cap = (size_t)((double)pr.last_kept * 1.25 * grow) + 16384;
    }
    cap = (cap + 4095) & ~(size_t)4095;
    if (cap > (size_t)0x7FFFF000u) cap = (size_t)0x7FFFF000u;
    return cap;
}

// ---------------------------------------------------------------------------------------------
// Stage profiler: CUDA events recorded on the launching stream around every stage, summed per
// stage by sgs_profile_read().  Off by default; bench.py turns it on to obtain the per-kernel
// durations behind its `roofline` object.
// ---------------------------------------------------------------------------------------------
constexpr int kStages = SGS_PROFILE_STAGES;
constexpr int kPool = 8192;
struct Profiler {
    bool enabled = false;
    cudaEvent_t ev[kPool][2];
    int stage_of[kPool];
    int created = 0;
    int used = 0;
    uint64_t own_launches = 0;  // hand-written kernels launched by this library since the last read
};
Profiler g_prof;
std::mutex g_prof_mu;

// SGS_NVTX=1: one NVTX range per stage (the reference has coarse timers only)
bool nvtx_on() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SGS_NVTX");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
const char* const kStageNames[kStages] = {"sgs/preprocess_fwd", "sgs/depth_sort_scan", "sgs/duplicate", "sgs/tile_sort",
                                          "sgs/tile_ranges",    "sgs/render_fwd",      "sgs/bwd_zero",  "sgs/render_bwd",
                                          "sgs/preprocess_bwd"};

struct StageScope {
    int slot = -1;
    bool range = false;
    cudaStream_t s;
    StageScope(int stage, cudaStream_t stream, int own_kernel_launches) : s(stream) {
        if (nvtx_on()) {
            nvtxRangePushA(kStageNames[stage]);
            range = true;
        }
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.own_launches += (uint64_t)own_kernel_launches;
        if (!g_prof.enabled || g_prof.used >= kPool) return;
        if (g_prof.used >= g_prof.created) {
            if (cudaEventCreate(&g_prof.ev[g_prof.created][0]) != cudaSuccess) return;
            if (cudaEventCreate(&g_prof.ev[g_prof.created][1]) != cudaSuccess) return;
            g_prof.created++;
        }
        slot = g_prof.used++;
        g_prof.stage_of[slot] = stage;
        cudaEventRecord(g_prof.ev[slot][0], s);
    }
    ~StageScope() {
        if (slot >= 0) cudaEventRecord(g_prof.ev[slot][1], s);
        if (range) nvtxRangePop();
    }
};

}  // namespace

extern "C" {

int sgs_abi_version(void) { return SGS_ABI_VERSION; }

const char* sgs_last_error(void) { return g_last_error.c_str(); }

int sgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present)))
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_mark_visible: null pointer");
    sgs::launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
    SGS_CUDA_OK(cudaGetLastError());
    return 0;
}

int64_t sgs_forward(sgs_resize_fn geometry_buffer, void* geometry_user, sgs_resize_fn binning_buffer,
                    void* binning_user, sgs_resize_fn image_buffer, void* image_user, int P, int D, int M,
                    const float* background, int width, int height, const float* means3D, const float* shs,
                    const float* colors_precomp, const float* opacities, const float* scales,
                    float scale_modifier, const float* rotations, const float* cov3D_precomp,
                    const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                    float tan_fovy, int prefiltered, float* out_color, float* out_depth, int* radii, int flags,
                    void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || width <= 0 || height <= 0 || width > 16 * 65535 || height > 16 * 65535)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: bad sizes");
    if (P == 0) return 0;  // reference binding skips the call entirely ($R/rasterize_points.cu:80)
    if (P >= (1 << 28)) return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: more than 2^28 - 1 Gaussians");
    if (!geometry_buffer || !binning_buffer || !image_buffer)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: null resize callback");
    if (!means3D || !opacities || !viewmatrix || !projmatrix || !cam_pos || !background || !out_color ||
        !out_depth || !radii)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: null required pointer");
    if (colors_precomp == nullptr && (shs == nullptr || M <= 0))
        // $R/cuda_rasterizer/rasterizer_impl.cu:242-245 analogue: colours must come from somewhere
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: provide SHs or precomputed colors");
    if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr))
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: provide scales+rotations or precomputed cov3D");
    if (colors_precomp == nullptr && (D < 0 || D > 3 || (D + 1) * (D + 1) > M))
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: SH degree/coefficient mismatch");

    const sgs::ViewParams vp = make_view(width, height, viewmatrix, projmatrix, cam_pos, background, tan_fovx,
                                         tan_fovy, scale_modifier, D, M, prefiltered);
    const size_t N = (size_t)width * height;
    const size_t tiles = (size_t)vp.tiles_x * vp.tiles_y;
    const bool tr = trace_on();
    const double t_enter = tr ? now_us() : 0;

    const size_t geom_bytes = required_bytes([&](char*& p) { carve_geom(p, (size_t)P); });
    char* gchunk = geometry_buffer(geometry_user, geom_bytes);
    if (!gchunk) return fail(SGS_ERR_ALLOC, "sgs_forward: geometry buffer allocation failed");
    sgs::GeomState g = carve_geom(gchunk, (size_t)P);

    const size_t supers = (size_t)sgs::binning_supertiles(vp.tiles_x, vp.tiles_y, nullptr);
    if (supers > 65536)     // the coarse sort carries the supertile id in 16 bits (images up to 16384 x 16384 pixels)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: image larger than 65536 supertiles of 64x64 pixels");
    const size_t img_bytes = required_bytes([&](char*& p) { carve_image(p, N, tiles, supers); });
    char* ichunk = image_buffer(image_user, img_bytes);
    if (!ichunk) return fail(SGS_ERR_ALLOC, "sgs_forward: image buffer allocation failed");
    sgs::ImageState img = carve_image(ichunk, N, tiles, supers);

    const int cull = (flags & SGS_FLAG_NO_TILE_CULL) ? 0 : 1;
    const bool keep = (flags & SGS_FLAG_KEEP_FOR_BACKWARD) != 0;
    if (sgs::binning_grid_blocks() <= 0) return fail(SGS_ERR_CUDA, "sgs_forward: no CUDA device");
    sgs::binning_set_mode(choose_binning_mode(P, supers));

    // Binning buffer sized from a PREDICTION (see predict_capacity): every kernel of the pass is queued before the
    // host learns the true instance count, so the GPU never idles while the host waits for it.
    size_t cap = predict_capacity(P);
    size_t bin_bytes = required_bytes([&](char*& p) { carve_binning(p, cap, keep); });
    char* bchunk = binning_buffer(binning_user, bin_bytes);
    if (!bchunk) return fail(SGS_ERR_ALLOC, "sgs_forward: binning buffer allocation failed");
    sgs::BinningState bin = carve_binning(bchunk, cap, keep);

    SlotRing* ring = slot_ring();
    if (!ring) return fail(SGS_ERR_ALLOC, "sgs_forward: pinned report slots unavailable");
    const unsigned long long ticket = ring->next++;
    const int slot_idx = (int)(ticket % kSlots);

    {
        StageScope sc(SGS_STAGE_PREPROCESS_FWD, s, 1);
        // ranges + tile_count + supertile buckets are zeroed by the preprocess kernel (one span, padding included)
        uint32_t* z0 = reinterpret_cast<uint32_t*>(img.ranges);
        uint32_t* z1 = reinterpret_cast<uint32_t*>(img.cranges + supers);
        sgs::launch_preprocess_fwd(P, vp, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp,
                                   radii, g, z0, (size_t)(z1 - z0), cull, s);
    }
    auto render = [&]() -> int {
        StageScope sc(SGS_STAGE_RENDER_FWD, s, 1);
        sgs::launch_render_fwd(P, vp, g, bin, img, bin.point_list, keep ? 1 : 0, cull, out_color, out_depth, s);
        SGS_CUDA_OK(cudaGetLastError());
        return 0;
    };
    auto bin_and_render = [&]() -> int {      // stand-alone binning stage: capacity re-launch and per-stage profiling
        {
            StageScope sc(SGS_STAGE_TILE_SORT, s, 3);
            SGS_CUDA_OK(sgs::launch_tile_binning(P, vp, g, bin, img, keep ? 1 : 0, s));
        }
        return render();
    };
    bool split_stages;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        split_stages = g_prof.enabled;            // the stage profiler wants one launch per stage
    }
    if (split_stages) {
        {
            StageScope sc(SGS_STAGE_DEPTH_SORT_SCAN, s, 1);
            SGS_CUDA_OK(sgs::launch_depth_sort(P, g, ring->dev + slot_idx, ticket, s));
        }
        if (int rc = bin_and_render()) return rc;
    } else {
        {
            StageScope sc(SGS_STAGE_DEPTH_SORT_SCAN, s, 3);
            SGS_CUDA_OK(sgs::launch_binning_fused(P, vp, g, bin, img, keep ? 1 : 0, ring->dev + slot_idx, ticket, s));
        }
        if (int rc = render()) return rc;
    }
    const double t_presync = tr ? now_us() : 0;

    // The one device->host dependency of the path: the instance counts (same information as the reference's read
    // at $R/cuda_rasterizer/rasterizer_impl.cu:281-282; `num_rendered` is what the reference returns to Python).
    if (int rc = wait_for_ticket(ring->host + slot_idx, ticket, s)) return rc;
    const unsigned long long Rk = ring->host[slot_idx].kept;       // kept instances (binned, sorted and rendered)
    const unsigned long long R = ring->host[slot_idx].touched;     // the reference's num_rendered
    const double t_synced = tr ? now_us() : 0;
    if (Rk > 0x7FFFF000ull || R > 0x7FFFFFFFFFFFull)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_forward: more than 2^31 tile instances");
    bool relaunched = false;
    if (Rk > cap) {
        relaunched = true;
        // prediction too small: the binning kernel left everything untouched (ranges still zero, barrier counter
        // unused) and the render kernel drew background only — redo both with the exact size
        cap = ((size_t)Rk + 4095) & ~(size_t)4095;
        bin_bytes = required_bytes([&](char*& p) { carve_binning(p, cap, keep); });
        bchunk = binning_buffer(binning_user, bin_bytes);
        if (!bchunk) return fail(SGS_ERR_ALLOC, "sgs_forward: binning buffer allocation failed");
        bin = carve_binning(bchunk, cap, keep);
        if (int rc = bin_and_render()) return rc;
    }
    CapPredictor& pr = cap_predictor();
    pr.last_kept = (size_t)Rk;
    pr.last_P = P;
    pr.kept = Rk;
    pr.touched = R;
    pr.visible = ring->host[slot_idx].visible;
    pr.cap = cap;
    pr.relaunched = relaunched ? 1 : 0;
    if (tr)
        fprintf(stderr, "[sgs_forward] launches %.1f us | wait %.1f us | kept %llu cap %zu\n", t_presync - t_enter,
                t_synced - t_presync, Rk, cap);
    return (int64_t)R;
}

// one-shot sink of the calling thread, armed by sgs_densify_attach and consumed by its next sgs_backward
struct PendingSink {
    int P = -1;
    sgs::DensifySink sink{nullptr, nullptr, nullptr};
};
static PendingSink& pending_sink() {
    thread_local PendingSink p;
    return p;
}

int sgs_densify_attach(int P, float* grad_sum, int* vis_count, int* radii_max) {
    PendingSink& p = pending_sink();
    if (!grad_sum && !vis_count && !radii_max) {   // detach
        p = PendingSink();
        return 0;
    }
    if (P < 0 || !grad_sum || !vis_count || !radii_max)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_densify_attach: bad size or null running buffer");
    p.P = P;
    p.sink = sgs::DensifySink{grad_sum, vis_count, radii_max};
    return 0;
}

int sgs_backward(int P, int D, int M, int64_t R, const float* background, int width, int height,
                 const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
                 float scale_modifier, const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                 float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                 const float* dL_dpix, float* dL_dmean2D, float* dL_dacc, float* dL_dopacity, float* dL_dcolor,
                 float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                 void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    // the densification sink is consumed by this call whatever its outcome (it must never leak into a later view)
    const PendingSink armed = pending_sink();
    pending_sink() = PendingSink();
    const sgs::DensifySink sink = armed.sink;
    if (sink.grad_sum && armed.P != P)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_backward: sgs_densify_attach was armed for a different number of Gaussians");
    if (P < 0 || R < 0 || width <= 0 || height <= 0) return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_backward: bad sizes");
    if (P == 0) return 0;
    if (!geom_buffer || !binning_buffer || !image_buffer)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_backward: null state buffer");
    (void)dL_dacc;   // ABI compatibility: the accumulator is part of the geometry state since round 2 (may be NULL)
    if (!means3D || !viewmatrix || !projmatrix || !campos || !background || !radii || !dL_dpix || !dL_dmean2D ||
        !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dcov3D || !dL_dscale || !dL_drot)
        return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_backward: null required pointer");
    if (shs != nullptr && M > 0 && !dL_dsh) return fail(SGS_ERR_INVALID_ARGUMENT, "sgs_backward: null dL_dsh");

    const sgs::ViewParams vp = make_view(width, height, viewmatrix, projmatrix, campos, background, tan_fovx,
                                         tan_fovy, scale_modifier, D, shs ? M : 0, 0);
    const size_t N = (size_t)width * height;
    const size_t tiles = (size_t)vp.tiles_x * vp.tiles_y;
    sgs::GeomState g = carve_geom(geom_buffer, (size_t)P);
    sgs::ImageState img = carve_image(image_buffer, N, tiles, (size_t)sgs::binning_supertiles(vp.tiles_x, vp.tiles_y, nullptr));
    // only the header + packed records are needed: their position does not depend on the kept count
    sgs::BinningState bin = carve_binning(binning_buffer, 0, true, /*header_and_packed_only=*/true);

    // No zero-fill here (round 1 issued a 14 MB memset): the moment accumulator g.acc was zeroed by the forward render
    // kernel and is left zeroed again by the backward-preprocess kernel below.
    if (R > 0) {
        StageScope sc(SGS_STAGE_RENDER_BWD, s, 1);
        sgs::launch_render_bwd(vp, bin, img, dL_dpix, g.acc, s);
    }
    const float* cov3D = cov3D_precomp ? cov3D_precomp : g.cov3D;
    {
        StageScope sc(SGS_STAGE_PREPROCESS_BWD, s, 1);
        sgs::launch_preprocess_bwd(P, vp, means3D, radii, shs, cov3D_precomp ? nullptr : scales,
                                   cov3D_precomp ? nullptr : rotations, cov3D, g, g.acc, dL_dmean2D, dL_dopacity,
                                   dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, sink, s);
    }
    (void)colors_precomp;
    SGS_CUDA_OK(cudaGetLastError());
    return 0;
}

int sgs_debug_export(int P, int width, int height, int64_t R, char* geom_buffer, char* binning_buffer,
                     char* image_buffer, uint32_t* tiles_touched, uint32_t* ranges, uint32_t* n_contrib,
                     float* final_T, float* means2D, float* conic_opacity, float* rgbd, float* cov3D,
                     uint32_t* tile_count, uint32_t* point_list, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return 0;
    const size_t N = (size_t)width * height;
    const int tx = (width + SGS_TILE_X - 1) / SGS_TILE_X, ty = (height + SGS_TILE_Y - 1) / SGS_TILE_Y;
    const size_t tiles = (size_t)tx * ty;
    const auto D2D = cudaMemcpyDeviceToDevice;
    if (geom_buffer) {
        sgs::GeomState g = carve_geom(geom_buffer, (size_t)P);
        if (tiles_touched) SGS_CUDA_OK(cudaMemcpyAsync(tiles_touched, g.tiles_touched, 4 * (size_t)P, D2D, s));
        if (means2D) SGS_CUDA_OK(cudaMemcpyAsync(means2D, g.means2D, 8 * (size_t)P, D2D, s));
        if (conic_opacity) SGS_CUDA_OK(cudaMemcpyAsync(conic_opacity, g.conic_opacity, 16 * (size_t)P, D2D, s));
        if (rgbd) SGS_CUDA_OK(cudaMemcpyAsync(rgbd, g.rgbd, 16 * (size_t)P, D2D, s));
        if (cov3D) SGS_CUDA_OK(cudaMemcpyAsync(cov3D, g.cov3D, 24 * (size_t)P, D2D, s));
    }
    if (image_buffer) {
        sgs::ImageState img = carve_image(image_buffer, N, tiles, (size_t)sgs::binning_supertiles(tx, ty, nullptr));
        if (ranges) SGS_CUDA_OK(cudaMemcpyAsync(ranges, img.ranges, 8 * tiles, D2D, s));
        if (n_contrib) SGS_CUDA_OK(cudaMemcpyAsync(n_contrib, img.n_contrib, 4 * N, D2D, s));
        if (final_T) SGS_CUDA_OK(cudaMemcpyAsync(final_T, img.final_T, 4 * N, D2D, s));
        if (tile_count) SGS_CUDA_OK(cudaMemcpyAsync(tile_count, img.tile_count, 4 * tiles, D2D, s));
    }
    if (binning_buffer && point_list && R > 0) {
        uint32_t hdr[4] = {0, 0, 0, 0};
        SGS_CUDA_OK(cudaMemcpyAsync(hdr, binning_buffer + ((128 - (reinterpret_cast<size_t>(binning_buffer) & 127)) & 127),
                                    sizeof(hdr), cudaMemcpyDeviceToHost, s));
        SGS_CUDA_OK(cudaStreamSynchronize(s));
        const size_t Rk = hdr[1], cap = hdr[3];
        if (Rk > 0) {
            sgs::BinningState b = carve_binning(binning_buffer, cap, hdr[2] != 0);
            SGS_CUDA_OK(cudaMemcpyAsync(point_list, b.point_list, 4 * Rk, D2D, s));
        }
    }
    return 0;
}

int64_t sgs_debug_kept(char* binning_buffer, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!binning_buffer) return 0;
    uint32_t hdr[3] = {0, 0, 0};
    SGS_CUDA_OK(cudaMemcpyAsync(hdr, binning_buffer + ((128 - (reinterpret_cast<size_t>(binning_buffer) & 127)) & 127),
                                sizeof(hdr), cudaMemcpyDeviceToHost, s));
    SGS_CUDA_OK(cudaStreamSynchronize(s));
    return (int64_t)hdr[1];
}

int sgs_last_forward_counts(int64_t* out5) {
    if (!out5) return SGS_ERR_INVALID_ARGUMENT;
    const CapPredictor& pr = cap_predictor();
    out5[0] = (int64_t)pr.kept;
    out5[1] = (int64_t)pr.touched;
    out5[2] = (int64_t)pr.visible;
    out5[3] = (int64_t)pr.cap;
    out5[4] = (int64_t)pr.relaunched;
    return 0;
}

void sgs_debug_set_binning_mode(int mode) { g_forced_bin_mode.store(mode < 0 ? -1 : (mode ? 1 : 0)); }

void sgs_debug_set_capacity(int64_t instances) { g_forced_cap.store(instances < 0 ? -1 : (long long)instances); }

int sgs_debug_binning_profile(int enable, uint64_t* out128) {
    sgs::binning_profile_enable(enable != 0);
    const unsigned long long* h = sgs::binning_profile(false);
    if (out128 && h) {
        if (cudaDeviceSynchronize() != cudaSuccess) return SGS_ERR_CUDA;
        for (int i = 0; i < 128; i++) out128[i] = h[i];
    }
    return 0;
}

void sgs_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.enabled = on != 0;
}

int sgs_profile_read(float* stage_ms, int* stage_calls, uint64_t* own_kernel_launches) {
    SGS_CUDA_OK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < kStages; i++) {
        if (stage_ms) stage_ms[i] = 0.f;
        if (stage_calls) stage_calls[i] = 0;
    }
    for (int k = 0; k < g_prof.used; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof.ev[k][0], g_prof.ev[k][1]) != cudaSuccess) continue;
        if (stage_ms) stage_ms[g_prof.stage_of[k]] += ms;
        if (stage_calls) stage_calls[g_prof.stage_of[k]] += 1;
    }
    if (own_kernel_launches) *own_kernel_launches = g_prof.own_launches;
    g_prof.used = 0;
    g_prof.own_launches = 0;
    return 0;
}

}  // extern "C"

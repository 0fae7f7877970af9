/* saro_gs_b200.h — C ABI of the B200-native differentiable Gaussian rasterizer.
 *
 * Drop-in boundary for the hot path of yjb6/SaRO-GS
 * (submodules/gaussian_rasterization_ch3, cited below as $R).  These entry points are
 * what the reference's binding layer ($R/rasterize_points.cu, pybind module `_C` in
 * $R/ext.cpp:15-19) calls into:
 *
 *   sgs_forward       replaces CudaRasterizer::Rasterizer::forward      $R/cuda_rasterizer/rasterizer.h:33-56
 *   sgs_backward      replaces CudaRasterizer::Rasterizer::backward     $R/cuda_rasterizer/rasterizer.h:58-87
 *   sgs_mark_visible  replaces CudaRasterizer::Rasterizer::markVisible  $R/cuda_rasterizer/rasterizer.h:26-31
 *
 * Conventions (identical to the reference unless stated):
 *   - every data pointer is a DEVICE pointer to float32 / int32 data on the current CUDA
 *     device, contiguous, row-major; optional inputs are signalled by NULL
 *     ($R/cuda_rasterizer/forward.cu:205,241);
 *   - viewmatrix / projmatrix are 4x4 in row-vector convention (world_view_transform =
 *     W2C^T), campos and background are 3 floats, all on the device;
 *   - the three state buffers are opaque byte buffers obtained through resize callbacks
 *     (the C form of the reference's std::function<char*(size_t)>,
 *     $R/cuda_rasterizer/rasterizer.h:34-36) and handed back unchanged to sgs_backward;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream, which is what the
 *     reference always uses); all work is enqueued on it.  sgs_forward needs the number of
 *     tile instances on the host like the reference does ($R/cuda_rasterizer/rasterizer_impl.cu:281-282),
 *     but it does NOT drain the stream for it: the binning buffer is sized from a prediction
 *     (previous call of the same thread), every kernel of the pass is queued, and only then
 *     the host waits for the count, which the binning kernel writes into pinned host memory.
 *     A prediction that turns out too small costs one re-launch of the binning + render
 *     kernels, never a wrong result;
 *   - alignment: every pointer must be 4-byte aligned.  16-byte alignment of shs / dL_dsh
 *     (with M == 16) and of rotations / dL_drot enables 128-bit accesses; otherwise scalar
 *     paths are taken.  The widened rows state their own (stricter) requirements.
 *
 * Differences from the reference ABI, all additive:
 *   - `stream` and `flags` parameters;
 *   - where the reference takes dL_dconic ([P][2][2], must be zero) sgs_backward keeps the parameter
 *     `dL_dacc` for ABI stability but ignores it (may be NULL): its [P][12] moment accumulator is
 *     part of the geometry state, zeroed by the forward pass (which therefore must have run with
 *     SGS_FLAG_KEEP_FOR_BACKWARD) and left zeroed again by the backward pass; and it WRITES every
 *     output element, so outputs need not be zero-initialised;
 *   - errors are reported by a negative return code + sgs_last_error() instead of C++
 *     exceptions.
 */
#ifndef SARO_GS_B200_H_
#define SARO_GS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGS_ABI_VERSION 1

/* sgs_forward flags */
#define SGS_FLAG_KEEP_FOR_BACKWARD 1 /* write the packed per-tile lists sgs_backward consumes */
#define SGS_FLAG_NO_TILE_CULL 2      /* validation only: disable the exact tile/quadrant culling, so that the
                                        internal lists (ranges, point_list, n_contrib) equal the reference's */

/* error codes (negative returns) */
#define SGS_ERR_INVALID_ARGUMENT -1
#define SGS_ERR_CUDA -2
#define SGS_ERR_ALLOC -3

/* Resize callback: must return a device pointer to at least `bytes` bytes that stays
 * valid until the matching sgs_backward has run (or is never called). */
typedef char* (*sgs_resize_fn)(void* user, size_t bytes);

int sgs_abi_version(void);
const char* sgs_last_error(void);

/* bool[P] <- (z_view > 0.2).  $R/cuda_rasterizer/rasterizer_impl.cu:54-66,141-153 */
int sgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* Forward: returns the number of rendered tile instances as the reference counts them (sum over Gaussians
 * of the tiles of their 3-sigma rect, >= 0) or a negative error code.  Internally only the instances that
 * can reach alpha >= 1/255 somewhere in their tile are binned, sorted and composited (bit-identical output).
 * D = active SH degree, M = SH coefficients per Gaussian (0 when shs == NULL).
 * out_color [3][H][W], out_depth [1][H][W], radii [P] int32 are fully written when P > 0. */
int64_t sgs_forward(sgs_resize_fn geometry_buffer, void* geometry_user,
                    sgs_resize_fn binning_buffer, void* binning_user,
                    sgs_resize_fn image_buffer, void* image_user,
                    int P, int D, int M,
                    const float* background, int width, int height,
                    const float* means3D, const float* shs, const float* colors_precomp,
                    const float* opacities, const float* scales, float scale_modifier,
                    const float* rotations, const float* cov3D_precomp,
                    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                    float tan_fovx, float tan_fovy, int prefiltered,
                    float* out_color, float* out_depth, int* radii,
                    int flags, void* stream);

/* Backward: R is the value sgs_forward returned.  Outputs (all fully written):
 *   dL_dmean2D [P][3] (z = 0), dL_dopacity [P], dL_dcolor [P][3], dL_dmean3D [P][3],
 *   dL_dcov3D [P][6], dL_dsh [P][M][3], dL_dscale [P][3], dL_drot [P][4].
 * dL_dacc: ignored (NULL allowed) — kept from ABI version 1, where it was a [P][12] float scratch. */
int sgs_backward(int P, int D, int M, int64_t R,
                 const float* background, int width, int height,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                 char* geom_buffer, char* binning_buffer, char* image_buffer,
                 const float* dL_dpix, float* dL_dmean2D, float* dL_dacc, float* dL_dopacity,
                 float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                 float* dL_dscale, float* dL_drot, void* stream);

/* Introspection used by the parity tests (bit-exact integer checks against the reference):
 * copies internal per-Gaussian / per-tile / per-pixel integer state to caller-provided
 * DEVICE buffers (any of which may be NULL). */
int sgs_debug_export(int P, int width, int height, int64_t R, char* geom_buffer, char* binning_buffer,
                     char* image_buffer, uint32_t* tiles_touched /*[P]*/, uint32_t* ranges /*[tiles][2]*/,
                     uint32_t* n_contrib /*[H*W]*/, float* final_T /*[H*W]*/, float* means2D /*[P][2]*/,
                     float* conic_opacity /*[P][4]*/, float* rgbd /*[P][4]*/, float* cov3D /*[P][6]*/,
                     uint32_t* tile_count /*[tiles]*/, uint32_t* point_list /*[>= sgs_debug_kept()]*/, void* stream);

/* Number of tile instances that survived the exact tile-level cull in the forward call that filled
 * `binning_buffer` (= length of point_list; equals the value sgs_forward returned under SGS_FLAG_NO_TILE_CULL).
 * Synchronises the stream. */
int64_t sgs_debug_kept(char* binning_buffer, void* stream);

/* Counts of the calling thread's most recent sgs_forward (no synchronisation: they were read during that call):
 * out5 = { kept tile instances (binned + rendered), num_rendered as the reference counts it, visible Gaussians
 * (radii > 0), capacity the binning buffer was sized for, 1 if the prediction was too small and the binning + render
 * kernels were launched a second time }. */
int sgs_last_forward_counts(int64_t* out5);

/* Validation only: force the predicted capacity of the binning buffer (in tile instances) for the following
 * sgs_forward calls, e.g. 0 to exercise the "prediction too small -> re-launch with the exact size" path;
 * a negative value restores the predictor. */
void sgs_debug_set_capacity(int64_t instances);
/* Tests / measurements: where the depth sort of the binning stage happens — 1 inside every supertile, 0 one global sort
 * of the Gaussians first, -1 (default) chosen per call from the previous frame's counts.  Both are exact. */
void sgs_debug_set_binning_mode(int mode);

/* Developer aid: phase timestamps (%globaltimer, ns) of the two persistent binning kernels, written by block 0 into
 * pinned memory while enabled.  Slots 0.. : depth-sort kernel, 64.. : tile-sort kernel.  out128 (host, 128 entries,
 * may be NULL) receives the timestamps of the most recent forward call; synchronises the device. */
int sgs_debug_binning_profile(int enable, uint64_t* out128);

/* Stage profiler (off by default): CUDA events on the launching stream around every stage.
 * sgs_profile_read synchronises the device, returns the summed milliseconds and call counts per
 * stage since the previous read, and the number of hand-written kernels this library launched. */
#define SGS_STAGE_PREPROCESS_FWD 0   /* per-Gaussian preprocess + exact tile-cull count (+ zeroing of ranges / control block) */
#define SGS_STAGE_DEPTH_SORT_SCAN 1 /* persistent kernel: radix sort of the P depth keys + scan of the kept tile counts */
#define SGS_STAGE_DUPLICATE 2       /* unused since round 2 (fused into SGS_STAGE_TILE_SORT) */
#define SGS_STAGE_TILE_SORT 3       /* persistent kernel: instance generation + radix sort by tile + tile ranges */
#define SGS_STAGE_TILE_RANGES 4     /* unused since round 2 (fused into SGS_STAGE_TILE_SORT) */
#define SGS_STAGE_RENDER_FWD 5
#define SGS_STAGE_BWD_ZERO 6        /* memset of the [P][12] accumulator */
#define SGS_STAGE_RENDER_BWD 7
#define SGS_STAGE_PREPROCESS_BWD 8
#define SGS_PROFILE_STAGES 9
void sgs_profile_enable(int on);
int sgs_profile_read(float* stage_ms /*[SGS_PROFILE_STAGES]*/, int* stage_calls /*[SGS_PROFILE_STAGES]*/,
                     uint64_t* own_kernel_launches);

/* ---------------------------------------------------------------------------------------------------
 * Photometric loss on the rasterizer's output — replaces, for CUDA float32 images,
 * the PyTorch code of the reference's utils/loss_utils.py:18-19 (l1_loss) and :38-68 (ssim: 11x11 Gaussian window,
 * sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2) as used by helper_train.py:50-53 / train.py:208-209.
 * Images are [B][C][H][W] float32, contiguous, on the device.
 *
 * sgs_l1_dssim_forward:  sums[b] = ( sum |img - gt| , sum ssim_map ) over the C planes of image b (device, [B][2]).
 *                        `dmaps` ([3][B*C][H][W] floats) receives what the backward pass needs, or NULL for
 *                        evaluation only.  `workspace`: sgs_loss_workspace_floats() floats of scratch.
 * sgs_l1_dssim_backward: dL_dimg = coef[b][0] * d(sum|.|)/dimg + coef[b][1] * d(sum ssim)/dimg, with coef ([B][2],
 *                        device) the gradient of the caller's scalar loss with respect to `sums`.
 * Both return 0 or a negative error code. */
size_t sgs_loss_workspace_floats(int B, int C, int H, int W);
int sgs_l1_dssim_forward(int B, int C, int H, int W, const float* img, const float* gt, float* dmaps, float* workspace,
                         float* sums, void* stream);
int sgs_l1_dssim_backward(int B, int C, int H, int W, const float* img, const float* gt, const float* dmaps,
                          const float* coef, float* dL_dimg, void* stream);
/* The training-step form in one call each way (helper_train.py:50-53 of the reference):
 *   loss[0] = (1 - lambda) * mean|img - gt| + lambda * (1 - mean ssim_map)      (device scalar)
 *   dL_dimg = grad_loss[0] * d loss / d img                                      (grad_loss: device scalar) */
int sgs_l1_dssim_loss_forward(int B, int C, int H, int W, const float* img, const float* gt, float lambda_dssim,
                              float* dmaps, float* workspace, float* loss, void* stream);
int sgs_l1_dssim_loss_backward(int B, int C, int H, int W, const float* img, const float* gt, float lambda_dssim,
                               const float* dmaps, const float* grad_loss, float* dL_dimg, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Per-frame deformation -> rasterizer hand-off — replaces, for the configuration every
 * shipped config uses (dx, drot, dopacity, dsh on; hidden width 128; time encoding 4; plane feature width 8/16/24/32),
 * GaussianModel.get_deformation_eval of the reference (scene/saro_gaussian.py:871-921): survival-state selection
 * (state = exp(-4 ((t - temporal_pos) / lifespan)^2) > 0.001, :757-759,:872-881), the three 3-layer MLPs
 * (:104,:108,:110) on [plane feature | time embedding], and the residual / activation epilogues that produce the
 * rasterizer's means3D, rotations, scales, opacities and shs for the selected Gaussians, in source order.
 *
 * sgs_deform_pack_mlp: converts one MLP's float32 nn.Linear parameters (W [out][in] row-major, device pointers) into
 *   the resident tensor-core image; call once per MLP (mlp = 0 motion [3 outputs], 1 rot+scale [7], 2 shs [48]) whenever
 *   the weights change.  `packed`: sgs_deform_packed_bytes() bytes on the device.  in_dim = feature width + 9.
 * sgs_deform_eval: all inputs are device float32, contiguous: xyz [N][3], rotation [N][4], scaling [N][3], opacity [N],
 *   features_dc [N][3], features_rest [N][45], temporal_pos [N], lifespan [N], hexplane_feature [N][feat_dim].
 *   Outputs have room for N rows; the first `return value` rows are written: out_means3D [.][3], out_rotations [.][4],
 *   out_scales [.][3], out_opacity [.], out_shs [.][48].  `workspace`: sgs_deform_workspace_bytes(N) bytes of device
 *   scratch.  rotation, hexplane_feature, out_rotations and out_shs must be 16-byte aligned (128-bit accesses;
 *   SGS_ERR_INVALID_ARGUMENT otherwise).  Returns the number of selected Gaussians (the host waits only for the selection pass; the MLP kernel is
 *   still running on `stream` when the call returns) or a negative error code.  The call keeps a small pinned
 *   read-back slot and an event per process: use it from one device per process (the deployment model of this
 *   library); calls are serialised by an internal mutex. */
size_t sgs_deform_packed_bytes(void);
size_t sgs_deform_workspace_bytes(int N);
int sgs_deform_pack_mlp(int mlp, int in_dim, const float* W1, const float* b1, const float* W2, const float* b2,
                        const float* W3, const float* b3, void* packed, void* stream);
int64_t sgs_deform_eval(int N, int feat_dim, float timestamp, const float* xyz, const float* rotation, const float* scaling,
                        const float* opacity, const float* features_dc, const float* features_rest,
                        const float* temporal_pos, const float* lifespan, const float* hexplane_feature,
                        const void* packed, void* workspace, size_t workspace_bytes, float* out_means3D,
                        float* out_rotations, float* out_scales, float* out_opacity, float* out_shs, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Training-time deformation (training half) — the MLP evaluations of
 * GaussianModel.get_deformation (scene/saro_gaussian.py:779-847) over ALL N Gaussians and their backward.  Each
 * evaluation is a job; all jobs of a call run in ONE persistent tcgen05 launch.
 *
 * sgs_deform_pack_general(backward = 0): image of an MLP  in_w -> 128 -> hid2 -> n_out  (nn.Linear parameters, W
 *   [out][in] row-major; in_w = feat_dim [opacity_mlp, :103] or feat_dim + 9 [motion/rot/shs, :102,:106,:108];
 *   hid2 <= 128; n_out <= 8 or == 48).  backward = 1: image of its data-gradient chain (transposed weights, no biases)
 *   n_out -> hid2 -> 128 -> feat_dim.  `image`: sgs_deform_image_bytes() device bytes.
 * sgs_deform_train_forward: per job, out[N][n_io] = MLP([feature | embedding(d)]) with d = timestamp - temporal_pos
 *   (zero_time = 0, :788-790) or d = 0 (zero_time = 1: base feature :793-794, and opacity_mlp whose image has zero
 *   weights on the time columns); writes the ReLU sign bits of both hidden layers (mask_a: layer 1, mask_b: layer 2;
 *   N x 16 bytes each, may be NULL) and, where given, operand planes for the weight-gradient GEMMs: save_a / save_b =
 *   the hidden activations h1 / h2 (16 column groups), save_in = the MLP input [feature | embedding | 0] (6 groups).
 * sgs_deform_train_backward: per job, in = dL/d out [N][n_io]; out = dL/d feature [N][feat_dim] of this job (the
 *   caller sums the jobs); mask_a = the forward's LAYER-2 bits, mask_b = its LAYER-1 bits; save_a / save_b receive
 *   dL/d(pre-activation) of layer 2 / layer 1 (16 groups), save_in receives dL/d out (2 groups for n_io <= 8, else 6).
 * Operand planes: a [rows][8 G] matrix as bf16 hi + lo (x = hi + lo), per 32-row tile a hi plane then a lo plane, each
 *   G groups of [32 rows][8 columns]: byte offset of (row R, group g, plane p) = ((R / 32) * 2 + p) * G * 512 + g * 512
 *   + (R % 32) * 16.  sgs_deform_planes_bytes(N, G) bytes (whole 128-row tiles; rows past N are written as zeros).
 *   This is the tensor-core operand form: sgs_deform_wgrad streams it from HBM to the MMA by TMA, no conversion.
 * All pointers are device memory, 16-byte aligned; the backward's out and 48-wide forward outs 32-byte aligned
 * (256-bit stores).  Return 0 or a negative error code. */
typedef struct sgs_mlp_job {
    const void* packed;
    const float* in;
    float* out;
    void* save_a;
    void* save_b;
    void* save_in;
    void* mask_a;
    void* mask_b;
    int n_io;
    int zero_time;
} sgs_mlp_job_t;
size_t sgs_deform_image_bytes(void);
size_t sgs_deform_planes_bytes(int N, int groups);
int sgs_deform_pack_general(int backward, int in_w, int hid2, int n_out, int feat_dim, const float* W1, const float* b1,
                            const float* W2, const float* b2, const float* W3, const float* b3, void* image, void* stream);
int sgs_deform_train_forward(int N, int feat_dim, float timestamp, const float* temporal_pos, const float* feature, int n_jobs,
                             const sgs_mlp_job_t* jobs, void* stream);
int sgs_deform_train_backward(int N, int feat_dim, int n_jobs, const sgs_mlp_job_t* jobs, void* stream);

/* Weight gradients of the training-time deformation: every task is one GEMM over the rows of two operand-plane
 * matrices (above),   D[m][n] = sum_r A[r][m] * B[r][n]   (A: 16 groups, m < 128; B: groups_b = 2, 6 or 16 groups) and,
 * when db is given,   db[m] = sum_r A[r][m],   written as dW[m][n] (row stride ldw; m < rows, n < cols) or, with
 * transposed = 1, dW[n][m] (the last layer: A = the hidden activation h2, B = dL/d out).  Layer 1: A = the backward's
 * save_b, B = the forward's save_in; layer 2: A = the backward's save_a, B = the forward's save_a; layer 3: A = the
 * forward's save_b, B = the backward's save_in.  accumulate = 1 adds to dW / db (the second evaluation of an MLP in the
 * same call).  `partials`: scratch of sgs_deform_wgrad_max_ctas() x sgs_deform_wgrad_partial_floats() floats.  One GEMM
 * launch + one or two reduction launches on `stream`; deterministic (fixed summation order).  N = the row count the
 * planes were produced for.  Returns 0 or a negative error code. */
typedef struct sgs_wgrad_task {
    const void* A;
    const void* B;
    int groups_b;
    float* dW;
    int ldw;
    int rows;
    int cols;
    int transposed;
    float* db;
    int accumulate;
} sgs_wgrad_task_t;
int sgs_deform_wgrad_max_ctas(void);
size_t sgs_deform_wgrad_partial_floats(void);
int sgs_deform_wgrad(int N, int n_tasks, const sgs_wgrad_task_t* tasks, float* partials, void* stream);

/* Elementwise epilogues of the training-time deformation (scene/saro_gaussian.py:782-831), forward and backward, one
 * launch each instead of ~25 PyTorch kernels: lifespan = (1 - min_scale) (1 - sigmoid(life_raw)) + min_scale;
 * opacity = sigmoid(opacity) exp(-4 ((timestamp - temporal_pos) / lifespan)^2); rotations = normalize(rotation +
 * rot_raw[:, :4]) (eps 1e-12); scales = exp(scaling + rot_raw[:, 4:]); means3D = xyz + motion_raw; real_xyz = xyz +
 * motion_base_raw (skipped when motion_base_raw is NULL).  Raw MLP outputs are [N][1], [N][3], [N][7], [N][3].
 * Backward: gradients of the five outputs (any g_* may be NULL = no gradient) -> d life_raw [N], d rot_raw [N][7],
 * d rotation [N][4], d scaling [N][3], d opacity [N], d temporal_pos [N]; the gradients of xyz and motion_raw are the
 * incoming means3D gradient itself.  Device float32, contiguous.  Return 0 or a negative error code. */
int sgs_deform_train_epilogue_forward(int N, float timestamp, float min_scale, const float* life_raw, const float* motion_raw,
                                      const float* rot_raw, const float* motion_base_raw, const float* xyz, const float* rotation,
                                      const float* scaling, const float* opacity, const float* temporal_pos, float* out_means3D,
                                      float* out_rotations, float* out_scales, float* out_opacity, float* out_lifespan,
                                      float* out_real_xyz, void* stream);
int sgs_deform_train_epilogue_backward(int N, float timestamp, float min_scale, const float* life_raw, const float* rot_raw,
                                       const float* rotation, const float* scaling, const float* opacity, const float* temporal_pos,
                                       const float* g_rotations, const float* g_scales, const float* g_opacity,
                                       const float* g_lifespan, float* d_life_raw, float* d_rot_raw, float* d_rotation,
                                       float* d_scaling, float* d_opacity, float* d_temporal_pos, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Densification statistics of one training iteration — replaces the per-view lists and
 * the batch reduction of the reference's train.py:192-218 and :281-292 (with scene/saro_gaussian.py:745-747).
 *
 * sgs_densify_add_view: after one view's backward pass: grad_sum[i] += |dL_dmeans2D[i][0:2]|, vis_count[i] += radii[i] > 0,
 *   radii_max[i] = max(radii_max[i], radii[i]).  dL_dmeans2D is the rasterizer's [P][3] means2D gradient, radii its
 *   int32 [P] output; the three running buffers ([P] float / int32 / int32, device) start at zero each iteration.
 * sgs_densify_commit: where vis_count > 0: max_radii2D = max(max_radii2D, radii_max); xyz_gradient_accum += grad_sum /
 *   vis_count; denom += 1   (all three are the model's float32 [P] statistics, updated in place).
 * Data-parallel training runs sgs_densify_add_view on each rank's view, all-reduces grad_sum (SUM), vis_count (SUM) and
 * radii_max (MAX), then commits on every rank.  Both return 0 or a negative error code. */
int sgs_densify_add_view(int P, const float* dL_dmeans2D, const int* radii, float* grad_sum, int* vis_count, int* radii_max,
                         void* stream);
/* Fused form of sgs_densify_add_view: arms the calling thread's NEXT sgs_backward call (of P Gaussians) to apply the
 * same update to the three running buffers in the epilogue of its last kernel, where dL_dmeans2D and radii are in
 * registers — no extra launch and no re-read (the reference appends to its lists right after loss.backward(),
 * train.py:211-215).  One-shot: the sink is cleared by that sgs_backward call whether it succeeds or not; all three
 * pointers NULL disarms it.  Call it from the thread that calls sgs_backward (in PyTorch: inside the autograd
 * function's backward — saro_gs_b200.densify.BatchDensifyStats.attach_next_backward does that).  Returns 0 or a
 * negative error code. */
int sgs_densify_attach(int P, float* grad_sum, int* vis_count, int* radii_max);
int sgs_densify_commit(int P, const float* grad_sum, const int* vis_count, const int* radii_max, float* max_radii2D,
                       float* xyz_gradient_accum, float* denom, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Scale-aware plane sampler — replaces ScaleAwareResField.forward of the reference
 * (scene/hexplane.py:258-286: normalize_aabb / normalize_time, get_level :231-242, interpolate_ms_features :91-137,
 * grid_sample_wrapper :26-60) including the third-party op it calls,
 *   nvdiffrast.torch.texture(tex, uv, mip_level_bias = min level of the plane's two axes, boundary_mode = "clamp",
 *                            max_mip_level = 7 for the three space planes, 0 for the three time planes),
 * rebuilt from its published algorithm (linear-mipmap-linear at an explicit level, texel centres at half-integers,
 * clamp-to-edge, 2x2 box mip stack; extents must be even at every level that is built).
 *
 * A plane is an nn.Parameter [1][C][H][W] (W follows coordinate dim_u, H follows dim_v; 0 x, 1 y, 2 z, 3 t).
 *   sgs_plane_levels / sgs_plane_pyramid_floats: mip levels (1..8) and floats of the channels-last pyramid of a plane.
 *   sgs_plane_build: NCHW parameter -> channels-last mip pyramid (run when the parameter changed, not per call).
 *   sgs_plane_sample_forward: out[n][out_offset + c] = sum over the n_planes planes of the sample at point n, for
 *     c < C; pts [N][3], timestamps [N], scales [N][3] (activated), aabb [2][3] (row 0 = xyz_max, row 1 = xyz_min) and
 *     base_scale [3] are DEVICE pointers (the module's buffers), time_scale = duration / (duration - 1), reso0 = HOST
 *     int[3], resolution of the coarsest grid; `planes` is a HOST array.  One call per resolution level of the field.
 *   sgs_plane_sample_backward: scatters w * dout[n][out_offset + c] into the GRADIENT pyramids (same layout, must be
 *     zero-initialised by the caller) given as `grad_planes`; positions and scales carry no gradient (the reference
 *     samples at detached inputs, scene/saro_gaussian.py:765,780,865).
 *   sgs_plane_fold: gradient pyramid -> dL/dplane [1][C][H][W] (fully written).
 * C must be a power of two below 32 or a multiple of 32 (and C * 132 bytes of shared memory <= 48 KB). */
typedef struct {
    float* pyramid;      /* device: channels-last levels, level l at the offset of all levels before it */
    int H, W;
    int dim_u, dim_v;
    int max_mip_level;   /* 7 or 0 in the reference */
} sgs_plane_t;
int sgs_plane_levels(int H, int W, int max_mip_level);
size_t sgs_plane_pyramid_floats(int C, int H, int W, int max_mip_level);
int sgs_plane_build(int C, int H, int W, int max_mip_level, const float* plane_nchw, float* pyramid, void* stream);
int sgs_plane_sample_forward(int N, int C, const float* pts, const float* timestamps, const float* scales,
                             const float* aabb, const float* base_scale, float time_scale, const int* reso0,
                             int n_planes, const sgs_plane_t* planes, int out_stride, int out_offset, float* out,
                             void* stream);
int sgs_plane_sample_backward(int N, int C, const float* pts, const float* timestamps, const float* scales,
                              const float* aabb, const float* base_scale, float time_scale, const int* reso0,
                              int n_planes, const sgs_plane_t* grad_planes, int out_stride, int out_offset,
                              const float* dout, void* stream);
int sgs_plane_fold(int C, int H, int W, int max_mip_level, const float* grad_pyramid, float* dplane_nchw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SARO_GS_B200_H_ */

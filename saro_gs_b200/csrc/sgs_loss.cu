// Fused photometric loss on the rasterizer's CHW output: per-image sums of |x - y| and of the SSIM map
// (11x11 Gaussian window, sigma 1.5, zero padding) and the gradient with respect to the rendered image.
//
// Restates, as two hand-written kernels, what the reference computes with five depthwise conv2d calls plus
// elementwise ops and autograd:
//   l1_loss                                      utils/loss_utils.py:18-19
//   gaussian / create_window (11 taps, 1.5)      utils/loss_utils.py:27-35
//   ssim / _ssim (C1 = 0.01^2, C2 = 0.03^2)      utils/loss_utils.py:38-68
//   loss = (1 - l) L1 + l (1 - ssim)             helper_train.py:50-53 (getloss), train.py:208-209
// this is the step that runs between the rasterizer's forward and backward in every
// training iteration and produces exactly the dL/dcolor image the backward rasterizer consumes.
//
// B200 design: HBM-bound streaming kernels.  One CTA = one 32x32 tile of one image plane; the two input tiles
// (+5 halo) are staged once in shared memory, the five windowed moments (x, y, xx, yy, xy) are produced by a
// separable pass (11 + 11 taps instead of 121), and the forward kernel also emits the three derivative maps
//   dS/dmu1, dS/dE[xx], dS/dE[xy]
// so that backward is a single separable convolution of those maps:
//   dL/dx(p) = cL1 sign(x - y) + cS [ (w * dS/dmu1)(p) + 2 x(p) (w * dS/dExx)(p) + y(p) (w * dS/dExy)(p) ].
// Sums are reduced deterministically (per-CTA partials, then one CTA per image in double precision).
// Algorithmic bytes per pixel-channel: forward 8 read + 12 written, backward 20 read + 4 written.
#include "../../include/saro_gs_b200.h"
#include <cuda_runtime.h>
#include <cstdint>

namespace sgs_loss {

#define SL_T 32          // tile edge (output pixels)
#define SL_R 5           // window radius
#define SL_E (SL_T + 2 * SL_R)   // staged edge: 42
#define SL_THREADS 256

// float32 weights exactly as the reference builds them: torch.Tensor([exp(-(x-5)^2 / (2*1.5^2))]) / sum
__constant__ float c_g[11] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f,
                              2.130055279e-01f, 2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f,
                              3.600077331e-02f, 7.598758209e-03f, 1.028380124e-03f};

__device__ __forceinline__ float tile_load(const float* __restrict__ p, int H, int W, int y, int x) {
    return (y >= 0 && y < H && x >= 0 && x < W) ? p[(size_t)y * W + x] : 0.f;   // zero padding (conv2d padding=5)
}

// sliding-window helper: out[j] = sum_k g[k] * v[j + k] for j = 0..3 from 14 consecutive samples
__device__ __forceinline__ void window4(const float (&v)[14], float (&out)[4]) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) a = fmaf(c_g[k], v[j + k], a);
        out[j] = a;
    }
}

__global__ void __launch_bounds__(SL_THREADS)
l1_dssim_fwd_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                    float* __restrict__ dm /*[3][N][H][W] or null*/, size_t plane_stride_all,
                    float* __restrict__ partials /*[N][tiles][2]*/) {
    __shared__ float sx[SL_E][SL_E + 1], sy[SL_E][SL_E + 1];
    __shared__ float sh[5][SL_E][SL_T + 1];
    __shared__ float s_red[2][SL_THREADS / 32];
    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * SL_T, y0 = blockIdx.y * SL_T;
    const float* px = img + (size_t)n * H * W;
    const float* py = gt + (size_t)n * H * W;

    {   // all global loads of the thread are issued before the first shared-memory store (latency paid once)
        constexpr int kIt = (SL_E * SL_E + SL_THREADS - 1) / SL_THREADS;
        float rx[kIt], ry[kIt];
#pragma unroll
        for (int it = 0; it < kIt; it++) {
            const int i = tid + it * SL_THREADS;
            const int r = i / SL_E, c = i - r * SL_E;
            const bool ok = i < SL_E * SL_E;
            rx[it] = ok ? tile_load(px, H, W, y0 + r - SL_R, x0 + c - SL_R) : 0.f;
            ry[it] = ok ? tile_load(py, H, W, y0 + r - SL_R, x0 + c - SL_R) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < kIt; it++) {
            const int i = tid + it * SL_THREADS;
            const int r = i / SL_E, c = i - r * SL_E;
            if (i < SL_E * SL_E) { sx[r][c] = rx[it]; sy[r][c] = ry[it]; }
        }
    }
    __syncthreads();
    // horizontal pass: 42 rows x 8 groups of 4 columns; every thread slides the 11-tap window over 14 samples held
    // in registers (products formed once per sample instead of once per tap)
    for (int i = tid; i < SL_E * (SL_T / 4); i += SL_THREADS) {
        const int r = i / (SL_T / 4), c0 = (i - r * (SL_T / 4)) * 4;
        float u[14], v[14], t[14], o[4];
#pragma unroll
        for (int k = 0; k < 14; k++) { u[k] = sx[r][c0 + k]; v[k] = sy[r][c0 + k]; }
        window4(u, o);
#pragma unroll
        for (int j = 0; j < 4; j++) sh[0][r][c0 + j] = o[j];
        window4(v, o);
#pragma unroll
        for (int j = 0; j < 4; j++) sh[1][r][c0 + j] = o[j];
#pragma unroll
        for (int k = 0; k < 14; k++) t[k] = u[k] * u[k];
        window4(t, o);
#pragma unroll
        for (int j = 0; j < 4; j++) sh[2][r][c0 + j] = o[j];
#pragma unroll
        for (int k = 0; k < 14; k++) t[k] = v[k] * v[k];
        window4(t, o);
#pragma unroll
        for (int j = 0; j < 4; j++) sh[3][r][c0 + j] = o[j];
#pragma unroll
        for (int k = 0; k < 14; k++) t[k] = u[k] * v[k];
        window4(t, o);
#pragma unroll
        for (int j = 0; j < 4; j++) sh[4][r][c0 + j] = o[j];
    }
    __syncthreads();
    // vertical pass: 32 columns x 8 groups of 4 rows = one item per thread; then SSIM + derivative maps
    float sum_abs = 0.f, sum_ssim = 0.f;
    const size_t HW = (size_t)H * W;
    {
        const int c = tid & (SL_T - 1), r0 = (tid / SL_T) * 4;
        float m[5][4];
#pragma unroll
        for (int p = 0; p < 5; p++) {
            float v[14];
#pragma unroll
            for (int k = 0; k < 14; k++) v[k] = sh[p][r0 + k][c];
            window4(v, m[p]);
        }
        const int gx = x0 + c;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gy = y0 + r0 + j;
            if (gy >= H || gx >= W) continue;
            const float mu1 = m[0][j], mu2 = m[1][j], e11 = m[2][j], e22 = m[3][j], e12 = m[4][j];
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
            const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2;
            const float B1 = mu1_sq + mu2_sq + C1, B2 = s1 + s2 + C2;
            const float inv = 1.f / (B1 * B2);
            const float S = A1 * A2 * inv;
            sum_ssim += S;
            sum_abs += fabsf(sx[r0 + j + SL_R][c + SL_R] - sy[r0 + j + SL_R][c + SL_R]);
            if (dm != nullptr) {
                // total derivative of S through mu1 (also inside sigma1^2 = Exx - mu1^2 and sigma12 = Exy - mu1 mu2)
                const float dS_dmu1 = 2.f * inv * (mu2 * (A2 - A1) - mu1 * S * (B2 - B1));
                const float dS_de11 = -S / B2;
                const float dS_de12 = 2.f * A1 * inv;
                const size_t o = (size_t)n * HW + (size_t)gy * W + gx;
                dm[o] = dS_dmu1;
                dm[plane_stride_all + o] = dS_de11;
                dm[2 * plane_stride_all + o] = dS_de12;
            }
        }
    }
    // deterministic block reduction
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        sum_abs += __shfl_xor_sync(0xFFFFFFFFu, sum_abs, d);
        sum_ssim += __shfl_xor_sync(0xFFFFFFFFu, sum_ssim, d);
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = sum_abs; s_red[1][tid >> 5] = sum_ssim; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b2 = 0.f;
        for (int w = 0; w < SL_THREADS / 32; w++) { a += s_red[0][w]; b2 += s_red[1][w]; }
        const size_t blk = ((size_t)n * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partials[2 * blk] = a;
        partials[2 * blk + 1] = b2;
    }
}

// one CTA per image: sums[b] = (sum |x-y|, sum ssim_map) over its C planes, accumulated in double
__global__ void __launch_bounds__(256)
reduce_partials_kernel(int blocks_per_image, const float* __restrict__ partials, float* __restrict__ sums) {
    __shared__ double s_a[256], s_b[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    double a = 0.0, c = 0.0;
    const float* p = partials + (size_t)b * blocks_per_image * 2;
    for (int i = tid; i < blocks_per_image; i += 256) { a += p[2 * i]; c += p[2 * i + 1]; }
    s_a[tid] = a; s_b[tid] = c;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (tid < d) { s_a[tid] += s_a[tid + d]; s_b[tid] += s_b[tid + d]; }
        __syncthreads();
    }
    if (tid == 0) { sums[2 * b] = (float)s_a[0]; sums[2 * b + 1] = (float)s_b[0]; }
}

// whole-batch scalar: loss = (1 - l) sum|x-y| / n + l (1 - sum ssim / n), one CTA, double accumulation
__global__ void __launch_bounds__(256)
reduce_loss_kernel(int n_blocks, const float* __restrict__ partials, double inv_numel, float lambda_dssim,
                   float* __restrict__ loss) {
    __shared__ double s_a[256], s_b[256];
    const int tid = threadIdx.x;
    double a = 0.0, c = 0.0;
    for (int i = tid; i < n_blocks; i += 256) { a += partials[2 * i]; c += partials[2 * i + 1]; }
    s_a[tid] = a; s_b[tid] = c;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (tid < d) { s_a[tid] += s_a[tid + d]; s_b[tid] += s_b[tid + d]; }
        __syncthreads();
    }
    if (tid == 0)
        *loss = (float)((1.0 - (double)lambda_dssim) * s_a[0] * inv_numel +
                        (double)lambda_dssim * (1.0 - s_b[0] * inv_numel));
}

// coef: per-image pairs (coef_stride = 2, scales = 1) or one upstream scalar (coef_stride = 0) times host scales
__global__ void __launch_bounds__(SL_THREADS)
l1_dssim_bwd_kernel(int C, int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                    const float* __restrict__ dm, size_t plane_stride_all, const float* __restrict__ coef,
                    int coef_stride, float scale_l1, float scale_ssim, float* __restrict__ dL_dimg) {
    __shared__ float sd[3][SL_E][SL_E + 1];
    __shared__ float sh[3][SL_E][SL_T + 1];
    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * SL_T, y0 = blockIdx.y * SL_T;
    const size_t HW = (size_t)H * W;
    const float cL1 = coef[coef_stride * (n / C)] * scale_l1;
    const float cS = coef[coef_stride * (n / C) + (coef_stride ? 1 : 0)] * scale_ssim;
    {   // all global loads of the thread are issued before the first shared-memory store
        constexpr int kIt = (SL_E * SL_E + SL_THREADS - 1) / SL_THREADS;
        float rv[3][kIt];
#pragma unroll
        for (int it = 0; it < kIt; it++) {
            const int i = tid + it * SL_THREADS;
            const int r = i / SL_E, c = i - r * SL_E;
            const bool ok = i < SL_E * SL_E;
#pragma unroll
            for (int m = 0; m < 3; m++)
                rv[m][it] = ok ? tile_load(dm + m * plane_stride_all + (size_t)n * HW, H, W, y0 + r - SL_R, x0 + c - SL_R)
                               : 0.f;
        }
#pragma unroll
        for (int it = 0; it < kIt; it++) {
            const int i = tid + it * SL_THREADS;
            const int r = i / SL_E, c = i - r * SL_E;
            if (i < SL_E * SL_E) {
#pragma unroll
                for (int m = 0; m < 3; m++) sd[m][r][c] = rv[m][it];
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < SL_E * (SL_T / 4); i += SL_THREADS) {
        const int r = i / (SL_T / 4), c0 = (i - r * (SL_T / 4)) * 4;
#pragma unroll
        for (int m = 0; m < 3; m++) {
            float v[14], o[4];
#pragma unroll
            for (int k = 0; k < 14; k++) v[k] = sd[m][r][c0 + k];
            window4(v, o);
#pragma unroll
            for (int j = 0; j < 4; j++) sh[m][r][c0 + j] = o[j];
        }
    }
    __syncthreads();
    {
        const int c = tid & (SL_T - 1), r0 = (tid / SL_T) * 4;
        float w[3][4];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            float v[14];
#pragma unroll
            for (int k = 0; k < 14; k++) v[k] = sh[m][r0 + k][c];
            window4(v, w[m]);
        }
        const int gx = x0 + c;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gy = y0 + r0 + j;
            if (gy >= H || gx >= W) continue;
            const size_t o = (size_t)n * HW + (size_t)gy * W + gx;
            const float x = img[o], y = gt[o];
            const float d = x - y;
            const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);   // torch.abs: zero gradient at 0
            dL_dimg[o] = cL1 * sgn + cS * (w[0][j] + 2.f * x * w[1][j] + y * w[2][j]);
        }
    }
}

static int tiles(int v) { return (v + SL_T - 1) / SL_T; }

}  // namespace sgs_loss

extern "C" {

size_t sgs_loss_workspace_floats(int B, int C, int H, int W) {
    return (size_t)B * C * sgs_loss::tiles(H) * sgs_loss::tiles(W) * 2;
}

int sgs_l1_dssim_forward(int B, int C, int H, int W, const float* img, const float* gt, float* dmaps, float* workspace,
                         float* sums, void* stream) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || !img || !gt || !workspace || !sums) return SGS_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = B * C;
    if (N > 65535) return SGS_ERR_INVALID_ARGUMENT;
    dim3 grid(sgs_loss::tiles(W), sgs_loss::tiles(H), N);
    sgs_loss::l1_dssim_fwd_kernel<<<grid, SL_THREADS, 0, s>>>(H, W, img, gt, dmaps, (size_t)N * H * W, workspace);
    sgs_loss::reduce_partials_kernel<<<B, 256, 0, s>>>(C * (int)grid.x * (int)grid.y, workspace, sums);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int sgs_l1_dssim_backward(int B, int C, int H, int W, const float* img, const float* gt, const float* dmaps,
                          const float* coef, float* dL_dimg, void* stream) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || !img || !gt || !dmaps || !coef || !dL_dimg) return SGS_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = B * C;
    if (N > 65535) return SGS_ERR_INVALID_ARGUMENT;
    dim3 grid(sgs_loss::tiles(W), sgs_loss::tiles(H), N);
    sgs_loss::l1_dssim_bwd_kernel<<<grid, SL_THREADS, 0, s>>>(C, H, W, img, gt, dmaps, (size_t)N * H * W, coef, 2, 1.f,
                                                             1.f, dL_dimg);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int sgs_l1_dssim_loss_forward(int B, int C, int H, int W, const float* img, const float* gt, float lambda_dssim,
                              float* dmaps, float* workspace, float* loss, void* stream) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || !img || !gt || !workspace || !loss) return SGS_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = B * C;
    if (N > 65535) return SGS_ERR_INVALID_ARGUMENT;
    dim3 grid(sgs_loss::tiles(W), sgs_loss::tiles(H), N);
    sgs_loss::l1_dssim_fwd_kernel<<<grid, SL_THREADS, 0, s>>>(H, W, img, gt, dmaps, (size_t)N * H * W, workspace);
    sgs_loss::reduce_loss_kernel<<<1, 256, 0, s>>>(N * (int)grid.x * (int)grid.y, workspace,
                                                  1.0 / ((double)N * H * W), lambda_dssim, loss);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int sgs_l1_dssim_loss_backward(int B, int C, int H, int W, const float* img, const float* gt, float lambda_dssim,
                               const float* dmaps, const float* grad_loss, float* dL_dimg, void* stream) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || !img || !gt || !dmaps || !grad_loss || !dL_dimg)
        return SGS_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = B * C;
    if (N > 65535) return SGS_ERR_INVALID_ARGUMENT;
    dim3 grid(sgs_loss::tiles(W), sgs_loss::tiles(H), N);
    const double inv = 1.0 / ((double)N * H * W);
    sgs_loss::l1_dssim_bwd_kernel<<<grid, SL_THREADS, 0, s>>>(C, H, W, img, gt, dmaps, (size_t)N * H * W, grad_loss, 0,
                                                             (float)((1.0 - lambda_dssim) * inv),
                                                             (float)(-(double)lambda_dssim * inv), dL_dimg);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

}  // extern "C"

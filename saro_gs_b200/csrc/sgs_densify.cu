// Densification statistics of one training iteration (the part that consumes the
// rasterizer's outputs directly): what the reference does with Python lists and ~12 PyTorch kernels per iteration.
//
// Reference (train.py):
//   :211      batch_point_grad.append(torch.norm(viewspace_point_tensor.grad[:, :2], dim=-1))     per view
//   :214-215  batch_radii.append(radii); batch_visibility_filter.append(radii > 0)                per view
//   :281-287  visibility_count = sum over views; visibility_filter = count > 0; radii = max over views;
//             grad = (sum over views) / visibility_count   where visible
//   :290      max_radii2D[visible] = max(max_radii2D[visible], radii[visible])
//   :291      add_densification_stats_grad (scene/saro_gaussian.py:745-747):
//                 xyz_gradient_accum[visible] += grad[visible];  denom[visible] += 1
//
// Two streaming kernels, one pass over the P Gaussians each (HBM-bound: 20 B read + 12 B updated per Gaussian and view,
// 32 B per Gaussian for the commit).  The per-view kernel keeps running sums instead of a list of per-view tensors, so
// the data-parallel form is three small all-reduces (SUM, SUM, MAX) of the running buffers between the two kernels.
#include "../../include/saro_gs_b200.h"
#include <cuda_runtime.h>
#include <cstdint>

namespace sgs_densify {

__global__ void __launch_bounds__(256) add_view_kernel(int P, const float* __restrict__ dmeans2D, const int* __restrict__ radii,
                                                       float* __restrict__ grad_sum, int* __restrict__ vis_count,
                                                       int* __restrict__ radii_max) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const float gx = dmeans2D[(size_t)i * 3], gy = dmeans2D[(size_t)i * 3 + 1];
        const int r = radii[i];
        grad_sum[i] += sqrtf(gx * gx + gy * gy);          // torch.norm(grad[:, :2], dim=-1), summed over the batch (:284)
        vis_count[i] += r > 0;                              // :281
        radii_max[i] = max(radii_max[i], r);                // :283
    }
}

__global__ void __launch_bounds__(256) commit_kernel(int P, const float* __restrict__ grad_sum, const int* __restrict__ vis_count,
                                                     const int* __restrict__ radii_max, float* __restrict__ max_radii2D,
                                                     float* __restrict__ xyz_gradient_accum, float* __restrict__ denom) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const int n = vis_count[i];
        if (n > 0) {                                                             // visibility_filter (:282)
            max_radii2D[i] = fmaxf(max_radii2D[i], (float)radii_max[i]);         // :290
            xyz_gradient_accum[i] += grad_sum[i] / (float)n;                     // :285, saro_gaussian.py:746
            denom[i] += 1.f;                                                     // saro_gaussian.py:747
        }
    }
}

inline int grid_for(int P) {
    int g = (P + 255) / 256;
    return g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g);
}

}  // namespace sgs_densify

extern "C" {

int sgs_densify_add_view(int P, const float* dL_dmeans2D, const int* radii, float* grad_sum, int* vis_count, int* radii_max,
                         void* stream) {
    if (P < 0) return SGS_ERR_INVALID_ARGUMENT;
    if (P == 0) return 0;
    if (!dL_dmeans2D || !radii || !grad_sum || !vis_count || !radii_max) return SGS_ERR_INVALID_ARGUMENT;
    sgs_densify::add_view_kernel<<<sgs_densify::grid_for(P), 256, 0, (cudaStream_t)stream>>>(P, dL_dmeans2D, radii, grad_sum,
                                                                                         vis_count, radii_max);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int sgs_densify_commit(int P, const float* grad_sum, const int* vis_count, const int* radii_max, float* max_radii2D,
                       float* xyz_gradient_accum, float* denom, void* stream) {
    if (P < 0) return SGS_ERR_INVALID_ARGUMENT;
    if (P == 0) return 0;
    if (!grad_sum || !vis_count || !radii_max || !max_radii2D || !xyz_gradient_accum || !denom) return SGS_ERR_INVALID_ARGUMENT;
    sgs_densify::commit_kernel<<<sgs_densify::grid_for(P), 256, 0, (cudaStream_t)stream>>>(P, grad_sum, vis_count, radii_max,
                                                                                       max_radii2D, xyz_gradient_accum, denom);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

}  // extern "C"

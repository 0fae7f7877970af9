"""Python surface of the rasterizer — mirrors, name for name, the module SaRO-GS imports:

    from diff_gaussian_rasterization_ch3 import GaussianRasterizationSettings, GaussianRasterizer
    (renderer/__init__.py:32 of the reference)

Interface mirrored from $R/diff_gaussian_rasterization_ch3/__init__.py:
    rasterize_gaussians(...)                 :17-38
    _RasterizeGaussians (autograd.Function)  :40-132   (argument reordering, saved tensors, grad order)
    GaussianRasterizationSettings            :134-145  (11 fields, positional order kept)
    GaussianRasterizer(nn.Module)            :147-196  (markVisible, forward, the two exception texts)

The classes are produced by `make_api(backend)` so that the *same* host layer can drive
either this repo's native backend (saro_gs_b200.backend, the product) or — in tests and in
bench.py's reference arm only — the compiled reference `_C` module.
"""
from typing import NamedTuple

import torch
import torch.nn as nn


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool


def make_api(backend, supports_keep_flag=False):
    """Build (rasterize_gaussians, GaussianRasterizer, _RasterizeGaussians) on top of `backend`,
    any object exposing rasterize_gaussians / rasterize_gaussians_backward / mark_visible with
    the reference `_C` signatures."""

    class _RasterizeGaussians(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    raster_settings):
            rs = raster_settings
            args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                    rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh,
                    rs.sh_degree, rs.campos, rs.prefiltered)
            if supports_keep_flag:
                # inference (no input needs a gradient): skip writing the per-tile lists backward would read
                keep = any(ctx.needs_input_grad[:8])
                out = backend.rasterize_gaussians(*args, keep_for_backward=keep)
            else:
                out = backend.rasterize_gaussians(*args)
            num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer, depth = out

            ctx.raster_settings = rs
            ctx.num_rendered = num_rendered
            # radii/depth never receive a gradient: do not let autograd zero-fill [P] + [1,H,W] for them
            ctx.set_materialize_grads(False)
            ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                                  binningBuffer, imgBuffer)
            return color, radii, depth

        @staticmethod
        def backward(ctx, grad_out_color, _grad_radii, _grad_depth):
            # depth and radii are non-differentiable in the reference (grads ignored, __init__.py:88)
            rs = ctx.raster_settings
            (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
             imgBuffer) = ctx.saved_tensors
            if grad_out_color is None:   # only depth was used downstream: depth carries no gradient
                grad_out_color = torch.zeros((3, rs.image_height, rs.image_width), dtype=means3D.dtype,
                                             device=means3D.device)
            args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                    rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, sh, rs.sh_degree, rs.campos,
                    geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer)
            (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
             grad_scales, grad_rotations) = backend.rasterize_gaussians_backward(*args)
            return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                    grad_rotations, grad_cov3Ds_precomp, None)

    def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                            raster_settings):
        return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                         cov3Ds_precomp, raster_settings)

    class GaussianRasterizer(nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            self.raster_settings = raster_settings

        def markVisible(self, positions):
            with torch.no_grad():
                rs = self.raster_settings
                return backend.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            rs = self.raster_settings
            if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
                raise Exception('Please provide excatly one of either SHs or precomputed colors!')
            if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                    ((scales is not None or rotations is not None) and cov3D_precomp is not None):
                raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
            # "not provided" travels as an empty tensor, as in the reference (:173-183)
            if shs is None:
                shs = torch.Tensor([])
            if colors_precomp is None:
                colors_precomp = torch.Tensor([])
            if scales is None:
                scales = torch.Tensor([])
            if rotations is None:
                rotations = torch.Tensor([])
            if cov3D_precomp is None:
                cov3D_precomp = torch.Tensor([])
            return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                       cov3D_precomp, rs)

    return rasterize_gaussians, GaussianRasterizer, _RasterizeGaussians

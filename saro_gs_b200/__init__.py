"""saro_gs_b200 — B200-native (sm_100a) differentiable 3D-Gaussian tile rasterizer.

Drop-in for the hot path of yjb6/SaRO-GS (``submodules/gaussian_rasterization_ch3``):
``GaussianRasterizationSettings`` / ``GaussianRasterizer`` keep the reference's Python
surface (see rasterizer.py) and run on hand-written CUDA kernels behind a C ABI
(include/saro_gs_b200.h).  There is no CPU or PyTorch fallback.
"""
from . import backend as _C  # same three functions as the reference's pybind module `_C`
from .rasterizer import GaussianRasterizationSettings, make_api

rasterize_gaussians, GaussianRasterizer, _RasterizeGaussians = make_api(_C, supports_keep_flag=True)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "_RasterizeGaussians", "_C"]
__version__ = "0.1.0"

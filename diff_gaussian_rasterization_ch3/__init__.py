"""Drop-in module name used by SaRO-GS (renderer/__init__.py:32 of the reference):

    from diff_gaussian_rasterization_ch3 import GaussianRasterizationSettings, GaussianRasterizer

Put this repository's root on PYTHONPATH instead of installing the reference submodule and the
import above resolves to the B200-native rasterizer.
"""
from saro_gs_b200 import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,  # noqa: F401
                          _RasterizeGaussians, _C)

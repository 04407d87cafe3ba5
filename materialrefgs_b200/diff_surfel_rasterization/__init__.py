"""Import-compatible stand-in for the reference package `diff_surfel_rasterization`
(gaussian_renderer/__init__.py:18 does `from diff_surfel_rasterization import
GaussianRasterizationSettings, GaussianRasterizer`)."""
from ..rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,
                          _RasterizeGaussians)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]

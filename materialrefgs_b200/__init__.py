"""B200-native (sm_100a) drop-in for MaterialRefGS's surfel-splatting render path.

Public surface mirrors the reference:
  materialrefgs_b200.diff_surfel_rasterization  -> GaussianRasterizationSettings, GaussianRasterizer
  materialrefgs_b200.shading                    -> EnvLight, get_specular_color_surfel, shade_surfel
Everything computes through libmrgs.so (include/mrgs.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

"""The CPU oracle rasterizer (oracle/surfel_oracle.cpp) behind the reference's Python API
(rast/diff_surfel_rasterization/__init__.py:167-235): GaussianRasterizationSettings / GaussianRasterizer with autograd.
TEST INFRASTRUCTURE ONLY (only tests/, tests/golden/make_golden_*.py, __graft_entry__.smoke() and bench.py's baseline
legs import it).

Why it exists: with this module standing in for `diff_surfel_rasterization`, the reference's OWN render functions
(gaussian_renderer/__init__.py: render_initial, render_surfel, render_volume) run on the CPU of the build container
(tests/golden/make_golden_render.py) and produce the vectors that pin oracle/render_oracle.py. The C++ oracle itself is
pinned bit-for-bit / to 1e-5 by vectors of the reference CUDA extension (tests/test_oracle_cpu.py).
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np
import torch

from . import surfel_oracle as so


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
    debug: bool


def _np(t):
    return None if t is None or t.numel() == 0 else t.detach().cpu().numpy()


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, features, opacities, scales, rotations, rs):
        S = 0 if features is None else features.shape[1]
        o = so.OracleRaster(
            means3D=_np(means3D), opacities=_np(opacities), viewmatrix=_np(rs.viewmatrix), projmatrix=_np(rs.projmatrix),
            campos=_np(rs.campos), W=rs.image_width, H=rs.image_height, tan_fovx=rs.tanfovx, tan_fovy=rs.tanfovy,
            background=_np(rs.bg), shs=_np(sh), colors_precomp=_np(colors_precomp),
            features=None if S == 0 else _np(features), scales=_np(scales), rotations=_np(rotations),
            sh_degree=rs.sh_degree, scale_modifier=rs.scale_modifier)
        o.preprocess()
        o.bin()
        color, feature, others = o.forward()
        ctx.oracle = o
        ctx.has = (sh is not None and sh.numel() > 0, colors_precomp is not None and colors_precomp.numel() > 0, S)
        radii = torch.from_numpy(o.geom["radii"].copy())
        ctx.mark_non_differentiable(radii)
        return (torch.zeros((1, rs.image_height, rs.image_width), dtype=torch.int32), torch.from_numpy(color.copy()),
                torch.from_numpy(feature.copy()), radii, torch.from_numpy(others.copy()))

    @staticmethod
    def backward(ctx, g_contrib, g_color, g_feature, g_radii, g_others):
        o = ctx.oracle
        H, W, S = o.H, o.W, o.S
        z = lambda g, c: np.zeros((c, H, W), np.float32) if g is None else g.detach().numpy()
        g = o.backward(z(g_color, 3), z(g_feature, S), z(g_others, 7))
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        has_sh, has_col, S = ctx.has
        return (t(g["dL_dmeans3D"]), t(g["dL_dmeans2D"]), t(g["dL_dsh"]) if has_sh else None,
                t(g["dL_dcolors"]) if has_col else None, t(g["dL_dfeatures"]) if S else None, t(g["dL_dopacity"]),
                t(g["dL_dscales"]), t(g["dL_drotations"]), None)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, features=None, scales=None,
                rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if cov3D_precomp is not None:
            raise NotImplementedError("the torch front end of the CPU oracle takes scales + rotations only")
        if scales is None or rotations is None:
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        return _Rasterize.apply(means3D, means2D, shs, colors_precomp, features, opacities, scales, rotations,
                                self.raster_settings)

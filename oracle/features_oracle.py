"""Torch restatement of the reference's per-surfel feature preparation.
TEST INFRASTRUCTURE ONLY (only tests/, __graft_entry__.smoke() and bench.py's baseline legs import it).

Pinned by tests/golden/features_*.npz: outputs AND gradients produced by the reference's OWN functions
(utils/sh_utils.py eval_sh, utils/general_utils.py build_scaling_rotation / flip_align_view /
safe_normalize, imported from /root/reference by tests/golden/make_golden_features.py) composed the way
render_surfel composes them; tests/test_features_cpu.py checks this file against those vectors.

What it follows:
  activations                      scene/gaussian_model.py:56-76, :236-266
  get_indirect (cat dc, rest)      scene/gaussian_model.py:299-303
  get_covariance / get_normal      scene/gaussian_model.py:48-54, :269-285, :346-347
  build_rotation                   utils/general_utils.py:78-99
  flip_align_view, safe_normalize  utils/general_utils.py:179-190
  eval_sh                          utils/sh_utils.py:57-112
  composition                      gaussian_renderer/__init__.py:259-266, :334-353
Differentiable torch: gradients come from autograd.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)

RAW_FIELDS = (("xyz", 3), ("scaling", 2), ("rotation", 4), ("opacity", 1), ("refl_strength", 1), ("roughness", 1),
              ("ori_color", 3), ("indirect_dc", 3), ("indirect_rest", 45))


def eval_sh3(sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """sh [P,3,16], dirs [P,3] -> [P,3]; degree 3, the reference's term order (sh_utils.py:79-104)."""
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    r = C0 * sh[..., 0]
    r = r - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    r = (r + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] +
         C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
    r = (r + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10] +
         C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
         C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14] +
         C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return r


def surfel_normal_raw(rotation: torch.Tensor) -> torch.Tensor:
    """splat2world[:, 2, :3] = third column of R(q / |q|) (the unit third scale leaves it unscaled)."""
    q = rotation / torch.sqrt((rotation * rotation).sum(-1))[:, None]
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack((2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y)), dim=-1)


def prepare_features(xyz, scaling, rotation, opacity, refl_strength, roughness, ori_color, indirect_dc, indirect_rest,
                     campos):
    """Raw parameters -> what render_surfel hands to the rasterizer:
    scales [P,2], rotations [P,4], opacities [P,1], features [P,8] = (refl, roughness, ori_color, indirect)."""
    scales = torch.exp(scaling)
    rotations = F.normalize(rotation)
    opacities = torch.sigmoid(opacity)
    dir_pp = xyz - campos
    dirn = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    nraw = surfel_normal_raw(rotation)
    non_flip = (nraw * -dirn).sum(-1, keepdim=True) >= 0
    nraw = nraw * torch.where(non_flip, 1.0, -1.0)
    normals = nraw / torch.clamp(torch.linalg.norm(nraw, dim=-1, keepdim=True), min=1e-20)
    w_o = -dirn
    reflection = 2 * (normals * w_o).sum(1, keepdim=True) * normals - w_o
    shs = torch.cat((indirect_dc.reshape(-1, 1, 3), indirect_rest.reshape(-1, 15, 3)), dim=1).transpose(1, 2)
    indirect = torch.clamp_min(eval_sh3(shs, reflection), 0.0)
    features = torch.cat((torch.sigmoid(refl_strength), torch.sigmoid(roughness), torch.sigmoid(ori_color), indirect), -1)
    return scales, rotations, opacities, features


def synthetic_params(P: int, seed: int = 0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, w in RAW_FIELDS:
        t = torch.randn(P, w, generator=g, dtype=torch.float64)
        if name == "scaling":
            t = t * 0.5 - 3.0
        if name == "indirect_rest":
            t = t * 0.3
        out[name] = t.to(dtype)
    campos = torch.tensor([0.3, -2.5, 1.1], dtype=dtype)
    return out, campos

"""Torch restatement of the reference's render() contract — render_initial, render_surfel, render_volume of
gaussian_renderer/__init__.py (FLAG "2dgs", SH indirect light, no tracer) — on top of ANY module with the reference's
rasterizer API (the reference CUDA extension in oracle/_ref on a GPU, oracle/raster_torch.py on the CPU) and the shading
restatement of oracle/shading_oracle.py. TEST INFRASTRUCTURE ONLY.

Pinned by tests/golden/render_*.npz: every map of the result dictionaries and the gradients of all leaves, produced by the
reference's OWN three functions run on the CPU with oracle/raster_torch.py as `diff_surfel_rasterization` and the
oracle's texel fetch as `dr.texture` (tests/golden/make_golden_render.py; tests/test_render_oracle_cpu.py).

What it follows: gaussian_renderer/__init__.py:42-90 (compute_2dgs_normal_and_regularizations), :94-222
(render_initial), :225-475 (render_surfel), :521-745 (render_volume); utils/sh_utils.py:57-112 (eval_sh).
"""
from __future__ import annotations

import math

import torch

from . import features_oracle as fo
from . import shading_oracle as so



class RawSurfelModel:
    """The getters the render functions use from scene/gaussian_model.py:236-303, over RAW parameters
    (keys of features_oracle.RAW_FIELDS) + SH colour coefficients [P,16,3] + the two environment lights."""
    active_sh_degree = 3
    max_sh_degree = 3
    ray_tracer = None

    def __init__(self, raw: dict, shs, env, env2=None):
        self.raw, self.shs, self.env, self.env2 = raw, shs, env, env2 if env2 is not None else env

    get_xyz = property(lambda s: s.raw["xyz"])
    get_scaling = property(lambda s: torch.exp(s.raw["scaling"]))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s.raw["rotation"]))
    get_opacity = property(lambda s: torch.sigmoid(s.raw["opacity"]))
    get_refl = property(lambda s: torch.sigmoid(s.raw["refl_strength"]))
    get_rough = property(lambda s: torch.sigmoid(s.raw["roughness"]))
    get_ori_color = property(lambda s: torch.sigmoid(s.raw["ori_color"]))
    get_features = property(lambda s: s.shs)
    get_indirect = property(lambda s: torch.cat((s.raw["indirect_dc"].reshape(-1, 1, 3),
                                                 s.raw["indirect_rest"].reshape(-1, 15, 3)), dim=1))
    get_envmap = property(lambda s: s.env)
    get_envmap_2 = property(lambda s: s.env2)

    def get_normal(self, scaling_modifier, dir_pp_normalized):   # gaussian_model.py:269-285 (return_delta False)
        nraw = fo.surfel_normal_raw(self.raw["rotation"])
        non_flip = (nraw * -dir_pp_normalized).sum(-1, keepdim=True) >= 0
        return so.safe_normalize(nraw * torch.where(non_flip, 1.0, -1.0))


eval_sh3 = fo.eval_sh3


def _rasterizer(raster, cam, pc, bg):
    rs = raster.GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=math.tan(cam.FoVx * 0.5),
        tanfovy=math.tan(cam.FoVy * 0.5), bg=torch.zeros_like(bg), scale_modifier=1.0, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=pc.active_sh_degree, campos=cam.camera_center, prefiltered=False,
        debug=False)
    return raster.GaussianRasterizer(raster_settings=rs)


def regularizations(allmap, cam, depth_ratio):
    """compute_2dgs_normal_and_regularizations (:42-90)."""
    out = {"rend_alpha": allmap[1:2], "rend_dist": allmap[6:7],
           "rend_normal": (allmap[2:5].permute(1, 2, 0) @ cam.world_view_transform[:3, :3].T).permute(2, 0, 1)}
    out["surf_depth"], out["surf_normal"] = so.surf_depth_normal(allmap, cam, depth_ratio)
    return out


def _indirect(pc, cam):
    d = pc.get_xyz - cam.camera_center
    d = d / d.norm(dim=1, keepdim=True)
    n = pc.get_normal(1.0, d)
    w_o = -d
    refl_dir = 2 * torch.sum(n * w_o, dim=1, keepdim=True) * n - w_o
    ind = torch.clamp_min(eval_sh3(pc.get_indirect.transpose(1, 2).reshape(-1, 3, 16), refl_dir), 0.0)
    return n, ind


def render_initial(raster, cam, pc, pipe, bg, srgb=False):
    """:94-222."""
    m2d = torch.zeros_like(pc.get_xyz, requires_grad=True)
    _, color, _, radii, allmap = _rasterizer(raster, cam, pc, bg)(
        means3D=pc.get_xyz, means2D=m2d, opacities=pc.get_opacity, shs=pc.get_features,
        features=torch.empty((pc.get_xyz.shape[0], 0)), scales=pc.get_scaling, rotations=pc.get_rotation)
    out = regularizations(allmap, cam, pipe.depth_ratio)
    if srgb:
        color = so.linear_to_srgb(color)
    out.update({"render": color + bg[:, None, None] * (1 - out["rend_alpha"]), "radii": radii, "viewspace_points": m2d})
    return out


def render_surfel(raster, cam, pc, pipe, bg, srgb=False):
    """:225-475 (opt.indirect False)."""
    _, ind = _indirect(pc, cam)
    feats = torch.cat((pc.get_refl, pc.get_rough, pc.get_ori_color, ind), -1)
    m2d = torch.zeros_like(pc.get_xyz, requires_grad=True)
    _, color, feat, radii, allmap = _rasterizer(raster, cam, pc, bg)(
        means3D=pc.get_xyz, means2D=m2d, opacities=pc.get_opacity, shs=pc.get_features, features=feats,
        scales=pc.get_scaling, rotations=pc.get_rotation)
    out = so.shade_surfel(pc.get_envmap, so.load_lut(color.device), color, feat, allmap, cam, bg, srgb=srgb)
    if srgb:
        out["base_color_map"] = so.linear_to_srgb(out["base_color_map"])
        out["specular_map"] = so.linear_to_srgb(out["specular_map"])
    out.update(regularizations(allmap, cam, pipe.depth_ratio))
    out.update({"radii": radii, "diffuse_map_ori": color, "viewspace_points": m2d})
    return out


def render_volume(raster, cam, pc, pipe, bg, srgb=False, indirect=False):
    """:521-745 (visibility = 1 when opt.indirect)."""
    n, ind = _indirect(pc, cam)
    lut = so.load_lut(pc.get_xyz.device)
    diffuse, specular = so.get_full_color_volume(pc.get_envmap_2, lut, pc.get_xyz, pc.get_ori_color, cam, n.contiguous(),
                                                 pc.get_refl, pc.get_rough)
    feats = [pc.get_rough, pc.get_refl, diffuse, specular, pc.get_ori_color]
    if indirect:   # refl_utils.py:450-490 without a tracer: specular_light = direct_light
        _, rays_o = so.sample_camera_rays(cam.HWK, cam.R, cam.T, pc.get_xyz.device)
        w_o = so.safe_normalize(rays_o.expand(n.shape[0], -1) - pc.get_xyz)
        rr = so.safe_normalize(2 * n * torch.sum(w_o * n, -1, keepdim=True) - w_o)
        feats += [torch.ones_like(pc.get_opacity), ind, pc.get_envmap_2(rr, roughness=pc.get_rough)]
    m2d = torch.zeros_like(pc.get_xyz, requires_grad=True)
    _, color, feat, radii, allmap = _rasterizer(raster, cam, pc, bg)(
        means3D=pc.get_xyz, means2D=m2d, opacities=pc.get_opacity, colors_precomp=specular + diffuse,
        features=torch.cat(feats, -1), scales=pc.get_scaling, rotations=pc.get_rotation)
    out = regularizations(allmap, cam, pipe.depth_ratio)
    full, dmap, smap = color, feat[2:5], feat[5:8]
    if srgb:
        full, dmap, smap = so.linear_to_srgb(full), so.linear_to_srgb(dmap), so.linear_to_srgb(smap)
    out.update({"render": full + bg[:, None, None] * (1 - out["rend_alpha"]), "roughness_map": feat[:1],
                "refl_strength_map": feat[1:2], "diffuse_map": dmap, "specular_map": smap, "base_color_map": feat[8:11],
                "radii": radii, "viewspace_points": m2d})
    if indirect:
        out.update({"visibility": feat[11:12], "indirect_light": feat[12:15], "direct_light": feat[15:18]})
    return out

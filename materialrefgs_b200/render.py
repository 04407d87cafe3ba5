"""The render() contract of gaussian_renderer/__init__.py — render_initial (:94-222), render_surfel (:225-469) and
render_volume (:521-745) with the reference's signatures and return dicts — running the rasterizer, the shading
(deferred per pixel for render_surfel, per surfel for render_volume) and the depth->normal regularisers on libmrgs kernels.

`pc` is duck-typed like scene/gaussian_model.py's GaussianModel (get_xyz, get_opacity, get_refl,
get_ori_color, get_rough, get_scaling, get_rotation, get_features, get_indirect, get_normal(), get_envmap,
active_sh_degree, max_sh_degree); `viewpoint_camera` like scene/cameras.py's Camera; `pipe` needs `debug`
and `depth_ratio`. When `pc` exposes the raw parameters (`_scaling`, `_rotation`, ... like GaussianModel) the
per-surfel feature preparation (:259-353) runs as one libmrgs kernel pair (features.py).
Unsupported reference branches raise: pipe.compute_cov3D_python, pipe.use_asg, opt.indirect (OptiX/mesh
visibility tracing stays on the reference).
"""
from __future__ import annotations

import math

import torch

from .diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from .features import MODEL_RAW_ATTRS, surfel_features_from_model
from .shading import (get_full_color_volume, get_full_color_volume_indirect, linear_to_srgb, shade_surfel,
                      surf_depth_normal)

_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def sh_basis(deg: int, dirs: torch.Tensor) -> torch.Tensor:
    """Real SH basis values [..., (deg+1)^2] with the constants/sign convention of utils/sh_utils.py:26-45."""
    x, y, z = dirs[..., 0], dirs[..., 1], dirs[..., 2]
    b = [torch.full_like(x, _C0)]
    if deg > 0:
        b += [-_C1 * y, _C1 * z, -_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [_C2[0] * xy, _C2[1] * yz, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy)]
        if deg > 2:
            b += [_C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
                  _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy), _C3[5] * z * (xx - yy),
                  _C3[6] * x * (xx - 3 * yy)]
    return torch.stack(b, dim=-1)


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """sh [..., C, >= (deg+1)^2], dirs [..., 3] -> [..., C] (utils/sh_utils.py:57-112)."""
    n = (deg + 1) ** 2
    return (sh[..., :n] * sh_basis(deg, dirs)[..., None, :]).sum(-1)


def render_surfel(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
                  srgb=False, opt=None, wo_render_img=False, normal_img_map=None):
    if getattr(pipe, "compute_cov3D_python", False) or getattr(pipe, "use_asg", False):
        raise NotImplementedError("compute_cov3D_python / use_asg paths stay on the reference")
    if opt is not None and getattr(opt, "indirect", False):
        raise NotImplementedError("opt.indirect needs the OptiX/mesh tracer, which stays on the reference")

    means3D = pc.get_xyz
    screenspace_points = torch.zeros_like(means3D, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    imH, imW = int(viewpoint_camera.image_height), int(viewpoint_camera.image_width)
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=imH, image_width=imW, tanfovx=tanfovx, tanfovy=tanfovy, bg=torch.zeros_like(bg_color),
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=getattr(pipe, "debug", False))
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    shs, colors_precomp = (pc.get_features, None) if override_color is None else (None, override_color)

    if all(hasattr(pc, n) for n in MODEL_RAW_ATTRS) and pc.max_sh_degree == 3:
        # a GaussianModel: activations, normal, reflection, indirect SH and the cat as one kernel pair
        scales, rotations, opacity, features = surfel_features_from_model(pc, viewpoint_camera.camera_center)
    else:
        # duck-typed model exposing only the activated getters: the reference's torch lines (:334-353)
        dir_pp = means3D - viewpoint_camera.camera_center
        dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
        normals = pc.get_normal(scaling_modifier, dir_pp_normalized)
        w_o = -dir_pp_normalized
        reflection = 2 * torch.sum(normals * w_o, dim=1, keepdim=True) * normals - w_o
        shs_indirect = pc.get_indirect.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
        indirect = torch.clamp_min(eval_sh(3, shs_indirect, reflection), 0.0)
        features = torch.cat((pc.get_refl, pc.get_rough, pc.get_ori_color, indirect), dim=-1)
        scales, rotations, opacity = pc.get_scaling, pc.get_rotation, pc.get_opacity

    contrib, rendered_image, rendered_features, radii, allmap = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp, features=features,
        opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=None)

    if wo_render_img:
        # the reference skips depth_to_normal here (return_depth_normal=False, gaussian_renderer/__init__.py:388,
        # :71-77): surf_normal is None and only the depth mix of :50-70 is evaluated
        ratio = getattr(pipe, "depth_ratio", 0.0)
        surf_depth = torch.nan_to_num(allmap[0:1] / allmap[1:2], 0, 0) * (1 - ratio) + ratio * torch.nan_to_num(allmap[5:6], 0, 0)
        surf_normal = None
        w2v_rot = viewpoint_camera.world_view_transform[:3, :3]
        render_normal = (allmap[2:5].permute(1, 2, 0) @ w2v_rot.T).permute(2, 0, 1)
        return {"refl_strength_map": rendered_features[:1], "base_color_map": rendered_features[2:5],
                "roughness_map": rendered_features[1:2], "viewspace_points": screenspace_points,
                "visibility_filter": radii > 0, "radii": radii, "rend_alpha": allmap[1:2],
                "rend_normal": render_normal, "rend_dist": allmap[6:7], "surf_depth": surf_depth,
                "surf_normal": surf_normal}

    surf_depth, surf_normal = surf_depth_normal(allmap, imH, imW, tanfovx, tanfovy, viewpoint_camera.R,
                                                viewpoint_camera.T, getattr(pipe, "depth_ratio", 0.0))
    maps = shade_surfel(pc.get_envmap, rendered_image, rendered_features, allmap, viewpoint_camera.HWK,
                        viewpoint_camera.R, bg_color, srgb=srgb)
    albedo, specular = maps["base_color_map"], maps["specular_map"]
    if srgb:
        from .shading import linear_to_srgb
        albedo, specular = linear_to_srgb(albedo), linear_to_srgb(specular)
    return {"render": maps["render"], "refl_strength_map": maps["refl_strength_map"],
            "diffuse_map": maps["diffuse_map"], "diffuse_map_ori": rendered_image, "specular_map": specular,
            "base_color_map": albedo, "roughness_map": maps["roughness_map"],
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "rend_alpha": maps["rend_alpha"], "rend_normal": maps["rend_normal"], "rend_dist": maps["rend_dist"],
            "surf_depth": surf_depth, "surf_normal": surf_normal, "direct_light": maps["direct_light"]}


def _raster_setup(viewpoint_camera, pc, pipe, bg_color, scaling_modifier):
    if getattr(pipe, "compute_cov3D_python", False):
        raise NotImplementedError("pipe.compute_cov3D_python stays on the reference")
    means3D = pc.get_xyz
    screenspace_points = torch.zeros_like(means3D, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    imH, imW = int(viewpoint_camera.image_height), int(viewpoint_camera.image_width)
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=imH, image_width=imW, tanfovx=tanfovx, tanfovy=tanfovy, bg=torch.zeros_like(bg_color),
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=getattr(pipe, "debug", False))
    return GaussianRasterizer(raster_settings=raster_settings), screenspace_points, (imH, imW, tanfovx, tanfovy)


def render_initial(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
                   srgb=False, opt=None):
    """gaussian_renderer/__init__.py:94-222 (FLAG "2dgs"): plain SH-colour surfel splatting, S = 0 feature channels."""
    rasterizer, screenspace_points, (imH, imW, tanfovx, tanfovy) = _raster_setup(viewpoint_camera, pc, pipe, bg_color,
                                                                               scaling_modifier)
    means3D = pc.get_xyz
    shs, colors_precomp = (pc.get_features, None) if override_color is None else (None, override_color)
    features = torch.empty((means3D.shape[0], 0))     # a CPU tensor, as in the reference (:174)
    contrib, rendered_image, rendered_features, radii, allmap = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp, features=features,
        opacities=pc.get_opacity, scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    surf_depth, surf_normal = surf_depth_normal(allmap, imH, imW, tanfovx, tanfovy, viewpoint_camera.R,
                                                viewpoint_camera.T, getattr(pipe, "depth_ratio", 0.0))
    render_alpha = allmap[1:2]
    w2v_rot = viewpoint_camera.world_view_transform[:3, :3]
    render_normal = (allmap[2:5].permute(1, 2, 0) @ w2v_rot.T).permute(2, 0, 1)
    if srgb:
        rendered_image = linear_to_srgb(rendered_image)
    final_image = rendered_image + bg_color[:, None, None] * (1 - render_alpha)
    return {"render": final_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "rend_alpha": render_alpha, "rend_normal": render_normal, "rend_dist": allmap[6:7], "surf_depth": surf_depth,
            "surf_normal": surf_normal}


def render_volume(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
                  srgb=False, opt=None):
    """gaussian_renderer/__init__.py:521-745 (FLAG "2dgs", SH indirect light): every surfel is shaded on its own
    (diffuse + specular split-sum colours from the two EnvLight queries), the shaded colours are splatted as
    colors_precomp and the material channels ride along as S = 11 (or 18 with opt.indirect) features."""
    if getattr(pipe, "use_asg", False):
        raise NotImplementedError("pipe.use_asg (anisotropic spherical Gaussians) stays on the reference")
    rasterizer, screenspace_points, (imH, imW, tanfovx, tanfovy) = _raster_setup(viewpoint_camera, pc, pipe, bg_color,
                                                                               scaling_modifier)
    means3D, opacity = pc.get_xyz, pc.get_opacity
    refl, ori_color, roughness = pc.get_refl, pc.get_ori_color, pc.get_rough
    dir_pp = means3D - viewpoint_camera.camera_center
    dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    normals = pc.get_normal(scaling_modifier, dir_pp_normalized)
    w_o = -dir_pp_normalized
    reflection = 2 * torch.sum(normals * w_o, dim=1, keepdim=True) * normals - w_o
    shs_indirect = pc.get_indirect.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
    indirect = torch.clamp_min(eval_sh(3, shs_indirect, reflection), 0.0)

    use_indirect = opt is not None and getattr(opt, "indirect", False)
    cam = viewpoint_camera
    if use_indirect:
        diffuse, specular, extra = get_full_color_volume_indirect(
            pc.get_envmap_2, means3D, ori_color, cam.HWK, cam.R, cam.T, normals.contiguous(), opacity,
            refl_strength=refl, roughness=roughness, pc=pc, indirect_light=indirect)
        features = torch.cat((roughness, refl, diffuse, specular, ori_color, extra["visibility"], indirect,
                              extra["direct_light"]), dim=-1)
    else:
        diffuse, specular = get_full_color_volume(pc.get_envmap_2, means3D, ori_color, cam.HWK, cam.R, cam.T,
                                                  normals.contiguous(), opacity, refl_strength=refl, roughness=roughness)
        features = torch.cat((roughness, refl, diffuse, specular, ori_color), dim=-1)
    colors_precomp = specular + diffuse

    contrib, rendered_image, rendered_features, radii, allmap = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=None, colors_precomp=colors_precomp, features=features,
        opacities=opacity, scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)

    full_color = rendered_image
    render_diffuse_color, render_specular_color = rendered_features[2:5], rendered_features[5:8]
    surf_depth, surf_normal = surf_depth_normal(allmap, imH, imW, tanfovx, tanfovy, cam.R, cam.T,
                                                getattr(pipe, "depth_ratio", 0.0))
    render_alpha = allmap[1:2]
    w2v_rot = cam.world_view_transform[:3, :3]
    render_normal = (allmap[2:5].permute(1, 2, 0) @ w2v_rot.T).permute(2, 0, 1)
    if srgb:
        render_diffuse_color = linear_to_srgb(render_diffuse_color)
        render_specular_color = linear_to_srgb(render_specular_color)
        full_color = linear_to_srgb(full_color)
    final_image = full_color + bg_color[:, None, None] * (1 - render_alpha)
    results = {"render": final_image, "refl_strength_map": rendered_features[1:2], "diffuse_map": render_diffuse_color,
               "specular_map": render_specular_color, "base_color_map": rendered_features[8:11],
               "roughness_map": rendered_features[:1], "viewspace_points": screenspace_points,
               "visibility_filter": radii > 0, "radii": radii, "rend_alpha": render_alpha, "rend_normal": render_normal,
               "rend_dist": allmap[6:7], "surf_depth": surf_depth, "surf_normal": surf_normal}
    if use_indirect:
        results.update({"visibility": rendered_features[11:12], "indirect_light": rendered_features[12:15],
                        "direct_light": rendered_features[15:18]})
    return results

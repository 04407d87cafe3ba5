"""Torch restatement of the reference's deferred split-sum shading. TEST INFRASTRUCTURE ONLY
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import it).

PARITY UNPINNED FOR THE TEXEL FETCH ONLY: the reference's shading runs through nvdiffrast's `dr.texture`, which
is neither vendored under /root/reference (requirements.txt:57 is a local file:// path without a version) nor
installed here, and the reference has no test or golden image for this path. Everything AROUND the
fetch is pinned: tests/golden/shading_*.npz hold outputs and gradients of the reference's OWN
get_specular_color_surfel / get_full_color_volume / sample_camera_rays / reflection / EnvLight.__call__ / get_mip,
run on the CPU with `dr.texture` supplied by lut_fetch / cube_texture below (tests/golden/make_golden_shading.py), and
tests/golden/depth_normal_*.npz those of utils/point_utils.py (make_golden_depth_normal.py); tests/test_shading_oracle_cpu.py
holds this file to them at 1e-6. This file restates (a) the reference's own torch code line by line and (b) nvdiffrast's
published texture semantics from memory of its texture.cu: texel centres at (i+0.5)/size,
`u*size-0.5` taps, clamp-to-edge for 2-D "clamp" mode, seamless cube faces with the missing corner
tap replaced by the mean of the other three, mip level = clamp(bias,0,L-1) blended linearly.
What it follows:
  sample_camera_rays / reflection          utils/refl_utils.py:54-73, :95-98
  get_specular_color_surfel (visibility=1) utils/refl_utils.py:364-419
  get_full_color_volume                    utils/refl_utils.py:426-447
  EnvLight.get_mip / __call__              scene/light.py:88-129
  render_surfel compositing                gaussian_renderer/__init__.py:372-376, :419-420, :433-445
  compute_2dgs_normal_and_regularizations  gaussian_renderer/__init__.py:42-48 (normal to world)
  linear_to_srgb                           utils/graphics_utils.py:102-110
  cubemap_mip forward                      scene/light_utils.py:66-70
Everything is differentiable torch, so gradients come from autograd.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

LUT_PATH = Path(__file__).resolve().parent.parent / "materialrefgs_b200" / "assets" / "bsdf_256_256.bin"


def load_lut(device="cpu") -> torch.Tensor:
    return torch.from_numpy(np.fromfile(LUT_PATH, dtype=np.float32).reshape(1, 256, 256, 2)).to(device)


def safe_normalize(x, eps=1e-20):  # utils/general_utils.py:179-182
    return x / torch.clamp(torch.linalg.norm(x, dim=-1, keepdim=True), min=eps)


def sample_camera_rays(HWK, R, T, device):
    """utils/refl_utils.py:54-73 (R is the camera's c2w rotation as stored by 3DGS, T its w2c translation)."""
    H, W, K = HWK
    R = torch.as_tensor(R, dtype=torch.float32, device=device).T
    T = torch.as_tensor(T, dtype=torch.float32, device=device)
    K = np.asarray(K).astype(np.float32)
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    xy1 = np.stack([i, j, np.ones_like(i)], axis=2)
    pixel_camera = torch.tensor(np.dot(xy1, np.linalg.inv(K).T), device=device)
    rays_o = (-R.T @ T.unsqueeze(-1)).flatten()
    pixel_world = (pixel_camera - T[None, None]).reshape(-1, 3) @ R
    rays_d = pixel_world - rays_o[None]
    rays_d = rays_d / torch.norm(rays_d, dim=1, keepdim=True)
    return rays_d.reshape(H, W, 3), rays_o


def lut_fetch(lut, uv):
    """dr.texture(FG_LUT, uv, filter_mode='linear', boundary_mode='clamp'); uv [N,2] in [0,1]."""
    grid = (uv * 2.0 - 1.0).reshape(1, -1, 1, 2)
    out = F.grid_sample(lut.permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="border",
                        align_corners=False)
    return out[0, :, :, 0].T  # [N,2]


# ---- cube map fetch ------------------------------------------------------------------------------
def _cube_index(d):
    ax, ay, az = d.abs().unbind(-1)
    is_z = az > torch.maximum(ax, ay)
    is_y = (~is_z) & (ay > ax)
    x, y, z = d.unbind(-1)
    c = torch.where(is_z, z, torch.where(is_y, y, x))
    a = torch.where(is_z | is_y, x, z)
    b = torch.where(is_y, z, y)
    face = torch.where(is_z, 4, torch.where(is_y, 2, 0)) + (c < 0).long()
    m = 0.5 / c.abs()
    su = torch.where((face == 0) | (face == 5), -1.0, 1.0)
    sv = torch.where(face == 2, 1.0, -1.0)
    u = (su * a * m + 0.5).clamp(0.0, 1.0)
    v = (sv * b * m + 0.5).clamp(0.0, 1.0)
    return face, u, v


def _face_to_dir(face, fx, fy):  # scene/light_utils.py:24-31 (cube_to_dir), unnormalised
    one = torch.ones_like(fx)
    cands = torch.stack([
        torch.stack([one, -fy, -fx], -1), torch.stack([-one, -fy, fx], -1),
        torch.stack([fx, one, fy], -1), torch.stack([fx, -one, -fy], -1),
        torch.stack([fx, -fy, one], -1), torch.stack([-fx, -fy, -one], -1)], 0)  # [6,N,3]
    return cands[face, torch.arange(face.shape[0], device=face.device)]


def _bilinear_cube(tex, face, u, v):
    """tex [6,res,res,C]; face [N] long; u,v [N] in [0,1] (differentiable). Returns [N,C]."""
    res = tex.shape[1]
    U = u * res - 0.5
    V = v * res - 0.5
    iu0 = torch.floor(U.detach()).long()
    iv0 = torch.floor(V.detach()).long()
    fu = U - iu0
    fv = V - iv0
    vals, valid = [], []
    for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)):
        ix, iy = iu0 + dx, iv0 + dy
        ox = (ix < 0) | (ix >= res)
        oy = (iy < 0) | (iy >= res)
        corner = ox & oy
        edge = ox ^ oy
        f2, x2, y2 = face.clone(), ix.clone(), iy.clone()
        if edge.any():
            fx = 2.0 * (ix[edge].float() + 0.5) / res - 1.0
            fy = 2.0 * (iy[edge].float() + 0.5) / res - 1.0
            nf, nu, nv = _cube_index(_face_to_dir(face[edge], fx, fy))
            f2[edge] = nf
            x2[edge] = torch.clamp((nu * res).long(), max=res - 1)
            y2[edge] = torch.clamp((nv * res).long(), max=res - 1)
        x2 = x2.clamp(0, res - 1)
        y2 = y2.clamp(0, res - 1)
        t = tex[f2, y2, x2]
        vals.append(torch.where(corner[:, None], torch.zeros_like(t), t))
        valid.append(~corner)
    any_corner = ~(valid[0] & valid[1] & valid[2] & valid[3])
    avg = (vals[0] + vals[1] + vals[2] + vals[3]) * 0.33333333
    vals = [torch.where((any_corner & ~ok)[:, None], avg, t) for t, ok in zip(vals, valid)]
    w = [(1 - fu) * (1 - fv), fu * (1 - fv), (1 - fu) * fv, fu * fv]
    return sum(wk[:, None] * tk for wk, tk in zip(w, vals))


def cube_texture(levels, dirs, mip_level=None):
    """dr.texture(levels[0], dirs, mip=levels[1:], mip_level_bias=mip_level,
    filter_mode='linear-mipmap-linear', boundary_mode='cube'); dirs [N,3]."""
    face, u, v = _cube_index(dirs)
    if mip_level is None:
        return _bilinear_cube(levels[0], face, u, v)
    L = len(levels)
    lvl = mip_level.clamp(0.0, float(L - 1))
    l0 = torch.floor(lvl.detach()).long()
    l1 = torch.clamp(l0 + 1, max=L - 1)
    f = torch.where(l1 == l0, torch.zeros_like(lvl), lvl - l0)
    out = torch.zeros(dirs.shape[0], levels[0].shape[-1], dtype=dirs.dtype, device=dirs.device)
    for l in range(L):
        m0 = l0 == l
        if m0.any():
            out[m0] = out[m0] + (1 - f[m0])[:, None] * _bilinear_cube(levels[l], face[m0], u[m0], v[m0])
        m1 = (l1 == l) & (l1 != l0)
        if m1.any():
            out[m1] = out[m1] + f[m1][:, None] * _bilinear_cube(levels[l], face[m1], u[m1], v[m1])
    return out


def cubemap_mip(cubemap):  # scene/light_utils.py:66-70 (forward)
    return F.avg_pool2d(cubemap.permute(0, 3, 1, 2), (2, 2)).permute(0, 2, 3, 1).contiguous()


class EnvLightOracle:
    """scene/light.py:21-129 with a caller-provided mip chain (logit space)."""

    def __init__(self, levels, min_roughness=0.08, max_roughness=0.5, diffuse=None):
        self.specular = list(levels)
        self.diffuse = diffuse          # the cosine-convolved last level (scene/light.py:79), [6,16,16,3]
        self.min_roughness, self.max_roughness = min_roughness, max_roughness

    def get_mip(self, roughness):  # scene/light.py:88-96
        L = len(self.specular)
        return torch.where(
            roughness < self.max_roughness,
            (torch.clamp(roughness, self.min_roughness, self.max_roughness) - self.min_roughness)
            / (self.max_roughness - self.min_roughness) * (L - 2),
            (torch.clamp(roughness, self.max_roughness, 1.0) - self.max_roughness)
            / (1.0 - self.max_roughness) + L - 2)

    def __call__(self, l, mode=None, roughness=None):  # scene/light.py:98-129
        prefix = l.shape[:-1]
        d = l.reshape(-1, 3)
        if mode == "diffuse" and self.diffuse is not None:   # scene/light.py:108-110
            light = cube_texture([self.diffuse], d)
        elif mode in ("diffuse", "pure_env") or roughness is None:
            light = cube_texture(self.specular, d)
        else:
            light = cube_texture(self.specular, d, self.get_mip(roughness.reshape(-1)))
        return torch.sigmoid(light.view(*prefix, -1))


def linear_to_srgb(linear):  # utils/graphics_utils.py:102-110
    eps = torch.finfo(linear.dtype).eps
    srgb0 = 323 / 25 * linear
    srgb1 = (211 * linear.clamp_min(eps) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, srgb0, srgb1)


def get_specular_color_surfel(envmap, lut, albedo, HWK, R, T, normal_map, render_alpha, refl_strength,
                              roughness):
    """utils/refl_utils.py:364-419 with pc.ray_tracer = None (visibility = 1)."""
    H, W, K = HWK
    rays_cam, _ = sample_camera_rays(HWK, R, T, normal_map.device)
    w_o = -rays_cam
    NdotV = torch.sum(w_o * normal_map, dim=-1, keepdim=True)
    rays_refl = safe_normalize(2 * normal_map * NdotV - w_o)
    fg_uv = torch.cat([NdotV, roughness], -1).clamp(0, 1)
    fg = lut_fetch(lut, fg_uv.reshape(-1, 2)).reshape(H, W, 2)
    direct_light = envmap(rays_refl, roughness=roughness)
    specular_weight = (0.04 * (1 - refl_strength) + albedo * refl_strength) * fg[..., 0:1] + fg[..., 1:2]
    specular = direct_light * render_alpha * specular_weight
    return specular.permute(2, 0, 1), {"direct_light": direct_light.permute(2, 0, 1),
                                       "specular_weight": specular_weight}


def get_full_color_volume(envmap, lut, xyz, albedo, cam, normal_map, refl_strength, roughness):
    """utils/refl_utils.py:426-447: per-surfel split-sum colours of the volume-rendering stage. The reference fetches
    the LUT for all N surfels, squeezes the result to [N,2] and then indexes `fg[0]` (:445): the FIRST surfel's
    (scale, bias) pair multiplies every surfel. Restated as written."""
    _, rays_o = sample_camera_rays(cam.HWK, cam.R, cam.T, xyz.device)
    w_o = safe_normalize(rays_o.expand(normal_map.shape[0], -1) - xyz)
    NdotV = torch.sum(w_o * normal_map, dim=-1, keepdim=True)          # reflection(), :95-98
    rays_refl = safe_normalize(2 * normal_map * NdotV - w_o)
    fg = lut_fetch(lut, torch.cat([NdotV, roughness], -1).clamp(0, 1))  # [N,2]
    diffuse = envmap(normal_map, mode="diffuse") * (1 - refl_strength) * albedo
    specular = envmap(rays_refl, roughness=roughness) * (
        (0.04 * (1 - refl_strength) + albedo * refl_strength) * fg[0][..., 0:1] + fg[0][..., 1:2])
    return diffuse, specular


def shade_surfel(envmap, lut, rendered_image, rendered_features, allmap, cam, bg_color, srgb=False):
    """The part of render_surfel after the rasterizer call (gaussian_renderer/__init__.py:372-469)."""
    base_color = rendered_image
    refl_strength = rendered_features[:1]
    roughness = rendered_features[1:2]
    albedo = rendered_features[2:5]
    render_alpha = allmap[1:2]
    render_normal = allmap[2:5]
    w2v = torch.as_tensor(cam.world_view_transform, device=allmap.device)
    render_normal = (render_normal.permute(1, 2, 0) @ (w2v[:3, :3].T)).permute(2, 0, 1)
    normal_map = render_normal.permute(1, 2, 0) / render_alpha.permute(1, 2, 0).clamp_min(1e-6)
    specular, extra = get_specular_color_surfel(
        envmap, lut, albedo.permute(1, 2, 0), cam.HWK, cam.R, cam.T, normal_map, render_alpha.permute(1, 2, 0),
        refl_strength.permute(1, 2, 0), roughness.permute(1, 2, 0))
    final_image = (1 - refl_strength) * base_color + specular
    if srgb:
        final_image = linear_to_srgb(final_image)
    final_image = final_image + bg_color[:, None, None] * (1 - render_alpha)
    return {"render": final_image, "specular_map": specular, "diffuse_map": (1 - refl_strength) * base_color,
            "rend_normal": render_normal, "rend_alpha": render_alpha, "direct_light": extra["direct_light"],
            "refl_strength_map": refl_strength, "roughness_map": roughness, "base_color_map": albedo}


def depths_to_points(cam, depthmap):
    """utils/point_utils.py:9-24 (all in float32 torch ops like the reference)."""
    dev = depthmap.device
    w2v = torch.as_tensor(cam.world_view_transform, device=dev)
    full = torch.as_tensor(cam.full_proj_transform, device=dev)
    c2w = (w2v.T).inverse()
    W, H = cam.image_width, cam.image_height
    ndc2pix = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]], device=dev).float().T
    projection_matrix = c2w.T @ full
    intrins = (projection_matrix @ ndc2pix)[:3, :3].T
    grid_x, grid_y = torch.meshgrid(torch.arange(W, device=dev).float(), torch.arange(H, device=dev).float(),
                                    indexing="xy")
    points = torch.stack([grid_x, grid_y, torch.ones_like(grid_x)], dim=-1).reshape(-1, 3)
    rays_d = points @ intrins.inverse().T @ c2w[:3, :3].T
    rays_o = c2w[:3, 3]
    return depthmap.reshape(-1, 1) * rays_d + rays_o


def depth_to_normal(cam, depth):
    """utils/point_utils.py:26-37."""
    points = depths_to_points(cam, depth).reshape(*depth.shape[1:], 3)
    output = torch.zeros_like(points)
    dx = points[2:, 1:-1] - points[:-2, 1:-1]
    dy = points[1:-1, 2:] - points[1:-1, :-2]
    output[1:-1, 1:-1, :] = F.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    return output


def surf_depth_normal(allmap, cam, depth_ratio=0.0):
    """gaussian_renderer/__init__.py:50-78 (FLAG == "2dgs")."""
    render_alpha = allmap[1:2]
    median = torch.nan_to_num(allmap[5:6], 0, 0)
    expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
    surf_depth = expected * (1 - depth_ratio) + depth_ratio * median
    surf_normal = depth_to_normal(cam, surf_depth).permute(2, 0, 1) * render_alpha.detach()
    return surf_depth, surf_normal


def synthetic_gbuffer(H, W, S=8, seed=5, device="cpu"):
    """Config C1 G-buffer: view-space normals scaled by alpha, alpha ~ U[0.5,1] with
    10 % zeros, albedo/roughness/refl/base ~ U[0,1]."""
    g = torch.Generator().manual_seed(seed)
    alpha = torch.rand(1, H, W, generator=g) * 0.5 + 0.5
    alpha = alpha * (torch.rand(1, H, W, generator=g) > 0.1)
    n = F.normalize(torch.randn(3, H, W, generator=g), dim=0)
    n[2] = -n[2].abs()  # face the camera (view space looks down +z)
    allmap = torch.zeros(7, H, W)
    allmap[1:2] = alpha
    allmap[2:5] = n * alpha
    feats = torch.rand(S, H, W, generator=g)
    base = torch.rand(3, H, W, generator=g)
    return base.to(device), feats.to(device), allmap.to(device)


def synthetic_chain(res=128, min_res=16, seed=9, device="cpu"):
    """A logit-space cubemap ~ N(0,1) and its plain 2x2-average mip chain (no GGX prefilter)."""
    g = torch.Generator().manual_seed(seed)
    levels = [torch.randn(6, res, res, 3, generator=g)]
    while levels[-1].shape[1] > min_res:
        levels.append(cubemap_mip(levels[-1]))
    return [l.to(device) for l in levels]

"""Host-side mirror of the reference's deferred split-sum shading, backed by the fused CUDA
kernels of libmrgs.so (mrgs_shade_forward / mrgs_shade_backward / mrgs_envlight_query).

Reference interfaces mirrored here (same names and argument meaning):
  EnvLight                      scene/light.py:21-129 (base parameter in logit space, build_mips,
                                get_mip, __call__(l, mode, roughness) -> sigmoid(texture fetch))
  get_specular_color_surfel     utils/refl_utils.py:364-419 (visibility branch excluded: OptiX/mesh
                                tracing stays on the reference)
  shade_surfel                  the part of render_surfel after the rasterizer call,
                                gaussian_renderer/__init__.py:372-469
There is no torch fallback: every function raises if libmrgs.so is unavailable.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
import torch

from . import _lib

_LUT_PATH = Path(__file__).resolve().parent / "assets" / "bsdf_256_256.bin"
_LUT_CACHE: dict = {}


def fg_lut(device) -> torch.Tensor:
    """The 256x256x2 split-sum DFG table (utils/refl_utils.py:9), cached per device."""
    key = str(device)
    if key not in _LUT_CACHE:
        arr = np.fromfile(_LUT_PATH, dtype=np.float32).reshape(256, 256, 2)
        _LUT_CACHE[key] = torch.from_numpy(arr).to(device).contiguous()
    return _LUT_CACHE[key]


def linear_to_srgb(linear, eps=None):
    """utils/graphics_utils.py:102-110."""
    if eps is None:
        eps = torch.finfo(linear.dtype).eps
    srgb0 = 323 / 25 * linear
    srgb1 = (211 * linear.clamp_min(eps) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, srgb0, srgb1)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _chain_args(levels, min_roughness, max_roughness, min_levels: int = 2) -> _lib.ShadeArgs:
    if not min_levels <= len(levels) <= _lib.MAX_MIP_LEVELS:
        raise RuntimeError(f"mip chain must have {min_levels}..{_lib.MAX_MIP_LEVELS} levels, got {len(levels)}")
    a = _lib.ShadeArgs()
    a.num_levels = len(levels)
    a.base_res = int(levels[0].shape[1])
    a.min_roughness, a.max_roughness = float(min_roughness), float(max_roughness)
    for i, l in enumerate(levels):
        if not l.is_cuda or l.dtype != torch.float32 or not l.is_contiguous():
            raise RuntimeError("mip levels must be contiguous float32 CUDA tensors")
        if tuple(l.shape) != (6, a.base_res >> i, a.base_res >> i, 3):
            raise RuntimeError(f"level {i} has shape {tuple(l.shape)}, expected (6,{a.base_res >> i},{a.base_res >> i},3)")
        a.levels[i] = l.data_ptr()
    return a


def ray_matrix(HWK, R) -> np.ndarray:
    """M with ray_dir ∝ M @ (x, y, 1): the pixel_camera -> world chain of sample_camera_rays
    (utils/refl_utils.py:54-73) collapsed to one 3x3 (R is the c2w rotation stored by 3DGS)."""
    _, _, K = HWK
    return (np.asarray(R, np.float64) @ np.linalg.inv(np.asarray(K, np.float64))).astype(np.float32)


class _ShadeSurfel(torch.autograd.Function):
    """(base_color[3,H,W], features[S,H,W], allmap[7,H,W], *levels) ->
    (final, specular, direct_light, normal_world, diffuse), each [3,H,W]."""

    @staticmethod
    def forward(ctx, base_color, features, allmap, bg, cfg, *levels):
        lib = _lib.load()
        dev = base_color.device
        M, Q, min_r, max_r, srgb = cfg[:5]
        ctx.set_materialize_grads(False)   # outputs nobody differentiates arrive as None, not as zero-filled maps
        base_color, features, allmap = base_color.contiguous(), features.contiguous(), allmap.contiguous()
        levels = tuple(l.contiguous() for l in levels)
        if features.shape[0] < 5:
            raise RuntimeError("features must hold at least refl, roughness and albedo planes")
        H, W = base_color.shape[1], base_color.shape[2]
        a = _chain_args(levels, min_r, max_r)
        a.width, a.height, a.srgb = W, H, int(bool(srgb))
        a.ray_matrix[:] = [float(v) for v in np.asarray(M, np.float32).reshape(-1)]
        a.normal_matrix[:] = [float(v) for v in np.asarray(Q, np.float32).reshape(-1)]
        lut = fg_lut(dev)
        bg = bg.to(device=dev, dtype=torch.float32).contiguous()
        outs = [torch.empty((3, H, W), dtype=torch.float32, device=dev) for _ in range(5)]
        a.background, a.base_color, a.features, a.allmap, a.lut = (
            bg.data_ptr(), base_color.data_ptr(), features.data_ptr(), allmap.data_ptr(), lut.data_ptr())
        a.out_final, a.out_specular, a.out_direct, a.out_normal, a.out_diffuse = (o.data_ptr() for o in outs)
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_shade_forward(C.byref(a), _stream(dev)), "mrgs_shade_forward")
        ctx.cfg = cfg
        ctx.save_for_backward(base_color, features, allmap, bg, *levels)
        ctx.mark_non_differentiable(outs[2])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_final, g_spec, g_direct, g_normal, g_diffuse):
        lib = _lib.load()
        base_color, features, allmap, bg, *levels = ctx.saved_tensors
        dev = base_color.device
        M, Q, min_r, max_r, srgb = ctx.cfg[:5]
        sink = ctx.cfg[5] if len(ctx.cfg) > 5 else None     # EnvLight.level_grad_sink: [texels, 4] accumulation buffer
        H, W = base_color.shape[1], base_color.shape[2]
        a = _chain_args(levels, min_r, max_r)
        a.width, a.height, a.srgb = W, H, int(bool(srgb))
        a.ray_matrix[:] = [float(v) for v in np.asarray(M, np.float32).reshape(-1)]
        a.normal_matrix[:] = [float(v) for v in np.asarray(Q, np.float32).reshape(-1)]
        lut = fg_lut(dev)
        a.background, a.base_color, a.features, a.allmap, a.lut = (
            bg.data_ptr(), base_color.data_ptr(), features.data_ptr(), allmap.data_ptr(), lut.data_ptr())
        keep = [None if g is None else g.contiguous() for g in (g_final, g_spec, g_normal, g_diffuse)]
        a.dL_dfinal, a.dL_dspecular, a.dL_dnormal, a.dL_ddiffuse = (None if g is None else g.data_ptr() for g in keep)
        d_base = torch.empty_like(base_color)
        d_feat = torch.zeros_like(features) if features.shape[0] > 5 else torch.empty_like(features)
        d_allmap = torch.zeros_like(allmap)
        # one flat [texels, 4] accumulation buffer for the whole chain (16-byte texels for vector atomics)
        counts = [l.shape[0] * l.shape[1] * l.shape[2] for l in levels]
        need_levels = any(ctx.needs_input_grad[5:])
        if sink is not None:
            if tuple(sink.shape) != (sum(counts), 4) or sink.device != dev or sink.dtype != torch.float32:
                raise RuntimeError(f"level_grad_sink must be a float32 [{sum(counts)}, 4] tensor on {dev}")
            flat4 = sink
        else:
            flat4 = torch.zeros((sum(counts), 4), dtype=torch.float32, device=dev) if need_levels else None
        a.dL_dbase_color, a.dL_dfeatures, a.dL_dallmap = d_base.data_ptr(), d_feat.data_ptr(), d_allmap.data_ptr()
        off = 0
        for i, n in enumerate(counts):
            if flat4 is not None:
                a.dL_dlevels[i] = flat4.data_ptr() + off * 16
            off += n
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_shade_backward(C.byref(a), _stream(dev)), "mrgs_shade_backward")
        if sink is not None:
            sink._mrgs_pending = True
            hook = getattr(ctx.cfg[6], "after_sink_backward", None) if len(ctx.cfg) > 6 else None
            if hook is not None:        # e.g. start the sink's allreduce under the rasterizer backward that follows
                hook()
        if sink is not None or flat4 is None:    # the sink owns the texel gradients (EnvLight.flush_level_grads)
            return (d_base, d_feat, d_allmap, None, None, *([None] * len(levels)))
        flat3 = flat4[:, :3].contiguous()
        d_levels, off = [], 0
        for l, n in zip(levels, counts):
            d_levels.append(flat3[off:off + n].view(l.shape))
            off += n
        return (d_base, d_feat, d_allmap, None, None, *d_levels)


def depth_ray_matrix(H, W, tanfovx, tanfovy, R) -> np.ndarray:
    """A with world ray = A @ (x, y, 1) as depths_to_points builds it (utils/point_utils.py:9-24): the
    pinhole recovered from full_proj_transform has its principal point at (W/2, H/2)."""
    K = np.array([[W / (2.0 * tanfovx), 0.0, W / 2.0], [0.0, H / (2.0 * tanfovy), H / 2.0], [0.0, 0.0, 1.0]])
    return (np.asarray(R, np.float64) @ np.linalg.inv(K)).astype(np.float32)


class _SurfDepthNormal(torch.autograd.Function):
    """allmap [7,H,W] -> (surf_depth [1,H,W], surf_normal [3,H,W]); compute_2dgs_normal_and_regularizations
    (gaussian_renderer/__init__.py:50-78) + depth_to_normal (utils/point_utils.py:26-37) in one kernel."""

    @staticmethod
    def forward(ctx, allmap, cfg):
        lib = _lib.load()
        A, origin, ratio = cfg
        allmap = allmap.contiguous()
        dev = allmap.device
        H, W = allmap.shape[1], allmap.shape[2]
        depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        normal = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        Ac = (C.c_float * 9)(*[float(v) for v in np.asarray(A, np.float32).reshape(-1)])
        oc = (C.c_float * 3)(*[float(v) for v in np.asarray(origin, np.float32).reshape(-1)])
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_depth_normal_forward(W, H, float(ratio), Ac, oc, allmap.data_ptr(), depth.data_ptr(),
                                                     normal.data_ptr(), _stream(dev)), "mrgs_depth_normal_forward")
        ctx.cfg = cfg
        ctx.save_for_backward(allmap)
        return depth, normal

    @staticmethod
    def backward(ctx, g_depth, g_normal):
        lib = _lib.load()
        (allmap,) = ctx.saved_tensors
        A, origin, ratio = ctx.cfg
        dev = allmap.device
        H, W = allmap.shape[1], allmap.shape[2]
        g_allmap = torch.zeros_like(allmap)
        gd, gn = g_depth.contiguous(), g_normal.contiguous()
        Ac = (C.c_float * 9)(*[float(v) for v in np.asarray(A, np.float32).reshape(-1)])
        oc = (C.c_float * 3)(*[float(v) for v in np.asarray(origin, np.float32).reshape(-1)])
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_depth_normal_backward(W, H, float(ratio), Ac, oc, allmap.data_ptr(), gd.data_ptr(),
                                                      gn.data_ptr(), g_allmap.data_ptr(), _stream(dev)),
                       "mrgs_depth_normal_backward")
        return g_allmap, None


def surf_depth_normal(allmap, H, W, tanfovx, tanfovy, R, T, depth_ratio: float = 0.0):
    """(surf_depth, surf_normal) of render_surfel's dict; R, T as stored on the reference's Camera."""
    origin = -(np.asarray(R, np.float64) @ np.asarray(T, np.float64))
    cfg = (depth_ray_matrix(H, W, tanfovx, tanfovy, R), origin.astype(np.float32), float(depth_ratio))
    return _SurfDepthNormal.apply(allmap, cfg)


def shade_surfel(envmap: "EnvLight", rendered_image, rendered_features, allmap, HWK, R, bg_color,
                 srgb: bool = False) -> dict:
    """Everything render_surfel does after the rasterizer call except depth_to_normal
    (gaussian_renderer/__init__.py:372-469). HWK = (H, W, K) and R (c2w rotation) come from the
    camera exactly as the reference passes `viewpoint_camera.HWK / .R`."""
    cfg = (ray_matrix(HWK, R), np.asarray(R, np.float32), envmap.min_roughness, envmap.max_roughness, srgb,
           getattr(envmap, "level_grad_sink", None), envmap)
    if getattr(envmap, "_bg_event", None) is not None:
        envmap.sync()
    levels = envmap.specular
    if cfg[5] is not None:
        # sink mode: the texel gradients go to the sink, not through autograd - the view's autograd graph need not (and,
        # when the view is captured in a CUDA graph, must not) reach back into the graph build_mips() recorded earlier
        levels = [l.detach() for l in levels]
    final, specular, direct, normal_w, diffuse = _ShadeSurfel.apply(
        rendered_image, rendered_features, allmap, bg_color, cfg, *levels)
    return {
        "render": final,
        "refl_strength_map": rendered_features[:1],
        "diffuse_map": diffuse,
        "diffuse_map_ori": rendered_image,
        "specular_map": specular,
        "base_color_map": rendered_features[2:5],
        "roughness_map": rendered_features[1:2],
        "rend_alpha": allmap[1:2],
        "rend_normal": normal_w,
        "rend_dist": allmap[6:7],
        "direct_light": direct,
    }


def get_specular_color_surfel(envmap: "EnvLight", albedo, HWK, R, T, normal_map, render_alpha,
                              scaling_modifier=1.0, refl_strength=None, roughness=None, pc=None,
                              surf_depth=None, indirect_light=None):
    """utils/refl_utils.py:364-419 with the reference's HWC tensors; returns (specular [3,H,W],
    {'direct_light', 'specular_weight'}). The ray-traced visibility branch is out of scope."""
    if pc is not None and getattr(pc, "ray_tracer", None) is not None and indirect_light is not None:
        raise NotImplementedError("mesh/OptiX visibility tracing stays on the reference")
    H, W, _ = HWK
    alpha = render_alpha.permute(2, 0, 1)
    allmap = torch.zeros((7, H, W), dtype=torch.float32, device=albedo.device)
    allmap[1:2] = alpha
    allmap[2:5] = (normal_map * render_alpha.clamp_min(1e-6)).permute(2, 0, 1)
    feats = torch.cat([refl_strength, roughness, albedo], -1).permute(2, 0, 1).contiguous()
    cfg = (ray_matrix(HWK, R), np.eye(3, dtype=np.float32), envmap.min_roughness, envmap.max_roughness, False)
    zero3 = torch.zeros((3, H, W), dtype=torch.float32, device=albedo.device)
    _, specular, direct, _, _ = _ShadeSurfel.apply(zero3, feats, allmap, torch.zeros(3, device=albedo.device),
                                                   cfg, *envmap.specular)
    with torch.no_grad():
        safe = (direct * alpha).clamp_min(1e-20)
    extra = {"direct_light": direct, "specular_weight": (specular / safe).permute(1, 2, 0)}
    return specular, extra


def safe_normalize(x, eps=1e-20):
    return x / torch.clamp(torch.linalg.norm(x, dim=-1, keepdim=True), min=eps)   # utils/general_utils.py:179-182


def _camera_origin(R, T, device):
    """rays_o of sample_camera_rays (utils/refl_utils.py:56, :68): -R @ T with the c2w rotation 3DGS cameras store."""
    R = torch.as_tensor(np.asarray(R.detach().cpu() if isinstance(R, torch.Tensor) else R), dtype=torch.float32, device=device)
    T = torch.as_tensor(np.asarray(T.detach().cpu() if isinstance(T, torch.Tensor) else T), dtype=torch.float32, device=device)
    return (-R @ T.unsqueeze(-1)).flatten()


def _fg_of_first_surfel(ndotv, roughness):
    """The FG pair get_full_color_volume applies to EVERY surfel: the reference fetches the LUT for all N surfels and
    then indexes `fg[0]` on the [N,2] result (utils/refl_utils.py:437-445, :474-481), i.e. the first surfel's pair
    (kept as is). dr.texture(filter_mode='linear', boundary_mode='clamp') == bilinear with border clamping."""
    uv = torch.cat([ndotv[0:1], roughness[0:1]], -1).clamp(0, 1)
    lut = fg_lut(uv.device)[None]                             # [1,256,256,2]
    grid = (uv * 2.0 - 1.0).reshape(1, 1, 1, 2)
    fg = torch.nn.functional.grid_sample(lut.permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="border",
                                         align_corners=False)
    return fg[0, :, 0, 0]                                     # [2]


class _SurfelShade(torch.autograd.Function):
    """(xyz, normals, albedo, refl, rough [P,*], fg [2], campos [3], diffuse_map, *levels) ->
    (diffuse, specular, direct_light) [P,3]; mrgs_surfel_shade_forward / _backward."""

    @staticmethod
    def forward(ctx, xyz, normals, albedo, refl, rough, fg, campos, cfg, diffuse_map, *levels):
        lib = _lib.load()
        if not xyz.is_cuda:
            raise RuntimeError("get_full_color_volume: tensors must be CUDA tensors")
        ctx.set_materialize_grads(False)
        t = [x.detach().contiguous().float() for x in (xyz, normals, albedo, refl, rough, fg, campos, diffuse_map)]
        lv = [l.detach().contiguous() for l in levels]
        P, dev = t[0].shape[0], t[0].device
        a = _lib.SurfelShadeArgs()
        a.chain = _chain_args(lv, cfg[0], cfg[1])
        a.P, a.diffuse_res = P, int(t[7].shape[1])
        (a.xyz, a.normals, a.albedo, a.refl_strength, a.roughness, a.fg, a.campos, a.diffuse_map) = (x.data_ptr() for x in t)
        outs = [torch.empty((P, 3), dtype=torch.float32, device=dev) for _ in range(3)]
        a.diffuse, a.specular, a.direct_light = (o.data_ptr() for o in outs)
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_surfel_shade_forward(C.byref(a), _stream(dev)), "mrgs_surfel_shade_forward")
        ctx.cfg = cfg
        ctx.save_for_backward(*t, *lv)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_diffuse, g_specular, g_direct):
        lib = _lib.load()
        xyz, normals, albedo, refl, rough, fg, campos, diffuse_map, *lv = ctx.saved_tensors
        P, dev = xyz.shape[0], xyz.device
        a = _lib.SurfelShadeArgs()
        a.chain = _chain_args(lv, ctx.cfg[0], ctx.cfg[1])
        a.P, a.diffuse_res = P, int(diffuse_map.shape[1])
        (a.xyz, a.normals, a.albedo, a.refl_strength, a.roughness, a.fg, a.campos, a.diffuse_map) = (
            x.data_ptr() for x in (xyz, normals, albedo, refl, rough, fg, campos, diffuse_map))
        keep = [None if g is None else g.contiguous().float() for g in (g_diffuse, g_specular, g_direct)]
        a.dL_ddiffuse, a.dL_dspecular, a.dL_ddirect = (None if g is None else g.data_ptr() for g in keep)
        need = ctx.needs_input_grad
        grads = [torch.empty_like(x) if need[i] else None for i, x in enumerate((xyz, normals, albedo, refl, rough))]
        a.dL_dxyz, a.dL_dnormals, a.dL_dalbedo, a.dL_drefl_strength, a.dL_droughness = (
            None if g is None else g.data_ptr() for g in grads)
        g_fg = torch.zeros(2, dtype=torch.float32, device=dev) if need[5] else None
        a.dL_dfg = None if g_fg is None else g_fg.data_ptr()
        n_d = diffuse_map.shape[0] * diffuse_map.shape[1] * diffuse_map.shape[2]
        d4 = torch.zeros((n_d, 4), dtype=torch.float32, device=dev) if need[8] else None
        a.dL_ddiffuse_map = None if d4 is None else d4.data_ptr()
        counts = [l.shape[0] * l.shape[1] * l.shape[2] for l in lv]
        need_l = [need[9 + i] for i in range(len(lv))]
        flat4 = torch.zeros((sum(counts), 4), dtype=torch.float32, device=dev) if any(need_l) else None
        off = 0
        for i, n in enumerate(counts):
            if need_l[i]:
                a.chain.dL_dlevels[i] = flat4.data_ptr() + off * 16
            off += n
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_surfel_shade_backward(C.byref(a), _stream(dev)), "mrgs_surfel_shade_backward")
        g_levels, off = [], 0
        for l, n, nd in zip(lv, counts, need_l):
            g_levels.append(flat4[off:off + n, :3].reshape(l.shape) if nd else None)
            off += n
        g_dmap = None if d4 is None else d4[:, :3].reshape(diffuse_map.shape)
        return (*grads, g_fg, None, None, g_dmap, *g_levels)


def _surfel_colours(envmap: "EnvLight", xyz, albedo, R, T, normal_map, refl_strength, roughness):
    """(diffuse, specular, direct_light) [N,3]: one fused kernel pair for everything per-surfel; the FG pair of the first
    surfel (the reference's `fg[0]`, see _fg_of_first_surfel) is evaluated in torch so its gradient reaches surfel 0."""
    if getattr(envmap, "_bg_event", None) is not None:
        envmap.sync()
    rays_o = _camera_origin(R, T, xyz.device)
    w_o0 = safe_normalize(rays_o[None] - xyz[0:1])
    fg = _fg_of_first_surfel(torch.sum(w_o0 * normal_map[0:1], dim=-1, keepdim=True), roughness[0:1])
    return _SurfelShade.apply(xyz, normal_map, albedo, refl_strength, roughness, fg, rays_o,
                              (envmap.min_roughness, envmap.max_roughness), envmap.diffuse, *envmap.specular)


def get_full_color_volume(envmap: "EnvLight", xyz, albedo, HWK, R, T, normal_map, render_alpha, scaling_modifier=1.0,
                          refl_strength=None, roughness=None):
    """utils/refl_utils.py:426-447 — per-SURFEL split-sum colours (the volume-rendering stage): (diffuse, specular) [N,3]."""
    diffuse, specular, _ = _surfel_colours(envmap, xyz, albedo, R, T, normal_map, refl_strength, roughness)
    return diffuse, specular


def get_full_color_volume_indirect(envmap: "EnvLight", xyz, albedo, HWK, R, T, normal_map, render_alpha,
                                   scaling_modifier=1.0, refl_strength=None, roughness=None, pc=None, indirect_light=None):
    """utils/refl_utils.py:450-490 with visibility = 1 (the mesh/OptiX tracer stays on the reference): the specular light
    is then the direct light alone, `indirect_light` only rides along as a feature."""
    if pc is not None and getattr(pc, "ray_tracer", None) is not None:
        raise NotImplementedError("mesh/OptiX visibility tracing stays on the reference")
    diffuse, specular, direct_light = _surfel_colours(envmap, xyz, albedo, R, T, normal_map, refl_strength, roughness)
    return diffuse, specular, {"visibility": torch.ones_like(render_alpha), "direct_light": direct_light}


class EnvLight(torch.nn.Module):
    """Trainable logit-space cubemap with a GGX-prefiltered mip chain (scene/light.py:21-129)."""

    def __init__(self, path=None, device=None, scale=1.0, min_res=16, max_res=128, min_roughness=0.08,
                 max_roughness=0.5, trainable=False):
        super().__init__()
        if path is not None:
            raise NotImplementedError("loading .hdr/.exr environment maps stays on the reference")
        self.device = device if device is not None else "cuda"
        self.scale = scale
        self.min_res, self.max_res = min_res, max_res
        self.min_roughness, self.max_roughness = min_roughness, max_roughness
        self.trainable = trainable
        self.base = torch.nn.Parameter(
            torch.zeros(6, max_res, max_res, 3, dtype=torch.float32, device=self.device),
            requires_grad=trainable)
        self.build_mips()

    _chain = None
    shard_world = None      # set by shard_build_mips(): (group,) - build_mips forward/backward split over its ranks

    def shard_build_mips(self, group=None):
        """View-sharded training replicates the cubemap on every rank. After this call every rank filters only its share
        of each level (forward) / of each level's gradient (backward) and one allreduce of the 25 MB chain assembles the
        result - identical numbers, 1/world of the work. Collective: every rank of `group` must call build_mips() and
        flush_level_grads() / backward together."""
        self.shard_world = (group,)

    def _shard(self):
        if self.shard_world is None:
            return None
        import torch.distributed as dist
        group = self.shard_world[0]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.get_rank(group), dist.get_world_size(group), group

    background = None       # (side stream, CTA cap) set by run_in_background()
    _bg_event = None

    def run_in_background(self, stream=None, ctas_per_sm: float = 1.0):
        """build_mips() and flush_level_grads() are HBM-bound gathers that leave most issue slots idle; the tile-blend
        kernels of the rasterizer are issue-bound and hardly touch HBM. After this call both run on `stream` with a
        grid capped at ctas_per_sm CTAs per SM, i.e. in the BACKGROUND of the rasterizer kernels the caller enqueues on
        its own stream meanwhile (forward: the rasterizer forwards of the step's views; backward: their rasterizer
        backwards). Everything that reads the chain (`shade_surfel`, `__call__`, ...) and the next `build_mips()` call
        sync() first; call sync() yourself before reading `base.grad`."""
        dev = self.base.device
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.background = (stream if stream is not None else torch.cuda.Stream(device=dev), max(1, int(ctas_per_sm * sms)))

    def sync(self):
        """Order the current stream after the background build_mips / flush_level_grads, if any."""
        if self._bg_event is not None:
            torch.cuda.current_stream(self.base.device).wait_event(self._bg_event)

    static_chain = False    # True: build_mips() rewrites the SAME level buffers every time (CUDA-graph consumers, graphs.py)

    def chain_roughnesses(self, n):
        """Roughness of every level as build_mips assigns them (scene/light.py:81-86)."""
        return [(idx / (n - 2)) * (self.max_roughness - self.min_roughness) + self.min_roughness
                for idx in range(n - 1)] + [1.0]

    def build_mips(self, cutoff=0.99):
        """scene/light.py:72-86. Power-of-two chains run as ONE autograd node over cached prefilter plans
        (prefilter.py: mip pyramid + one gather launch forward, one gather + the mip backward chain backward);
        other shapes, and chains whose plans exceed the HBM budget, compose the per-level ops like the reference."""
        from . import cubemap as cm
        from . import prefilter as pf
        self.sync()
        sink = self.level_grad_sink
        if sink is not None and getattr(sink, "_mrgs_pending", False):
            raise RuntimeError("EnvLight.build_mips(): the level-gradient sink holds gradients of the previous chain; "
                               "call flush_level_grads() first")
        n, r = 1, self.max_res
        while r > self.min_res:
            r //= 2
            n += 1
        self._chain = None
        if self.base.is_cuda and n != 2 and n <= _lib.MAX_MIP_LEVELS and self.max_res % (1 << (n - 1)) == 0:
            try:
                self._chain = pf.get_chain(self.max_res, n, self.chain_roughnesses(n), cutoff, self.base.device)
            except pf.PrefilterTooLarge:
                self._chain = None
        if self._chain is not None and self.background is not None:
            # in the background of whatever the current stream does next (see run_in_background): consumers call sync()
            side, ctas = self.background
            main = torch.cuda.current_stream(self.base.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.specular, self.diffuse = pf.build_mips(self.base, self._chain, self.static_chain, self._shard(), ctas)
                self._bg_event = torch.cuda.Event()
                self._bg_event.record(side)
            for t in (*self.specular, self.diffuse):
                t.record_stream(main)
        elif self._chain is not None:
            self.specular, self.diffuse = pf.build_mips(self.base, self._chain, self.static_chain, self._shard())
        else:
            self.specular = [self.base]
            while self.specular[-1].shape[1] > self.min_res:
                self.specular += [cm.cubemap_mip(self.specular[-1])]
            self.diffuse = cm.diffuse_cubemap(self.specular[-1])
            n = len(self.specular)
            for idx in range(n - 1):
                roughness = (idx / (n - 2)) * (self.max_roughness - self.min_roughness) + self.min_roughness
                self.specular[idx] = cm.specular_cubemap(self.specular[idx], roughness, cutoff)
            self.specular[-1] = cm.specular_cubemap(self.specular[-1], 1.0, cutoff)
        if sink is not None:
            self.enable_level_grad_sink()     # re-sized if the chain's texel count changed

    def set_chain(self, levels):
        """Install an externally built mip chain (tests / benchmarks)."""
        self.specular = list(levels)
        self._chain = None

    level_grad_sink = None

    def enable_level_grad_sink(self):
        """Multi-view steps: let the shading backward ADD the texel gradients of every view into ONE persistent
        [texels, 4] buffer instead of returning them through autograd (per view that costs a zero-fill of the buffer, a
        strided copy to [.,3] and one accumulation pass per level). Call flush_level_grads() once per step, before the
        next build_mips(): with a fused chain it runs the prefilter backward straight from the buffer into `base.grad`
        (no autograd graph involved, so chains that were ALSO differentiated through autograd in the same step simply
        add up in `base.grad`); otherwise it feeds the sums into the levels' autograd graph (which must then not have
        been freed by an earlier backward) or the levels' .grad when they are leaves. The sink serves shade_surfel
        only; EnvLight.__call__ and the per-surfel shading return their texel gradients through autograd."""
        n = sum(l.shape[0] * l.shape[1] * l.shape[2] for l in self.specular)
        dev = self.specular[0].device
        if self.level_grad_sink is None or self.level_grad_sink.shape[0] != n or self.level_grad_sink.device != dev:
            self.level_grad_sink = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        return self.level_grad_sink

    after_sink_backward = None     # optional callable, run right after a shading backward has added into the sink

    def use_level_grad_sink(self, buffer):
        """Adopt a caller-owned [texels, 4] fp32 buffer as the sink (e.g. the tail of parallel.GradArena's flat buffer, so
        that one allreduce carries the surfel gradients and the cubemap gradients)."""
        n = sum(l.shape[0] * l.shape[1] * l.shape[2] for l in self.specular)
        if tuple(buffer.shape) != (n, 4) or buffer.dtype != torch.float32 or not buffer.is_contiguous() or \
                buffer.device != self.specular[0].device:
            raise RuntimeError(f"level-gradient sink must be a contiguous float32 [{n}, 4] tensor on {self.specular[0].device}")
        self.level_grad_sink = buffer
        return buffer

    def flush_level_grads(self):
        sink = self.level_grad_sink
        if sink is None:
            return
        if self._chain is not None and self.background is not None:
            side, ctas = self.background
            main = torch.cuda.current_stream(self.base.device)
            side.wait_stream(main)       # every shading backward enqueued so far has added into the sink
            with torch.cuda.stream(side):
                if self.base.requires_grad:
                    g = self._chain.backward(sink, None, self._shard(), ctas)
                    self.base.grad = g if self.base.grad is None else self.base.grad.add_(g)
                sink.zero_()
                self._bg_event = torch.cuda.Event()
                self._bg_event.record(side)
            sink._mrgs_pending = False
            return
        if self._chain is not None:
            if self.base.requires_grad:
                g = self._chain.backward(sink, None, self._shard())
                self.base.grad = g if self.base.grad is None else self.base.grad.add_(g)
        else:
            flat3 = sink[:, :3].contiguous()
            grads, off = [], 0
            for l in self.specular:
                n = l.shape[0] * l.shape[1] * l.shape[2]
                grads.append(flat3[off:off + n].view(l.shape))
                off += n
            live = [(l, g) for l, g in zip(self.specular, grads) if l.requires_grad]
            if live:
                torch.autograd.backward([l for l, _ in live], grad_tensors=[g for _, g in live])
        sink.zero_()
        sink._mrgs_pending = False

    def get_mip(self, roughness):
        n = len(self.specular)
        return torch.where(
            roughness < self.max_roughness,
            (torch.clamp(roughness, self.min_roughness, self.max_roughness) - self.min_roughness)
            / (self.max_roughness - self.min_roughness) * (n - 2),
            (torch.clamp(roughness, self.max_roughness, 1.0) - self.max_roughness)
            / (1.0 - self.max_roughness) + n - 2)

    def forward(self, l, mode=None, roughness=None):
        """Query the environment light for directions l[..., 3] (scene/light.py:98-129): "diffuse" samples the
        cosine-convolved 16^2 map, "pure_env" the base level, otherwise the GGX chain at get_mip(roughness).
        Differentiable like dr.texture: gradients reach the mip levels (and through build_mips the base cubemap),
        the directions and the roughness."""
        prefix = l.shape[:-1]
        self.sync()
        if mode == "diffuse":
            levels, rough = [self.diffuse], None
        elif mode == "pure_env":
            levels, rough = [self.base], None
        else:
            levels, rough = self.specular, roughness
        out = _EnvQuery.apply(l.reshape(-1, 3), None if rough is None else rough.reshape(-1),
                              (self.min_roughness, self.max_roughness), *levels)
        return out.view(*prefix, 3)


class _EnvQuery(torch.autograd.Function):
    """(dirs [n,3], roughness [n] or None, *levels) -> sigmoid(cube fetch) [n,3]; mrgs_envlight_query(+_backward)."""

    @staticmethod
    def forward(ctx, dirs, roughness, cfg, *levels):
        lib = _lib.load()
        if not dirs.is_cuda:
            raise RuntimeError("EnvLight: directions must be a CUDA tensor")
        d = dirs.detach().contiguous().float()
        r = None if roughness is None else roughness.detach().contiguous().float()
        lv = [x.detach().contiguous() for x in levels]
        a = _chain_args(lv, cfg[0], cfg[1], min_levels=1 if r is None else 2)
        out = torch.empty_like(d)
        with torch.cuda.device(d.device):
            _lib.check(lib.mrgs_envlight_query(C.byref(a), d.shape[0], d.data_ptr(), None if r is None else r.data_ptr(),
                                               out.data_ptr(), _stream(d.device)), "mrgs_envlight_query")
        ctx.cfg, ctx.has_rough = cfg, r is not None
        ctx.save_for_backward(d, *([r] if r is not None else []), *lv)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        d, *rest = ctx.saved_tensors
        r = rest.pop(0) if ctx.has_rough else None
        lv = rest
        dev = d.device
        a = _chain_args(lv, ctx.cfg[0], ctx.cfg[1], min_levels=1 if r is None else 2)
        need_d, need_r = ctx.needs_input_grad[0], ctx.has_rough and ctx.needs_input_grad[1]
        need_l = [ctx.needs_input_grad[3 + i] for i in range(len(lv))]
        g = g_out.contiguous().float()
        g_d = torch.empty_like(d) if need_d else None
        g_r = torch.empty_like(r) if need_r else None
        counts = [x.shape[0] * x.shape[1] * x.shape[2] for x in lv]
        flat4 = torch.zeros((sum(counts), 4), dtype=torch.float32, device=dev) if any(need_l) else None
        off = 0
        for i, n in enumerate(counts):
            if need_l[i]:
                a.dL_dlevels[i] = flat4.data_ptr() + off * 16
            off += n
        with torch.cuda.device(dev):
            _lib.check(lib.mrgs_envlight_query_backward(
                C.byref(a), d.shape[0], d.data_ptr(), None if r is None else r.data_ptr(), g.data_ptr(),
                None if g_d is None else g_d.data_ptr(), None if g_r is None else g_r.data_ptr(), _stream(dev)),
                "mrgs_envlight_query_backward")
        g_levels, off = [], 0
        for x, n, need in zip(lv, counts, need_l):
            g_levels.append(flat4[off:off + n, :3].reshape(x.shape) if need else None)
            off += n
        return (g_d, g_r, None, *g_levels)

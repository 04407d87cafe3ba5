"""GPU: the whole render() contract of gaussian_renderer/__init__.py (render_initial, render_surfel, render_volume) — our
kernels against oracle/render_oracle.py driven by the reference rasterizer extension. That restatement is pinned on the
CPU by vectors of the reference's OWN three functions (tests/test_render_oracle_cpu.py)."""
import math
import types

import pytest
import torch
import torch.nn.functional as F

from materialrefgs_b200 import synthetic
from materialrefgs_b200.render import render_surfel
from oracle import render_oracle as ro
from oracle import shading_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


_MEASURED = []


def _close_grads(ga, gb, name):
    """End-to-end gradient check of a whole render() call. The shaded image is only piecewise differentiable, so the
    bar is the relative L1 error plus a bound on the fraction of elements that are off by more than 1 % of the max-norm
    (see the comment in test_render_surfel_contract); the measured values are logged to gpurun_out/ for the record."""
    l1 = ((ga - gb).abs().sum() / gb.abs().sum().clamp_min(1e-20)).item()
    out_frac = ((ga - gb).abs() > 1e-2 * gb.abs().max()).float().mean().item()
    mx = ((ga - gb).abs().max() / gb.abs().max().clamp_min(1e-20)).item()
    _MEASURED.append({"name": name, "rel_l1": l1, "outlier_fraction": out_frac, "rel_max": mx})
    try:
        import json
        import os
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/render_contract_errors.json", "w") as fh:
            json.dump(_MEASURED, fh, indent=1)
    except OSError:
        pass
    assert l1 <= L1_BAR and out_frac <= 1e-3, (name, l1, out_frac, mx)


L1_BAR = 3e-3   # measured (gpurun_out/render_contract_errors.json): <= 2.5e-3 on the FakeModel case, <= 2e-4 elsewhere


class FakeModel:
    """The getters render_surfel uses from scene/gaussian_model.py, over a synthetic cloud."""
    active_sh_degree = 3
    max_sh_degree = 3

    def __init__(self, cloud, env):
        self.c, self.env = cloud, env
        g = torch.Generator().manual_seed(4)
        self.ind = (0.2 * torch.randn(cloud.P, 16, 3, generator=g)).to(cloud.means3D.device).requires_grad_(True)
        self.leaves = {k: getattr(cloud, k).clone().requires_grad_(True)
                       for k in ("means3D", "scales", "rotations", "opacities", "shs", "features")}

    get_xyz = property(lambda s: s.leaves["means3D"])
    get_opacity = property(lambda s: s.leaves["opacities"])
    get_scaling = property(lambda s: s.leaves["scales"])
    get_rotation = property(lambda s: s.leaves["rotations"])
    get_features = property(lambda s: s.leaves["shs"])
    get_refl = property(lambda s: s.leaves["features"][:, 0:1])
    get_rough = property(lambda s: s.leaves["features"][:, 1:2])
    get_ori_color = property(lambda s: s.leaves["features"][:, 2:5])
    get_indirect = property(lambda s: s.ind)
    get_envmap = property(lambda s: s.env)

    def get_normal(self, scaling_modifier, dir_pp_normalized):
        q = F.normalize(self.leaves["rotations"], dim=-1)
        w, x, y, z = q.unbind(-1)
        n = torch.stack([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], -1)
        flip = (n * -dir_pp_normalized).sum(-1, keepdim=True) >= 0
        return n * torch.where(flip, 1.0, -1.0)


class OracleSide:
    """The same model seen by the oracle: every getter of `pc`, but the environment light is the torch restatement."""
    def __init__(self, pc, env):
        self._pc, self._env = pc, env

    def __getattr__(self, k):
        if k in ("get_envmap", "get_envmap_2"):
            return self._env
        return getattr(self._pc, k)


def reference_pipeline(ref_ext, cam, pc, pipe, bg, levels):
    """render_surfel with the reference rasterizer + the pinned torch restatement of everything after it."""
    return ro.render_surfel(ref_ext, cam, OracleSide(pc, so.EnvLightOracle(levels)), pipe, bg)


def test_render_surfel_contract(ref_ext):
    from materialrefgs_b200.shading import EnvLight
    W, H, P = 400, 304, 60_000
    cloud = synthetic.make_cloud(P, S=5, seed=31).to(DEV)
    cam = synthetic.orbit_camera(2, 8, W, H).to(DEV)
    cam.FoVx, cam.FoVy = cam.FoVx, cam.FoVy
    env = EnvLight(device=DEV, max_res=64, min_res=16, trainable=True)
    with torch.no_grad():
        env.base.copy_(torch.randn(6, 64, 64, 3, generator=torch.Generator().manual_seed(2)).to(DEV))
    env.build_mips()
    pipe = types.SimpleNamespace(debug=False, depth_ratio=0.25, compute_cov3D_python=False, use_asg=False)
    bg = torch.tensor([0.2, 0.3, 0.4], device=DEV)
    g = torch.Generator().manual_seed(12)
    wts = {k: (torch.randn(c, H, W, generator=g) / (H * W)).to(DEV)
           for k, c in (("render", 3), ("rend_normal", 3), ("surf_normal", 3), ("rend_dist", 1), ("rend_alpha", 1))}

    def loss_of(o):
        return sum((o[k] * w).sum() for k, w in wts.items())

    pc = FakeModel(cloud, env)
    out = render_surfel(cam, pc, pipe, bg)
    assert set(out) >= {"render", "refl_strength_map", "diffuse_map", "diffuse_map_ori", "specular_map", "base_color_map",
                        "roughness_map", "viewspace_points", "visibility_filter", "radii", "rend_alpha", "rend_normal",
                        "rend_dist", "surf_depth", "surf_normal"}
    loss_of(out).backward()
    g_ours = {k: v.grad.clone() for k, v in pc.leaves.items()}
    g_env = env.base.grad.clone()
    g_ind = pc.ind.grad.clone()

    pc2 = FakeModel(cloud, env)
    env.base.grad = None
    env.build_mips()
    levels = [l for l in env.specular]
    ref = reference_pipeline(ref_ext, cam, pc2, pipe, bg, levels)
    loss_of(ref).backward()

    assert torch.equal(out["radii"], ref["radii"])
    for k in ("render", "specular_map", "diffuse_map", "rend_normal", "rend_alpha", "surf_depth"):
        assert (out[k] - ref[k]).abs().max().item() <= 1e-4, k
    assert (out["surf_normal"] - ref["surf_normal"]).abs().max().item() <= 5e-4
    # The shaded image is only piecewise differentiable (bilinear texel cells of the LUT / cube levels): a pixel
    # whose lookup lands within float rounding of a cell boundary takes the left derivative in one implementation
    # and the right one in the other (measured: 2 of 121 600 pixels, central finite differences sit exactly between
    # the two analytic values). Such a pixel can dominate the max-norm of ONE surfel's gradient, so this end-to-end
    # check uses the relative L1 error plus a bound on the fraction of outliers; the per-kernel tests keep the
    # max-norm bars (1e-3 rasterizer with identical upstream gradients, 1e-3 shading on a smooth G-buffer).
    close = _close_grads
    for k in g_ours:
        close(g_ours[k], pc2.leaves[k].grad, k)
    close(g_ind, pc2.ind.grad, "indirect")
    close(g_env, env.base.grad, "envmap")
    vis = out["visibility_filter"]
    assert out["viewspace_points"].grad is not None and out["viewspace_points"].grad[vis].abs().sum() > 0


class RawModel:
    """A GaussianModel-like object holding RAW parameters (scene/gaussian_model.py attribute names): render_surfel
    then prepares the rasterizer inputs with the fused feature kernel."""
    active_sh_degree = 3
    max_sh_degree = 3

    def __init__(self, cloud, env, seed=5):
        from oracle import features_oracle as fo
        raw, _ = fo.synthetic_params(cloud.P, seed=seed)
        dev = cloud.means3D.device
        raw["xyz"] = cloud.means3D.cpu()
        raw["scaling"] = torch.log(cloud.scales.cpu())
        raw["rotation"] = cloud.rotations.cpu() * 1.7           # un-normalised on purpose
        raw["opacity"] = torch.logit(cloud.opacities.cpu().clamp(1e-4, 1 - 1e-4))
        self.raw = {k: v.to(dev).requires_grad_(True) for k, v in raw.items()}
        self.shs = cloud.shs.clone().requires_grad_(True)
        self.env = env
        r = self.raw
        self._xyz, self._scaling, self._rotation, self._opacity = r["xyz"], r["scaling"], r["rotation"], r["opacity"]
        self._refl_strength, self._roughness, self._ori_color = r["refl_strength"], r["roughness"], r["ori_color"]
        self._indirect_dc, self._indirect_rest = r["indirect_dc"], r["indirect_rest"]

    get_xyz = property(lambda s: s._xyz)
    get_features = property(lambda s: s.shs)
    get_envmap = property(lambda s: s.env)


def test_render_surfel_contract_raw_parameters(ref_ext):
    """render_surfel on a raw-parameter model (fused feature preparation + rasterizer + shading + depth normals)
    against: features oracle -> reference rasterizer -> shading oracle; gradients w.r.t. the RAW parameters."""
    from materialrefgs_b200.shading import EnvLight
    from oracle import features_oracle as fo
    W, H, P = 320, 240, 40_000
    cloud = synthetic.make_cloud(P, S=8, seed=33).to(DEV)
    cam = synthetic.orbit_camera(5, 8, W, H).to(DEV)
    env = EnvLight(device=DEV, max_res=64, min_res=16, trainable=False)
    with torch.no_grad():
        env.base.copy_(torch.randn(6, 64, 64, 3, generator=torch.Generator().manual_seed(3)).to(DEV))
    env.build_mips()
    pipe = types.SimpleNamespace(debug=False, depth_ratio=0.0, compute_cov3D_python=False, use_asg=False)
    bg = torch.tensor([0.1, 0.1, 0.1], device=DEV)
    g = torch.Generator().manual_seed(13)
    wts = {k: (torch.randn(c, H, W, generator=g) / (H * W)).to(DEV)
           for k, c in (("render", 3), ("rend_normal", 3), ("surf_normal", 3), ("rend_alpha", 1))}
    loss_of = lambda o: sum((o[k] * w).sum() for k, w in wts.items())

    pc = RawModel(cloud, env)
    out = render_surfel(cam, pc, pipe, bg)
    loss_of(out).backward()

    pc2 = RawModel(cloud, env)
    scales, rots, opac, feats = fo.prepare_features(*[pc2.raw[k] for k, _ in fo.RAW_FIELDS], cam.camera_center)
    m2d = torch.zeros_like(pc2._xyz, requires_grad=True)
    rs = ref_ext.GaussianRasterizationSettings(
        H, W, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), torch.zeros_like(bg), 1.0,
        cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center, False, False)
    _, color, feat, radii, allmap = ref_ext.GaussianRasterizer(rs)(
        means3D=pc2._xyz, means2D=m2d, opacities=opac, shs=pc2.shs, features=feats, scales=scales, rotations=rots)
    ref = so.shade_surfel(so.EnvLightOracle([l for l in env.specular]), so.load_lut(DEV), color, feat, allmap, cam, bg)
    ref["surf_depth"], ref["surf_normal"] = so.surf_depth_normal(allmap, cam, pipe.depth_ratio)
    loss_of(ref).backward()

    assert torch.equal(out["radii"], radii)
    for k in ("render", "specular_map", "diffuse_map", "rend_normal", "rend_alpha", "surf_depth", "roughness_map"):
        assert (out[k] - ref[k]).abs().max().item() <= 1e-4, k
    for k in pc.raw:
        _close_grads(pc.raw[k].grad, pc2.raw[k].grad, k)


# ---- render_initial / render_volume (gaussian_renderer/__init__.py:94-222, :521-745) -----------------------------


def test_render_initial_contract(ref_ext):
    from materialrefgs_b200.render import render_initial
    W, H, P = 333, 200, 30_000
    cloud = synthetic.make_cloud(P, S=5, seed=41).to(DEV)
    cam = synthetic.orbit_camera(3, 8, W, H).to(DEV)
    pipe = types.SimpleNamespace(debug=False, depth_ratio=1.0, compute_cov3D_python=False)
    bg = torch.tensor([0.0, 0.5, 1.0], device=DEV)
    g = torch.Generator().manual_seed(14)
    wts = {k: (torch.randn(c, H, W, generator=g) / (H * W)).to(DEV)
           for k, c in (("render", 3), ("rend_normal", 3), ("surf_normal", 3), ("rend_dist", 1), ("surf_depth", 1))}
    loss_of = lambda o: sum((o[k] * w).sum() for k, w in wts.items())

    pc = FakeModel(cloud, None)
    out = render_initial(cam, pc, pipe, bg, srgb=True)
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii", "rend_alpha", "rend_normal", "rend_dist",
                        "surf_depth", "surf_normal"}
    loss_of(out).backward()

    pc2 = FakeModel(cloud, None)
    ref = ro.render_initial(ref_ext, cam, pc2, pipe, bg, srgb=True)
    radii = ref["radii"]
    loss_of(ref).backward()
    assert torch.equal(out["radii"], radii)
    for k in wts:
        assert (out[k] - ref[k]).abs().max().item() <= (5e-4 if k == "surf_normal" else 1e-4), k
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        _close_grads(pc.leaves[k].grad, pc2.leaves[k].grad, k)


@pytest.mark.parametrize("indirect", [False, True])
def test_render_volume_contract(ref_ext, indirect):
    """Per-surfel shading through the differentiable EnvLight queries + S = 11 / 18 feature channels."""
    from materialrefgs_b200.render import render_volume
    from materialrefgs_b200.shading import EnvLight
    W, H, P = 256, 192, 25_000
    cloud = synthetic.make_cloud(P, S=5, seed=43).to(DEV)
    cam = synthetic.orbit_camera(6, 8, W, H).to(DEV)
    env = EnvLight(device=DEV, max_res=64, min_res=16, trainable=True)
    with torch.no_grad():
        env.base.copy_(torch.randn(6, 64, 64, 3, generator=torch.Generator().manual_seed(4)).to(DEV))
    pipe = types.SimpleNamespace(debug=False, depth_ratio=0.0, compute_cov3D_python=False, use_asg=False)
    opt = types.SimpleNamespace(indirect=indirect)
    bg = torch.tensor([0.3, 0.2, 0.1], device=DEV)
    g = torch.Generator().manual_seed(15)
    keys = [("render", 3), ("diffuse_map", 3), ("specular_map", 3), ("base_color_map", 3), ("roughness_map", 1),
            ("refl_strength_map", 1), ("rend_normal", 3), ("surf_normal", 3)] + ([("direct_light", 3), ("indirect_light", 3)] if indirect else [])
    wts = {k: (torch.randn(c, H, W, generator=g) / (H * W)).to(DEV) for k, c in keys}
    loss_of = lambda o: sum((o[k] * w).sum() for k, w in wts.items())

    class VolumeModel(FakeModel):
        get_envmap_2 = property(lambda s: s.env)
        ray_tracer = None

    env.build_mips()
    pc = VolumeModel(cloud, env)
    out = render_volume(cam, pc, pipe, bg, opt=opt)
    loss_of(out).backward()
    g_env = env.base.grad.clone()

    env.base.grad = None
    env.build_mips()
    pc2 = VolumeModel(cloud, env)
    oracle_env = so.EnvLightOracle(list(env.specular), diffuse=env.diffuse)
    ref = ro.render_volume(ref_ext, cam, OracleSide(pc2, oracle_env), pipe, bg, indirect=indirect)
    radii = ref["radii"]
    loss_of(ref).backward()

    assert torch.equal(out["radii"], radii)
    for k in wts:
        assert (out[k] - ref[k]).abs().max().item() <= (5e-4 if k == "surf_normal" else 1e-4), k
    if indirect:
        assert (out["visibility"] - ref["visibility"]).abs().max().item() <= 1e-4
    for k in ("means3D", "scales", "rotations", "opacities", "features"):
        _close_grads(pc.leaves[k].grad, pc2.leaves[k].grad, k)
    _close_grads(g_env, env.base.grad, "envmap")
    if indirect:
        _close_grads(pc.ind.grad, pc2.ind.grad, "indirect")

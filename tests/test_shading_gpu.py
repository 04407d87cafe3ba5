"""GPU: the fused deferred-shading kernels against the torch restatement (oracle/shading_oracle.py,
parity with real nvdiffrast unpinned — see its header). Tolerances: 1e-4 absolute on images,
1e-3 relative (to the max-norm) on gradients."""
import numpy as np
import pytest
import torch

from materialrefgs_b200 import synthetic
from oracle import shading_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _env(levels, min_r=0.08, max_r=0.5):
    from materialrefgs_b200.shading import EnvLight
    env = EnvLight.__new__(EnvLight)
    torch.nn.Module.__init__(env)
    env.min_roughness, env.max_roughness = min_r, max_r
    env.set_chain(levels)
    return env


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


@pytest.mark.parametrize("H,W,res,srgb,view", [(120, 160, 64, False, 1), (97, 131, 128, True, 4), (64, 64, 32, False, 6),
                                               (800, 800, 512, False, 2)])   # BASELINE C1/C3: 800^2, 6 x 512^2, 6 levels
def test_shade_forward_backward(H, W, res, srgb, view):
    from materialrefgs_b200.shading import shade_surfel
    cam = synthetic.orbit_camera(view, 8, W, H)
    base, feats, allmap = so.synthetic_gbuffer(H, W, device=DEV, seed=view)
    if view == 6:   # drive roughness outside [min_r, 1] and N.V outside [0,1] to hit the clamps
        feats[1] = feats[1] * 1.6 - 0.3
        allmap[2:5] = -allmap[2:5]
    if res == 512:
        # BASELINE size: the chain the shader is fed in C1/C3 = EnvLight.build_mips of a 6x512^2 logit cubemap ~ N(0,1)
        # (GGX-prefiltered levels; white-noise texels at 512^2 would make the fp32 rounding of u*512 itself worth 1e-4)
        from materialrefgs_b200.shading import EnvLight
        env0 = EnvLight(device=DEV, max_res=512, min_res=16, trainable=False)
        with torch.no_grad():
            env0.base.copy_(torch.randn(6, 512, 512, 3, generator=torch.Generator().manual_seed(1234)).to(DEV))
        env0.build_mips()
        levels = [l.detach().clone() for l in env0.specular]
        assert len(levels) == 6
    else:
        levels = so.synthetic_chain(res, 16, device=DEV)
    bg = torch.tensor([0.2, 0.4, 0.6], device=DEV)
    g = torch.Generator().manual_seed(3)
    wts = {k: torch.randn(3, H, W, generator=g).to(DEV) for k in ("render", "specular_map", "diffuse_map", "rend_normal")}

    def loss_of(d):
        return sum((d[k] * w).sum() for k, w in wts.items())

    # oracle (autograd)
    lv_o = [l.clone().requires_grad_(True) for l in levels]
    b_o, f_o, a_o = (t.clone().requires_grad_(True) for t in (base, feats, allmap))
    ref = so.shade_surfel(so.EnvLightOracle(lv_o), so.load_lut(DEV), b_o, f_o, a_o, cam, bg, srgb=srgb)
    loss_of(ref).backward()

    lv_m = [l.clone().requires_grad_(True) for l in levels]
    b_m, f_m, a_m = (t.clone().requires_grad_(True) for t in (base, feats, allmap))
    out = shade_surfel(_env(lv_m), b_m, f_m, a_m, cam.HWK, cam.R, bg, srgb=srgb)
    loss_of(out).backward()

    for k in ("render", "specular_map", "diffuse_map", "rend_normal", "direct_light"):
        assert (out[k] - ref[k]).abs().max().item() <= 1e-4, k

    def close(a, b):
        if res < 512:
            return _rel(a, b) <= 1e-3
        # 640 000 pixels looking up a 512^2 chain: a few dozen land within fp32 rounding of a texel-cell / mip-level
        # boundary, where the fetch has two one-sided derivatives and either implementation may take either
        # (tests/test_texture_properties_gpu.py checks the derivative away from the kinks). Bar: 1e-3 of the max-norm on
        # all but 0.05 % of the elements.
        bad = ((a - b).abs() > 1e-3 * b.abs().max()).float().mean().item()
        return bad <= 5e-4
    assert close(b_m.grad, b_o.grad)
    assert close(f_m.grad[:5], f_o.grad[:5])
    assert not f_m.grad[5:].any()
    assert close(a_m.grad[1:5], a_o.grad[1:5])
    for lm, lo in zip(lv_m, lv_o):
        assert _rel(lm.grad, lo.grad) <= 1e-3


def test_cube_fetch_edges_and_corners():
    """Directions on face edges / cube corners: the seamless wrap and the 3-texel corner average."""
    levels = so.synthetic_chain(16, 4, device=DEV)
    env = _env(levels)
    g = torch.Generator().manual_seed(1)
    d = torch.randn(20000, 3, generator=g)
    d[:5000] = torch.sign(d[:5000]) * (1 + 0.02 * torch.rand(5000, 3, generator=g))     # near corners
    d[5000:10000, 0] = torch.sign(d[5000:10000, 0]) * d[5000:10000, 1].abs()            # exact |x| == |y| edges
    d = d.to(DEV)
    rough = torch.rand(20000, 1, generator=g).to(DEV)
    ref = so.EnvLightOracle(levels)(d, roughness=rough)
    out = env(d, roughness=rough)
    assert (out - ref).abs().max().item() <= 1e-5
    ref0 = so.EnvLightOracle(levels)(d, mode="pure_env")
    env.base = levels[0]
    out0 = env(d, mode="pure_env")
    assert (out0 - ref0).abs().max().item() <= 1e-5


def test_get_specular_color_surfel_api():
    from materialrefgs_b200.shading import get_specular_color_surfel
    H, W = 80, 112
    cam = synthetic.orbit_camera(2, 8, W, H)
    base, feats, allmap = so.synthetic_gbuffer(H, W, device=DEV)
    levels = so.synthetic_chain(64, 16, device=DEV)
    alpha = allmap[1:2].permute(1, 2, 0)
    normal_map = allmap[2:5].permute(1, 2, 0) / alpha.clamp_min(1e-6)
    args = dict(refl_strength=feats[0:1].permute(1, 2, 0), roughness=feats[1:2].permute(1, 2, 0))
    ref, ref_extra = so.get_specular_color_surfel(so.EnvLightOracle(levels), so.load_lut(DEV), feats[2:5].permute(1, 2, 0),
                                                  cam.HWK, cam.R, cam.T, normal_map, alpha, **args)
    out, extra = get_specular_color_surfel(_env(levels), feats[2:5].permute(1, 2, 0), cam.HWK, cam.R, cam.T, normal_map,
                                           alpha, **args)
    assert (out - ref).abs().max().item() <= 1e-4
    assert (extra["direct_light"] - ref_extra["direct_light"]).abs().max().item() <= 1e-4


@pytest.mark.parametrize("ratio,H,W", [(0.0, 96, 128), (0.35, 75, 101)])
def test_surf_depth_and_depth_to_normal(ratio, H, W):
    from materialrefgs_b200.shading import surf_depth_normal
    cam = synthetic.orbit_camera(3, 8, W, H)
    g = torch.Generator().manual_seed(17)
    allmap = torch.zeros(7, H, W)
    alpha = torch.rand(1, H, W, generator=g) * 0.6 + 0.4
    alpha = alpha * (torch.rand(1, H, W, generator=g) > 0.1)          # holes: 0/0 -> nan_to_num
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    depth = 3.0 + 0.5 * torch.sin(6 * xx) * torch.cos(4 * yy) + 0.02 * torch.rand(H, W, generator=g)
    allmap[0] = depth * alpha[0]
    allmap[1:2] = alpha
    allmap[5] = depth + 0.05
    allmap = allmap.to(DEV)
    wd = torch.randn(1, H, W, generator=g).to(DEV)
    wn = torch.randn(3, H, W, generator=g).to(DEV)
    a_o = allmap.clone().requires_grad_(True)
    d_o, n_o = so.surf_depth_normal(a_o, cam.to(DEV), ratio)
    ((d_o * wd).sum() + (n_o * wn).sum()).backward()
    a_m = allmap.clone().requires_grad_(True)
    d_m, n_m = surf_depth_normal(a_m, H, W, cam.tanfovx, cam.tanfovy, cam.R, cam.T, ratio)
    ((d_m * wd).sum() + (n_m * wn).sum()).backward()
    assert (d_m - d_o).abs().max().item() <= 1e-5
    assert (n_m - n_o).abs().max().item() <= 2e-4          # normals of nearly flat 3x3 stencils amplify 1e-7 depth noise
    assert not n_m[:, 0].any() and not n_m[:, :, -1].any()  # border stays zero
    # where alpha == 0 the reference's autograd yields NaN (0/0 in the division backward); those pixels have no
    # contributors, so the rasterizer backward never reads them — we return 0 there
    ok = (allmap[1] > 0)
    assert torch.isfinite(a_m.grad).all()
    for pl in (0, 1, 5):
        ref = a_o.grad[pl][ok]
        assert ((a_m.grad[pl][ok] - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item() <= 2e-3, pl
    assert not a_m.grad[[2, 3, 4, 6]].any()


# ---- differentiable EnvLight.__call__ (mrgs_envlight_query / _backward) -----------------------------------------
def _query_dirs(n, seed):
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    # a share of directions hugging cube edges and corners (seamless taps, 3-texel corner average)
    k = n // 8
    d[:k, 0] = d[:k, 1] * (1 + 1e-3 * torch.randn(k, generator=g))
    d[k:2 * k] = torch.sign(d[k:2 * k]) * (1 + 2e-2 * torch.rand(k, 3, generator=g))
    return d * (0.5 + torch.rand(n, 1, generator=g))          # un-normalised: dr.texture does not normalise either


@pytest.mark.parametrize("mode,res", [(None, 64), ("diffuse", 16), ("pure_env", 64), ("diffuse", 32)])
def test_envlight_query_forward_backward(mode, res):
    n = 20_000
    levels = so.synthetic_chain(64, 16, device=DEV)
    diffuse = 0.7 * torch.randn(6, res, res, 3, generator=torch.Generator().manual_seed(8)).to(DEV)
    d0 = _query_dirs(n, 5).to(DEV)
    r0 = (torch.rand(n, 1, generator=torch.Generator().manual_seed(6)) * 1.3 - 0.15).to(DEV)   # beyond both clamps
    w = torch.randn(n, 3, generator=torch.Generator().manual_seed(7)).to(DEV)

    def run(make_env):
        lv = [l.clone().requires_grad_(True) for l in levels]
        df = diffuse.clone().requires_grad_(True)
        d, r = d0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
        out = make_env(lv, df)(d, mode=mode, roughness=None if mode else r)
        (out * w).sum().backward()
        return out.detach(), d.grad, r.grad, [l.grad for l in lv], df.grad

    def ours(lv, df):
        env = _env(lv)
        env.diffuse, env.base = df, lv[0]
        return env
    o_out, o_d, o_r, o_lv, o_df = run(ours)
    r_out, r_d, r_r, r_lv, r_df = run(lambda lv, df: so.EnvLightOracle(lv, diffuse=df))
    assert o_out.shape == (n, 3) and (o_out - r_out).abs().max().item() <= 1e-5
    # direction gradients are piecewise (texel cells): compare by relative L1 with a bound on outliers
    l1 = ((o_d - r_d).abs().sum() / r_d.abs().sum()).item()
    assert l1 <= 2e-3 and ((o_d - r_d).abs() > 1e-2 * r_d.abs().max()).float().mean().item() <= 1e-3
    if mode is None:
        l1r = ((o_r - r_r).abs().sum() / r_r.abs().sum()).item()
        assert l1r <= 2e-3
        for a, b in zip(o_lv, r_lv):
            assert _rel(a, b) <= 1e-3
        assert o_df is None
    elif mode == "diffuse":
        assert _rel(o_df, r_df) <= 1e-3 and all(g is None for g in o_lv) and o_r is None
    else:
        assert _rel(o_lv[0], r_lv[0]) <= 1e-3 and all(g is None for g in o_lv[1:])


def test_envlight_query_keeps_leading_shape_and_reaches_the_base_cubemap():
    from materialrefgs_b200.shading import EnvLight
    env = EnvLight(device=DEV, max_res=64, min_res=16, trainable=True)   # 3 levels (2 would divide by zero, as in light.py:82)
    with torch.no_grad():
        env.base.copy_(torch.randn(6, 64, 64, 3, generator=torch.Generator().manual_seed(1)).to(DEV))
    env.build_mips()
    d = _query_dirs(4 * 5 * 6, 2).view(4, 5, 6, 3).to(DEV)
    out = env(d, roughness=torch.full((4, 5, 6, 1), 0.3, device=DEV)) + env(d, mode="diffuse")
    assert out.shape == (4, 5, 6, 3)
    out.sum().backward()
    assert env.base.grad is not None and env.base.grad.abs().sum().item() > 0


def test_level_grad_sink_equals_per_view_autograd_accumulation():
    """EnvLight.enable_level_grad_sink(): the texel gradients of several views summed in the persistent buffer and
    flushed once equal what autograd accumulates view by view; outputs nobody differentiates cost no zero maps."""
    from materialrefgs_b200.shading import shade_surfel
    H, W = 96, 128
    levels = so.synthetic_chain(64, 16, device=DEV)
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    views = []
    for v in (1, 3, 5):
        base, feats, allmap = so.synthetic_gbuffer(H, W, device=DEV, seed=v)
        views.append((synthetic.orbit_camera(v, 8, W, H), base, feats, allmap,
                      torch.randn(3, H, W, generator=torch.Generator().manual_seed(v)).to(DEV)))

    def run(use_sink):
        lv = [l.clone().requires_grad_(True) for l in levels]
        env = _env(lv)
        if use_sink:
            env.enable_level_grad_sink()
        feat_grads = []
        for cam, base, feats, allmap, w in views:
            f = feats.clone().requires_grad_(True)
            out = shade_surfel(env, base, f, allmap, cam.HWK, cam.R, bg)
            (out["render"] * w).sum().backward()          # only ONE of the five outputs carries a gradient
            feat_grads.append(f.grad)
            if use_sink:
                assert all(l.grad is None for l in lv)     # nothing reaches the levels before the flush
        if use_sink:
            env.flush_level_grads()
            assert not env.level_grad_sink.any()
        return [l.grad for l in lv], feat_grads

    a_lv, a_f = run(True)
    b_lv, b_f = run(False)
    for a, b in zip(a_lv, b_lv):
        assert _rel(a, b) <= 1e-5
    for a, b in zip(a_f, b_f):
        assert torch.equal(a, b)


def test_per_surfel_volume_colours_forward_backward():
    """get_full_color_volume (utils/refl_utils.py:426-447) as the fused mrgs_surfel_shade_* pair against the torch
    restatement: values 1e-5, gradients of every per-surfel input, of the diffuse map and of the chain; includes the
    reference's `fg[0]` quirk (the FIRST surfel's LUT pair multiplies every surfel, its gradient lands on surfel 0)."""
    from materialrefgs_b200.shading import get_full_color_volume
    N = 30_000
    g = torch.Generator().manual_seed(11)
    cam = synthetic.orbit_camera(2, 8, 64, 64)
    levels = so.synthetic_chain(64, 16, device=DEV)
    diffuse_map = 0.7 * torch.randn(6, 16, 16, 3, generator=g).to(DEV)
    base = dict(xyz=1.3 * (2 * torch.rand(N, 3, generator=g) - 1), n=torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1),
                albedo=torch.rand(N, 3, generator=g), rs=torch.rand(N, 1, generator=g), ro=torch.rand(N, 1, generator=g) * 1.2 - 0.1)
    w_d, w_s = torch.randn(N, 3, generator=g).to(DEV), torch.randn(N, 3, generator=g).to(DEV)

    def run(fn):
        t = {k: v.to(DEV).clone().requires_grad_(True) for k, v in base.items()}
        lv = [l.clone().requires_grad_(True) for l in levels]
        dm = diffuse_map.clone().requires_grad_(True)
        d, s = fn(t, lv, dm)
        ((d * w_d).sum() + (s * w_s).sum()).backward()
        return d.detach(), s.detach(), {k: v.grad for k, v in t.items()}, [l.grad for l in lv], dm.grad

    def ours(t, lv, dm):
        env = _env(lv)
        env.diffuse = dm
        return get_full_color_volume(env, t["xyz"], t["albedo"], cam.HWK, cam.R, cam.T, t["n"], None,
                                     refl_strength=t["rs"], roughness=t["ro"])

    def oracle(t, lv, dm):
        return so.get_full_color_volume(so.EnvLightOracle(lv, diffuse=dm), so.load_lut(DEV), t["xyz"], t["albedo"], cam, t["n"],
                                        t["rs"], t["ro"])
    od, os_, og, olv, odm = run(ours)
    rd, rs_, rg, rlv, rdm = run(oracle)
    assert (od - rd).abs().max().item() <= 1e-5 and (os_ - rs_).abs().max().item() <= 1e-5
    for k in base:
        l1 = ((og[k] - rg[k]).abs().sum() / rg[k].abs().sum()).item()
        frac = ((og[k] - rg[k]).abs() > 1e-2 * rg[k].abs().max()).float().mean().item()
        assert l1 <= 2e-3 and frac <= 1e-3, (k, l1, frac)
    # surfel 0 carries the gradient of the shared FG pair: compare it on its own
    for k in ("n", "ro", "xyz"):
        assert _rel(og[k][0], rg[k][0]) <= 2e-3, k
    assert _rel(odm, rdm) <= 1e-3
    for a, b in zip(olv, rlv):
        assert _rel(a, b) <= 1e-3


def test_cubemap_gradient_through_build_mips_sink_equals_autograd():
    """The trainable base cubemap receives its gradient through EnvLight.build_mips either by autograd (texel gradients
    returned by the shading backward -> _BuildMips.backward) or through the multi-view sink (shading backward adds into
    one [texels,4] buffer, flush_level_grads runs the prefilter backward once): same numbers, and twice the gradient for
    two identical views."""
    from materialrefgs_b200.shading import EnvLight, shade_surfel
    H, W = 120, 160
    cam = synthetic.orbit_camera(3, 8, W, H)
    base, feats, allmap = so.synthetic_gbuffer(H, W, device=DEV, seed=2)
    bg = torch.zeros(3, device=DEV)
    env = EnvLight(device=DEV, max_res=128, min_res=16, trainable=True)
    with torch.no_grad():
        env.base.copy_(torch.randn(6, 128, 128, 3, generator=torch.Generator().manual_seed(21)).to(DEV))
    w = torch.randn(3, H, W, generator=torch.Generator().manual_seed(22)).to(DEV)
    env.build_mips()
    assert env._chain is not None
    (shade_surfel(env, base, feats, allmap, cam.HWK, cam.R, bg)["render"] * w).sum().backward()
    auto = env.base.grad.clone()
    assert float(auto.abs().max()) > 0
    env.base.grad = None
    env.build_mips()
    env.enable_level_grad_sink()
    base = base.requires_grad_(True)    # (in sink mode the texel gradients do not travel through autograd)
    for _ in range(2):
        (shade_surfel(env, base, feats, allmap, cam.HWK, cam.R, bg)["render"] * w).sum().backward()
    assert env.base.grad is None            # nothing reaches the base before the flush
    env.flush_level_grads()
    assert _rel(env.base.grad, 2 * auto) <= 1e-5


def test_background_build_mips_and_flush_equal_foreground():
    """EnvLight.run_in_background(): build_mips / flush_level_grads on a second stream with a capped grid (the
    grid-stride path of the gather) give the same chain and the same cubemap gradient as the foreground calls."""
    from materialrefgs_b200.shading import EnvLight, shade_surfel
    H, W = 96, 128
    cam = synthetic.orbit_camera(5, 8, W, H)
    base, feats, allmap = so.synthetic_gbuffer(H, W, device=DEV, seed=4)
    base = base.requires_grad_(True)
    bg = torch.zeros(3, device=DEV)
    w = torch.randn(3, H, W, generator=torch.Generator().manual_seed(23)).to(DEV)
    res = {}
    for mode in ("foreground", "background"):
        env = EnvLight(device=DEV, max_res=128, min_res=16, trainable=True)
        with torch.no_grad():
            env.base.copy_(torch.randn(6, 128, 128, 3, generator=torch.Generator().manual_seed(31)).to(DEV))
        if mode == "background":
            env.run_in_background(ctas_per_sm=0.25)      # 37 CTAs: every warp strides over many patches
        env.build_mips()
        env.enable_level_grad_sink()
        for _ in range(2):     # two steps: the second build_mips must wait for the first flush
            env.base.grad = None
            env.build_mips()
            out = shade_surfel(env, base, feats, allmap, cam.HWK, cam.R, bg)
            (out["render"] * w).sum().backward()
            env.flush_level_grads()
        env.sync()
        torch.cuda.synchronize()
        res[mode] = ([l.detach().clone() for l in env.specular], out["render"].detach().clone(), env.base.grad.clone())
    for a, b in zip(res["foreground"][0], res["background"][0]):
        assert torch.equal(a, b)
    assert torch.equal(res["foreground"][1], res["background"][1])
    assert _rel(res["background"][2], res["foreground"][2]) <= 1e-5     # (texel-gradient atomics are order-dependent)

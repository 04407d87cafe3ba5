"""GPU: EnvLight.build_mips through prefilter plans (csrc/prefilter.cu) against the UNMODIFIED reference
renderutils plugin (oracle/_ref) at the chain's real resolutions and roughnesses, up to the 6x512^2 chain of
BASELINE.json's C1/C3 configs. Bars: bounds equal (test_cubemap_gpu.py), forward 1e-5, backward 1e-4 of the
tensor's max-norm."""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import cubemap_oracle as co
from oracle import shading_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def ref_plugin():
    d = ROOT / "oracle" / "_ref" / "renderutils_plugin"
    if not (d / "renderutils_plugin.so").exists():
        pytest.skip("oracle/_ref/renderutils_plugin is not built")
    sys.path.insert(0, str(d))
    return importlib.import_module("renderutils_plugin")


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def _reference_chain(ref_plugin, base, min_res, min_r, max_r, cutoff=0.99):
    """EnvLight.build_mips (scene/light.py:72-86) composed from the reference plugin's ops; returns the raw levels,
    the prefiltered levels, the diffuse map and per-level (roughness, cos cutoff, bounds)."""
    from materialrefgs_b200 import prefilter as pf
    raw = [base]
    while raw[-1].shape[1] > min_res:
        raw.append(so.cubemap_mip(raw[-1]))
    n = len(raw)
    rough = [(i / (n - 2)) * (max_r - min_r) + min_r for i in range(n - 1)] + [1.0]
    keys, spec = [], []
    for lvl, r in zip(raw, rough):
        ct = pf.cutoff_costheta(r, cutoff)
        b = ref_plugin.specular_bounds(lvl.shape[1], ct)
        out4 = ref_plugin.specular_cubemap_fwd(lvl.contiguous(), b, r, ct)
        spec.append(out4[..., :3] / out4[..., 3:])
        keys.append((r, ct, b, out4[..., 3:]))
    diffuse = ref_plugin.diffuse_cubemap_fwd(raw[-1].contiguous())
    return raw, spec, diffuse, keys


def _reference_chain_backward(ref_plugin, raw, keys, g_levels, g_diffuse):
    """Gradient of the base cubemap: plugin backward per level + the reference's cubemap_mip backward
    (scene/light_utils.py:72-80: bilinear cube fetch of 0.25 * dout at the fine texel directions)."""
    n = len(raw)
    G = None
    for l in range(n - 1, -1, -1):
        r, ct, b, wsum = keys[l]
        d4 = torch.cat([g_levels[l] / wsum, torch.zeros_like(wsum)], -1).contiguous()   # d(rgb / wsum) w.r.t. rgb
        g = ref_plugin.specular_cubemap_bwd(raw[l].contiguous(), b, d4, r, ct)
        if l == n - 1:
            g = g + ref_plugin.diffuse_cubemap_bwd(raw[l].contiguous(), g_diffuse.contiguous())
        else:
            res = raw[l].shape[1]
            dirs = torch.from_numpy(co.texel_dirs(res)).to(DEV).reshape(-1, 3)
            g = g + so.cube_texture([G * 0.25], dirs).reshape(6, res, res, 3)
        G = g
    return G


@pytest.mark.parametrize("res,min_res,shape", [(128, 16, "32x1"), (128, 16, "32x2"), (128, 16, "8x1"), (128, 16, "16x2"),
                                               (64, 8, "8x2"), (64, 8, None), (512, 16, None)])
def test_build_mips_chain_vs_reference_plugin(ref_plugin, res, min_res, shape, monkeypatch):
    from materialrefgs_b200 import prefilter as pf
    from materialrefgs_b200.shading import EnvLight
    if shape:
        monkeypatch.setenv("MRGS_PREFILTER_SHAPE", shape)
    g = torch.Generator().manual_seed(res + len(shape or ""))
    env = EnvLight(device=DEV, max_res=res, min_res=min_res, trainable=True)
    with torch.no_grad():
        env.base.copy_(torch.randn(6, res, res, 3, generator=g).to(DEV))
    env.build_mips()
    assert env._chain is not None, "the fused prefilter chain must be the path taken"
    if shape:
        want = tuple(int(v) for v in shape.split("x"))
        for (p, _), r in zip(env._chain.spec, env._chain.sizes):
            tiles = r % want[0] == 0 and r % (32 // want[0] * want[1]) == 0
            assert (p.patch_width, p.rows_per_lane) == (want if tiles else (32, 1))
    raw, spec, diffuse, keys = _reference_chain(ref_plugin, env.base.detach(), min_res, env.min_roughness, env.max_roughness)
    assert len(spec) == len(env.specular)
    for l, (a, b) in enumerate(zip(env.specular, spec)):
        fin = torch.isfinite(b)
        assert torch.equal(torch.isfinite(a), fin), f"level {l}: NaN pattern differs"
        assert _rel(a.detach()[fin], b[fin]) <= 1e-5, f"level {l} forward"
        assert torch.equal(env._chain.spec[l][0].wsum.view_as(keys[l][3]), keys[l][3]) or \
            _rel(env._chain.spec[l][0].wsum.view_as(keys[l][3]), keys[l][3]) <= 1e-6, f"level {l} weight sum"
    assert _rel(env.diffuse.detach(), diffuse) <= 1e-5
    # backward through autograd: random upstream gradients on every level and on the diffuse map
    g_levels = [torch.randn(l.shape, generator=g).to(DEV) for l in env.specular]
    g_diffuse = torch.randn(env.diffuse.shape, generator=g).to(DEV)
    if min_res < 16:   # empty cones (NaN outputs) carry no gradient in either implementation
        g_levels = [torch.where(torch.isfinite(s), gl, torch.zeros_like(gl)) for s, gl in zip(spec, g_levels)]
    loss = sum((l.nan_to_num() * gl).sum() for l, gl in zip(env.specular, g_levels)) + (env.diffuse * g_diffuse).sum()
    loss.backward()
    if min_res >= 16:
        ref = _reference_chain_backward(ref_plugin, raw, keys, g_levels, g_diffuse)
        assert _rel(env.base.grad, ref) <= 1e-4
    # the sink route (shade_surfel's multi-view accumulation) must give the same gradient without autograd
    auto = env.base.grad.clone()
    env.base.grad = None
    env.build_mips()
    sink = env.enable_level_grad_sink()
    off = 0
    for gl in g_levels:
        k = gl.numel() // 3
        sink[off:off + k, :3] = gl.reshape(-1, 3)
        off += k
    sink._mrgs_pending = True
    with pytest.raises(RuntimeError):
        env.build_mips()
    env.flush_level_grads()
    assert float(sink.abs().max()) == 0.0
    direct = env.base.grad + env._chain.backward(None, g_diffuse)
    assert _rel(direct, auto) <= 1e-6


def test_transposed_plan_is_the_exact_transpose():
    """W^T built by swapping the roles in the reference's expressions must be the transpose of W: <W x, y> = <x, W^T y>
    up to fp32 summation order, for the specular and the diffuse operators, both patch shapes."""
    from materialrefgs_b200 import prefilter as pf
    g = torch.Generator().manual_seed(3)
    for res, rough, shape in ((32, 0.3, (32, 1)), (64, 0.12, (32, 2)), (16, 1.0, (32, 1)), (64, 0.2, (8, 2)), (32, 0.4, (16, 1))):
        ct = pf.cutoff_costheta(rough, 0.99)
        fwd, bwd = pf.plan_pair("specular", res, rough, ct, DEV, shape)
        assert fwd.taps == bwd.taps and fwd.taps > 0
        x = torch.randn(6, res, res, 3, generator=g).to(DEV)
        y = torch.randn(6, res, res, 3, generator=g).to(DEV)
        Wx, Wty = torch.empty_like(x), torch.empty_like(x)
        pf.apply_jobs([(fwd, x, 3, Wx, 3, None)], False, DEV)
        pf.apply_jobs([(bwd, y, 3, Wty, 3, None)], True, DEV)
        a, b = (Wx.double() * y.double()).sum().item(), (x.double() * Wty.double()).sum().item()
        assert abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1.0), (res, rough, shape, a, b)
    fwd, bwd = pf.plan_pair("diffuse", 16, 0.0, None, DEV)
    x = torch.randn(6, 16, 16, 3, generator=g).to(DEV)
    y = torch.randn(6, 16, 16, 3, generator=g).to(DEV)
    Wx, Wty = torch.empty_like(x), torch.empty_like(x)
    pf.apply_jobs([(fwd, x, 3, Wx, 3, None)], False, DEV)
    pf.apply_jobs([(bwd, y, 3, Wty, 3, None)], True, DEV)
    a, b = (Wx.double() * y.double()).sum().item(), (x.double() * Wty.double()).sum().item()
    assert abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1.0)


def test_padded_sources_and_job_batches():
    """float4-padded sources/destinations and more jobs than one launch holds give the same numbers."""
    from materialrefgs_b200 import prefilter as pf
    res, rough = 32, 0.25
    ct = pf.cutoff_costheta(rough, 0.99)
    fwd, _ = pf.plan_pair("specular", res, rough, ct, DEV)
    g = torch.Generator().manual_seed(11)
    x3 = torch.randn(6, res, res, 3, generator=g).to(DEV)
    x4 = torch.cat([x3, torch.full((6, res, res, 1), 7.0, device=DEV)], -1).contiguous()
    ref = torch.empty_like(x3)
    pf.apply_jobs([(fwd, x3, 3, ref, 3, None)], False, DEV)
    outs = [torch.zeros(6, res, res, 4, device=DEV) for _ in range(11)]
    pf.apply_jobs([(fwd, x4, 4, o, 4, None) for o in outs], False, DEV)
    for o in outs:
        assert torch.equal(o[..., :3], ref) and float(o[..., 3].abs().max()) == 0.0


def test_plan_budget_falls_back_to_direct_kernels(monkeypatch):
    from materialrefgs_b200 import cubemap as cm
    from materialrefgs_b200 import prefilter as pf
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 32, 32, 3, generator=g).to(DEV).requires_grad_(True)
    a = cm.specular_cubemap(x, 0.31)
    a.sum().backward()
    ga = x.grad.clone()
    x.grad = None
    monkeypatch.setenv("MRGS_PREFILTER_BUDGET_GB", "0")
    with pytest.raises(pf.PrefilterTooLarge):
        pf.plan_pair("specular", 32, 0.32, pf.cutoff_costheta(0.32, 0.99), DEV)
    for k in [k for k in pf._plan_cache if k[:3] == ("specular", 32, (0.31, pf.cutoff_costheta(0.31, 0.99)))]:
        pf._plan_cache.pop(k)
    b = cm.specular_cubemap(x, 0.31)     # direct per-texel kernels
    b.sum().backward()
    assert _rel(a.detach(), b.detach()) <= 1e-5
    assert _rel(ga, x.grad) <= 1e-4

"""GPU: EnvLight.build_mips kernels against (a) the UNMODIFIED reference renderutils_plugin in
oracle/_ref (when built) and (b) the numpy oracle."""
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


@pytest.mark.parametrize("N,rough", [(16, 1.0), (32, 0.3), (64, 0.08), (128, 0.2)])
def test_specular_and_diffuse_vs_reference_plugin(ref_plugin, N, rough):
    from materialrefgs_b200 import cubemap as cm
    g = torch.Generator().manual_seed(N)
    cube = torch.randn(6, N, N, 3, generator=g).to(DEV)
    dout = torch.randn(6, N, N, 4, generator=g).to(DEV)
    ct = cm.ndf_cutoff_costheta(rough, 0.99)
    b_ref = ref_plugin.specular_bounds(N, ct)
    b = cm.specular_bounds(N, ct, DEV)
    assert torch.equal(b.view(6, N, N, 24).float(), b_ref)
    s_ref = ref_plugin.specular_cubemap_fwd(cube, b_ref, rough, ct)
    s = cm._specular_cubemap.apply(cube, rough, ct, b)
    assert _rel(s, s_ref) <= 1e-5
    ds_ref = ref_plugin.specular_cubemap_bwd(cube, b_ref, dout, rough, ct)
    lib_in = cube.clone().requires_grad_(True)
    cm._specular_cubemap.apply(lib_in, rough, ct, b).backward(dout)
    assert _rel(lib_in.grad, ds_ref) <= 1e-4
    if N <= 32:
        d_ref = ref_plugin.diffuse_cubemap_fwd(cube)
        x = cube.clone().requires_grad_(True)
        d = cm.diffuse_cubemap(x)
        assert _rel(d, d_ref) <= 1e-5
        d.backward(dout[..., :3].contiguous())
        assert _rel(x.grad, ref_plugin.diffuse_cubemap_bwd(cube, dout[..., :3].contiguous())) <= 1e-4


def test_against_numpy_oracle():
    from materialrefgs_b200 import cubemap as cm
    N, rough = 8, 0.4
    g = torch.Generator().manual_seed(2)
    cube = torch.randn(6, N, N, 3, generator=g)
    rgb_ref, out4_ref, ct = co.specular_cubemap(cube.numpy(), rough)
    out = cm.specular_cubemap(cube.to(DEV), rough).cpu().numpy()
    # N = 8 < the 16x16 culling tile: the reference's interval test empties some bounds -> 0/0
    assert np.array_equal(np.isnan(out), np.isnan(rgb_ref))
    assert np.nanmax(np.abs(out - rgb_ref)) <= 1e-5
    assert np.array_equal(cm.specular_bounds(N, ct, DEV).cpu().numpy(), co.specular_bounds(N, ct))
    assert np.abs(cm.diffuse_cubemap(cube.to(DEV)).cpu().numpy() - co.diffuse_cubemap(cube.numpy())).max() <= 1e-5


def test_cubemap_mip_forward_and_reference_backward():
    from materialrefgs_b200 import cubemap as cm
    g = torch.Generator().manual_seed(4)
    x = torch.randn(6, 32, 32, 3, generator=g).to(DEV).requires_grad_(True)
    y = cm.cubemap_mip(x)
    assert torch.allclose(y, so.cubemap_mip(x.detach()), atol=1e-6)
    dout = torch.randn(6, 16, 16, 3, generator=g).to(DEV)
    y.backward(dout)
    # scene/light_utils.py:72-80: bilinear cube fetch of dout*0.25 at the fine texel directions
    dirs = torch.from_numpy(co.texel_dirs(32)).to(DEV).reshape(-1, 3)
    ref = so.cube_texture([dout * 0.25], dirs).reshape(6, 32, 32, 3)
    assert (x.grad - ref).abs().max().item() <= 1e-5


def test_envlight_build_mips_end_to_end():
    from materialrefgs_b200.shading import EnvLight
    env = EnvLight(device=DEV, max_res=64, min_res=16, trainable=True)
    with torch.no_grad():
        env.base.copy_(torch.randn(6, 64, 64, 3, generator=torch.Generator().manual_seed(8)).to(DEV))
    env.build_mips()
    assert [tuple(l.shape) for l in env.specular] == [(6, 64, 64, 3), (6, 32, 32, 3), (6, 16, 16, 3)]
    assert tuple(env.diffuse.shape) == (6, 16, 16, 3)
    loss = sum((l ** 2).sum() for l in env.specular) + env.diffuse.sum()
    loss.backward()
    assert torch.isfinite(env.base.grad).all() and env.base.grad.abs().sum() > 0

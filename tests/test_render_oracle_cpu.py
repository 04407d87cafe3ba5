"""CPU: oracle/render_oracle.py (the restated render() contract) against vectors produced by the reference's OWN
render_initial / render_surfel / render_volume run on the CPU (tests/golden/make_golden_render.py: the CPU oracle rasterizer
stands in for the CUDA extension, the oracle's texel fetch for nvdiffrast's dr.texture). Both sides share those two
pieces, so what is pinned here is the reference's glue around them — and the GPU contract tests
(tests/test_render_contract_gpu.py) hold the product to the same restatement with the reference CUDA rasterizer."""
import glob
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from materialrefgs_b200 import synthetic
from oracle import features_oracle as fo
from oracle import raster_torch, render_oracle as ro
from oracle import shading_oracle as so

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "render_*.npz")))


def test_golden_files_present():
    assert len(GOLDEN) >= 5


def build_scene(P, W, H, view, seed, res):
    """Same construction as make_golden_render.scene()."""
    cloud = synthetic.make_cloud(P, S=8, seed=seed)
    cam = synthetic.orbit_camera(view, 8, W, H)
    raw, _ = fo.synthetic_params(P, seed=seed + 1)
    raw["xyz"] = cloud.means3D.clone()
    raw["scaling"] = torch.log(cloud.scales)
    raw["rotation"] = cloud.rotations * 1.3
    raw["opacity"] = torch.logit(cloud.opacities.clamp(1e-4, 1 - 1e-4))
    levels = so.synthetic_chain(res, 16, seed=seed + 2)
    diffuse = 0.7 * torch.randn(6, 16, 16, 3, generator=torch.Generator().manual_seed(seed + 3))
    return raw, cloud.shs.clone(), cam, levels, diffuse


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_render_oracle_matches_reference_functions(path):
    z = np.load(path)
    P, W, H, view, seed, res, srgb, indirect = (int(v) for v in z["cfg"])
    fn = str(z["fn"])
    raw, shs, cam, levels, diffuse = build_scene(P, W, H, view, seed, res)
    raw = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    shs = shs.clone().requires_grad_(True)
    levels = [l.clone().requires_grad_(True) for l in levels]
    diffuse = diffuse.clone().requires_grad_(True)
    pc = ro.RawSurfelModel(raw, shs, so.EnvLightOracle(levels, diffuse=diffuse))
    pipe = types.SimpleNamespace(depth_ratio=float(z["depth_ratio"]))
    bg = torch.tensor([0.2, 0.5, 0.8])
    kw = dict(srgb=bool(srgb))
    if fn == "render_volume":
        kw["indirect"] = bool(indirect)
    out = getattr(ro, fn)(raster_torch, cam, pc, pipe, bg, **kw)
    out["visibility_filter"] = out["radii"] > 0
    keys = [str(k) for k in z["keys"]]
    assert set(keys) <= set(out), sorted(set(keys) - set(out))
    for k in keys:
        if k == "viewspace_points":
            continue
        got, ref = out[k].detach().numpy(), z["out_" + k]
        if got.dtype == bool or np.issubdtype(got.dtype, np.integer):
            assert np.array_equal(got, ref), k
        else:
            assert np.abs(got - ref).max() <= 2e-6, (k, np.abs(got - ref).max())
    wts = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_")}
    sum((out[k] * w).sum() for k, w in wts.items()).backward()

    def close(got, ref, name):
        got = np.zeros_like(ref) if got is None else got.numpy()
        assert np.abs(got - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-12), (name, np.abs(got - ref).max(), np.abs(ref).max())
    for k, v in raw.items():
        close(v.grad, z["grad_" + k], k)
    close(shs.grad, z["grad_shs"], "shs")
    close(out["viewspace_points"].grad, z["grad_viewspace"], "viewspace_points")
    for i, l in enumerate(levels):
        close(l.grad, z[f"grad_level{i}"], f"level{i}")
    close(diffuse.grad, z["grad_diffuse_map"], "diffuse_map")

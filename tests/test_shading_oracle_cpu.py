"""CPU: the parts of oracle/shading_oracle.py that the reference's own code CAN pin here (everything that does not go
through nvdiffrast): depths_to_points / depth_to_normal against vectors from utils/point_utils.py
(tests/golden/make_golden_depth_normal.py)."""
import glob
from pathlib import Path

import numpy as np
import pytest
import torch

from materialrefgs_b200 import synthetic
from oracle import shading_oracle as so

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "depth_normal_*.npz")))


def test_golden_files_present():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_depth_to_normal_matches_reference_functions(path):
    z = np.load(path)
    view, W, H, radius = z["view"]
    cam = synthetic.orbit_camera(int(view), 8, int(W), int(H), radius=float(radius))
    depth = torch.from_numpy(z["depth"]).requires_grad_(True)
    points = so.depths_to_points(cam, depth)
    normal = so.depth_to_normal(cam, depth)
    assert np.abs(points.detach().numpy() - z["points"]).max() <= 1e-6
    assert np.abs(normal.detach().numpy() - z["normal"]).max() <= 1e-6
    (normal * torch.from_numpy(z["w"])).sum().backward()
    ref = z["grad_depth"]
    assert np.abs(depth.grad.numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    # border pixels carry no normal (point_utils.py:33-36)
    n = normal.detach()
    assert not n[0].any() and not n[-1].any() and not n[:, 0].any() and not n[:, -1].any()


# ---- the reference's own shading composition (refl_utils.py, light.py) with the texture fetch injected ------------
def _leaves(z, names):
    return {k: torch.from_numpy(z[k]).requires_grad_(True) for k in names}


@pytest.mark.parametrize("name", ["surfel_a", "surfel_b"])
def test_specular_colour_matches_reference_composition(name):
    """oracle get_specular_color_surfel + EnvLightOracle against vectors from the reference's OWN get_specular_color_surfel,
    sample_camera_rays, reflection and EnvLight.__call__ / get_mip (tests/golden/make_golden_shading.py). Both sides use
    the same restated texel fetch, so this pins everything around it: rays, clamps, mip mapping, sigmoid, weights."""
    z = np.load(ROOT / "tests" / "golden" / f"shading_{name}.npz")
    view, W, H, res = (int(v) for v in z["view"])
    cam = synthetic.orbit_camera(view, 8, W, H)
    rays_d, rays_o = so.sample_camera_rays(cam.HWK, cam.R, cam.T, "cpu")
    assert np.abs(rays_d.numpy() - z["rays_d"]).max() <= 1e-6 and np.abs(rays_o.numpy() - z["rays_o"]).max() <= 1e-6
    levels = [l.clone().requires_grad_(True) for l in so.synthetic_chain(res, 16, seed=view)]
    t = _leaves(z, ("albedo", "normal", "alpha", "refl", "rough"))
    spec, extra = so.get_specular_color_surfel(so.EnvLightOracle(levels), so.load_lut(), t["albedo"], cam.HWK, cam.R, cam.T,
                                               t["normal"], t["alpha"], t["refl"], t["rough"])
    assert np.abs(spec.detach().numpy() - z["specular"]).max() <= 1e-6
    assert np.abs(extra["direct_light"].detach().numpy() - z["direct_light"]).max() <= 1e-6
    assert np.abs(extra["specular_weight"].detach().numpy() - z["specular_weight"]).max() <= 1e-6
    (spec * torch.from_numpy(z["w"])).sum().backward()
    for k, v in t.items():
        ref = z["grad_" + k]
        assert np.abs(v.grad.numpy() - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-12), k
    for i, l in enumerate(levels):
        ref = z[f"grad_level{i}"]
        got = l.grad.numpy() if l.grad is not None else np.zeros_like(ref)
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-12), i


def test_volume_colours_match_reference_composition():
    """oracle get_full_color_volume against the reference's OWN get_full_color_volume, including its `fg[0]` indexing
    (the first surfel's LUT pair multiplies every surfel, refl_utils.py:445)."""
    z = np.load(ROOT / "tests" / "golden" / "shading_volume.npz")
    view, W, H, res = (int(v) for v in z["view"])
    cam = synthetic.orbit_camera(view, 8, W, H)
    levels = [l.clone().requires_grad_(True) for l in so.synthetic_chain(res, 16, seed=4)]
    dm = torch.from_numpy(z["diffuse_map"]).requires_grad_(True)
    t = _leaves(z, ("xyz", "normal", "albedo", "refl", "rough"))
    d, s = so.get_full_color_volume(so.EnvLightOracle(levels, diffuse=dm), so.load_lut(), t["xyz"], t["albedo"], cam, t["normal"],
                                    t["refl"], t["rough"])
    assert np.abs(d.detach().numpy() - z["diffuse"]).max() <= 1e-6 and np.abs(s.detach().numpy() - z["specular"]).max() <= 1e-6
    ((d * torch.from_numpy(z["wd"])).sum() + (s * torch.from_numpy(z["ws"])).sum()).backward()
    for k, v in t.items():
        ref = z["grad_" + k]
        assert np.abs(v.grad.numpy() - ref).max() <= 1e-5 * np.abs(ref).max(), k
    # the shared pair routes a gradient to surfel 0's roughness that no other surfel's roughness path has
    assert np.abs(dm.grad.numpy() - z["grad_diffuse_map"]).max() <= 1e-5 * np.abs(z["grad_diffuse_map"]).max()
    for i, l in enumerate(levels):
        ref = z[f"grad_level{i}"]
        assert np.abs(l.grad.numpy() - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-12), i


# ---- the PRODUCT's host-side matrices (what its kernels are fed) against the reference's own ray / point construction --
@pytest.mark.parametrize("name", ["surfel_a", "surfel_b"])
def test_product_ray_matrix_reproduces_reference_camera_rays(name):
    """materialrefgs_b200.shading.ray_matrix collapses sample_camera_rays (utils/refl_utils.py:54-73, the transposed-R
    convention, integer pixel centres) into one 3x3: normalising M @ (x, y, 1) must give the reference's rays_d, and
    _camera_origin its rays_o (vectors from the reference's own function, make_golden_shading.py)."""
    from materialrefgs_b200 import shading
    z = np.load(ROOT / "tests" / "golden" / f"shading_{name}.npz")
    view, W, H, _ = (int(v) for v in z["view"])
    cam = synthetic.orbit_camera(view, 8, W, H)
    M = shading.ray_matrix(cam.HWK, cam.R).astype(np.float64)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    d = np.stack([xs, ys, np.ones_like(xs)], -1) @ M.T
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    assert np.abs(d - z["rays_d"]).max() <= 2e-6
    assert np.abs(shading._camera_origin(cam.R, cam.T, "cpu").numpy() - z["rays_o"]).max() <= 1e-6


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_product_depth_ray_matrix_reproduces_reference_points(path):
    """materialrefgs_b200.shading.depth_ray_matrix + the origin surf_depth_normal hands to the kernel: depth * (A @ (x, y, 1))
    + origin must give the reference's depths_to_points (utils/point_utils.py:9-24, make_golden_depth_normal.py)."""
    from materialrefgs_b200 import shading
    z = np.load(path)
    view, W, H, radius = z["view"]
    W, H = int(W), int(H)
    cam = synthetic.orbit_camera(int(view), 8, W, H, radius=float(radius))
    A = shading.depth_ray_matrix(H, W, cam.tanfovx, cam.tanfovy, cam.R).astype(np.float64)
    origin = -(np.asarray(cam.R, np.float64) @ np.asarray(cam.T, np.float64))
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    rays = np.stack([xs, ys, np.ones_like(xs)], -1).reshape(-1, 3) @ A.T
    points = z["depth"].reshape(-1, 1) * rays + origin
    assert np.abs(points - z["points"]).max() <= 2e-5


def test_linear_to_srgb_matches_reference_function():
    """Product (host-side torch) and oracle linear_to_srgb against the reference's utils/graphics_utils.py:102-110,
    values and derivative, across the branch point and below zero."""
    from materialrefgs_b200 import shading
    z = np.load(ROOT / "tests" / "golden" / "linear_to_srgb.npz")
    for fn in (shading.linear_to_srgb, so.linear_to_srgb):
        x = torch.from_numpy(z["x"]).requires_grad_(True)
        y = fn(x)
        y.sum().backward()
        assert np.abs(y.detach().numpy() - z["y"]).max() <= 1e-6, fn.__module__
        assert np.abs(x.grad.numpy() - z["dy"]).max() <= 1e-4 * np.abs(z["dy"]).max(), fn.__module__

"""CPU-only: the oracle (oracle/surfel_oracle.cpp) against golden vectors produced by the
unmodified reference CUDA extension (tests/golden/make_golden.py), plus host-side logic that
needs no GPU. Tolerances: integer artefacts exact on these fixtures; floats as measured
(<= 1e-5 abs on images, <= 2e-4 relative to max-norm on gradients) with head-room."""
import glob
import re
from pathlib import Path

import numpy as np
import pytest

from materialrefgs_b200 import synthetic
from oracle import surfel_oracle as so

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "raster_*.npz")))
GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dfeatures", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat",
              "dL_dsh", "dL_dscales", "dL_drotations"]


def oracle_from_golden(z):
    return so.OracleRaster(
        means3D=z["means3D"], opacities=z["opacities"], viewmatrix=z["viewmatrix"], projmatrix=z["projmatrix"],
        campos=z["campos"], W=int(z["W"]), H=int(z["H"]), tan_fovx=float(z["tan_fovx"]),
        tan_fovy=float(z["tan_fovy"]), background=z["bg"], shs=z["shs"], features=z["features"],
        scales=z["scales"], rotations=z["rotations"], sh_degree=int(z["sh_degree"]),
        scale_modifier=float(z["scale_modifier"]))


def test_golden_files_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_oracle_matches_reference_golden(path):
    z = np.load(path)
    o = oracle_from_golden(z)
    R = o.preprocess()
    vis = z["radii"] > 0
    assert R == int(z["R"])
    assert np.array_equal(o.geom["radii"], z["radii"])
    assert np.array_equal(o.geom["tiles_touched"], z["tiles_touched"].view(np.uint32))
    assert np.array_equal(o.geom["depths"][vis].view(np.uint32), z["depths"][vis].view(np.uint32))
    assert np.abs(o.geom["means2D"][vis] - z["means2D"][vis]).max() <= 1e-4
    tnorm = np.abs(z["transMat"][vis]).max()
    assert np.abs(o.geom["transMat"][vis] - z["transMat"][vis]).max() <= 1e-6 * tnorm
    assert np.abs(o.geom["normal_opacity"][vis] - z["normal_opacity"][vis]).max() <= 1e-6
    assert np.abs(o.geom["rgb"][vis] - z["rgb"][vis]).max() <= 1e-6
    assert np.array_equal(o.geom["clamped"][vis], z["clamped"][vis])

    keys, plist, ranges = o.bin()
    assert np.array_equal(keys, z["keys"].view(np.uint64))
    assert np.array_equal(plist, z["point_list"].view(np.uint32))
    assert np.array_equal(ranges, z["ranges"].view(np.uint32))

    color, feat, others = o.forward()
    assert np.array_equal(o.n_contrib[0], z["n_contrib"].view(np.uint32)[0])
    H, W = int(z["H"]), int(z["W"])
    gx = (W + 15) // 16
    lens = (ranges[:, 1] - ranges[:, 0]).reshape(-1, gx)
    live = np.repeat(np.repeat(lens > 0, 16, 0), 16, 1)[:H, :W]   # see tests/test_raster_gpu.py
    assert np.array_equal(o.n_contrib[1][live], z["n_contrib"].view(np.uint32)[1][live])
    assert np.abs(o.final_T - z["final_T"]).max() <= 1e-5
    assert np.abs(color - z["color"]).max() <= 1e-5
    assert np.abs(others - z["others"]).max() <= 5e-5
    if feat.size:
        assert np.abs(feat - z["feature"]).max() <= 1e-5

    g = o.backward(z["dL_dcolor"], z["dL_dfeature"], z["dL_dothers"])
    for k in GRAD_NAMES:
        ref = z[k]
        if ref.size == 0:
            continue
        err = np.abs(g[k].reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err <= 5e-4, (k, err)


def test_oracle_tile_sampling_is_a_subset():
    z = np.load(GOLDEN[0])
    o = oracle_from_golden(z)
    o.preprocess(); o.bin()
    full = [a.copy() for a in o.forward()]
    part = o.forward(tile_step=3)
    H, W = int(z["H"]), int(z["W"])
    gx = (W + 15) // 16
    tile_id = (np.arange(H)[:, None] // 16) * gx + (np.arange(W)[None, :] // 16)
    sel = (tile_id % 3) == 0
    assert np.array_equal(part[0][:, sel], full[0][:, sel])
    assert not part[0][:, ~sel].any()


# ---- host-side logic that needs no GPU -----------------------------------------------------------
def test_c_abi_exports_every_declared_symbol():
    from materialrefgs_b200 import _lib
    header = (ROOT / "include" / "mrgs.h").read_text()
    declared = set(re.findall(r"MRGS_API\s+[\w\s\*]+?\b(mrgs_\w+)\s*\(", header))
    assert declared, "no MRGS_API declarations found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/mrgs.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.mrgs_abi_version() == _lib.MRGS_ABI_VERSION


def test_layout_queries_and_error_codes():
    import ctypes as C
    from materialrefgs_b200 import _lib
    lib = _lib.load()
    gl = _lib.GeomLayout()
    assert lib.mrgs_geom_layout(1000, 8, C.byref(gl)) == 0
    assert gl.cf_stride == 12 and gl.rec % 256 == 0 and gl.cf >= 1000 * 64
    assert gl.total == lib.mrgs_geom_bytes(1000, 8)
    assert lib.mrgs_geom_layout(10, 25, C.byref(gl)) != 0            # S > MAX_FEATURES
    assert b"bad arguments" in lib.mrgs_last_error()
    il = _lib.ImageLayout()
    assert lib.mrgs_image_layout(800, 800, C.byref(il)) == 0
    assert il.total >= 2500 * 5 * 256 * 4 + 2500 * 8
    assert lib.mrgs_image_layout(0, 5, C.byref(il)) != 0
    assert lib.mrgs_grad_arena_stride(8) == 28 and lib.mrgs_grad_arena_stride(0) == 20
    assert lib.mrgs_grad_arena_stride(2) >= 15 + 8                    # padded channels fit the row
    slots = {lib.mrgs_tile_slot(x, y) for x in range(16) for y in range(16)}
    assert slots == set(range(256))
    # null / inconsistent arguments are rejected before anything is launched
    a = _lib.ForwardArgs()
    a.P, a.width, a.height = 10, 64, 64
    assert lib.mrgs_forward(C.byref(a), None) == 1
    assert b"exactly one of" in lib.mrgs_last_error()


def test_api_argument_validation_without_gpu():
    import torch
    from materialrefgs_b200.diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    rs = GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3,
                                       torch.zeros(3), False, False)
    r = GaussianRasterizer(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1))
    with pytest.raises(Exception, match="scale/rotation pair"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 16, 3))
    with pytest.raises(RuntimeError, match="CUDA tensor"):   # no silent CPU fallback
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 16, 3),
          scales=torch.ones(4, 2), rotations=torch.ones(4, 4))


def test_synthetic_generator_is_deterministic():
    from materialrefgs_b200 import synthetic
    a = synthetic.make_cloud(1000, S=8)
    b = synthetic.make_cloud(1000, S=8)
    assert all((getattr(a, k) == getattr(b, k)).all() for k in ("means3D", "scales", "rotations", "shs"))
    cam = synthetic.orbit_camera(1, 8, 800, 800)
    assert abs(cam.K[0, 0] - 800 / (2 * np.tan(0.6911 / 2))) < 1e-3
    assert np.allclose(cam.world_view_transform.numpy()[:3, :3].T @ cam.R, np.eye(3), atol=1e-5)


def test_synthetic_cameras_follow_the_reference_conventions():
    """tests/golden/camera_matrices.npz comes from the reference's own utils/graphics_utils.py composed as
    scene/cameras.py:70-84 does (make_golden_cameras.py): transposed world-to-view, projection, their product, the
    camera centre and the pinhole focal lengths of the synthetic views used everywhere in tests / bench / tools."""
    z = np.load(ROOT / "tests" / "golden" / "camera_matrices.npz")
    k = 0
    while f"view{k}" in z.files:
        i, n, W, H, radius = z[f"view{k}"]
        cam = synthetic.orbit_camera(int(i), int(n), int(W), int(H), radius=float(radius))
        # getWorld2View2 inverts the pose twice (translate = 0, scale = 1): equal up to float32 rounding of that round trip
        assert np.allclose(cam.world_view_transform.numpy(), z[f"w2v{k}"], atol=1e-6)
        assert np.allclose(cam.full_proj_transform.numpy(), z[f"full{k}"], atol=1e-5)
        assert np.allclose(cam.camera_center.numpy(), z[f"center{k}"], atol=1e-5)
        K = np.asarray(cam.HWK[2], np.float64)
        assert np.allclose([K[0, 0], K[1, 1]], z[f"focal{k}"], rtol=1e-6)
        # principal point at the image centre: the K-based projection the reference uses when HWK is given is the same
        assert np.allclose(z[f"projK{k}"], z[f"proj{k}"], atol=1e-6)
        k += 1
    assert k == 3

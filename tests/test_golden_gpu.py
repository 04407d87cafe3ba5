"""GPU: libmrgs through the C ABI against (a) the committed golden vectors of the reference and
(b) the CPU oracle on fresh seeded inputs the golden files do not cover."""
import glob
from pathlib import Path

import numpy as np
import pytest
import torch

from materialrefgs_b200 import synthetic
from tests import refimpl
from tests.test_oracle_cpu import GRAD_NAMES

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "raster_*.npz")))


def _fwd_bwd(z, dev):
    import materialrefgs_b200.rasterizer as ours
    t = lambda k: torch.from_numpy(np.ascontiguousarray(z[k])).to(dev)
    e = torch.empty(0, device=dev)
    H, W = int(z["H"]), int(z["W"])
    out = ours.rasterize_forward_raw(
        t("bg"), t("means3D"), e, t("features"), t("opacities"), t("scales"), t("rotations"),
        float(z["scale_modifier"]), e, t("viewmatrix"), t("projmatrix"), float(z["tan_fovx"]),
        float(z["tan_fovy"]), H, W, t("shs"), int(z["sh_degree"]), t("campos"), False, True)
    R, contrib, color, feat, others, radii, geom, binning, img = out
    grads = ours.rasterize_backward_raw(
        t("bg"), t("means3D"), radii, e, t("features"), t("scales"), t("rotations"), float(z["scale_modifier"]),
        e, t("viewmatrix"), t("projmatrix"), float(z["tan_fovx"]), float(z["tan_fovy"]), t("dL_dcolor"),
        t("dL_dfeature"), t("dL_dothers"), t("shs"), int(z["sh_degree"]), t("campos"), geom, R, binning, img,
        contrib, True)
    return out, grads


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_against_reference_golden(path):
    dev = torch.device("cuda:0")
    z = np.load(path)
    (R, contrib, color, feat, others, radii, geom, binning, img), grads = _fwd_bwd(z, dev)
    P, S, H, W = z["means3D"].shape[0], int(z["S"]), int(z["H"]), int(z["W"])
    assert R == int(z["R"])
    assert np.array_equal(radii.cpu().numpy(), z["radii"])
    vis = torch.from_numpy(z["radii"] > 0).to(dev)
    gm = refimpl.decode_mrgs_geom(geom, P, S)
    n = lambda x: x.cpu().numpy()
    assert np.array_equal(n(gm["tiles_touched"]), z["tiles_touched"])
    v = n(vis)
    assert np.array_equal(n(gm["depths"])[v].view(np.uint32), z["depths"][v].view(np.uint32))
    assert np.array_equal(n(gm["means2D"])[v], z["means2D"][v])
    assert np.array_equal(n(gm["transMat"])[v], z["transMat"][v])
    bm = refimpl.decode_mrgs_binning(binning, R, gm["depths"])
    assert np.array_equal(n(bm["keys"]), z["keys"])
    assert np.array_equal(n(bm["point_list"]), z["point_list"])
    im = refimpl.decode_mrgs_image(img, H, W)
    assert np.array_equal(n(im["ranges"]), z["ranges"])
    assert np.array_equal(n(im["n_contrib"]), z["n_contrib"][0])
    assert np.array_equal(n(im["final_T"]).view(np.uint32), z["final_T"][0].view(np.uint32))
    assert np.abs(n(color) - z["color"]).max() <= 1e-4
    assert np.abs(n(others) - z["others"]).max() <= 1e-4
    if S:
        assert np.abs(n(feat) - z["feature"]).max() <= 1e-4
    for k, g in zip(GRAD_NAMES, grads):
        ref = z[k]
        if ref.size == 0:
            continue
        err = np.abs(n(g).reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err <= 1e-3, (k, err)


@pytest.mark.parametrize("P,S,W,H,opacity,seed", [
    (3000, 8, 128, 96, "trained", 11),
    (1200, 0, 70, 45, "init", 12),       # ragged, empty corners
    (1, 3, 33, 33, "trained", 13),       # a single surfel
])
def test_against_cpu_oracle(P, S, W, H, opacity, seed):
    from oracle import surfel_oracle as so
    import materialrefgs_b200.rasterizer as ours
    dev = torch.device("cuda:0")
    cloud = synthetic.make_cloud(P, S=S, opacity=opacity, seed=seed, scale_mult=2.0 if P < 10 else 1.0)
    if P == 1:
        cloud.means3D[:] = 0.0
        cloud.scales[:] = 0.4
    cam = synthetic.orbit_camera(2, 8, W, H)
    gc, gf, go = synthetic.upstream_grads(S, H, W, seed=seed)
    bg = np.array([0.3, 0.6, 0.1], np.float32)
    o = so.from_synthetic(cloud, cam, bg=bg)
    R_cpu = o.preprocess(); o.bin()
    c_cpu, f_cpu, a_cpu = o.forward()
    g_cpu = o.backward(gc.numpy(), gf.numpy(), go.numpy())

    cl, cm = cloud.to(dev), cam.to(dev)
    e = torch.empty(0, device=dev)
    bgt = torch.from_numpy(bg).to(dev)
    R, contrib, color, feat, others, radii, geom, binning, img = ours.rasterize_forward_raw(
        bgt, cl.means3D, e, cl.features, cl.opacities, cl.scales, cl.rotations, 1.0, e, cm.world_view_transform,
        cm.full_proj_transform, cm.tanfovx, cm.tanfovy, H, W, cl.shs, 3, cm.camera_center, False, False)
    grads = ours.rasterize_backward_raw(
        bgt, cl.means3D, radii, e, cl.features, cl.scales, cl.rotations, 1.0, e, cm.world_view_transform,
        cm.full_proj_transform, cm.tanfovx, cm.tanfovy, gc.to(dev), gf.to(dev), go.to(dev), cl.shs, 3,
        cm.camera_center, geom, R, binning, img, contrib, False)
    # rsqrt / exp differ between host libm and the GPU: allow a vanishing fraction of off-by-one radii
    mism = (radii.cpu().numpy() != o.geom["radii"]).mean()
    assert mism <= 1e-3
    if mism == 0:
        assert R == R_cpu
        gm = refimpl.decode_mrgs_geom(geom, P, S)
        bm = refimpl.decode_mrgs_binning(binning, R, gm["depths"])
        assert np.array_equal(bm["keys"].cpu().numpy().view(np.uint64), o.keys)
        assert np.array_equal(bm["point_list"].cpu().numpy().view(np.uint32), o.point_list)
        im = refimpl.decode_mrgs_image(img, H, W)
        assert (im["n_contrib"].cpu().numpy().view(np.uint32) != o.n_contrib[0]).mean() <= 1e-3
    assert np.abs(color.cpu().numpy() - c_cpu).max() <= 1e-4
    assert np.abs(others.cpu().numpy() - a_cpu).max() <= 1e-4
    for k, g in zip(GRAD_NAMES, grads):
        ref = g_cpu[k]
        if ref.size == 0:
            continue
        err = np.abs(g.cpu().numpy().reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err <= 1e-3, (k, err)

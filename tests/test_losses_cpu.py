"""CPU: oracle/losses_oracle.py against vectors produced by the reference's own l1_loss / ssim."""
import glob
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import losses_oracle as lo

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "losses_*.npz")))


def test_golden_files_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_oracle_matches_reference_functions(path):
    z = np.load(path)
    img = torch.from_numpy(z["img"]).requires_grad_(True)
    gt = torch.from_numpy(z["gt"])
    l1, s = lo.l1_loss(img, gt), lo.ssim(img, gt)
    assert abs(l1.item() - float(z["l1"])) <= 1e-7 and abs(s.item() - float(z["ssim"])) <= 1e-6
    g_l1, = torch.autograd.grad(l1, img, retain_graph=True)
    g_s, = torch.autograd.grad(s, img)
    assert np.abs(g_l1.numpy() - z["grad_l1"]).max() <= 1e-9
    assert np.abs(g_s.numpy() - z["grad_ssim"]).max() <= 1e-6 * np.abs(z["grad_ssim"]).max() + 1e-9


# ---- geometric regularisers (calculate_loss, first_order_edge_aware_loss, get_img_grad_weight) ----------------
GEOM = sorted(glob.glob(str(ROOT / "tests" / "golden" / "geomloss_*.npz")))
LEAVES = ("render", "rend_normal", "surf_normal", "surf_depth", "rend_dist")


def golden_opt(z):
    import types
    o = types.SimpleNamespace(lambda_dssim=0.2, lambda_dist=0.0, lambda_normal_render_depth=0.05, lambda_normal_smooth=0.0,
                              lambda_depth_smooth=0.0, normal_loss_start=0, dist_loss_start=3000, normal_smooth_from_iter=0,
                              normal_smooth_until_iter=18000, use_perceptual_loss=False, perceptual_loss_start_iter=18000)
    for k, v in zip(z["opt_keys"], z["opt_vals"]):
        setattr(o, str(k), float(v))
    return o


def test_geomloss_golden_files_present():
    assert len(GEOM) >= 3


@pytest.mark.parametrize("path", GEOM, ids=lambda p: Path(p).stem)
def test_oracle_matches_reference_calculate_loss(path):
    z = np.load(path)
    gt = torch.from_numpy(z["gt"])
    leaves = {k: torch.from_numpy(z[k]).requires_grad_(True) for k in LEAVES}
    w = lo.get_img_grad_weight(gt)
    assert np.abs(w.numpy() - z["grad_weight"]).max() <= 1e-7
    iw = (1.0 - w).clamp(0, 1) ** 2 if bool(z["weighted"]) else None
    loss = lo.calculate_loss(gt, leaves, golden_opt(z), int(z["iteration"]), iw)
    assert abs(loss.item() - float(z["loss"])) <= 1e-6
    loss.backward()
    for k, v in leaves.items():
        ref = z["grad_" + k]
        got = v.grad.numpy() if v.grad is not None else np.zeros_like(ref)
        assert np.abs(got - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1e-12) + 1e-12, k
    assert abs(lo.first_order_edge_aware_loss(leaves["rend_normal"].detach(), gt).item() - float(z["edge_normal"])) <= 1e-6
    assert abs(lo.first_order_edge_aware_loss(leaves["surf_depth"].detach(), gt).item() - float(z["edge_depth"])) <= 1e-6


def test_spatial_gradient_restates_a_replicate_padded_normalised_sobel():
    # independent check of the kornia restatement: on a linear ramp a*x + b*y the normalised Sobel pair is (a, b) in
    # the interior and half of it across a replicated border
    H, W = 7, 9
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    g = lo.spatial_gradient((0.5 * xx - 2.0 * yy)[None, None])[0, 0]
    assert torch.allclose(g[0, 1:-1, 1:-1], torch.full((H - 2, W - 2), 0.5))
    assert torch.allclose(g[1, 1:-1, 1:-1], torch.full((H - 2, W - 2), -2.0))
    assert torch.allclose(g[0, 1:-1, 0], torch.full((H - 2,), 0.25)) and torch.allclose(g[1, 0, 1:-1], torch.full((W - 2,), -1.0))

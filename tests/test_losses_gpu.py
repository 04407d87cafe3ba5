"""GPU: fused L1 + SSIM (mrgs_photometric_*) through the C ABI against (a) vectors produced by the
reference's own l1_loss / ssim and (b) the torch oracle at BASELINE image size."""
import glob
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import losses_oracle as lo

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(glob.glob(str(ROOT / "tests" / "golden" / "losses_*.npz")))
VAL_ATOL = 2e-6
GRAD_RTOL = 1e-4     # relative to the gradient's max-norm


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: Path(p).stem)
def test_against_reference_function_vectors(path):
    from materialrefgs_b200 import losses
    dev = torch.device("cuda:0")
    z = np.load(path)
    gt = torch.from_numpy(z["gt"]).to(dev)
    img = torch.from_numpy(z["img"]).to(dev).requires_grad_(True)
    l1, s = losses.l1_ssim(img, gt)
    assert abs(l1.item() - float(z["l1"])) <= VAL_ATOL and abs(s.item() - float(z["ssim"])) <= VAL_ATOL
    g_l1, = torch.autograd.grad(l1, img, retain_graph=True)
    g_s, = torch.autograd.grad(s, img)
    assert np.abs(g_l1.cpu().numpy() - z["grad_l1"]).max() <= 1e-9
    ref = z["grad_ssim"]
    assert np.abs(g_s.cpu().numpy() - ref).max() <= GRAD_RTOL * np.abs(ref).max()


def test_full_size_loss_and_gradient_vs_oracle():
    from materialrefgs_b200 import losses
    dev = torch.device("cuda:0")
    img, gt = lo.synthetic_pair(3, 800, 800, seed=9)
    a = img.to(dev).requires_grad_(True)
    b = img.clone().requires_grad_(True)      # the oracle runs on the CPU: cuDNN convolutions default to TF32
    gtd = gt.to(dev)
    ours = losses.photometric_loss(a, gtd, 0.2)
    ref = lo.photometric_loss(b, gt, 0.2)
    assert abs(ours.item() - ref.item()) <= VAL_ATOL
    ours.backward()
    ref.backward()
    assert ((a.grad.cpu() - b.grad).abs().max() / b.grad.abs().max()).item() <= GRAD_RTOL
    # same-named wrappers, no-grad path, determinism
    with torch.no_grad():
        s1 = losses.ssim(a, gtd)
        s2 = losses.ssim(a, gtd)
        assert torch.equal(s1, s2) and abs(s1.item() - lo.ssim(b.detach(), gt).item()) <= VAL_ATOL
        assert abs(losses.l1_loss(a, gtd).item() - lo.l1_loss(b.detach(), gt).item()) <= VAL_ATOL
    assert abs(losses.ssim(gtd, gtd).item() - 1.0) <= 1e-6


def test_argument_checks():
    from materialrefgs_b200 import losses
    dev = torch.device("cuda:0")
    x = torch.rand(3, 20, 20)
    with pytest.raises(RuntimeError):
        losses.l1_ssim(x, x)
    with pytest.raises(RuntimeError):
        losses.l1_ssim(x.to(dev), torch.rand(3, 20, 21, device=dev))
    with pytest.raises(NotImplementedError):
        losses.ssim(x.to(dev), x.to(dev), window_size=7)


# ---- geometric regularisers (mrgs_geometry_loss_*, mrgs_img_grad_weight) --------------------------------------
GEOM = sorted(glob.glob(str(ROOT / "tests" / "golden" / "geomloss_*.npz")))
LEAVES = ("render", "rend_normal", "surf_normal", "surf_depth", "rend_dist")


class _Cam:
    def __init__(self, gt):
        self.original_image = gt


@pytest.mark.parametrize("path", GEOM, ids=lambda p: Path(p).stem)
def test_calculate_loss_against_reference_vectors(path):
    import types
    from materialrefgs_b200 import losses
    from tests.test_losses_cpu import golden_opt
    dev = torch.device("cuda:0")
    z = np.load(path)
    gt = torch.from_numpy(z["gt"]).to(dev)
    leaves = {k: torch.from_numpy(z[k]).to(dev).requires_grad_(True) for k in LEAVES}
    w = losses.get_img_grad_weight(gt)
    assert np.abs(w.cpu().numpy() - z["grad_weight"]).max() <= 1e-6
    iw = (1.0 - w).clamp(0, 1) ** 2 if bool(z["weighted"]) else None
    pc = types.SimpleNamespace(get_xyz=torch.zeros(4, 3))
    loss, tb = losses.calculate_loss(_Cam(gt), pc, leaves, golden_opt(z), int(z["iteration"]), iw)
    assert abs(loss.item() - float(z["loss"])) <= 5e-6
    assert abs(tb["loss"] - float(z["loss"])) <= 5e-6 and tb["num_points"] == 4
    loss.backward()
    for k, v in leaves.items():
        ref = z["grad_" + k]
        got = v.grad.cpu().numpy() if v.grad is not None else np.zeros_like(ref)
        assert np.abs(got - ref).max() <= GRAD_RTOL * max(np.abs(ref).max(), 1e-12) + 1e-12, k
    with torch.no_grad():
        assert abs(losses.first_order_edge_aware_loss(leaves["rend_normal"], gt).item() - float(z["edge_normal"])) <= VAL_ATOL
        assert abs(losses.first_order_edge_aware_loss(leaves["surf_depth"], gt).item() - float(z["edge_depth"])) <= VAL_ATOL


def test_geometry_losses_full_size_vs_oracle():
    import types
    from materialrefgs_b200 import losses
    dev = torch.device("cuda:0")
    pkg, gt = lo.synthetic_render_pkg(800, 800, seed=21)
    opt = types.SimpleNamespace(lambda_dssim=0.2, lambda_dist=100.0, lambda_normal_render_depth=0.05, lambda_normal_smooth=0.01,
                                lambda_depth_smooth=0.02, normal_loss_start=0, dist_loss_start=3000, normal_smooth_from_iter=0,
                                normal_smooth_until_iter=18000, use_perceptual_loss=False, perceptual_loss_start_iter=18000)
    a = {k: pkg[k].to(dev).requires_grad_(True) for k in LEAVES}
    b = {k: pkg[k].clone().requires_grad_(True) for k in LEAVES}
    gtd = gt.to(dev)
    wd = losses.get_img_grad_weight(gtd)
    wc = lo.get_img_grad_weight(gt)
    assert (wd.cpu() - wc).abs().max().item() <= 1e-6
    pc = types.SimpleNamespace(get_xyz=torch.zeros(4, 3))
    ours, tb = losses.calculate_loss(_Cam(gtd), pc, a, opt, 5000, (1.0 - wd).clamp(0, 1) ** 2)
    ref = lo.calculate_loss(gt, b, opt, 5000, (1.0 - wc).clamp(0, 1) ** 2)
    assert abs(ours.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    ours.backward()
    ref.backward()
    for k in LEAVES:
        assert ((a[k].grad.cpu() - b[k].grad).abs().max() / b[k].grad.abs().max()).item() <= GRAD_RTOL, k
    # determinism of the fused sums, ragged image whose edges are not multiples of the CTA block
    with torch.no_grad():
        t1 = losses.geometry_losses(a["rend_normal"], a["surf_normal"], a["rend_dist"], a["surf_depth"], gtd, None,
                                    normal=True, dist=True, normal_smooth=True, depth_smooth=True)
        t2 = losses.geometry_losses(a["rend_normal"], a["surf_normal"], a["rend_dist"], a["surf_depth"], gtd, None,
                                    normal=True, dist=True, normal_smooth=True, depth_smooth=True)
        assert torch.equal(t1, t2)
    with pytest.raises(NotImplementedError):
        opt.use_perceptual_loss = True
        losses.calculate_loss(_Cam(gtd), pc, a, opt, 20000, None)
    with pytest.raises(RuntimeError):
        losses.get_img_grad_weight(torch.rand(3, 2, 9, device=dev))

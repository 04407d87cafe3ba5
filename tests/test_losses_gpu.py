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

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

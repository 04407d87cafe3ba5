"""Torch restatement of the reference's photometric loss terms.
TEST INFRASTRUCTURE ONLY (only tests/, __graft_entry__.smoke() and bench.py's baseline legs import it).

Pinned by tests/golden/losses_*.npz: values AND image gradients produced by the reference's OWN
utils/loss_utils.py l1_loss / ssim, imported from /root/reference by tests/golden/make_golden_losses.py
(kornia and matplotlib, which that module imports but these functions do not use, are stubbed there).

What it follows: utils/loss_utils.py:22-23 (l1_loss), :28-30 (gaussian), :77-81 (create_window),
:83-119 (ssim / _ssim), :155-157 (how calculate_loss combines them).
"""
from __future__ import annotations

from math import exp

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt):
    return torch.abs(network_output - gt).mean()


def gaussian(window_size: int, sigma: float) -> torch.Tensor:
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return g / g.sum()


def ssim(img1, img2, window_size: int = 11):
    """Mean SSIM of two [C,H,W] images: zero-padded depthwise 11x11 Gaussian filtering of the five moments."""
    C = img1.shape[-3]
    g = gaussian(window_size, 1.5).unsqueeze(1)
    window = (g @ g.t()).float()[None, None].expand(C, 1, window_size, window_size).contiguous().to(img1)
    conv = lambda t: F.conv2d(t, window, padding=window_size // 2, groups=C)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = conv(img1 * img1) - mu1_sq
    s2 = conv(img2 * img2) - mu2_sq
    s12 = conv(img1 * img2) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


def photometric_loss(image, gt, lambda_dssim: float):
    """loss0 of calculate_loss (loss_utils.py:155-157)."""
    return (1.0 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1.0 - ssim(image, gt))


def synthetic_pair(C: int, H: int, W: int, seed: int = 0):
    """A smooth-ish 'render' and a 'ground truth' that differs by structured noise (SSIM away from 0 and 1)."""
    g = torch.Generator().manual_seed(seed)
    base = F.interpolate(torch.rand(1, C, H // 8 + 2, W // 8 + 2, generator=g), size=(H, W), mode="bilinear",
                         align_corners=False)[0]
    gt = (base + 0.05 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    img = (base * 0.9 + 0.05 + 0.08 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    return img, gt

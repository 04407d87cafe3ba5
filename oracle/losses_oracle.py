"""Torch restatement of the reference's photometric loss terms.
TEST INFRASTRUCTURE ONLY (only tests/, __graft_entry__.smoke() and bench.py's baseline legs import it).

Pinned by tests/golden/losses_*.npz: values AND image gradients produced by the reference's OWN
utils/loss_utils.py l1_loss / ssim, imported from /root/reference by tests/golden/make_golden_losses.py
(kornia and matplotlib, which that module imports but these functions do not use, are stubbed there).

What it follows: utils/loss_utils.py:22-23 (l1_loss), :28-30 (gaussian), :77-81 (create_window),
:83-119 (ssim / _ssim), :155-157 (how calculate_loss combines them); :121-122 (first_order_edge_aware_loss),
:127-139 (get_img_grad_weight), :142-228 (calculate_loss, everything but the lpips branch).

kornia is NOT installed here and is not part of /root/reference (requirements.txt:45 pins kornia==0.7.3):
`spatial_gradient` below restates that release's published kornia/filters/sobel.py (mode='sobel', order=1,
normalized=True): kernels [[-1,0,1],[-2,0,2],[-1,0,1]] and its transpose, each divided by the sum of absolute values
(8), replicate padding, plain cross-correlation, output [B,C,2,H,W] = (d/dx, d/dy). PARITY UNPINNED for that one
function; everything AROUND it (abs / exp / sum / mean, the term selection and lambdas of calculate_loss,
get_img_grad_weight) is pinned by tests/golden/geomloss_*.npz, produced by the reference's own calculate_loss with this
spatial_gradient injected for the missing kornia symbol (tests/golden/make_golden_geomloss.py).
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


def spatial_gradient(inp: torch.Tensor) -> torch.Tensor:
    """kornia 0.7.3 spatial_gradient(input, mode='sobel', order=1, normalized=True): [B,C,H,W] -> [B,C,2,H,W]."""
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=inp.dtype, device=inp.device)
    kernel = torch.stack([kx, kx.t()])
    kernel = kernel / kernel.abs().sum(dim=-1).sum(dim=-1)[..., None, None]
    b, c, h, w = inp.shape
    padded = F.pad(inp.reshape(b * c, 1, h, w), [1, 1, 1, 1], mode="replicate")
    out = F.conv2d(padded, kernel[:, None], groups=1, padding=0, stride=1)
    return out.reshape(b, c, 2, h, w)


def first_order_edge_aware_loss(data, img):
    """loss_utils.py:121-122."""
    return (spatial_gradient(data[None])[0].abs() * torch.exp(-spatial_gradient(img[None])[0].abs())).sum(1).mean()


def get_img_grad_weight(img):
    """loss_utils.py:127-139."""
    _, hd, wd = img.shape
    bottom = img[..., 2:hd, 1:wd - 1]
    top = img[..., 0:hd - 2, 1:wd - 1]
    right = img[..., 1:hd - 1, 2:wd]
    left = img[..., 1:hd - 1, 0:wd - 2]
    gx = torch.mean(torch.abs(right - left), 0, keepdim=True)
    gy = torch.mean(torch.abs(top - bottom), 0, keepdim=True)
    g, _ = torch.max(torch.cat((gx, gy), dim=0), dim=0)
    g = (g - g.min()) / (g.max() - g.min())
    return F.pad(g[None, None], (1, 1, 1, 1), mode="constant", value=1.0).squeeze()


def calculate_loss(gt_image, render_pkg, opt, iteration, image_weight=None):
    """The value of loss_utils.py:142-228 (without the lpips branch and the logging dictionary)."""
    rendered_image = render_pkg["render"]
    rendered_normal = render_pkg["rend_normal"]
    loss = photometric_loss(rendered_image, gt_image, opt.lambda_dssim)
    if opt.lambda_normal_render_depth > 0 and iteration > opt.normal_loss_start:
        surf_normal = render_pkg["surf_normal"]
        if image_weight is not None:
            ln = (image_weight * (surf_normal - rendered_normal).abs().sum(0)).mean()
        else:
            ln = (1 - (rendered_normal * surf_normal).sum(dim=0))[None].mean()
        loss = loss + opt.lambda_normal_render_depth * ln
    if opt.lambda_dist > 0 and iteration > opt.dist_loss_start:
        loss = loss + opt.lambda_dist * render_pkg["rend_dist"].mean()
    if opt.lambda_normal_smooth > 0 and opt.normal_smooth_from_iter < iteration < opt.normal_smooth_until_iter:
        loss = loss + opt.lambda_normal_smooth * first_order_edge_aware_loss(rendered_normal, gt_image)
    if opt.lambda_depth_smooth > 0 and iteration > 3000:
        loss = loss + opt.lambda_depth_smooth * first_order_edge_aware_loss(render_pkg["surf_depth"], gt_image)
    return loss


def synthetic_render_pkg(H: int, W: int, seed: int = 0):
    """Smooth-ish maps shaped like render_surfel's dictionary entries that calculate_loss reads."""
    g = torch.Generator().manual_seed(seed)
    img, gt = synthetic_pair(3, H, W, seed)

    def smooth(c, amp=1.0, noise=0.05):
        base = F.interpolate(torch.randn(1, c, H // 6 + 2, W // 6 + 2, generator=g), size=(H, W), mode="bilinear",
                             align_corners=False)[0]
        return amp * base + noise * torch.randn(c, H, W, generator=g)

    rn = F.normalize(smooth(3), dim=0) * torch.rand(1, H, W, generator=g)
    sn = F.normalize(rn + 0.2 * smooth(3), dim=0)
    depth = 3.0 + smooth(1, 0.5, 0.01)
    dist = smooth(1, 0.01, 0.001).abs()
    return dict(render=img, rend_normal=rn, surf_normal=sn, surf_depth=depth, rend_dist=dist,
                rend_alpha=torch.rand(1, H, W, generator=g),
                visibility_filter=torch.ones(4, dtype=torch.bool)), gt

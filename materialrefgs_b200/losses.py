"""Photometric loss terms of the reference's calculate_loss on libmrgs kernels.

Same names and meaning as utils/loss_utils.py: l1_loss (:22-23), ssim (:83-119, size_average=True),
plus photometric_loss = (1 - lambda_dssim) * l1 + lambda_dssim * (1 - ssim) (:155-157). Both terms come from
ONE forward kernel (and one backward), whichever of the two a caller asks for.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _img(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    if t.dim() == 4 and t.shape[0] == 1:
        t = t[0]
    if t.dim() != 3:
        raise RuntimeError(f"{name} must have shape [C,H,W] (or [1,C,H,W])")
    return t.contiguous()


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, gt):
        lib = _lib.load()
        img, gt = _img(img, "img"), _img(gt.detach(), "gt")
        if img.shape != gt.shape:
            raise RuntimeError(f"img {tuple(img.shape)} and gt {tuple(gt.shape)} differ in shape")
        Cn, H, W = img.shape
        dev = img.device
        need_grad = ctx.needs_input_grad[0]
        maps = torch.empty((3, Cn, H, W), dtype=torch.float32, device=dev) if need_grad else None
        partials = torch.empty(lib.mrgs_photometric_partials_bytes(Cn, H, W), dtype=torch.uint8, device=dev)
        out2 = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mrgs_photometric_forward(img.data_ptr(), gt.data_ptr(), Cn, H, W,
                                                    maps.data_ptr() if maps is not None else None,
                                                    partials.data_ptr(), out2.data_ptr(), C.c_void_p(stream)),
                       "mrgs_photometric_forward")
        if need_grad:
            ctx.save_for_backward(img, gt, maps)
        return out2[0], out2[1]

    @staticmethod
    def backward(ctx, g_l1, g_ssim):
        lib = _lib.load()
        img, gt, maps = ctx.saved_tensors
        Cn, H, W = img.shape
        dev = img.device
        up = torch.stack((g_l1.reshape(()), g_ssim.reshape(()))).to(torch.float32).contiguous()
        dimg = torch.empty_like(img)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mrgs_photometric_backward(img.data_ptr(), gt.data_ptr(), maps.data_ptr(), Cn, H, W,
                                                     up.data_ptr(), dimg.data_ptr(), C.c_void_p(stream)),
                       "mrgs_photometric_backward")
        return dimg, None


def l1_ssim(img: torch.Tensor, gt: torch.Tensor):
    """(mean |img - gt|, mean SSIM) of two [C,H,W] images as 0-dim tensors; differentiable w.r.t. img."""
    return _L1SSIM.apply(img, gt)


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    return l1_ssim(network_output, gt)[0]


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    if window_size != 11 or not size_average:
        raise NotImplementedError("only the reference's default window (11, sigma 1.5) with size_average=True")
    return l1_ssim(img1, img2)[1]


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float) -> torch.Tensor:
    """loss0 of calculate_loss: (1 - lambda_dssim) * Ll1 + lambda_dssim * (1 - ssim)."""
    l1, s = l1_ssim(image, gt)
    return (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - s)

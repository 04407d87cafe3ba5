"""The reference's calculate_loss on libmrgs kernels.

Same names and meaning as utils/loss_utils.py: l1_loss (:22-23), ssim (:83-119, size_average=True),
photometric_loss = (1 - lambda_dssim) * l1 + lambda_dssim * (1 - ssim) (:155-157) — both terms from ONE forward
kernel (and one backward); first_order_edge_aware_loss (:121-122), get_img_grad_weight (:127-139) and
calculate_loss (:142-228) — the normal-consistency, distortion and the two edge-aware smoothness terms from ONE more
forward/backward kernel pair (mrgs_geometry_loss_*). The lpips branch (an external VGG network) stays on the reference.
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


# ---- geometric regularisers ---------------------------------------------------------------------------------
def _map(t, name: str, channels: int):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float32 CUDA tensor")
    if t.dim() == 2 and channels == 1:
        t = t[None]
    if t.dim() != 3 or t.shape[0] != channels:
        raise RuntimeError(f"{name} must have shape [{channels},H,W], got {tuple(t.shape)}")
    return t.contiguous()


class _GeometryLoss(torch.autograd.Function):
    """out4 = (normal consistency, mean distortion, edge-aware normal smoothness, edge-aware depth smoothness)."""

    @staticmethod
    def forward(ctx, terms, rend_normal, surf_normal, rend_dist, surf_depth, gt_image, image_weight):
        lib = _lib.load()
        maps = dict(rend_normal=_map(rend_normal, "rend_normal", 3), surf_normal=_map(surf_normal, "surf_normal", 3),
                    rend_dist=_map(rend_dist, "rend_dist", 1), surf_depth=_map(surf_depth, "surf_depth", 1),
                    gt_image=_map(gt_image, "gt_image", 3), image_weight=_map(image_weight, "image_weight", 1))
        first = next((m for m in maps.values() if m is not None), None)
        if first is None:
            raise RuntimeError("geometry_losses needs at least one map")
        H, W = first.shape[-2:]
        for k, m in maps.items():
            if m is not None and tuple(m.shape[-2:]) != (H, W):
                raise RuntimeError(f"{k} is {tuple(m.shape)}, expected [*,{H},{W}]")
        dev = first.device
        need_grad = any(ctx.needs_input_grad[1:5])
        smooth = terms & (_lib.GEOM_NORMAL_SMOOTH | _lib.GEOM_DEPTH_SMOOTH)
        coef = torch.empty((8, H, W), dtype=torch.float32, device=dev) if (need_grad and smooth) else None
        partials = torch.empty(lib.mrgs_geometry_loss_partials_bytes(H, W), dtype=torch.uint8, device=dev)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        a = _lib.GeometryLossArgs()
        a.height, a.width, a.terms = H, W, terms
        for k, m in maps.items():
            setattr(a, k, m.data_ptr() if m is not None else None)
        a.coef = coef.data_ptr() if coef is not None else None
        a.partials, a.out4 = partials.data_ptr(), out4.data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mrgs_geometry_loss_forward(C.byref(a), C.c_void_p(stream)), "mrgs_geometry_loss_forward")
        ctx.terms, ctx.shape = terms, (H, W)
        ctx.present = [maps[k] is not None for k in ("rend_normal", "surf_normal", "rend_dist", "surf_depth")]
        ctx.save_for_backward(*[t for t in (*maps.values(), coef) if t is not None])
        ctx.layout = [k for k, t in (*maps.items(), ("coef", coef)) if t is not None]
        return out4

    @staticmethod
    def backward(ctx, g_out4):
        lib = _lib.load()
        saved = dict(zip(ctx.layout, ctx.saved_tensors))
        H, W = ctx.shape
        dev = g_out4.device
        a = _lib.GeometryLossArgs()
        a.height, a.width, a.terms = H, W, ctx.terms
        for k in ("rend_normal", "surf_normal", "rend_dist", "surf_depth", "gt_image", "image_weight", "coef"):
            setattr(a, k, saved[k].data_ptr() if k in saved else None)
        up = g_out4.to(torch.float32).contiguous()
        a.upstream = up.data_ptr()
        grads = []
        for i, (k, ch) in enumerate((("rend_normal", 3), ("surf_normal", 3), ("rend_dist", 1), ("surf_depth", 1))):
            g = None
            if ctx.needs_input_grad[1 + i] and ctx.present[i]:
                g = torch.empty((ch, H, W), dtype=torch.float32, device=dev)
                setattr(a, "dL_d" + k, g.data_ptr())
            grads.append(g)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mrgs_geometry_loss_backward(C.byref(a), C.c_void_p(stream)), "mrgs_geometry_loss_backward")
        return (None, *grads, None, None)


def geometry_losses(rend_normal=None, surf_normal=None, rend_dist=None, surf_depth=None, gt_image=None, image_weight=None,
                    normal: bool = False, dist: bool = False, normal_smooth: bool = False, depth_smooth: bool = False):
    """The four geometric regularisers of calculate_loss as a [4] tensor (terms not selected are 0); one kernel."""
    terms = ((_lib.GEOM_NORMAL if normal else 0) | (_lib.GEOM_DIST if dist else 0)
             | (_lib.GEOM_NORMAL_SMOOTH if normal_smooth else 0) | (_lib.GEOM_DEPTH_SMOOTH if depth_smooth else 0))
    if gt_image is not None:
        gt_image = gt_image.detach()
    if image_weight is not None:
        image_weight = image_weight.detach()
    return _GeometryLoss.apply(terms, rend_normal, surf_normal, rend_dist, surf_depth, gt_image, image_weight)


def first_order_edge_aware_loss(data: torch.Tensor, img: torch.Tensor) -> torch.Tensor:
    """loss_utils.py:121-122 for data of 3 channels (a normal map) or 1 channel (a depth map) and a 3-channel image."""
    if data.dim() == 3 and data.shape[0] == 3:
        return geometry_losses(rend_normal=data, gt_image=img, normal_smooth=True)[2]
    if (data.dim() == 3 and data.shape[0] == 1) or data.dim() == 2:
        return geometry_losses(surf_depth=data, gt_image=img, depth_smooth=True)[3]
    raise NotImplementedError("first_order_edge_aware_loss: data must be [3,H,W] or [1,H,W]")


def get_img_grad_weight(img: torch.Tensor, beta: float = 2.0) -> torch.Tensor:
    """loss_utils.py:127-139 (beta is unused there as well): [C,H,W] -> [H,W], no gradient."""
    lib = _lib.load()
    img = _img(img.detach(), "img")
    Cn, H, W = img.shape
    out = torch.empty((H, W), dtype=torch.float32, device=img.device)
    scratch = torch.empty(2, dtype=torch.int32, device=img.device)
    with torch.cuda.device(img.device):
        stream = torch.cuda.current_stream(img.device).cuda_stream
        _lib.check(lib.mrgs_img_grad_weight(img.data_ptr(), Cn, H, W, out.data_ptr(), scratch.data_ptr(),
                                            C.c_void_p(stream)), "mrgs_img_grad_weight")
    return out


def calculate_loss(viewpoint_camera, pc, render_pkg, opt, iteration, image_weight=None, bg_mask=None):
    """utils/loss_utils.py:142-228: (loss, tb_dict) with the same term selection, lambdas and dictionary keys.

    Two fused kernel pairs instead of ~60 eager launches; every logged scalar reaches the host in ONE copy (the
    reference calls .item() six times). The lpips branch (opt.use_perceptual_loss past perceptual_loss_start_iter)
    needs the external VGG network and is not part of this path."""
    rendered_image = render_pkg["render"]
    gt_image = viewpoint_camera.original_image.cuda()
    if getattr(opt, "use_perceptual_loss", False) and iteration > opt.perceptual_loss_start_iter:
        raise NotImplementedError("the lpips term of calculate_loss stays on the reference (external VGG network)")

    Ll1, ssim_val = l1_ssim(rendered_image, gt_image)
    loss0 = (1.0 - opt.lambda_dssim) * Ll1 + opt.lambda_dssim * (1.0 - ssim_val)
    use_normal = opt.lambda_normal_render_depth > 0 and iteration > opt.normal_loss_start
    use_dist = opt.lambda_dist > 0 and iteration > opt.dist_loss_start
    use_nsmooth = (opt.lambda_normal_smooth > 0 and iteration > opt.normal_smooth_from_iter
                   and iteration < opt.normal_smooth_until_iter)
    use_dsmooth = opt.lambda_depth_smooth > 0 and iteration > 3000
    loss = loss0
    zero = torch.zeros_like(loss0.detach())
    terms = zero.new_zeros(4)
    if use_normal or use_dist or use_nsmooth or use_dsmooth:
        terms = geometry_losses(
            rend_normal=render_pkg["rend_normal"] if (use_normal or use_nsmooth) else None,
            surf_normal=render_pkg["surf_normal"] if use_normal else None,
            rend_dist=render_pkg["rend_dist"] if use_dist else None,
            surf_depth=render_pkg["surf_depth"] if use_dsmooth else None,
            gt_image=gt_image if (use_nsmooth or use_dsmooth) else None,
            image_weight=image_weight if use_normal else None,
            normal=use_normal, dist=use_dist, normal_smooth=use_nsmooth, depth_smooth=use_dsmooth)
        if use_normal:
            loss = loss + opt.lambda_normal_render_depth * terms[0]
        if use_dist:
            loss = loss + opt.lambda_dist * terms[1]
        if use_nsmooth:
            loss = loss + opt.lambda_normal_smooth * terms[2]
        if use_dsmooth:
            loss = loss + opt.lambda_depth_smooth * terms[3]

    with torch.no_grad():
        mse = ((rendered_image - gt_image) ** 2).reshape(rendered_image.shape[0], -1).mean(1)
        psnr = (20 * torch.log10(1.0 / torch.sqrt(mse))).mean()
        host = torch.stack((Ll1, psnr, ssim_val, loss0, terms[2], terms[3], loss)).tolist()
    t = terms.detach()
    tb_dict = {
        "num_points": pc.get_xyz.shape[0],
        "loss_l1": host[0], "psnr": host[1], "ssim": host[2], "loss0": host[3],
        "loss_normal_render_depth": t[0] if use_normal else zero,
        "loss_dist": opt.lambda_dist * t[1] if use_dist else zero,
        "loss_normal_smooth": host[4] if use_nsmooth else zero,
        "loss_depth_smooth": host[5] if use_dsmooth else zero,
        "loss": host[6],
    }
    return loss, tb_dict

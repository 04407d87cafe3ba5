"""Host-side mirror of the reference rasterizer API, backed by libmrgs.so.

Same names, argument meaning, return tuple and error behaviour as
submodules/diff-surfel-rasterization/diff_surfel_rasterization/__init__.py:
  GaussianRasterizationSettings (:167-179), GaussianRasterizer (:181-235),
  _RasterizeGaussians.forward/backward (:47-165), rasterize_gaussians (:22-45).
Differences that a caller cannot observe: kernels run on the CURRENT torch stream (the
reference uses the legacy default stream), scratch buffers have this library's own layout,
and S > 24 raises instead of silently overflowing MAX_FEATURES.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib


def _ptr(t: torch.Tensor | None) -> int | None:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.numel() == 0:
        return t
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    return t.contiguous()


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class ForwardState(NamedTuple):
    """What the forward keeps for the backward (the reference's geomBuffer / binningBuffer /
    imgBuffer triple, rast/rasterize_points.cu:96-101)."""
    num_rendered: int
    geom: torch.Tensor
    binning: torch.Tensor
    image: torch.Tensor


# per-device instance capacity for the optimistic binning buffer: 1.25 x the largest R seen so far
_capacity_hint = {}
_OPTIMISTIC = os.environ.get("MRGS_OPTIMISTIC_BINNING", "1") != "0"


def rasterize_forward_raw(bg, means3D, colors_precomp, features, opacities, scales, rotations,
                          scale_modifier, transMat_precomp, viewmatrix, projmatrix, tanfovx,
                          tanfovy, image_height, image_width, sh, sh_degree, campos, prefiltered,
                          debug):
    """Equivalent of _C.rasterize_gaussians (rast/rasterize_points.cu:41-144): returns
    (num_rendered, contrib, color, feature, others, radii, geomBuffer, binningBuffer, imgBuffer)."""
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    dev = means3D.device
    P = means3D.shape[0]
    if features.dim() != 2:
        raise RuntimeError("features must have dimensions (num_points, S)")
    S = features.shape[1]
    if S > _lib.MAX_FEATURES:
        raise RuntimeError(f"features has {S} channels; at most {_lib.MAX_FEATURES} are supported")
    H, W = int(image_height), int(image_width)

    bg = _f32c(bg, "background"); means3D = _f32c(means3D, "means3D")
    colors_precomp = _f32c(colors_precomp, "colors"); opacities = _f32c(opacities, "opacity")
    scales = _f32c(scales, "scales"); rotations = _f32c(rotations, "rotations")
    transMat_precomp = _f32c(transMat_precomp, "transMat_precomp")
    viewmatrix = _f32c(viewmatrix, "viewmatrix"); projmatrix = _f32c(projmatrix, "projmatrix")
    sh = _f32c(sh, "sh"); campos = _f32c(campos, "campos")
    if features.numel() != 0:
        features = _f32c(features, "features")

    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    contrib = torch.zeros((1, H, W), **i32)  # dead output in the reference as well (always 0)
    if P == 0:
        z = lambda c: torch.zeros((c, H, W), **f32)
        e = torch.empty(0, dtype=torch.uint8, device=dev)
        return 0, contrib, z(3), z(S), z(7), torch.zeros((0,), **i32), e, e.clone(), e.clone()

    color = torch.empty((3, H, W), **f32)
    feature = torch.empty((S, H, W), **f32)
    others = torch.empty((7, H, W), **f32)
    radii = torch.empty((P,), **i32)
    geom = torch.empty(lib.mrgs_geom_bytes(P, S), dtype=torch.uint8, device=dev)
    image = torch.empty(lib.mrgs_image_bytes(W, H), dtype=torch.uint8, device=dev)

    holder = {}

    def _alloc(_ctx, nbytes):
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        holder["t"] = t
        return t.data_ptr()

    M = sh.shape[1] if sh.numel() != 0 else 0
    a = _lib.ForwardArgs()
    a.P, a.S, a.sh_degree, a.sh_coeffs, a.width, a.height = P, S, int(sh_degree), int(M), W, H
    a.tan_fovx, a.tan_fovy, a.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
    a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
    a.background, a.means3D, a.shs = _ptr(bg), _ptr(means3D), _ptr(sh)
    a.colors_precomp, a.features, a.opacities = _ptr(colors_precomp), _ptr(features), _ptr(opacities)
    a.scales, a.rotations, a.transMat_precomp = _ptr(scales), _ptr(rotations), _ptr(transMat_precomp)
    a.viewmatrix, a.projmatrix, a.campos = _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos)
    a.out_color, a.out_feature, a.out_others, a.radii = _ptr(color), _ptr(feature), _ptr(others), _ptr(radii)
    a.geom_buffer, a.geom_bytes = geom.data_ptr(), geom.numel()
    a.image_buffer, a.image_bytes = image.data_ptr(), image.numel()
    cb = _lib.alloc_fn(_alloc)
    a.binning_alloc, a.binning_ctx = cb, None
    # optimistic binning: lend a buffer sized from the instance counts seen so far on this device, so the
    # library can enqueue the whole forward before it waits for R (include/mrgs.h, MrgsForwardArgs)
    scratch = None
    cap = _capacity_hint.get(dev.index, 0) if _OPTIMISTIC else 0
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing:
        # CUDA-graph capture of a view (graphs.py): no host wait is possible, so the call runs in the library's no_wait
        # mode against the capacity learnt from earlier frames; the real instance count lands in a pinned int the
        # graph's owner checks after a replay (captured_counts_ok)
        if cap <= 0:
            raise RuntimeError("rasterizer: render at least one frame on this device before capturing a CUDA graph "
                               "(the capture needs an instance-capacity estimate)")
        count = torch.zeros(1, dtype=torch.int32).pin_memory()
        _captured_counts.append((count, cap))
        a.no_wait, a.count_out = 1, count.data_ptr()
    if cap > 0:
        scratch = torch.empty(lib.mrgs_binning_bytes(cap), dtype=torch.uint8, device=dev)
        a.binning_scratch, a.binning_scratch_bytes, a.binning_capacity = scratch.data_ptr(), scratch.numel(), cap

    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mrgs_forward(C.byref(a), C.c_void_p(stream)), "mrgs_forward")
    R = int(a.num_rendered)
    _last_num_rendered[dev.index] = R
    binning = holder.get("t")
    if binning is None:
        binning = scratch if (scratch is not None and a.binning_buffer == scratch.data_ptr()) else \
            torch.empty(0, dtype=torch.uint8, device=dev)
    binning.mrgs_capacity = int(a.binning_capacity_used)   # layout key for the debug decoders
    if _OPTIMISTIC and not capturing:
        want = (R + R // 4 + 65535) & ~65535
        if want > cap:
            _capacity_hint[dev.index] = want
    return R, contrib, color, feature, others, radii, geom, binning, image


_last_num_rendered: dict = {}
_captured_counts: list = []      # (pinned int32 [1] written by every replay, capacity) per captured forward


def captured_counts_ok() -> bool:
    """After a synchronisation: did every captured forward's last replay fit its binning capacity? If not, that
    frame's outputs are invalid: raise the capacity (render the view eagerly once) and capture again."""
    return all(int(c[0]) <= cap for c, cap in _captured_counts)


def reserve_capacity(device, instances: int) -> None:
    """Make the optimistic / captured binning buffers of `device` hold at least `instances` (tile, surfel) pairs."""
    dev = torch.device(device)
    want = (int(instances) + 65535) & ~65535
    if want > _capacity_hint.get(dev.index, 0):
        _capacity_hint[dev.index] = want


def rasterize_backward_raw(bg, means3D, radii, colors_precomp, features, scales, rotations,
                           scale_modifier, transMat_precomp, viewmatrix, projmatrix, tanfovx,
                           tanfovy, dL_dout_color, dL_dout_feature, dL_dout_others, sh, sh_degree,
                           campos, geom, num_rendered, binning, image, contrib, debug, need_colors=True,
                           need_transmat=True, accumulate_into=None):
    """Equivalent of _C.rasterize_gaussians_backward (rast/rasterize_points.cu:146-252): returns
    (dL_dmeans2D, dL_dcolors, dL_dfeatures, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh,
    dL_dscales, dL_drotations).

    accumulate_into: optional {"means3D","shs","features","opacities","scales","rotations"} -> fp32 CUDA
    buffers; the parameter gradients are then ADDED to those buffers by the kernel itself (and the same
    tensors are returned) instead of being written to fresh tensors (MrgsBackwardArgs.accumulate)."""
    lib = _lib.load()
    dev = means3D.device
    P = means3D.shape[0]
    S = features.shape[1]
    H, W = dL_dout_color.shape[1], dL_dout_color.shape[2]
    M = sh.shape[1] if sh.numel() != 0 else 0
    f32 = dict(dtype=torch.float32, device=dev)

    out_shapes = [(P, 3), (P, 3), (P, S), (P, 1), (P, 3), (P, 9), (P, M, 3), (P, 2), (P, 4)]
    if P == 0:
        return tuple(torch.zeros(s, **f32) for s in out_shapes)
    (dL_dmeans2D, dL_dcolors, dL_dfeatures, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh,
     dL_dscales, dL_drotations) = (torch.empty(s, **f32) for s in out_shapes)
    has_sr = scales.numel() != 0
    if accumulate_into is not None:
        if not has_sr or M == 0:
            raise RuntimeError("accumulate_into needs the scales/rotations + SH input combination")

        def _sink(name, shape):
            t = accumulate_into[name]
            n = 1
            for d in shape:
                n *= d
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
                raise RuntimeError(f"accumulate_into[{name!r}] must be a contiguous float32 CUDA tensor of {n} elements")
            return t
        dL_dmeans3D, dL_dsh = _sink("means3D", (P, 3)), _sink("shs", (P, M, 3))
        dL_dopacity, dL_dscales, dL_drotations = _sink("opacities", (P, 1)), _sink("scales", (P, 2)), _sink("rotations", (P, 4))
        if S:
            dL_dfeatures = _sink("features", (P, S))
        need_colors = need_transmat = False
    if not has_sr:  # never written for precomputed transforms; the reference returns zeros
        dL_dscales.zero_(); dL_drotations.zero_()
    if M == 0:
        pass  # dL_dsh is (P,0,3)

    arena = torch.empty(lib.mrgs_grad_arena_bytes(P, S), dtype=torch.uint8, device=dev)

    a = _lib.BackwardArgs()
    a.P, a.S, a.sh_degree, a.sh_coeffs, a.width, a.height = P, S, int(sh_degree), int(M), W, H
    a.tan_fovx, a.tan_fovy, a.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
    a.debug, a.num_rendered = int(bool(debug)), int(num_rendered)
    keep = [_f32c(t, n) for t, n in (
        (bg, "background"), (means3D, "means3D"), (sh, "sh"), (colors_precomp, "colors"),
        (features, "features") if features.numel() else (features, "features"),
        (scales, "scales"), (rotations, "rotations"), (transMat_precomp, "transMat_precomp"),
        (viewmatrix, "viewmatrix"), (projmatrix, "projmatrix"), (campos, "campos"),
        (dL_dout_color, "dL_dout_color"), (dL_dout_feature, "dL_dout_feature"),
        (dL_dout_others, "dL_dout_others"))]
    (bg_, means_, sh_, col_, feat_, sc_, rot_, tm_, vm_, pm_, cam_, gc_, gf_, go_) = keep
    a.background, a.means3D, a.shs, a.colors_precomp = _ptr(bg_), _ptr(means_), _ptr(sh_), _ptr(col_)
    a.features, a.scales, a.rotations, a.transMat_precomp = _ptr(feat_), _ptr(sc_), _ptr(rot_), _ptr(tm_)
    a.viewmatrix, a.projmatrix, a.campos = _ptr(vm_), _ptr(pm_), _ptr(cam_)
    a.radii = _ptr(radii.contiguous())
    a.geom_buffer, a.binning_buffer, a.image_buffer = _ptr(geom), _ptr(binning), _ptr(image)
    a.dL_dout_color, a.dL_dout_feature, a.dL_dout_others = _ptr(gc_), _ptr(gf_), _ptr(go_)
    a.dL_dmeans2D, a.dL_dfeatures = _ptr(dL_dmeans2D), _ptr(dL_dfeatures)
    a.dL_dcolors = _ptr(dL_dcolors) if need_colors else None      # NULL = not wanted, not written
    a.dL_dtransMat = _ptr(dL_dtransMat) if need_transmat else None
    a.dL_dopacity, a.dL_dmeans3D = _ptr(dL_dopacity), _ptr(dL_dmeans3D)
    a.dL_dsh = _ptr(dL_dsh)
    a.dL_dscales = _ptr(dL_dscales) if has_sr else None
    a.dL_drotations = _ptr(dL_drotations) if has_sr else None
    a.grad_arena, a.grad_arena_bytes = arena.data_ptr(), arena.numel()
    a.accumulate = int(accumulate_into is not None)

    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mrgs_backward(C.byref(a), C.c_void_p(stream)), "mrgs_backward")
    return (dL_dmeans2D, dL_dcolors, dL_dfeatures, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh,
            dL_dscales, dL_drotations)


def mark_visible_raw(positions, viewmatrix, projmatrix):
    lib = _lib.load()
    P = positions.shape[0]
    present = torch.zeros((P,), dtype=torch.bool, device=positions.device)
    if P != 0:
        pos = _f32c(positions, "means3D"); vm = _f32c(viewmatrix, "viewmatrix"); pm = _f32c(projmatrix, "projmatrix")
        with torch.cuda.device(positions.device):
            stream = torch.cuda.current_stream(positions.device).cuda_stream
            _lib.check(lib.mrgs_mark_visible(P, pos.data_ptr(), vm.data_ptr(), pm.data_ptr(),
                                             present.data_ptr(), C.c_void_p(stream)), "mrgs_mark_visible")
    return present


def _cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, features, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings, grad_sink=None):
    if grad_sink is not None:
        for name, t in (("means3D", means3D), ("shs", sh), ("features", features), ("opacities", opacities),
                        ("scales", scales), ("rotations", rotations)):
            if t.requires_grad and not t.is_leaf:
                raise RuntimeError(f"grad_sink: {name} is not a leaf tensor; fused gradient accumulation bypasses "
                                   "autograd and is only valid for parameters")
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, features, opacities,
                                     scales, rotations, cov3Ds_precomp, raster_settings, grad_sink)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, features, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings, grad_sink=None):
        rs = raster_settings
        ctx.grad_sink = grad_sink
        args = (rs.bg, means3D, colors_precomp, features, opacities, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        if rs.debug:  # same failure artefact as the reference (rast/diff_surfel_rasterization/__init__.py:87-94)
            cpu_args = _cpu_deep_copy_tuple(args)
            try:
                outs = rasterize_forward_raw(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            outs = rasterize_forward_raw(*args)
        (num_rendered, contrib, color, feature, depth, radii, geom, binning, image) = outs
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, features, means3D, scales, rotations, cov3Ds_precomp,
                              radii, sh, geom, binning, image, contrib)
        ctx.mark_non_differentiable(contrib, radii)
        return contrib, color, feature, radii, depth

    @staticmethod
    def backward(ctx, grad_out_contrib, grad_out_color, grad_out_feature, grad_radii, grad_depth):
        rs = ctx.raster_settings
        (colors_precomp, features, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom,
         binning, image, contrib) = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, features, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color,
                grad_out_feature, grad_depth, sh, rs.sh_degree, rs.campos, geom, ctx.num_rendered,
                binning, image, contrib, rs.debug)
        want = dict(need_colors=ctx.needs_input_grad[3], need_transmat=ctx.needs_input_grad[8],
                    accumulate_into=ctx.grad_sink)
        if rs.debug:  # same failure artefact as the reference (rast/diff_surfel_rasterization/__init__.py:141-148)
            cpu_args = _cpu_deep_copy_tuple(args)
            try:
                grads = rasterize_backward_raw(*args, **want)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            grads = rasterize_backward_raw(*args, **want)
        (grad_means2D, grad_colors_precomp, grad_features, grad_opacities, grad_means3D,
         grad_cov3Ds_precomp, grad_sh, grad_scales, grad_rotations) = grads
        if not ctx.needs_input_grad[3]:
            grad_colors_precomp = None
        if not ctx.needs_input_grad[8]:
            grad_cov3Ds_precomp = None
        if ctx.grad_sink is not None:
            # the kernel has already added these into the sink: nothing for autograd to accumulate
            return (None, grad_means2D, None, None, None, None, None, None, None, None, None)
        # one gradient per forward input (the reference returns a surplus trailing None)
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_features,
                grad_opacities, grad_scales, grad_rotations, grad_cov3Ds_precomp, None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings, grad_sink=None):
        """grad_sink (extension, optional): {"means3D","shs","features","opacities","scales","rotations"} ->
        buffers the backward ADDS the parameter gradients into (e.g. parallel.GradArena.views) instead of
        returning them to autograd; only valid when those inputs are leaf parameters."""
        super().__init__()
        self.raster_settings = raster_settings
        self.grad_sink = grad_sink

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return mark_visible_raw(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, features=None,
                scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = lambda: torch.empty(0, dtype=torch.float32, device=means3D.device)
        if shs is None:
            shs = empty()
        if colors_precomp is None:
            colors_precomp = empty()
        if features is None:
            features = torch.empty_like(means3D[..., :0])
        if scales is None:
            scales = empty()
        if rotations is None:
            rotations = empty()
        if cov3D_precomp is None:
            cov3D_precomp = empty()
        out = rasterize_gaussians(means3D, means2D, shs, colors_precomp, features, opacities, scales,
                                  rotations, cov3D_precomp, rs, self.grad_sink)
        # instances (tile, surfel) of this view, already on the host (the reference's `rendered`): a free cost estimate
        self.num_rendered = _last_num_rendered.get(means3D.device.index)
        return out

"""Host-side mirror of the cubemap ops EnvLight.build_mips needs, backed by libmrgs.so.

specular_cubemap / diffuse_cubemap apply a cached prefilter plan (prefilter.py: the op as a precomputed sparse
operator, gather forward and backward); only plans beyond the HBM budget fall back to the direct per-texel kernels.

Same names / argument meaning as the reference:
  cubemap_mip        scene/light_utils.py:66-80 (autograd.Function with the non-adjoint backward)
  diffuse_cubemap    scene/renderutils/ops.py:391-411
  specular_cubemap   scene/renderutils/ops.py:413-458 (+ __ndfBounds :428-443, cached per key)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import prefilter as pf


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check_cube(cubemap, channels=None):
    if cubemap.dim() != 4 or cubemap.shape[0] != 6 or cubemap.shape[1] != cubemap.shape[2]:
        raise AssertionError("Bad shape for cubemap tensor: %s" % str(tuple(cubemap.shape)))
    if not cubemap.is_cuda or cubemap.dtype != torch.float32:
        raise RuntimeError("cubemap must be a float32 CUDA tensor")
    if channels is not None and cubemap.shape[3] != channels:
        raise RuntimeError(f"cubemap must have {channels} channels")


class _cubemap_mip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cubemap):
        _check_cube(cubemap)
        lib = _lib.load()
        x = cubemap.contiguous()
        res, ch = x.shape[1], x.shape[3]
        out = torch.empty((6, res // 2, res // 2, ch), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.mrgs_cubemap_mip_forward(x.data_ptr(), out.data_ptr(), res, ch, _stream(x.device)),
                       "mrgs_cubemap_mip_forward")
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        d = dout.contiguous()
        _check_cube(d, 3)
        res = d.shape[1]
        out = torch.empty((6, res * 2, res * 2, 3), dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            _lib.check(lib.mrgs_cubemap_mip_backward(d.data_ptr(), out.data_ptr(), res, _stream(d.device)),
                       "mrgs_cubemap_mip_backward")
        return out


def cubemap_mip(cubemap):
    return _cubemap_mip.apply(cubemap)


class _diffuse_cubemap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cubemap):
        _check_cube(cubemap, 3)
        x = cubemap.contiguous()
        fwd, bwd = pf.plan_pair("diffuse", x.shape[1], 0.0, None, x.device)
        out = torch.empty_like(x)
        pf.apply_jobs([(fwd, x, 3, out, 3, None)], False, x.device)
        ctx.plan = bwd
        return out

    @staticmethod
    def backward(ctx, dout):
        d = dout.contiguous()
        g = torch.empty_like(d)
        pf.apply_jobs([(ctx.plan, d, 3, g, 3, None)], True, d.device)
        return g


class _diffuse_cubemap_direct(torch.autograd.Function):
    """The per-texel loop kernels (cold path: maps too large for a plan)."""

    @staticmethod
    def forward(ctx, cubemap):
        _check_cube(cubemap, 3)
        lib = _lib.load()
        x = cubemap.contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.mrgs_diffuse_cubemap_forward(x.data_ptr(), x.shape[1], out.data_ptr(), _stream(x.device)),
                       "mrgs_diffuse_cubemap_forward")
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        lib = _lib.load()
        d = dout.contiguous()
        g = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.mrgs_diffuse_cubemap_backward(x.data_ptr(), x.shape[1], d.data_ptr(), g.data_ptr(),
                                                         _stream(x.device)), "mrgs_diffuse_cubemap_backward")
        return g


def diffuse_cubemap(cubemap, use_python=False):
    assert not use_python
    _check_cube(cubemap, 3)
    if pf.estimate_bytes(cubemap.shape[1], None) > pf.budget_bytes():
        return _diffuse_cubemap_direct.apply(cubemap)
    return _diffuse_cubemap.apply(cubemap)


ndf_cutoff_costheta = pf.ndf_cutoff_costheta
specular_bounds = pf.specular_bounds

_bounds_cache: dict = {}


def _ndf_bounds(res, roughness, cutoff, device):
    key = (res, roughness, cutoff, str(device))
    if key not in _bounds_cache:
        ct = pf.cutoff_costheta(roughness, cutoff)
        _bounds_cache[key] = (ct, specular_bounds(res, ct, device))
    return _bounds_cache[key]


class _specular_cubemap(torch.autograd.Function):
    """The plugin-level op pair specular_cubemap_fwd / _bwd (scene/renderutils/ops.py:413-426): out4 = (sum w rgb,
    sum w). `bounds` is accepted for signature parity; the plan derives the same bounds itself."""

    @staticmethod
    def forward(ctx, cubemap, roughness, costheta_cutoff, bounds):
        x = cubemap.contiguous()
        res = x.shape[1]
        fwd, bwd = pf.plan_pair("specular", res, float(roughness), float(costheta_cutoff), x.device)
        rgb = torch.empty_like(x)
        pf.apply_jobs([(fwd, x, 3, rgb, 3, None)], False, x.device)
        wsum = fwd.wsum.view(6, res, res, 1)
        ctx.plan, ctx.wsum = bwd, wsum
        return torch.cat([rgb * wsum, wsum], dim=-1)

    @staticmethod
    def backward(ctx, dout):
        d = (dout[..., :3] * ctx.wsum).contiguous()     # the plan's weights carry 1 / wsum
        g = torch.empty_like(d)
        pf.apply_jobs([(ctx.plan, d, 3, g, 3, None)], True, d.device)
        return g, None, None, None


class _specular_cubemap_normalised(torch.autograd.Function):
    """specular_cubemap incl. the division by the weight sum (ops.py:445-458) as one gather each way."""

    @staticmethod
    def forward(ctx, cubemap, roughness, costheta_cutoff):
        x = cubemap.contiguous()
        fwd, bwd = pf.plan_pair("specular", x.shape[1], float(roughness), float(costheta_cutoff), x.device)
        rgb = torch.empty_like(x)
        pf.apply_jobs([(fwd, x, 3, rgb, 3, fwd.wsum)], False, x.device)
        ctx.plan = bwd
        return rgb

    @staticmethod
    def backward(ctx, dout):
        d = dout.contiguous()
        g = torch.empty_like(d)
        pf.apply_jobs([(ctx.plan, d, 3, g, 3, None)], True, d.device)
        return g, None, None


class _specular_cubemap_direct(torch.autograd.Function):
    """The per-texel cone-loop kernels (cold path: cones too large for a plan)."""

    @staticmethod
    def forward(ctx, cubemap, roughness, costheta_cutoff, bounds):
        lib = _lib.load()
        x = cubemap.contiguous()
        res = x.shape[1]
        out = torch.empty((6, res, res, 4), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.mrgs_specular_cubemap_forward(x.data_ptr(), bounds.data_ptr(), res, float(roughness),
                                                         float(costheta_cutoff), out.data_ptr(), _stream(x.device)),
                       "mrgs_specular_cubemap_forward")
        ctx.save_for_backward(x, bounds)
        ctx.roughness, ctx.cutoff = float(roughness), float(costheta_cutoff)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, bounds = ctx.saved_tensors
        lib = _lib.load()
        d = dout.contiguous()
        g = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.mrgs_specular_cubemap_backward(x.data_ptr(), bounds.data_ptr(), x.shape[1], ctx.roughness,
                                                          ctx.cutoff, d.data_ptr(), g.data_ptr(), _stream(x.device)),
                       "mrgs_specular_cubemap_backward")
        return g, None, None, None


def specular_cubemap(cubemap, roughness, cutoff=0.99, use_python=False):
    assert not use_python
    _check_cube(cubemap, 3)
    ct = pf.cutoff_costheta(roughness, cutoff)
    try:
        return _specular_cubemap_normalised.apply(cubemap, roughness, ct)
    except pf.PrefilterTooLarge:
        ct, bounds = _ndf_bounds(cubemap.shape[1], roughness, cutoff, cubemap.device)
        out = _specular_cubemap_direct.apply(cubemap, roughness, ct, bounds)
        return out[..., 0:3] / out[..., 3:]

"""Host-side mirror of the cubemap ops EnvLight.build_mips needs, backed by libmrgs.so.

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
    return _diffuse_cubemap.apply(cubemap)


def ndf_cutoff_costheta(roughness: float, cutoff: float) -> float:
    """cos of the cone angle that keeps `cutoff` of the GGX lobe's energy: the host-side search of
    __ndfBounds (scene/renderutils/ops.py:428-441), same 1e6-sample cumsum in float64."""
    def ndfGGX(alphaSqr, costheta):
        costheta = np.clip(costheta, 0.0, 1.0)
        d = (costheta * alphaSqr - costheta) * costheta + 1.0
        return alphaSqr / (d * d * np.pi)
    nSamples = 1000000
    costheta = np.cos(np.linspace(0, np.pi / 2.0, nSamples))
    D = np.cumsum(ndfGGX(roughness ** 4, costheta))
    idx = np.argmax(D >= D[..., -1] * cutoff)
    return float(costheta[idx])


_bounds_cache: dict = {}


def specular_bounds(res: int, costheta_cutoff: float, device) -> torch.Tensor:
    lib = _lib.load()
    bounds = torch.empty((6, res, res, 6, 4), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.mrgs_specular_bounds(res, float(costheta_cutoff), bounds.data_ptr(), _stream(device)),
                   "mrgs_specular_bounds")
    return bounds


def _ndf_bounds(res, roughness, cutoff, device):
    key = (res, roughness, cutoff, str(device))
    if key not in _bounds_cache:
        ct = ndf_cutoff_costheta(roughness, cutoff)
        _bounds_cache[key] = (ct, specular_bounds(res, ct, device))
    return _bounds_cache[key]


class _specular_cubemap(torch.autograd.Function):
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
    ct, bounds = _ndf_bounds(cubemap.shape[1], roughness, cutoff, cubemap.device)
    out = _specular_cubemap.apply(cubemap, roughness, ct, bounds)
    return out[..., 0:3] / out[..., 3:]

"""Per-surfel feature preparation in front of the rasterizer as ONE kernel pair of
libmrgs.so: activations, surfel normal, reflection direction, degree-3 indirect SH, concatenation.

Replaces, for the default pipeline (FLAG "2dgs", no ASG), gaussian_renderer/__init__.py:259-266 and :334-353
together with the GaussianModel getters they call (scene/gaussian_model.py:236-303): get_scaling,
get_rotation, get_opacity, get_refl, get_rough, get_ori_color, get_indirect, get_normal.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

RAW_NAMES = ("xyz", "scaling", "rotation", "opacity", "refl_strength", "roughness", "ori_color", "indirect_dc",
             "indirect_rest")
_WIDTH = dict(xyz=3, scaling=2, rotation=4, opacity=1, refl_strength=1, roughness=1, ori_color=3, indirect_dc=3,
              indirect_rest=45)


def _check(t: torch.Tensor, name: str, P: int) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    if t.shape[0] != P or t.numel() != P * _WIDTH[name]:
        raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {P} x {_WIDTH[name]} values")
    return t.contiguous()


class _SurfelFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, campos, *raw):
        lib = _lib.load()
        P = raw[0].shape[0]
        raw = tuple(_check(t, n, P) for t, n in zip(raw, RAW_NAMES))
        campos = campos.detach().to(device=raw[0].device, dtype=torch.float32).contiguous()
        dev = raw[0].device
        f32 = dict(dtype=torch.float32, device=dev)
        scales, rotations = torch.empty((P, 2), **f32), torch.empty((P, 4), **f32)
        opacities, features = torch.empty((P, 1), **f32), torch.empty((P, 8), **f32)
        a = _lib.SurfelFeatureArgs()
        a.P, a.campos = P, campos.data_ptr()
        for n, t in zip(RAW_NAMES, raw):
            setattr(a, n, t.data_ptr())
        a.scales, a.rotations, a.opacities, a.features = (scales.data_ptr(), rotations.data_ptr(),
                                                          opacities.data_ptr(), features.data_ptr())
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mrgs_surfel_features_forward(C.byref(a), C.c_void_p(stream)), "mrgs_surfel_features_forward")
        ctx.save_for_backward(campos, *raw)
        return scales, rotations, opacities, features

    @staticmethod
    def backward(ctx, g_scales, g_rotations, g_opacities, g_features):
        lib = _lib.load()
        campos, *raw = ctx.saved_tensors
        P = raw[0].shape[0]
        dev = raw[0].device
        a = _lib.SurfelFeatureArgs()
        a.P, a.campos = P, campos.data_ptr()
        for n, t in zip(RAW_NAMES, raw):
            setattr(a, n, t.data_ptr())
        ups = [g.to(torch.float32).contiguous() for g in (g_scales, g_rotations, g_opacities, g_features)]
        a.dL_dscales, a.dL_drotations, a.dL_dopacities, a.dL_dfeatures = (u.data_ptr() for u in ups)
        grads = [torch.empty_like(t) for t in raw]
        for n, gt in zip(("dL_dxyz", "dL_dscaling", "dL_drotation", "dL_dopacity", "dL_drefl_strength",
                          "dL_droughness", "dL_dori_color", "dL_dindirect_dc", "dL_dindirect_rest"), grads):
            setattr(a, n, gt.data_ptr())
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.mrgs_surfel_features_backward(C.byref(a), C.c_void_p(stream)), "mrgs_surfel_features_backward")
        return (None, *grads)


def surfel_features(campos, xyz, scaling, rotation, opacity, refl_strength, roughness, ori_color, indirect_dc,
                    indirect_rest):
    """RAW (pre-activation) parameters -> (scales [P,2], rotations [P,4], opacities [P,1], features [P,8]),
    the tensors render_surfel passes to GaussianRasterizer. Gradients flow to all nine parameters
    (xyz through the view-dependent reflection direction only; the rasterizer adds its own dL/dmeans3D)."""
    return _SurfelFeatures.apply(campos, xyz, scaling, rotation, opacity, refl_strength, roughness, ori_color,
                                 indirect_dc, indirect_rest)


def surfel_features_from_model(pc, campos):
    """The same for a scene/gaussian_model.py GaussianModel-like object exposing the raw parameters."""
    return surfel_features(campos, pc._xyz, pc._scaling, pc._rotation, pc._opacity, pc._refl_strength, pc._roughness,
                           pc._ori_color, pc._indirect_dc, pc._indirect_rest)


MODEL_RAW_ATTRS = ("_xyz", "_scaling", "_rotation", "_opacity", "_refl_strength", "_roughness", "_ori_color",
                   "_indirect_dc", "_indirect_rest")

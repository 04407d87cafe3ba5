"""ctypes binding of libmrgs.so — the only place Python touches the C ABI (include/mrgs.h).

There is deliberately no CPU or torch fallback: if the CUDA library is missing the import of
the product path fails loudly (set MRGS_AUTOBUILD=0 to forbid the in-tree nvcc build).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmrgs.so"

MRGS_ABI_VERSION = 10
MAX_FEATURES = 24
TILE = 16

alloc_fn = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class GeomLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "rec", "cf", "clamped", "tiles_touched", "point_offsets", "rect", "depth", "bbox", "sort_keys", "sort_vals",
        "scan_temp",
        "scan_temp_bytes", "total")] + [("cf_stride", C.c_int32)]


class ImageLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("state", "ranges", "total")]


class BinningLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "point_list", "point_list_unsorted", "keys", "keys_unsorted", "sort_temp",
        "sort_temp_bytes", "total")]


_fp = C.c_void_p  # device pointers travel as integers


class ForwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("S", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int32), ("debug", C.c_int32),
        ("background", _fp), ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp),
        ("features", _fp), ("opacities", _fp), ("scales", _fp), ("rotations", _fp),
        ("transMat_precomp", _fp), ("viewmatrix", _fp), ("projmatrix", _fp), ("campos", _fp),
        ("out_color", _fp), ("out_feature", _fp), ("out_others", _fp), ("radii", _fp),
        ("geom_buffer", _fp), ("geom_bytes", C.c_size_t),
        ("image_buffer", _fp), ("image_bytes", C.c_size_t),
        ("binning_alloc", alloc_fn), ("binning_ctx", C.c_void_p),
        ("binning_scratch", _fp), ("binning_scratch_bytes", C.c_size_t), ("binning_capacity", C.c_int64),
        ("num_rendered", C.c_int32), ("binning_buffer", _fp), ("binning_capacity_used", C.c_int64),
        ("no_wait", C.c_int32), ("count_out", _fp),
    ]


class BackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("S", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("debug", C.c_int32), ("num_rendered", C.c_int32),
        ("background", _fp), ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp),
        ("features", _fp), ("scales", _fp), ("rotations", _fp), ("transMat_precomp", _fp),
        ("viewmatrix", _fp), ("projmatrix", _fp), ("campos", _fp), ("radii", _fp),
        ("geom_buffer", _fp), ("binning_buffer", _fp), ("image_buffer", _fp),
        ("dL_dout_color", _fp), ("dL_dout_feature", _fp), ("dL_dout_others", _fp),
        ("dL_dmeans2D", _fp), ("dL_dcolors", _fp), ("dL_dfeatures", _fp), ("dL_dopacity", _fp),
        ("dL_dmeans3D", _fp), ("dL_dtransMat", _fp), ("dL_dsh", _fp), ("dL_dscales", _fp),
        ("dL_drotations", _fp),
        ("grad_arena", _fp), ("grad_arena_bytes", C.c_size_t), ("accumulate", C.c_int32),
    ]


MAX_MIP_LEVELS = 12


class ShadeArgs(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("num_levels", C.c_int32), ("base_res", C.c_int32),
        ("srgb", C.c_int32), ("min_roughness", C.c_float), ("max_roughness", C.c_float),
        ("ray_matrix", C.c_float * 9), ("normal_matrix", C.c_float * 9),
        ("background", _fp), ("base_color", _fp), ("features", _fp), ("allmap", _fp), ("lut", _fp),
        ("levels", _fp * MAX_MIP_LEVELS),
        ("out_final", _fp), ("out_specular", _fp), ("out_direct", _fp), ("out_normal", _fp),
        ("out_diffuse", _fp),
        ("dL_dfinal", _fp), ("dL_dspecular", _fp), ("dL_ddiffuse", _fp), ("dL_dnormal", _fp),
        ("dL_dbase_color", _fp), ("dL_dfeatures", _fp), ("dL_dallmap", _fp),
        ("dL_dlevels", _fp * MAX_MIP_LEVELS),
    ]


class SurfelShadeArgs(C.Structure):
    """MrgsSurfelShadeArgs (include/mrgs.h)."""
    _fields_ = [("P", C.c_int32), ("diffuse_res", C.c_int32), ("chain", ShadeArgs)] + [(n, _fp) for n in (
        "diffuse_map", "campos", "fg", "xyz", "normals", "albedo", "refl_strength", "roughness", "diffuse", "specular",
        "direct_light", "dL_ddiffuse", "dL_dspecular", "dL_ddirect", "dL_dxyz", "dL_dnormals", "dL_dalbedo",
        "dL_drefl_strength", "dL_droughness", "dL_ddiffuse_map", "dL_dfg")]


# every symbol include/mrgs.h declares: name -> (restype, argtypes)
class SurfelFeatureArgs(C.Structure):
    _fields_ = [("P", C.c_int32), ("campos", _fp)] + [(n, _fp) for n in (
        "xyz", "scaling", "rotation", "opacity", "refl_strength", "roughness", "ori_color", "indirect_dc",
        "indirect_rest", "scales", "rotations", "opacities", "features", "dL_dscales", "dL_drotations",
        "dL_dopacities", "dL_dfeatures", "dL_dxyz", "dL_dscaling", "dL_drotation", "dL_dopacity",
        "dL_drefl_strength", "dL_droughness", "dL_dori_color", "dL_dindirect_dc", "dL_dindirect_rest")]



GEOM_NORMAL, GEOM_DIST, GEOM_NORMAL_SMOOTH, GEOM_DEPTH_SMOOTH = 1, 2, 4, 8


class GeometryLossArgs(C.Structure):
    """MrgsGeometryLossArgs (include/mrgs.h)."""
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("terms", C.c_uint32)] + [(n, _fp) for n in (
        "rend_normal", "surf_normal", "rend_dist", "surf_depth", "gt_image", "image_weight", "coef", "partials", "out4",
        "upstream", "dL_drend_normal", "dL_dsurf_normal", "dL_drend_dist", "dL_dsurf_depth")]


PREFILTER_SPECULAR, PREFILTER_SPECULAR_T, PREFILTER_DIFFUSE, PREFILTER_DIFFUSE_T = 0, 1, 2, 3
PREFILTER_MAX_JOBS = 8


class PrefilterPlan(C.Structure):
    """MrgsPrefilterPlan (include/mrgs.h)."""
    _fields_ = [("res", C.c_int32), ("rows_per_lane", C.c_int32), ("patch_width", C.c_int32)] + [(n, _fp) for n in (
        "patch_seg_begin", "patch_slot_begin", "seg_desc", "spans", "weights")]


class PrefilterBuildArgs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("res", C.c_int32), ("rows_per_lane", C.c_int32), ("patch_width", C.c_int32),
                ("full_search", C.c_int32),
                ("roughness", C.c_float), ("costheta_cutoff", C.c_float)] + [(n, _fp) for n in (
        "texel_table", "bounds", "wsum", "seg_count", "slot_count", "tap_count")] + [("plan", PrefilterPlan)]


class PrefilterJob(C.Structure):
    _fields_ = [("plan", PrefilterPlan), ("src", _fp), ("dst", _fp), ("nan_where_zero", _fp),
                ("src_stride", C.c_int32), ("dst_stride", C.c_int32), ("patch_begin", C.c_int32), ("patch_end", C.c_int32)]


SYMBOLS = {
    "mrgs_abi_version": (C.c_int, []),
    "mrgs_last_error": (C.c_char_p, []),
    "mrgs_geom_layout": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(GeomLayout)]),
    "mrgs_image_layout": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(ImageLayout)]),
    "mrgs_binning_layout": (C.c_int, [C.c_int64, C.POINTER(BinningLayout)]),
    "mrgs_geom_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "mrgs_image_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "mrgs_binning_bytes": (C.c_size_t, [C.c_int64]),
    "mrgs_tile_slot": (C.c_int, [C.c_int32, C.c_int32]),
    "mrgs_grad_arena_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "mrgs_grad_arena_stride": (C.c_int32, [C.c_int32]),
    "mrgs_profile_enable": (None, [C.c_int32]),
    "mrgs_profile_reset": (None, []),
    "mrgs_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "mrgs_launch_count": (C.c_int64, []),
    "mrgs_profile_collect_captured": (None, []),
    "mrgs_shade_forward": (C.c_int, [C.POINTER(ShadeArgs), C.c_void_p]),
    "mrgs_shade_backward": (C.c_int, [C.POINTER(ShadeArgs), C.c_void_p]),
    "mrgs_envlight_query": (C.c_int, [C.POINTER(ShadeArgs), C.c_int64, _fp, _fp, _fp, C.c_void_p]),
    "mrgs_surfel_shade_forward": (C.c_int, [C.POINTER(SurfelShadeArgs), C.c_void_p]),
    "mrgs_surfel_shade_backward": (C.c_int, [C.POINTER(SurfelShadeArgs), C.c_void_p]),
    "mrgs_envlight_query_backward": (C.c_int, [C.POINTER(ShadeArgs), C.c_int64, _fp, _fp, _fp, _fp, _fp, C.c_void_p]),
    "mrgs_depth_normal_forward": (C.c_int, [C.c_int32, C.c_int32, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                            _fp, _fp, _fp, C.c_void_p]),
    "mrgs_depth_normal_backward": (C.c_int, [C.c_int32, C.c_int32, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                             _fp, _fp, _fp, _fp, C.c_void_p]),
    "mrgs_cubemap_mip_forward": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, C.c_void_p]),
    "mrgs_cubemap_mip_backward": (C.c_int, [_fp, _fp, C.c_int32, C.c_void_p]),
    "mrgs_specular_bounds": (C.c_int, [C.c_int32, C.c_float, _fp, C.c_void_p]),
    "mrgs_specular_cubemap_forward": (C.c_int, [_fp, _fp, C.c_int32, C.c_float, C.c_float, _fp, C.c_void_p]),
    "mrgs_specular_cubemap_backward": (C.c_int, [_fp, _fp, C.c_int32, C.c_float, C.c_float, _fp, _fp, C.c_void_p]),
    "mrgs_diffuse_cubemap_forward": (C.c_int, [_fp, C.c_int32, _fp, C.c_void_p]),
    "mrgs_diffuse_cubemap_backward": (C.c_int, [_fp, C.c_int32, _fp, _fp, C.c_void_p]),
    "mrgs_prefilter_patch_count": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "mrgs_prefilter_texel_table": (C.c_int, [C.c_int32, _fp, C.c_void_p]),
    "mrgs_prefilter_plan_count": (C.c_int, [C.POINTER(PrefilterBuildArgs), C.c_void_p]),
    "mrgs_prefilter_plan_fill": (C.c_int, [C.POINTER(PrefilterBuildArgs), C.c_void_p]),
    "mrgs_prefilter_apply": (C.c_int, [C.POINTER(PrefilterJob), C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mrgs_mip_pyramid_forward": (C.c_int, [_fp, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p]),
    "mrgs_mip_chain_backward": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_void_p), _fp, C.c_void_p]),
    "mrgs_forward": (C.c_int, [C.POINTER(ForwardArgs), C.c_void_p]),
    "mrgs_backward": (C.c_int, [C.POINTER(BackwardArgs), C.c_void_p]),
    "mrgs_mark_visible": (C.c_int, [C.c_int32, _fp, _fp, _fp, _fp, C.c_void_p]),
    "mrgs_surfel_features_forward": (C.c_int, [C.POINTER(SurfelFeatureArgs), C.c_void_p]),
    "mrgs_surfel_features_backward": (C.c_int, [C.POINTER(SurfelFeatureArgs), C.c_void_p]),
    "mrgs_photometric_partials_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "mrgs_photometric_forward": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, C.c_int32, _fp, _fp, _fp, C.c_void_p]),
    "mrgs_photometric_backward": (C.c_int, [_fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, _fp, _fp, C.c_void_p]),
    "mrgs_geometry_loss_partials_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "mrgs_geometry_loss_forward": (C.c_int, [C.POINTER(GeometryLossArgs), C.c_void_p]),
    "mrgs_geometry_loss_backward": (C.c_int, [C.POINTER(GeometryLossArgs), C.c_void_p]),
    "mrgs_img_grad_weight": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, _fp, _fp, C.c_void_p]),
    "mrgs_densify_stats": (C.c_int, [C.c_int32, _fp, _fp, _fp, _fp, C.c_void_p]),
}

_lib = None


class MrgsError(RuntimeError):
    """Raised for any non-zero status returned by libmrgs (the reference raises RuntimeError
    from AT_ERROR / std::runtime_error, rast/rasterize_points.cu:62-64)."""


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("MRGS_AUTOBUILD", "1") != "0":
        from . import build as _build
        if not _build.is_current():
            try:
                _build.build()
            except Exception as ex:  # stale-but-present library is still usable on a box without nvcc
                if not LIB_PATH.exists():
                    raise ImportError(f"libmrgs.so is missing and could not be built: {ex}") from ex
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found. Build it with `python -m materialrefgs_b200.build`; "
            "there is no CPU fallback for the render path.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mrgs_abi_version() != MRGS_ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI version {lib.mrgs_abi_version()}, this package needs {MRGS_ABI_VERSION}: "
                          "rebuild it with `python -m materialrefgs_b200.build --force`")
    _lib = lib
    return lib


STAGES = ("preprocess_fwd", "scan", "duplicate", "sort", "ranges", "render_fwd", "render_bwd",
          "preprocess_bwd", "shade_fwd", "shade_bwd", "cubemap", "depth_sort", "prefilter_fwd", "prefilter_bwd")


def profile_read() -> dict:
    """{stage: (total_ms, calls)} since the last mrgs_profile_reset()."""
    n = len(STAGES)
    ms = (C.c_double * n)()
    calls = (C.c_int64 * n)()
    load().mrgs_profile_read(ms, calls, n)
    return {s: (ms[i], calls[i]) for i, s in enumerate(STAGES)}


def last_error() -> str:
    return load().mrgs_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != 0:
        raise MrgsError(f"{what} failed (status {status}): {last_error()}")

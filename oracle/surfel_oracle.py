"""numpy/ctypes front end of oracle/surfel_oracle.cpp (see that file's header for what each
function restates and how the oracle is pinned). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libsurfel_oracle.so"
_f = C.POINTER(C.c_float)
_d = C.POINTER(C.c_double)
_u32 = C.POINTER(C.c_uint32)


class Scene(C.Structure):
    _fields_ = [("P", C.c_int32), ("S", C.c_int32), ("D", C.c_int32), ("M", C.c_int32),
                ("W", C.c_int32), ("H", C.c_int32), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("scale_modifier", C.c_float)] + [(n, _f) for n in (
                    "background", "means3D", "shs", "colors_precomp", "features", "opacities", "scales",
                    "rotations", "transMat_precomp", "viewmatrix", "projmatrix", "campos")]


class Geom(C.Structure):
    _fields_ = [("radii", C.POINTER(C.c_int32)), ("depths", _f), ("means2D", _f), ("transMat", _f),
                ("normal_opacity", _f), ("rgb", _f), ("clamped", C.POINTER(C.c_uint8)),
                ("tiles_touched", _u32)]


class RawGrads(C.Structure):
    _fields_ = [(n, _d) for n in ("dT", "dmean2D", "dopacity", "dnormal", "dcolor", "dfeature")]


class Grads(C.Structure):
    _fields_ = [(n, _f) for n in ("dL_dmeans2D", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales",
                                  "dL_drotations")]


def build(force: bool = False) -> Path:
    src = HERE / "surfel_oracle.cpp"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "_build/libsurfel_oracle.so"], check=True,
                       capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.oracle_preprocess.restype = C.c_int64
        _lib.oracle_preprocess.argtypes = [C.POINTER(Scene), C.POINTER(Geom)]
        _lib.oracle_bin.argtypes = [C.POINTER(Scene), C.POINTER(Geom), C.c_int64,
                                    C.POINTER(C.c_uint64), _u32, _u32]
        _lib.oracle_render_forward.argtypes = [C.POINTER(Scene), C.POINTER(Geom), _u32, _u32, C.c_int,
                                               _f, _f, _f, _f, _u32]
        _lib.oracle_render_backward.argtypes = [C.POINTER(Scene), C.POINTER(Geom), _u32, _u32, C.c_int,
                                                _f, _u32, _f, _f, _f, C.POINTER(RawGrads)]
        _lib.oracle_preprocess_backward.argtypes = [C.POINTER(Scene), C.POINTER(Geom),
                                                    C.POINTER(RawGrads), C.POINTER(Grads)]
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _c(a, dtype=np.float32):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class OracleRaster:
    """One view of one cloud. Inputs are numpy arrays with the reference's shapes."""

    def __init__(self, *, means3D, opacities, viewmatrix, projmatrix, campos, W, H, tan_fovx, tan_fovy,
                 background=None, shs=None, colors_precomp=None, features=None, scales=None,
                 rotations=None, transMat_precomp=None, sh_degree=3, scale_modifier=1.0):
        self.P = int(means3D.shape[0])
        self.W, self.H = int(W), int(H)
        self.S = 0 if features is None else int(features.shape[1])
        self.M = 0 if shs is None else int(shs.shape[1])
        self.keep = dict(
            background=_c(np.zeros(3) if background is None else background), means3D=_c(means3D),
            shs=_c(shs), colors_precomp=_c(colors_precomp),
            features=_c(np.zeros((self.P, 0)) if features is None else features),
            opacities=_c(opacities).reshape(-1), scales=_c(scales), rotations=_c(rotations),
            transMat_precomp=_c(transMat_precomp), viewmatrix=_c(viewmatrix).reshape(-1),
            projmatrix=_c(projmatrix).reshape(-1), campos=_c(campos).reshape(-1))
        sc = Scene()
        sc.P, sc.S, sc.D, sc.M, sc.W, sc.H = self.P, self.S, int(sh_degree), self.M, self.W, self.H
        sc.tan_fovx, sc.tan_fovy, sc.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
        for k, v in self.keep.items():
            setattr(sc, k, _p(v, _f))
        self.sc = sc
        P = self.P
        self.geom = dict(radii=np.zeros(P, np.int32), depths=np.zeros(P, np.float32),
                         means2D=np.zeros((P, 2), np.float32), transMat=np.zeros((P, 9), np.float32),
                         normal_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
                         clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32))
        g = Geom()
        g.radii = _p(self.geom["radii"], C.POINTER(C.c_int32))
        g.clamped = _p(self.geom["clamped"], C.POINTER(C.c_uint8))
        g.tiles_touched = _p(self.geom["tiles_touched"], _u32)
        for k in ("depths", "means2D", "transMat", "normal_opacity", "rgb"):
            setattr(g, k, _p(self.geom[k], _f))
        self.g = g
        self.R = None

    @property
    def tiles(self):
        return ((self.W + 15) // 16) * ((self.H + 15) // 16)

    def preprocess(self):
        self.R = int(lib().oracle_preprocess(C.byref(self.sc), C.byref(self.g)))
        return self.R

    def bin(self):
        if self.R is None:
            self.preprocess()
        R = self.R
        self.keys = np.zeros(R, np.uint64)
        self.point_list = np.zeros(max(R, 1), np.uint32)[:R]
        self.ranges = np.zeros((self.tiles, 2), np.uint32)
        rc = lib().oracle_bin(C.byref(self.sc), C.byref(self.g), R, _p(self.keys, C.POINTER(C.c_uint64)),
                              _p(self.point_list, _u32), _p(self.ranges, _u32))
        assert rc == 0
        return self.keys, self.point_list, self.ranges

    def forward(self, tile_step: int = 1):
        if not hasattr(self, "ranges"):
            self.bin()
        H, W, S = self.H, self.W, self.S
        self.out_color = np.zeros((3, H, W), np.float32)
        self.out_feature = np.zeros((S, H, W), np.float32)
        self.out_others = np.zeros((7, H, W), np.float32)
        self.final_T = np.zeros((3, H, W), np.float32)
        self.n_contrib = np.zeros((2, H, W), np.uint32)
        lib().oracle_render_forward(C.byref(self.sc), C.byref(self.g), _p(self.point_list, _u32),
                                    _p(self.ranges, _u32), int(tile_step), _p(self.out_color, _f),
                                    _p(self.out_feature, _f), _p(self.out_others, _f), _p(self.final_T, _f),
                                    _p(self.n_contrib, _u32))
        return self.out_color, self.out_feature, self.out_others

    def backward(self, dL_dcolor, dL_dfeature, dL_dothers, tile_step: int = 1):
        P, S, M = self.P, self.S, self.M
        raw = dict(dT=np.zeros((P, 9)), dmean2D=np.zeros((P, 2)), dopacity=np.zeros(P),
                   dnormal=np.zeros((P, 3)), dcolor=np.zeros((P, 3)), dfeature=np.zeros((P, max(S, 1)))[:, :S].copy())
        rg = RawGrads()
        for k, v in raw.items():
            setattr(rg, k, _p(v, _d))
        gc, gf, go = _c(dL_dcolor), _c(dL_dfeature), _c(dL_dothers)
        lib().oracle_render_backward(C.byref(self.sc), C.byref(self.g), _p(self.point_list, _u32),
                                     _p(self.ranges, _u32), int(tile_step), _p(self.final_T, _f),
                                     _p(self.n_contrib, _u32), _p(gc, _f), _p(gf, _f), _p(go, _f), C.byref(rg))
        out = dict(dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dmeans3D=np.zeros((P, 3), np.float32),
                   dL_dtransMat=np.zeros((P, 9), np.float32), dL_dsh=np.zeros((P, M, 3), np.float32),
                   dL_dscales=np.zeros((P, 2), np.float32), dL_drotations=np.zeros((P, 4), np.float32))
        gr = Grads()
        for k, v in out.items():
            setattr(gr, k, _p(v, _f))
        lib().oracle_preprocess_backward(C.byref(self.sc), C.byref(self.g), C.byref(rg), C.byref(gr))
        out["dL_dopacity"] = raw["dopacity"].astype(np.float32).reshape(P, 1)
        out["dL_dcolors"] = raw["dcolor"].astype(np.float32)
        out["dL_dfeatures"] = raw["dfeature"].astype(np.float32)
        return out


def from_synthetic(cloud, cam, bg=None, sh_degree=3, scale_modifier=1.0):
    """Build an OracleRaster from materialrefgs_b200.synthetic objects (CPU tensors)."""
    n = lambda t: t.detach().cpu().numpy()
    return OracleRaster(means3D=n(cloud.means3D), opacities=n(cloud.opacities), viewmatrix=n(cam.world_view_transform),
                        projmatrix=n(cam.full_proj_transform), campos=n(cam.camera_center), W=cam.image_width,
                        H=cam.image_height, tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, background=bg,
                        shs=n(cloud.shs), features=n(cloud.features), scales=n(cloud.scales),
                        rotations=n(cloud.rotations), sh_degree=sh_degree, scale_modifier=scale_modifier)

"""Test-only access to the oracle: the reference CUDA extension in oracle/_ref and decoders for
its opaque scratch buffers (layout: rast/cuda_rasterizer/rasterizer_impl.cu:157-196, obtain()
in rasterizer_impl.h:23-29 = align the running pointer up to 128 B, then take count*sizeof(T))."""
from __future__ import annotations

import ctypes as C
import importlib
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF_DIR = ROOT / "oracle" / "_ref"


def load_reference():
    if not (REF_DIR / "diff_surfel_rasterization").exists():
        return None
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    try:
        import torch  # noqa: F401  (the extension links against libtorch)
        return importlib.import_module("diff_surfel_rasterization")
    except Exception as ex:  # pragma: no cover
        print(f"[refimpl] cannot import oracle/_ref: {ex}")
        return None


def _carve(buf: torch.Tensor, specs):
    """specs: list of (name, dtype, count). Returns dict of typed views following obtain()."""
    out = {}
    off = 0
    for name, dtype, count in specs:
        off = (off + 127) // 128 * 128
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        out[name] = buf[off:off + nbytes].view(dtype)
        off += nbytes
    return out


def decode_ref_geom(geom: torch.Tensor, P: int):
    d = _carve(geom, [("depths", torch.float32, P), ("clamped", torch.uint8, 3 * P),
                      ("internal_radii", torch.int32, P), ("means2D", torch.float32, 2 * P),
                      ("transMat", torch.float32, 9 * P), ("normal_opacity", torch.float32, 4 * P),
                      ("rgb", torch.float32, 3 * P), ("tiles_touched", torch.int32, P)])
    d["clamped"] = d["clamped"].view(P, 3)
    d["means2D"] = d["means2D"].view(P, 2)
    d["transMat"] = d["transMat"].view(P, 9)
    d["normal_opacity"] = d["normal_opacity"].view(P, 4)
    d["rgb"] = d["rgb"].view(P, 3)
    return d


def decode_ref_image(img: torch.Tensor, H: int, W: int):
    N = H * W
    d = _carve(img, [("accum_alpha", torch.float32, 3 * N), ("n_contrib", torch.int32, 2 * N),
                     ("ranges", torch.int32, 2 * N)])
    d["accum_alpha"] = d["accum_alpha"].view(3, H, W)
    d["n_contrib"] = d["n_contrib"].view(2, H, W)
    d["ranges"] = d["ranges"].view(N, 2)
    return d


def decode_ref_binning(binning: torch.Tensor, R: int):
    return _carve(binning, [("point_list", torch.int32, R), ("point_list_unsorted", torch.int32, R),
                            ("keys", torch.int64, R), ("keys_unsorted", torch.int64, R)])


# ---- decoders for libmrgs' own buffers (include/mrgs.h layouts) --------------------------------
def decode_mrgs_geom(geom: torch.Tensor, P: int, S: int):
    from materialrefgs_b200 import _lib
    lib = _lib.load()
    gl = _lib.GeomLayout()
    assert lib.mrgs_geom_layout(P, S, C.byref(gl)) == 0
    rec = geom[gl.rec:gl.rec + P * 64].view(torch.float32).view(P, 16)
    cf = geom[gl.cf:gl.cf + P * gl.cf_stride * 4].view(torch.float32).view(P, gl.cf_stride)
    d = {
        "transMat": torch.cat([rec[:, 0:3], rec[:, 4:7], rec[:, 3:4], rec[:, 7:8], rec[:, 8:9]], 1),
        "means2D": rec[:, 9:11], "opacity": rec[:, 11], "normal": rec[:, 12:15], "tau": rec[:, 15],
        "depths": geom[gl.depth:gl.depth + 4 * P].view(torch.float32),
        "bbox": geom[gl.bbox:gl.bbox + 32 * P].view(torch.float32).view(P, 8),
        "rgb": cf[:, 0:3], "features": cf[:, 3:3 + S],
        "clamped": geom[gl.clamped:gl.clamped + P],
        "tiles_touched": geom[gl.tiles_touched:gl.tiles_touched + 4 * P].view(torch.int32),
        "point_offsets": geom[gl.point_offsets:gl.point_offsets + 4 * P].view(torch.int32),
    }
    return d


def decode_mrgs_image(img: torch.Tensor, H: int, W: int):
    from materialrefgs_b200 import _lib
    lib = _lib.load()
    il = _lib.ImageLayout()
    assert lib.mrgs_image_layout(W, H, C.byref(il)) == 0
    gx, gy = (W + 15) // 16, (H + 15) // 16
    tiles = gx * gy
    state = img[il.state:il.state + tiles * 5 * 256 * 4].view(torch.float32).view(gy, gx, 5, 256)
    # slot -> (x,y) inside the tile
    slot = torch.arange(256, device=img.device)
    sx = ((slot >> 5) & 1) * 8 + (slot & 7)
    sy = (slot >> 6) * 4 + ((slot >> 3) & 3)
    planar = torch.zeros((5, gy * 16, gx * 16), dtype=torch.float32, device=img.device)
    ty = torch.arange(gy, device=img.device)[:, None, None] * 16 + sy[None, None, :]
    tx = torch.arange(gx, device=img.device)[None, :, None] * 16 + sx[None, None, :]
    ty = ty.expand(gy, gx, 256)
    tx = tx.expand(gy, gx, 256)
    for pl in range(5):
        planar[pl, ty, tx] = state[:, :, pl, :]
    planar = planar[:, :H, :W]
    return {
        "final_T": planar[0], "M1": planar[1], "M2": planar[2],
        "n_contrib": planar[3].contiguous().view(torch.int32),
        "median_contrib": planar[4].contiguous().view(torch.int32),
        "ranges": img[il.ranges:il.ranges + tiles * 8].view(torch.int32).view(tiles, 2),
    }


def decode_mrgs_binning(binning: torch.Tensor, R: int, depths: torch.Tensor = None):
    """libmrgs keeps the sorted surfel ids and the 16-bit tile id of every sorted instance; the
    reference's 64-bit key (tile << 32 | depth bits) is rebuilt from them for bit-exact comparison."""
    from materialrefgs_b200 import _lib
    lib = _lib.load()
    bl = _lib.BinningLayout()
    assert lib.mrgs_binning_layout(int(getattr(binning, "mrgs_capacity", 0)) or R, C.byref(bl)) == 0
    point_list = binning[bl.point_list:bl.point_list + 4 * R].view(torch.int32)
    tiles = binning[bl.keys:bl.keys + 2 * R].view(torch.int16).to(torch.int64) & 0xffff
    out = {"point_list": point_list, "tile_ids": tiles}
    if depths is not None:
        bits = depths.view(torch.int32)[point_list.long()].to(torch.int64) & 0xffffffff
        out["keys"] = (tiles << 32) | bits
    return out

"""Builds libmrgs.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m materialrefgs_b200.build [--force] [--verbose]

The .so lands next to this file so that it travels with a gpurun snapshot; it is git-ignored.
nvcc cross-compiles without a GPU. No -use_fast_math: expf / division / sqrtf must be the same
IEEE-grade routines the reference binary uses (bit-exact contributor counts depend on it).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libmrgs.so"

SOURCES = [
    "api.cu",
    "preprocess.cu",
    "binning.cu",
    "render_fwd.cu",
    "render_bwd.cu",
    "preprocess_bwd.cu",
    "shade.cu",
    "cubemap.cu",
    "prefilter.cu",
    "features.cu",
    "losses.cu",
    "geomloss.cu",
]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-diag-suppress", "177",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libmrgs.so cannot be built")


def _sources() -> list[Path]:
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def _stamp() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "mrgs.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp_file = BUILD_DIR / "stamp"
    return LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == _stamp()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    BUILD_DIR.mkdir(exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: Path) -> Path:
        obj = BUILD_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-gencode",
            "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lcudart"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    (BUILD_DIR / "stamp").write_text(_stamp())
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))

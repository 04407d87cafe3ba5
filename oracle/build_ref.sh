#!/usr/bin/env bash
# Build the UNMODIFIED reference CUDA rasterizer (submodules/diff-surfel-rasterization of
# /root/reference) for sm_100a into oracle/_ref/ (git-ignored, travels to the GPU box).
# Test infrastructure only: nothing under materialrefgs_b200/ may import it.
#
# The sources are compiled from a scratch copy under /tmp because setuptools writes build/
# next to setup.py and /root/reference is read-only. No source file is edited; the only
# workaround is `-include cstdint` (gcc 13 needs it for rasterizer_impl.h:26).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${MRGS_REFERENCE_ROOT:-/root/reference}/submodules/diff-surfel-rasterization"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF not present; keeping prebuilt oracle/_ref (if any)"; exit 0
fi
if ls "$OUT"/diff_surfel_rasterization/_C*.so >/dev/null 2>&1 && [ -z "${MRGS_REF_REBUILD:-}" ]; then
  echo "[build_ref] oracle/_ref already built"; exit 0
fi
TMP="$(mktemp -d /tmp/mrgs_ref.XXXXXX)"
cp -r "$REF" "$TMP/src"
mkdir -p "$OUT"
export TORCH_CUDA_ARCH_LIST="10.0a"
export NVCC_PREPEND_FLAGS="-include cstdint"
export MAX_JOBS="${MAX_JOBS:-8}"
python -m pip install --no-index --no-build-isolation --no-deps --upgrade \
    --target "$OUT" "$TMP/src" 2>&1 | tail -n 5
rm -rf "$TMP"
ls -la "$OUT"/diff_surfel_rasterization/

"""Builds the UNMODIFIED nvdiffrec `renderutils_plugin` of the reference (scene/renderutils/c_src)
for sm_100a into oracle/_ref/renderutils_plugin/ (git-ignored; travels to the GPU box).
TEST INFRASTRUCTURE ONLY. Sources are compiled where they lie under /root/reference; only object
files and the .so are written, all under oracle/_ref/. Used to pin the cubemap prefilter kernels
(specular_cubemap / diffuse_cubemap / specular_bounds, ops.py:391-458)."""
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref" / "renderutils_plugin"
REF = Path(os.environ.get("MRGS_REFERENCE_ROOT", "/root/reference")) / "scene" / "renderutils" / "c_src"


def main():
    if not REF.exists():
        print(f"[build_ref_renderutils] {REF} not present; keeping prebuilt plugin (if any)")
        return 0
    if (OUT / "renderutils_plugin.so").exists() and not os.environ.get("MRGS_REF_REBUILD"):
        print("[build_ref_renderutils] already built")
        return 0
    OUT.mkdir(parents=True, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    import torch.utils.cpp_extension as ext
    srcs = [str(REF / f) for f in ("mesh.cu", "loss.cu", "bsdf.cu", "normal.cu", "cubemap.cu", "common.cpp",
                                   "torch_bindings.cpp")]
    stubs = "/usr/local/cuda/lib64/stubs"
    ext.load(name="renderutils_plugin", sources=srcs, extra_cflags=["-DNVDR_TORCH"],
             extra_cuda_cflags=["-DNVDR_TORCH"], extra_ldflags=[f"-L{stubs}", "-lcuda", "-lnvrtc"],
             with_cuda=True, verbose=False, build_directory=str(OUT), is_python_module=False)
    print(sorted(p.name for p in OUT.iterdir()))
    return 0


if __name__ == "__main__":
    sys.exit(main())

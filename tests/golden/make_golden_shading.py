"""Generates tests/golden/shading_*.npz with the REFERENCE's own shading code — utils/refl_utils.py
(sample_camera_rays, reflection, get_specular_color_surfel, get_full_color_volume), scene/light.py (EnvLight.get_mip,
EnvLight.__call__), utils/general_utils.py (safe_normalize) — imported from /root/reference and run on the CPU of the
build container.

nvdiffrast (requirements.txt:57, a local unversioned path) is neither installed nor vendored, so `dr.texture`, the ONE
symbol this path needs from it, is supplied by the restatement in oracle/shading_oracle.py (lut_fetch / cube_texture).
These vectors therefore pin everything the reference itself does AROUND the texture fetch — ray construction, the
transposed-R convention, reflection, clamps, the roughness -> mip mapping, how `__call__` reshapes and applies the
sigmoid, the specular weight, the `fg[0]` indexing of get_full_color_volume — but not the texel addressing inside
dr.texture (see the oracle's header: that part stays "parity unpinned").

Arrangements for the run (no reference source is edited or copied): absent modules are empty stubs; Tensor.cuda is the
identity; the working directory is /root/reference while utils/refl_utils.py is imported (it reads
assets/bsdf_256_256.bin by a relative path at import, :9)."""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, str(ROOT))
from oracle import shading_oracle as so  # noqa: E402
from materialrefgs_b200 import synthetic  # noqa: E402


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **kw):
        return None


MISSING = {"plyfile", "cubemapencoder", "simple_knn", "raytracing_brdf", "kornia", "matplotlib", "open3d", "imageio",
           "lpips", "trimesh", "diff_surfel_tracing", "diff_surfel_rasterization", "diff_surfel_rasterization2",
           "_raytracing_brdf", "ipdb", "tinycudann", "pytorch3d", "skimage", "mediapy", "pyexr", "OpenEXR", "Imath", "xatlas"}


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] not in MISSING:
            return None
        return importlib.machinery.ModuleSpec(name, self)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def texture(tex, uv, mip=None, mip_level_bias=None, filter_mode="linear", boundary_mode="wrap", **kw):
    """dr.texture as the reference calls it (refl_utils.py:374, :439; light.py:110, :113, :118-125), by the oracle."""
    if boundary_mode == "clamp":                      # FG LUT [1,256,256,2], uv [1,N,1,2]
        out = so.lut_fetch(tex, uv.reshape(-1, 2))
        return out.reshape(*uv.shape[:-1], tex.shape[-1])
    assert boundary_mode == "cube"
    levels = [tex[0]] + ([m[0] for m in mip] if mip is not None else [])
    d = uv.reshape(-1, 3)
    lvl = None if mip_level_bias is None else mip_level_bias.reshape(-1)
    out = so.cube_texture(levels, d, lvl)
    return out.reshape(*uv.shape[:-1], tex.shape[-1])


sys.meta_path.append(_StubFinder())
sys.modules["raytracing_brdf"] = _Stub("raytracing_brdf")   # a package of the reference that needs its CUDA extension
nv = types.ModuleType("nvdiffrast")
nvt = types.ModuleType("nvdiffrast.torch")
nvt.texture = texture
nv.torch = nvt
sys.modules["nvdiffrast"], sys.modules["nvdiffrast.torch"] = nv, nvt
torch.Tensor.cuda = lambda self, *a, **k: self
_cwd = os.getcwd()
os.chdir("/root/reference")
from utils import refl_utils as ru  # noqa: E402  (the reference's)
from scene.light import EnvLight  # noqa: E402  (the reference's)
os.chdir(_cwd)


def reference_envlight(levels, diffuse, min_r=0.08, max_r=0.5):
    env = object.__new__(EnvLight)            # __init__ would build the mips with the CUDA plugin
    torch.nn.Module.__init__(env)
    env.min_roughness, env.max_roughness = min_r, max_r
    env.base = levels[0]
    env.specular = list(levels)
    env.diffuse = diffuse
    return env


def main():
    g = torch.Generator().manual_seed(21)
    # ---- deferred: get_specular_color_surfel on a G-buffer (render_surfel, gaussian_renderer/__init__.py:372-445)
    for name, (view, W, H, res) in {"surfel_a": (2, 40, 28, 32), "surfel_b": (5, 33, 47, 64)}.items():
        cam = synthetic.orbit_camera(view, 8, W, H)
        levels = [l.clone().requires_grad_(True) for l in so.synthetic_chain(res, 16, seed=view)]
        env = reference_envlight(levels, None)
        leaf = dict(albedo=torch.rand(H, W, 3, generator=g), normal=torch.nn.functional.normalize(torch.randn(H, W, 3, generator=g), dim=-1)
                    * (0.6 + 0.8 * torch.rand(H, W, 1, generator=g)), alpha=torch.rand(H, W, 1, generator=g),
                    refl=torch.rand(H, W, 1, generator=g), rough=torch.rand(H, W, 1, generator=g) * 1.3 - 0.15)
        leaf = {k: v.requires_grad_(True) for k, v in leaf.items()}
        ru.pixel_camera = None               # the reference caches the pixel grid per H (:58-59)
        R, T = torch.tensor(cam.R), torch.tensor(cam.T)
        spec, extra = ru.get_specular_color_surfel(env, leaf["albedo"], cam.HWK, R, T, leaf["normal"], leaf["alpha"],
                                                   refl_strength=leaf["refl"], roughness=leaf["rough"],
                                                   pc=types.SimpleNamespace(ray_tracer=None))
        w = torch.randn(3, H, W, generator=g)
        (spec * w).sum().backward()
        rays_d, rays_o = ru.sample_camera_rays(cam.HWK, R, T)
        out = dict(view=np.array([view, W, H, res]), w=w.numpy(), specular=spec.detach().numpy(),
                   direct_light=extra["direct_light"].detach().numpy(), specular_weight=extra["specular_weight"].detach().numpy(),
                   rays_d=rays_d.numpy(), rays_o=rays_o.numpy())
        for k, v in leaf.items():
            out[k] = v.detach().numpy()
            out["grad_" + k] = v.grad.numpy()
        for i, l in enumerate(levels):
            out[f"grad_level{i}"] = l.grad.numpy()
        np.savez_compressed(ROOT / "tests" / "golden" / f"shading_{name}.npz", **out)
        print("wrote", name)

    # ---- per surfel: get_full_color_volume (render_volume, gaussian_renderer/__init__.py:640-645)
    N, res = 600, 32
    cam = synthetic.orbit_camera(3, 8, 64, 64)
    levels = [l.clone().requires_grad_(True) for l in so.synthetic_chain(res, 16, seed=4)]
    diffuse = (0.7 * torch.randn(6, 16, 16, 3, generator=g)).requires_grad_(True)
    env = reference_envlight(levels, diffuse)
    leaf = dict(xyz=1.3 * (2 * torch.rand(N, 3, generator=g) - 1), normal=torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1),
                albedo=torch.rand(N, 3, generator=g), refl=torch.rand(N, 1, generator=g), rough=torch.rand(N, 1, generator=g))
    leaf = {k: v.requires_grad_(True) for k, v in leaf.items()}
    R, T = torch.tensor(cam.R), torch.tensor(cam.T)
    d, s = ru.get_full_color_volume(env, leaf["xyz"], leaf["albedo"], cam.HWK, R, T, leaf["normal"], torch.ones(N, 1),
                                    refl_strength=leaf["refl"], roughness=leaf["rough"])
    wd, ws = torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g)
    ((d * wd).sum() + (s * ws).sum()).backward()
    out = dict(view=np.array([3, 64, 64, res]), wd=wd.numpy(), ws=ws.numpy(), diffuse=d.detach().numpy(), specular=s.detach().numpy(),
               diffuse_map=diffuse.detach().numpy(), grad_diffuse_map=diffuse.grad.numpy())
    for k, v in leaf.items():
        out[k] = v.detach().numpy()
        out["grad_" + k] = v.grad.numpy()
    for i, l in enumerate(levels):
        out[f"grad_level{i}"] = l.grad.numpy()
    np.savez_compressed(ROOT / "tests" / "golden" / "shading_volume.npz", **out)
    print("wrote volume")


if __name__ == "__main__":
    main()

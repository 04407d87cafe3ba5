"""Generates tests/golden/densify_*.npz with the REFERENCE's own GaussianModel bookkeeping (scene/gaussian_model.py:
training_setup, add_densification_stats, densify_and_prune -> densify_and_clone / densify_and_split / prune_points /
cat_tensors_to_optimizer / _prune_optimizer, reset_opacity0), imported from /root/reference and run on the CPU of
the build container.

What had to be arranged for that (no reference source is edited or copied):
  * modules the file imports but this path never calls (plyfile, cubemapencoder, simple_knn, raytracing_brdf,
    nvdiffrast, kornia, ...) are absent here: the names in MISSING are replaced by empty stub modules;
  * the reference hard-codes device="cuda" in torch.zeros calls (gaussian_model.py:938,961,979-981 and
    general_utils.py:83): torch.zeros is wrapped to drop the device argument while the script runs;
  * GaussianModel.__init__ builds CUDA environment grids, so the object is made with __new__ and given exactly the
    attributes these methods read.
The random split offsets come from torch.normal under torch.manual_seed: the product draws the same shapes in the
same order, so on the CPU the results are comparable bit for bit."""
import importlib.abc
import importlib.machinery
import sys
import types
from pathlib import Path

import numpy as np
import torch
from torch import nn

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, str(ROOT))


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **kw):
        return None


MISSING = {"plyfile", "cubemapencoder", "simple_knn", "raytracing_brdf", "nvdiffrast", "kornia", "matplotlib", "open3d",
           "imageio", "lpips", "trimesh", "diff_surfel_tracing", "diff_surfel_rasterization", "diff_surfel_rasterization2",
           "_raytracing_brdf", "tinycudann", "pytorch3d", "skimage", "mediapy", "pyexr", "OpenEXR", "Imath", "xatlas"}


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] not in MISSING:
            return None
        return importlib.machinery.ModuleSpec(name, self)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


sys.meta_path.append(_StubFinder())          # last resort only: real modules win
# utils/refl_utils.py:9 loads the LUT from a relative path straight onto a CUDA device at import; this path never
# calls into it
sys.modules["utils.refl_utils"] = _Stub("utils.refl_utils")
sys.modules["raytracing_brdf"] = _Stub("raytracing_brdf")     # same at raytracing_brdf/raytracer.py:13
_zeros = torch.zeros
torch.zeros = lambda *a, **kw: _zeros(*a, **{k: v for k, v in kw.items() if k != "device"})
torch.Tensor.cuda = lambda self, *a, **k: self      # reset_opacity1 builds its constant with .cuda() (:541)

from scene.gaussian_model import GaussianModel  # noqa: E402  (the reference's)
from utils.general_utils import inverse_sigmoid  # noqa: E402

FIELDS = {  # attribute -> trailing shape (scene/gaussian_model.py:379-410)
    "_xyz": (3,), "_features_dc": (1, 3), "_features_rest": (15, 3), "_opacity": (1,), "_scaling": (2,), "_rotation": (4,),
    "_refl_strength": (1,), "_ori_color": (3,), "_diffuse_color": (3,), "_roughness": (1,), "_metalness": (1,),
    "_normal1": (3,), "_normal2": (3,), "_indirect_dc": (1, 3), "_indirect_rest": (15, 3), "_indirect_asg": (32, 5),
}
GROUP_OF = {"_xyz": "xyz", "_features_dc": "f_dc", "_features_rest": "f_rest", "_opacity": "opacity", "_scaling": "scaling",
            "_rotation": "rotation", "_refl_strength": "refl_strength", "_ori_color": "ori_color",
            "_diffuse_color": "diffuse_color", "_roughness": "roughness", "_metalness": "metalness", "_normal1": "normal1",
            "_normal2": "normal2", "_indirect_dc": "ind_dc", "_indirect_rest": "ind_rest", "_indirect_asg": "ind_asg"}


def initial_fields(P, seed):
    g = torch.Generator().manual_seed(seed)
    f = {k: 0.3 * torch.randn((P, *shp), generator=g) for k, shp in FIELDS.items()}
    f["_scaling"] = torch.log(0.02 * torch.exp(0.8 * torch.randn((P, 2), generator=g)))   # around percent_dense * extent
    f["_opacity"] = 2.0 * torch.randn((P, 1), generator=g) - 1.0
    return f


def make_reference_model(fields, percent_dense):
    m = object.__new__(GaussianModel)
    for k, v in fields.items():
        setattr(m, k, nn.Parameter(v.clone().requires_grad_(True)))
    m.scaling_activation, m.scaling_inverse_activation = torch.exp, torch.log
    m.opacity_activation, m.inverse_opacity_activation = torch.sigmoid, inverse_sigmoid
    m.rotation_activation = torch.nn.functional.normalize
    m.max_sh_degree = m.active_sh_degree = 3
    m.spatial_lr_scale = 1.0
    m.env_map, m.env_map_2 = nn.Linear(1, 1), nn.Linear(1, 1)
    m.max_radii2D = torch.zeros((fields["_xyz"].shape[0]))
    args = types.SimpleNamespace(percent_dense=percent_dense, position_lr_init=1.6e-4, position_lr_final=1.6e-6,
                                 position_lr_delay_mult=0.01, position_lr_max_steps=30000, features_lr=0.0025,
                                 opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001, envmap_cubemap_lr=0.01,
                                 refl_strength_lr=0.002, ori_color_lr=0.002, roughness_lr=0.002, metalness_lr=0.002,
                                 normal_lr=0.006, indirect_lr=0.0025, asg_lr=0.001)
    m.training_setup(args)
    return m, args


def adam_warmup(model_params, optimizer, steps, seed):
    g = torch.Generator().manual_seed(seed)
    for _ in range(steps):
        for p in model_params:
            if p.requires_grad:
                p.grad = 0.01 * torch.randn(p.shape, generator=g)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)


def snapshot(m):
    out = {}
    for attr, grp in GROUP_OF.items():
        p = getattr(m, attr)
        out["p_" + grp] = p.detach().numpy().copy()
        st = m.optimizer.state.get(p, None)
        if st is not None and "exp_avg" in st:
            out["m_" + grp] = st["exp_avg"].numpy().copy()
            out["v_" + grp] = st["exp_avg_sq"].numpy().copy()
    out["xyz_gradient_accum"] = m.xyz_gradient_accum.numpy().copy()
    out["denom"] = m.denom.numpy().copy()
    out["max_radii2D"] = m.max_radii2D.numpy().copy()
    return out


CASES = {  # name: (P, seed, views, max_grad, min_opacity, extent, max_screen_size, percent_dense)
    "clone_split_prune": (200, 5, 3, 0.0002, 0.05, 2.5, 20, 0.01),
    "no_screen_size": (96, 6, 2, 0.0004, 0.005, 4.0, None, 0.01),
}


def main():
    for name, (P, seed, views, max_grad, min_opacity, extent, max_screen, percent_dense) in CASES.items():
        fields = initial_fields(P, seed)
        m, args = make_reference_model(fields, percent_dense)
        adam_warmup([getattr(m, a) for a in FIELDS], m.optimizer, 2, seed + 100)
        before = snapshot(m)
        g = torch.Generator().manual_seed(seed + 200)
        view_grads, view_filters, view_radii = [], [], []
        for _ in range(views):   # train_refnerf.py:1416-1418
            vs = types.SimpleNamespace(grad=4e-4 * torch.randn((P, 3), generator=g) * (torch.rand((P, 1), generator=g) > 0.3))
            radii = torch.randint(0, 40, (P,), generator=g) * (torch.rand((P,), generator=g) > 0.25)
            filt = radii > 0
            m.max_radii2D[filt] = torch.max(m.max_radii2D[filt], radii[filt].float())
            m.add_densification_stats(vs, filt)
            view_grads.append(vs.grad.numpy().copy()); view_filters.append(filt.numpy().copy()); view_radii.append(radii.numpy().copy())
        stats = dict(xyz_gradient_accum=m.xyz_gradient_accum.numpy().copy(), denom=m.denom.numpy().copy(),
                     max_radii2D=m.max_radii2D.numpy().copy())
        torch.manual_seed(seed + 300)
        m.densify_and_prune(max_grad, min_opacity, extent, max_screen)
        after = snapshot(m)
        m.reset_opacity0()
        after_reset = {"p_opacity": m._opacity.detach().numpy().copy(),
                       "m_opacity": m.optimizer.state[m._opacity]["exp_avg"].numpy().copy()}
        out = dict(P=P, seed=seed, max_grad=max_grad, min_opacity=min_opacity, extent=extent,
                   max_screen_size=-1 if max_screen is None else max_screen, percent_dense=percent_dense,
                   view_grads=np.stack(view_grads), view_filters=np.stack(view_filters), view_radii=np.stack(view_radii))
        out.update({"before_" + k: v for k, v in before.items()})
        out.update({"stats_" + k: v for k, v in stats.items()})
        out.update({"after_" + k: v for k, v in after.items()})
        out.update({"reset_" + k: v for k, v in after_reset.items()})
        path = ROOT / "tests" / "golden" / f"densify_{name}.npz"
        np.savez_compressed(path, **out)
        print("wrote", path, "P", P, "->", after["p_xyz"].shape[0])


def checkpoint_and_schedule():
    """A small checkpoint written the way train_refnerf.py does it (torch.save((gaussians.capture(), iteration), path),
    scene/gaussian_model.py:124-148) and samples of the reference's learning-rate schedule (general_utils.py:29-63)."""
    from utils.general_utils import get_expon_lr_func
    P, seed = 12, 9
    m, args = make_reference_model(initial_fields(P, seed), 0.01)
    adam_warmup([getattr(m, a) for a in FIELDS], m.optimizer, 3, seed + 1)
    m.active_sh_degree = 2
    m.spatial_lr_scale = 3.5
    m.xyz_gradient_accum += 0.25
    m.denom += 2
    m.max_radii2D += 7
    torch.save((m.capture(), 1234), ROOT / "tests" / "golden" / "chkpnt_reference_small.pth")
    steps = np.array([-1, 0, 1, 10, 500, 7000, 29999, 30000, 45000], dtype=np.int64)
    f1 = get_expon_lr_func(lr_init=1.6e-4 * 3.5, lr_final=1.6e-6 * 3.5, lr_delay_mult=0.01, max_steps=30000)
    f2 = get_expon_lr_func(lr_init=0.01, lr_final=0.001, lr_delay_steps=1000, lr_delay_mult=0.1, max_steps=20000)
    np.savez(ROOT / "tests" / "golden" / "lr_schedule.npz", steps=steps, plain=np.array([f1(int(s)) for s in steps]),
             delayed=np.array([f2(int(s)) for s in steps]))
    print("wrote chkpnt_reference_small.pth, lr_schedule.npz")


def resets():
    """reset_opacity1 / reset_refl / dist_color / reset_scale as the training loop chains them after an opacity reset
    (train_refnerf.py:1439-1455), with and without an exclusion mask."""
    P, seed = 150, 12
    m, args = make_reference_model(initial_fields(P, seed), 0.01)
    m.init_refl_value, m.enlarge_scale, m.refl_msk_thr, m.rough_msk_thr = 0.1, 1.5, 0.02, 0.1
    m.refl_activation = m.roughness_activation = torch.sigmoid
    m.inverse_refl_activation = inverse_sigmoid
    with torch.no_grad():     # spread the material parameters over both sides of the thresholds
        g = torch.Generator().manual_seed(seed + 5)
        m._refl_strength.copy_(torch.randn(P, 1, generator=g) * 3 - 3)
        m._roughness.copy_(torch.randn(P, 1, generator=g) * 2 - 2)
        m._opacity.copy_(torch.randn(P, 1, generator=g) * 3 + 1)
    adam_warmup([getattr(m, a) for a in FIELDS], m.optimizer, 2, seed + 1)
    out = {"P": P, "seed": seed}
    out.update({"before_" + k: v for k, v in snapshot(m).items()})
    msk = torch.rand(P, generator=torch.Generator().manual_seed(seed + 6)) < 0.3
    out["mask"] = msk.numpy()
    for tag, mask in (("plain", None), ("masked", msk)):
        m.reset_opacity0()
        m.reset_refl(exclusive_msk=mask, rst_value=0.1 if mask is not None else None)
        m.reset_opacity1(exclusive_msk=mask)
        torch.manual_seed(seed + 7)
        m.dist_color(exclusive_msk=mask)
        m.reset_scale(exclusive_msk=mask)
        out.update({f"{tag}_" + k: v for k, v in snapshot(m).items()
                    if k[:2] in ("p_", "m_", "v_") and k[2:] in ("opacity", "refl_strength", "f_dc", "scaling")})
    np.savez_compressed(ROOT / "tests" / "golden" / "resets.npz", **out)
    print("wrote resets.npz")


if __name__ == "__main__":
    main()
    checkpoint_and_schedule()
    resets()

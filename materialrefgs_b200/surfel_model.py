"""Surfel parameter store: the data formats and bookkeeping on either side of the render path.

Mirrors the part of the reference's GaussianModel (scene/gaussian_model.py) that a view-sharded training step needs
in order to run an actual loop around the rasterizer:
  * the parameter fields, their shapes and optimizer group names                (:379-410, :422-447)
  * the `.ply` attribute order and per-field transposes                         (:462-523 save_ply, :725-836 load_ply)
  * densification statistics, clone / split / prune, Adam-state surgery         (:856-1061)
  * reset_opacity0 and the reflection-aware resets the loop chains after it      (:531-545, :558-565, :601-611, :626-671)
  * the checkpoint tuple of capture() / restore() (`chkpnt<iter>.pth`)           (:124-172)
  * the position learning-rate schedule                                         (:455-460, utils/general_utils.py:29-63)
Pinned by tests/golden/densify_*.npz, produced by the reference's own methods (tests/golden/make_golden_densify.py).

Multi-GPU: every rank holds the full store. After `GradArena.allreduce()` the statistics are identical on all ranks;
the only random draw (the split offsets) comes from a generator the caller seeds identically on every rank, so all
ranks densify to the same cloud without a broadcast (tests/test_parallel_cpu.py checks it over gloo).
This is host-side plumbing in PyTorch; the per-view statistics kernel is `mrgs_densify_stats` (parallel.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch
from torch import nn


# optimizer group order of the reference's training_setup (gaussian_model.py:422-447); "env"/"env2" are the two EnvLight
# modules' parameters, which live outside this store
REFERENCE_GROUP_ORDER = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "env", "env2", "refl_strength", "ori_color",
                         "diffuse_color", "roughness", "metalness", "normal1", "normal2", "ind_dc", "ind_rest", "ind_asg")


@dataclass(frozen=True)
class Field:
    group: str            # optimizer group name (gaussian_model.py:422-447)
    shape: tuple          # trailing shape of the parameter
    ply: str              # attribute prefix in the .ply
    transposed: bool      # stored as transpose(1, 2).flatten(1) (save_ply :494-498)
    numbered: bool = True # attributes are `<prefix>_<i>`; False: a single attribute called `<prefix>`


# Order = the column order of save_ply's `attributes` (gaussian_model.py:517) after x, y, z
FIELDS = {
    "xyz": Field("xyz", (3,), "", False),
    "normal1": Field("normal1", (3,), "n", False),
    "normal2": Field("normal2", (3,), "n2", False),
    "features_dc": Field("f_dc", (1, 3), "f_dc", True),
    "features_rest": Field("f_rest", (15, 3), "f_rest", True),
    "indirect_dc": Field("ind_dc", (1, 3), "ind_dc", True),
    "indirect_rest": Field("ind_rest", (15, 3), "ind_rest", True),
    "indirect_asg": Field("ind_asg", (32, 5), "ind_asg", True),
    "opacity": Field("opacity", (1,), "opacity", False, False),
    "refl_strength": Field("refl_strength", (1,), "refl_strength", False, False),
    "metalness": Field("metalness", (1,), "metalness", False, False),
    "roughness": Field("roughness", (1,), "roughness", False, False),
    "ori_color": Field("ori_color", (3,), "ori_color", False),
    "diffuse_color": Field("diffuse_color", (3,), "diffuse_color", False),
    "scaling": Field("scaling", (2,), "scale", False),
    "rotation": Field("rotation", (4,), "rot", False),
}


def _attribute_names(name: str, f: Field, shape: tuple) -> list[str]:
    n = int(np.prod(shape))
    if name == "xyz":
        return ["x", "y", "z"]
    if name == "normal1":
        return ["nx", "ny", "nz"]
    if name == "normal2":
        return ["nx2", "ny2", "nz2"]
    if not f.numbered:
        return [f.ply]
    return [f"{f.ply}_{i}" for i in range(n)]


def construct_list_of_attributes(shapes: dict | None = None) -> list[str]:
    """gaussian_model.py:462-487 — the .ply property names in file order."""
    out = []
    for name, f in FIELDS.items():
        out += _attribute_names(name, f, (shapes or {}).get(name, f.shape))
    return out


def _to_numpy(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float32)


def save_ply(path, fields: dict) -> None:
    """gaussian_model.py:489-522: one `vertex` element of float32 properties, binary little endian, in the
    reference's attribute order (the header is what plyfile's PlyData([el]).write produces)."""
    P = _to_numpy(fields["xyz"]).shape[0]
    cols, shapes = [], {}
    for name, f in FIELDS.items():
        a = _to_numpy(fields[name])
        if a.shape[0] != P:
            raise ValueError(f"{name}: {a.shape[0]} rows, expected {P}")
        shapes[name] = a.shape[1:]
        if f.transposed:
            a = np.transpose(a, (0, 2, 1))
        cols.append(a.reshape(P, int(np.prod(a.shape[1:]))))   # also for an empty cloud (P = 0)
    names = construct_list_of_attributes(shapes)
    table = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype="<f4")
    if table.shape[1] != len(names):
        raise ValueError("attribute list and data disagree")
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {P}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(("\n".join(header) + "\n").encode("ascii"))
        fh.write(table.tobytes())


_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def read_ply_vertices(path) -> dict:
    """{property name: column} of the first element of a binary-little-endian or ascii .ply."""
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a ply file")
        fmt, elements = None, []          # elements: [name, count, [(property, dtype)]]
        while True:
            line = fh.readline()
            if not line:
                raise ValueError(f"{path}: header without end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append([tok[1], int(tok[2]), []])
            elif tok[0] == "property":
                if tok[1] == "list":
                    if len(elements) == 1:
                        raise ValueError(f"{path}: list properties are not supported in the vertex element")
                    continue
                elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if not elements:
            raise ValueError(f"{path}: no element")
        _, count, props = elements[0]     # later elements (if any) follow the vertex table and are not read
        dtype = np.dtype(props)
        if fmt == "binary_little_endian":
            data = np.frombuffer(fh.read(count * dtype.itemsize), dtype=dtype, count=count)
        elif fmt == "ascii":
            data = np.loadtxt(fh, dtype=np.float64, max_rows=count, ndmin=2)
            data = np.rec.fromarrays([data[:, i].astype(t) for i, (_, t) in enumerate(props)], names=[n for n, _ in props])
        else:
            raise ValueError(f"{path}: unsupported ply format {fmt}")
    return {n: np.asarray(data[n]) for n, _ in props}


def load_ply(path, max_sh_degree: int = 3) -> dict:
    """gaussian_model.py:725-836: the parameter arrays (float32, the reference's parameter shapes) of a saved cloud.
    Attributes are looked up BY NAME, numbered families sorted by their index, as the reference does."""
    v = read_ply_vertices(path)

    def family(prefix):
        names = sorted((n for n in v if n.startswith(prefix + "_") and n[len(prefix) + 1:].isdigit()),
                       key=lambda n: int(n.split("_")[-1]))
        return np.stack([v[n] for n in names], axis=1).astype(np.float32) if names else np.zeros((len(v["x"]), 0), np.float32)

    P = len(v["x"])
    n_rest = (max_sh_degree + 1) ** 2 - 1
    out = {
        "xyz": np.stack((v["x"], v["y"], v["z"]), axis=1),
        "normal1": np.stack((v["nx"], v["ny"], v["nz"]), axis=1),
        "normal2": np.stack((v["nx2"], v["ny2"], v["nz2"]), axis=1),
        "opacity": v["opacity"][:, None], "refl_strength": v["refl_strength"][:, None],
        "metalness": v["metalness"][:, None], "roughness": v["roughness"][:, None],
        "ori_color": family("ori_color"), "diffuse_color": family("diffuse_color"),
        "scaling": family("scale"), "rotation": family("rot"),
    }
    for name, prefix in (("features_rest", "f_rest"), ("indirect_rest", "ind_rest")):
        a = family(prefix)
        if a.shape[1] != 3 * n_rest:   # the reference asserts (:760, :775)
            raise ValueError(f"{prefix}: {a.shape[1]} attributes, expected {3 * n_rest} for SH degree {max_sh_degree}")
        out[name] = np.transpose(a.reshape(P, 3, n_rest), (0, 2, 1))
    out["features_dc"] = np.transpose(family("f_dc").reshape(P, 3, 1), (0, 2, 1))
    out["indirect_dc"] = np.transpose(family("ind_dc").reshape(P, 3, 1), (0, 2, 1))
    asg = family("ind_asg")
    out["indirect_asg"] = np.transpose(asg.reshape(P, 5, asg.shape[1] // 5), (0, 2, 1))   # (an empty cloud has no -1 to infer)
    return {k: np.ascontiguousarray(a, dtype=np.float32) for k, a in out.items()}


# ---- densification ---------------------------------------------------------------------------------------------
def inverse_sigmoid(x: torch.Tensor) -> torch.Tensor:
    return torch.log(x / (1 - x))   # utils/general_utils.py:18-19


def build_rotation(r: torch.Tensor) -> torch.Tensor:
    """utils/general_utils.py:78-99: rotation matrices of (w, x, y, z) quaternions, normalised first."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device, dtype=r.dtype)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - w * z)
    R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """utils/general_utils.py:29-63: log-linear interpolation from lr_init to lr_final, optionally eased in."""
    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        return delay_rate * np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
    return helper


# order of the parameter entries in the checkpoint tuple (gaussian_model.py:124-148), after active_sh_degree
_CAPTURE_ORDER = ("xyz", "refl_strength", "metalness", "roughness", "ori_color", "diffuse_color", "features_dc",
                  "features_rest", "indirect_dc", "indirect_rest", "indirect_asg", "scaling", "rotation", "opacity",
                  "normal1", "normal2")


class SurfelStore:
    """Parameters, Adam optimizer and densification statistics of a surfel cloud (GaussianModel's bookkeeping)."""
    max_sh_degree = 3
    # thresholds of the reflection-aware resets (gaussian_model.py:106-112; train_refnerf.py:1510 sets enlarge_scale)
    init_refl_value = 0.1
    enlarge_scale = 1.5
    refl_msk_thr = 0.02
    rough_msk_thr = 0.1

    def __init__(self, fields: dict, lrs: dict | None = None, percent_dense: float = 0.01, frozen=("normal1", "normal2"),
                 spatial_lr_scale: float = 1.0, active_sh_degree: int = 0, xyz_schedule: dict | None = None):
        self.spatial_lr_scale = spatial_lr_scale
        self.active_sh_degree = active_sh_degree
        self.xyz_scheduler_args = get_expon_lr_func(**xyz_schedule) if xyz_schedule else None
        dev = fields["xyz"].device if isinstance(fields["xyz"], torch.Tensor) else torch.device("cpu")
        self.params: dict[str, nn.Parameter] = {}
        for name in FIELDS:
            t = torch.as_tensor(fields[name], dtype=torch.float32, device=dev).clone()
            self.params[name] = nn.Parameter(t.requires_grad_(name not in frozen))
        self.percent_dense = percent_dense
        lrs = lrs or {}
        groups = [{"params": [self.params[n]], "lr": lrs.get(FIELDS[n].group, 0.0), "name": FIELDS[n].group}
                  for n in FIELDS]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)   # gaussian_model.py:448
        P = self.num_points
        self.xyz_gradient_accum = torch.zeros((P, 1), device=dev)
        self.denom = torch.zeros((P, 1), device=dev)
        self.max_radii2D = torch.zeros((P,), device=dev)

    # -- accessors (gaussian_model.py:236-285)
    @property
    def num_points(self) -> int:
        return self.params["xyz"].shape[0]

    @property
    def device(self):
        return self.params["xyz"].device

    def __getitem__(self, name: str) -> nn.Parameter:
        return self.params[name]

    @property
    def get_scaling(self):
        return torch.exp(self.params["scaling"])

    @property
    def get_opacity(self):
        return torch.sigmoid(self.params["opacity"])

    # -- statistics (gaussian_model.py:1059-1061, train_refnerf.py:1416-1418)
    def add_densification_stats(self, viewspace_grad: torch.Tensor, update_filter: torch.Tensor, radii: torch.Tensor | None = None):
        self.xyz_gradient_accum[update_filter] += torch.norm(viewspace_grad[update_filter], dim=-1, keepdim=True)
        self.denom[update_filter] += 1
        if radii is not None:
            self.max_radii2D[update_filter] = torch.max(self.max_radii2D[update_filter], radii[update_filter].float())

    def load_reduced_stats(self, stats: torch.Tensor, max_radii: torch.Tensor):
        """Takes the all-reduced tail of a GradArena: stats [P,2] = (sum of per-view norms, view count), max radii [P]."""
        self.xyz_gradient_accum += stats[:, 0:1]
        self.denom += stats[:, 1:2]
        self.max_radii2D = torch.max(self.max_radii2D, max_radii.to(self.max_radii2D.dtype))

    # -- optimizer surgery (gaussian_model.py:839-925)
    def _group(self, name: str):
        g = FIELDS[name].group
        return next(gr for gr in self.optimizer.param_groups if gr["name"] == g)

    def _replace(self, name: str, new: torch.Tensor, state_fn):
        group = self._group(name)
        old = group["params"][0]
        stored = self.optimizer.state.get(old, None)
        p = nn.Parameter(new.requires_grad_(True))   # as the reference: every rebuilt parameter requires grad (:868)
        if stored is not None and "exp_avg" in stored:
            stored["exp_avg"] = state_fn(stored["exp_avg"])
            stored["exp_avg_sq"] = state_fn(stored["exp_avg_sq"])
            del self.optimizer.state[old]
            self.optimizer.state[p] = stored
        group["params"][0] = p
        self.params[name] = p

    def replace_tensor_to_optimizer(self, name: str, tensor: torch.Tensor):
        """:839-854 — new values, moments zeroed (a group without state is left alone, as in the reference)."""
        if self.optimizer.state.get(self._group(name)["params"][0], None) is None:
            return False
        self._replace(name, tensor, lambda m: torch.zeros_like(tensor))
        return True

    def prune_points(self, mask: torch.Tensor):
        """:875-902 — drops the surfels where `mask` is True."""
        keep = ~mask
        for name in FIELDS:
            self._replace(name, self.params[name].detach()[keep], lambda m: m[keep])
        self.xyz_gradient_accum = self.xyz_gradient_accum[keep]
        self.denom = self.denom[keep]
        self.max_radii2D = self.max_radii2D[keep]

    def densification_postfix(self, new: dict):
        """:904-981 — appends surfels (zero moments) and resets the statistics."""
        for name in FIELDS:
            ext = new[name]
            self._replace(name, torch.cat((self.params[name].detach(), ext), dim=0),
                          lambda m, ext=ext: torch.cat((m, torch.zeros_like(ext)), dim=0))
        P, dev = self.num_points, self.device
        self.xyz_gradient_accum = torch.zeros((P, 1), device=dev)
        self.denom = torch.zeros((P, 1), device=dev)
        self.max_radii2D = torch.zeros((P,), device=dev)

    # -- clone / split / prune (:983-1057)
    def densify_and_clone(self, grads, grad_threshold, scene_extent):
        sel = torch.norm(grads, dim=-1) >= grad_threshold
        sel = torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values <= self.percent_dense * scene_extent)
        self.densification_postfix({n: self.params[n].detach()[sel] for n in FIELDS})

    def densify_and_split(self, grads, grad_threshold, scene_extent, N: int = 2, generator=None):
        n_init = self.num_points
        padded = torch.zeros((n_init,), device=self.device)
        padded[:grads.shape[0]] = grads.reshape(-1)
        sel = padded >= grad_threshold
        sel = torch.logical_and(sel, torch.max(self.get_scaling, dim=1).values > self.percent_dense * scene_extent)
        stds = self.get_scaling[sel].repeat(N, 1)
        stds = torch.cat([stds, 0 * torch.ones_like(stds[:, :1])], dim=-1)
        if generator is not None and generator.device != stds.device:
            # a host generator with device-resident parameters: the draw happens where the generator lives, so every rank
            # (and a CPU run of the same step) sees the same offsets whatever device holds its surfels
            samples = torch.normal(mean=torch.zeros(stds.shape, device=generator.device), std=stds.to(generator.device),
                                   generator=generator).to(stds.device)
        else:
            samples = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=generator)
        rots = build_rotation(self.params["rotation"].detach()[sel]).repeat(N, 1, 1)
        new = {n: self.params[n].detach()[sel].repeat(N, *([1] * len(FIELDS[n].shape))) for n in FIELDS}
        new["xyz"] = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self.params["xyz"].detach()[sel].repeat(N, 1)
        new["scaling"] = torch.log(self.get_scaling[sel].repeat(N, 1) / (0.8 * N))
        self.densification_postfix(new)
        prune = torch.cat((sel, torch.zeros(N * int(sel.sum()), device=self.device, dtype=torch.bool)))
        self.prune_points(prune)

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, generator=None):
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent, generator=generator)
        prune = (self.get_opacity < min_opacity).squeeze(-1)   # the reference's bare squeeze() breaks at one surfel
        if max_screen_size:
            big_vs = self.max_radii2D > max_screen_size
            big_ws = self.get_scaling.max(dim=1).values > 0.1 * extent
            prune = torch.logical_or(torch.logical_or(prune, big_vs), big_ws)
        self.prune_points(prune)

    def reset_opacity0(self):
        """:531-534."""
        new = inverse_sigmoid(torch.min(self.get_opacity, torch.ones_like(self.get_opacity) * 0.01)).detach()
        self.replace_tensor_to_optimizer("opacity", new)

    # -- the resets the training loop applies after an opacity reset (train_refnerf.py:1439-1455)
    @property
    def get_refl(self):
        return torch.sigmoid(self.params["refl_strength"])

    @property
    def get_rough(self):
        return torch.sigmoid(self.params["roughness"])

    def reset_opacity1(self, exclusive_msk=None):
        """:536-545 — opacities above 0.9 (or excluded) keep their value, the rest restart at 0.9."""
        RESET_V = 0.9
        opacity_old = self.get_opacity
        o_msk = (opacity_old > RESET_V).flatten()
        if exclusive_msk is not None:
            o_msk = torch.logical_or(o_msk, exclusive_msk)
        new = torch.ones_like(opacity_old) * inverse_sigmoid(torch.tensor([RESET_V], device=self.device))
        new[o_msk] = self.params["opacity"].detach()[o_msk]
        self.replace_tensor_to_optimizer("opacity", new.detach())

    def reset_refl(self, exclusive_msk=None, rst_value=None):
        """:558-565 — the reflection strength is raised to at least rst_value (excluded surfels keep theirs)."""
        rst_value = self.init_refl_value if rst_value is None else rst_value
        new = inverse_sigmoid(torch.max(self.get_refl, torch.ones_like(self.get_refl) * rst_value)).detach()
        if exclusive_msk is not None:
            new[exclusive_msk] = self.params["refl_strength"].detach()[exclusive_msk]
        self.replace_tensor_to_optimizer("refl_strength", new)

    def dist_color(self, exclusive_msk=None):
        """:601-611 — the DC colour of NON-reflective surfels is perturbed by U(-0.4, 0.4) (global torch RNG)."""
        DIST_RANGE = 0.4
        refl_msk = self.get_refl.flatten() > self.refl_msk_thr
        if exclusive_msk is not None:
            refl_msk = torch.logical_or(refl_msk, exclusive_msk)
        dcc = self.params["features_dc"].detach().clone()
        dist = dcc + (torch.rand_like(dcc) * DIST_RANGE * 2 - DIST_RANGE)
        dist[refl_msk] = dcc[refl_msk]
        self.replace_tensor_to_optimizer("features_dc", dist)

    def enlarge_refl_scales(self, exclusive_msk=None):
        """:626-645 (ret_raw=True) — reflective, smooth surfels grow by enlarge_scale; the others keep their raw scale."""
        refl_msk = self.get_refl.flatten() < self.refl_msk_thr
        rough_msk = self.get_rough.flatten() > self.rough_msk_thr
        combined = torch.logical_or(refl_msk, rough_msk)
        if exclusive_msk is not None:
            combined = torch.logical_or(combined, exclusive_msk)
        scales = self.get_scaling
        new = torch.log(scales * (torch.ones_like(scales) * self.enlarge_scale)).detach()
        new[combined] = self.params["scaling"].detach()[combined]
        return new

    def reset_scale(self, exclusive_msk=None):
        """:667-671."""
        self.replace_tensor_to_optimizer("scaling", self.enlarge_refl_scales(exclusive_msk=exclusive_msk))

    # -- schedule (gaussian_model.py:455-460, :311-313)
    def update_learning_rate(self, iteration):
        if self.xyz_scheduler_args is None:
            return None
        for group in self.optimizer.param_groups:
            if group["name"] == "xyz":
                group["lr"] = self.xyz_scheduler_args(iteration)
                return group["lr"]

    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    # -- checkpoint tuple (gaussian_model.py:124-172); train_refnerf.py saves torch.save((capture(), iteration), path)
    def reference_optimizer_state(self, env_groups: dict | None = None) -> dict:
        """An Adam state dict laid out like the reference's optimizer (training_setup, gaussian_model.py:422-447): its 18
        groups in ITS order — the reference's restore() loads the dict positionally (:170) — including the two
        environment-map groups this store does not own. Those come from `env_groups` ({"env": (group_dict, [state per
        parameter]), "env2": ...}; group_dict = the param-group hyper-parameters without "params") or, failing that, from
        what restore() read out of a reference checkpoint. Without them the dict would mis-assign every moment, so that
        case raises instead of writing a checkpoint the reference silently loads wrongly."""
        own = {g["name"]: g for g in self.optimizer.param_groups}
        foreign = dict(getattr(self, "_foreign_groups", {}))
        foreign.update(env_groups or {})
        groups, state, idx = [], {}, 0
        for name in REFERENCE_GROUP_ORDER:
            if name in own:
                g = own[name]
                hyper = {k: v for k, v in g.items() if k != "params"}
                states = [self.optimizer.state.get(p) for p in g["params"]]
            elif name in foreign:
                hyper, states = foreign[name]
                hyper = dict(hyper)
            else:
                raise ValueError(f"capture(): optimizer group {name!r} is neither owned by the store nor supplied through "
                                 "env_groups; the reference loads optimizer state positionally and needs all 18 groups")
            hyper["params"] = list(range(idx, idx + len(states)))
            for k, st in enumerate(states):
                if st:
                    state[idx + k] = {a: (b.detach().clone() if isinstance(b, torch.Tensor) else b) for a, b in st.items()}
            idx += len(states)
            groups.append(hyper)
        return {"state": state, "param_groups": groups}

    def capture(self, env_groups: dict | None = None):
        """The reference's checkpoint tuple (gaussian_model.py:124-148) with the optimizer state in the reference's group
        layout (see reference_optimizer_state): `torch.save((store.capture(), iteration), path)` is loadable by the
        reference's restore()."""
        return (self.active_sh_degree, *[self.params[n] for n in _CAPTURE_ORDER], self.max_radii2D,
                self.xyz_gradient_accum, self.denom, self.reference_optimizer_state(env_groups), self.spatial_lr_scale)

    @classmethod
    def restore(cls, model_args, lrs: dict | None = None, device=None, **kw):
        """A store from the reference's checkpoint tuple. The optimizer state is loaded as saved; like the reference
        (:169) the anisotropic-Gaussian parameters are re-created as zeros."""
        if len(model_args) != len(_CAPTURE_ORDER) + 6:
            raise ValueError(f"checkpoint tuple has {len(model_args)} entries, expected {len(_CAPTURE_ORDER) + 6}")
        active_sh_degree, *rest = model_args
        tensors = dict(zip(_CAPTURE_ORDER, rest[:len(_CAPTURE_ORDER)]))
        max_radii2D, xyz_gradient_accum, denom, opt_dict, spatial_lr_scale = rest[len(_CAPTURE_ORDER):]
        device = device if device is not None else tensors["xyz"].device
        fields = {n: t.detach().to(device) for n, t in tensors.items()}
        fields["indirect_asg"] = torch.zeros((fields["rotation"].shape[0], 32, 5), device=device)
        st = cls(fields, lrs=lrs, spatial_lr_scale=spatial_lr_scale, active_sh_degree=active_sh_degree, **kw)
        st.max_radii2D = max_radii2D.to(device)
        st.xyz_gradient_accum = xyz_gradient_accum.to(device)
        st.denom = denom.to(device)
        # the reference's optimizer also holds the two environment maps (groups "env", "env2"): take the per-surfel
        # groups by NAME, whatever their position in the saved state dict
        by_name = {g["name"]: g for g in opt_dict["param_groups"]}
        own_names = {g["name"] for g in st.optimizer.param_groups}
        st._foreign_groups = {      # kept so that capture() can write them back where the reference expects them
            n: ({k: v for k, v in g.items() if k != "params"}, [opt_dict["state"].get(i) for i in g["params"]])
            for n, g in by_name.items() if n not in own_names}
        for group in st.optimizer.param_groups:
            saved = by_name.get(group["name"])
            if saved is None:
                continue
            group["lr"] = saved["lr"]
            state = opt_dict["state"].get(saved["params"][0])
            if state is not None:
                st.optimizer.state[group["params"][0]] = {
                    k: (v.detach().clone().to(device) if isinstance(v, torch.Tensor) else v) for k, v in state.items()}
        return st

    # -- formats
    def save_ply(self, path):
        save_ply(path, {n: p for n, p in self.params.items()})

    @classmethod
    def from_ply(cls, path, device="cpu", max_sh_degree: int = 3, **kw):
        f = load_ply(path, max_sh_degree)
        return cls({k: torch.from_numpy(v).to(device) for k, v in f.items()}, **kw)

    def checksum(self) -> torch.Tensor:
        """Order-sensitive digest of every parameter and moment (ranks compare it after densifying)."""
        acc = torch.zeros(2, dtype=torch.float64, device=self.device)
        for i, name in enumerate(FIELDS):
            p = self.params[name].detach().double().reshape(-1)
            w = torch.arange(1, p.numel() + 1, device=self.device, dtype=torch.float64)
            acc[0] += (p * torch.sin(w * (i + 1))).sum()
            acc[1] += float(p.numel())
        return acc

"""Generates tests/golden/features_*.npz by running the REFERENCE's own Python functions
(/root/reference/utils/sh_utils.py, utils/general_utils.py) composed as render_surfel composes them
(gaussian_renderer/__init__.py:259-266, :334-353; scene/gaussian_model.py:48-54, :236-303). Run in the build
container (where /root/reference is mounted); the vectors travel, the reference does not.

build_scaling_rotation allocates with device='cuda'; torch.zeros is wrapped for the duration of the call so
the unmodified function runs on the CPU."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, str(ROOT))
from utils import general_utils as gu   # noqa: E402  (the reference's)
from utils import sh_utils as su        # noqa: E402
from oracle import features_oracle as fo  # noqa: E402  (only for the synthetic inputs)


class _ZerosOnCpu:
    def __enter__(self):
        self.orig = torch.zeros

        def zeros(*a, **k):
            k.pop("device", None)
            return self.orig(*a, **k)
        torch.zeros = zeros

    def __exit__(self, *exc):
        torch.zeros = self.orig


def reference_chain(p, campos):
    scales = torch.exp(p["scaling"])
    rotations = torch.nn.functional.normalize(p["rotation"])
    opacities = torch.sigmoid(p["opacity"])
    dir_pp = p["xyz"] - campos
    dirn = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    with _ZerosOnCpu():
        RS = gu.build_scaling_rotation(torch.cat([scales * 1.0, torch.ones_like(scales)], dim=-1)[:, :3],
                                       p["rotation"]).permute(0, 2, 1)
    normals_raw, _ = gu.flip_align_view(RS[:, 2, :3], dirn)
    normals = gu.safe_normalize(normals_raw)
    w_o = -dirn
    reflection = 2 * torch.sum(normals * w_o, dim=1, keepdim=True) * normals - w_o
    get_indirect = torch.cat((p["indirect_dc"].reshape(-1, 1, 3), p["indirect_rest"].reshape(-1, 15, 3)), dim=1)
    shs_indirect = get_indirect.transpose(1, 2).view(-1, 3, 16)
    indirect = torch.clamp_min(su.eval_sh(3, shs_indirect, reflection), 0.0)
    features = torch.cat((torch.sigmoid(p["refl_strength"]), torch.sigmoid(p["roughness"]),
                          torch.sigmoid(p["ori_color"]), indirect), dim=-1)
    return scales, rotations, opacities, features


def main():
    for P, seed in ((257, 1), (64, 2)):
        p, campos = fo.synthetic_params(P, seed)
        p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        outs = reference_chain(p, campos)
        g = torch.Generator().manual_seed(100 + seed)
        ups = [torch.randn(o.shape, generator=g) for o in outs]
        sum((o * u).sum() for o, u in zip(outs, ups)).backward()
        z = {"campos": campos.numpy()}
        for k, v in p.items():
            z["in_" + k] = v.detach().numpy()
            z["grad_" + k] = v.grad.numpy()
        for n, o, u in zip(("scales", "rotations", "opacities", "features"), outs, ups):
            z["out_" + n] = o.detach().numpy()
            z["up_" + n] = u.numpy()
        path = ROOT / "tests" / "golden" / f"features_P{P}_s{seed}.npz"
        np.savez_compressed(path, **z)
        print("wrote", path, {k: v.shape for k, v in z.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()

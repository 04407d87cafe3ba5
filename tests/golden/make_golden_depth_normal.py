"""Generates tests/golden/depth_normal_*.npz with the REFERENCE's own utils/point_utils.py depths_to_points /
depth_to_normal (imported from /root/reference; run in the build container, on the CPU). The two functions hard-code the
CUDA device (`.cuda()`, `device='cuda'`): for the run, Tensor.cuda is the identity and torch.arange drops its device
argument; matplotlib (imported, unused) is an empty stub. No reference source is edited or copied."""
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, str(ROOT))
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
torch.Tensor.cuda = lambda self, *a, **k: self
_arange = torch.arange
torch.arange = lambda *a, **kw: _arange(*a, **{k: v for k, v in kw.items() if k != "device"})

from utils import point_utils as ref  # noqa: E402  (the reference's)
from materialrefgs_b200 import synthetic  # noqa: E402


def main():
    for name, (view, W, H, radius, seed) in {"a": (1, 96, 64, 4.0, 1), "b": (6, 75, 113, 3.0, 2)}.items():
        cam = synthetic.orbit_camera(view, 8, W, H, radius=radius)
        g = torch.Generator().manual_seed(seed)
        depth = 3.0 + F.interpolate(torch.randn(1, 1, H // 8 + 2, W // 8 + 2, generator=g), size=(H, W), mode="bilinear",
                                    align_corners=False)[0] * 0.4 + 0.01 * torch.randn(1, H, W, generator=g)
        depth = depth.clone().requires_grad_(True)
        points = ref.depths_to_points(cam, depth)
        normal = ref.depth_to_normal(cam, depth)
        w = torch.randn(H, W, 3, generator=g)
        (normal * w).sum().backward()
        np.savez_compressed(ROOT / "tests" / "golden" / f"depth_normal_{name}.npz", view=np.array([view, W, H, radius]),
                            depth=depth.detach().numpy(), points=points.detach().numpy(), normal=normal.detach().numpy(),
                            w=w.numpy(), grad_depth=depth.grad.numpy())
        print("wrote depth_normal_%s.npz" % name, tuple(normal.shape))


if __name__ == "__main__":
    main()

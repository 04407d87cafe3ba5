"""Generates tests/golden/linear_to_srgb.npz with the REFERENCE's own utils/graphics_utils.py linear_to_srgb (imported from
/root/reference; pure torch). Pins the product's host-side materialrefgs_b200.shading.linear_to_srgb (render_initial /
render_volume) and the oracle's."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
from utils.graphics_utils import linear_to_srgb  # noqa: E402  (the reference's)


def main():
    x = torch.cat([torch.linspace(-0.2, 1.5, 1001), torch.tensor([0.0, 0.0031308, 0.00313081, 1e-9, 1.0])])
    x = x.clone().requires_grad_(True)
    y = linear_to_srgb(x)
    y.sum().backward()
    np.savez(ROOT / "tests" / "golden" / "linear_to_srgb.npz", x=x.detach().numpy(), y=y.detach().numpy(), dy=x.grad.numpy())
    print("wrote linear_to_srgb.npz")


if __name__ == "__main__":
    main()

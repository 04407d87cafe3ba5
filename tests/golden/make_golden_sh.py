"""Generates tests/golden/eval_sh.npz with the REFERENCE's own utils/sh_utils.py eval_sh at degrees 0..3 (imported from
/root/reference; pure torch, runs anywhere). Pins materialrefgs_b200/render.py's eval_sh / sh_basis, which the
duck-typed-model paths of the render functions use for the indirect light."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
from utils.sh_utils import eval_sh  # noqa: E402  (the reference's)


def main():
    g = torch.Generator().manual_seed(17)
    sh = torch.randn(257, 3, 16, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(257, 3, generator=g), dim=-1)
    out = {"sh": sh.numpy(), "dirs": dirs.numpy()}
    for deg in range(4):
        out[f"deg{deg}"] = eval_sh(deg, sh, dirs).numpy()
    np.savez_compressed(ROOT / "tests" / "golden" / "eval_sh.npz", **out)
    print("wrote eval_sh.npz")


if __name__ == "__main__":
    main()

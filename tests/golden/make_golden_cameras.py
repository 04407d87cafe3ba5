"""Generates tests/golden/camera_matrices.npz with the REFERENCE's own utils/graphics_utils.py (getWorld2View2,
getProjectionMatrix, getProjectionMatrixCorrect, fov2focal, focal2fov), composed as scene/cameras.py:70-84 composes them
(row-vector convention: transposed matrices, full_proj = W2V^T @ P^T, camera centre = inverse(W2V^T)[3,:3]); run in the
build container. Pins the synthetic cameras every test, the bench and the tools feed to both rasterizers."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, str(ROOT))
from utils import graphics_utils as gu  # noqa: E402  (the reference's)
from materialrefgs_b200 import synthetic  # noqa: E402  (only for R, T, fov of the test views)


def main():
    out = {}
    views = [(0, 8, 800, 800, 4.0), (3, 8, 1920, 1080, 3.0), (5, 7, 333, 517, 2.5)]
    for k, (i, n, W, H, radius) in enumerate(views):
        cam = synthetic.orbit_camera(i, n, W, H, radius=radius)
        w2v = torch.tensor(gu.getWorld2View2(cam.R, cam.T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
        proj = gu.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=cam.FoVx, fovY=cam.FoVy).transpose(0, 1)
        full = (w2v.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
        projK = gu.getProjectionMatrixCorrect(0.01, 100.0, H, W, np.asarray(cam.HWK[2], np.float64)).transpose(0, 1)
        out.update({f"view{k}": np.array([i, n, W, H, radius]), f"w2v{k}": w2v.numpy(), f"proj{k}": proj.numpy(),
                    f"full{k}": full.numpy(), f"center{k}": w2v.inverse()[3, :3].numpy(), f"projK{k}": projK.numpy(),
                    f"focal{k}": np.array([gu.fov2focal(cam.FoVx, W), gu.fov2focal(cam.FoVy, H)])})
    np.savez(ROOT / "tests" / "golden" / "camera_matrices.npz", **out)
    print("wrote camera_matrices.npz")


if __name__ == "__main__":
    main()

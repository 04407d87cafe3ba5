"""Generates tests/golden/cube_dirs.npz with the REFERENCE's own cube addressing — scene/light_utils.py:24-31
(cube_to_dir) and the texel-centre grid + safe_normalize its cubemap_mip backward builds (:72-79) — imported from
/root/reference and run on the CPU of the build container (nvdiffrast, absent, is stubbed: these functions do not use
it). The vectors pin the face / u / v conventions every cube fetch of the product must agree with: fetching at a
texel-centre direction must return exactly that texel (tests/test_texture_properties_gpu.py)."""
import sys
import types
from pathlib import Path

import numpy as np
import torch

sys.modules.setdefault("nvdiffrast", types.ModuleType("nvdiffrast"))
sys.modules.setdefault("nvdiffrast.torch", types.ModuleType("nvdiffrast.torch"))
sys.path.insert(0, "/root/reference/scene")
import light_utils as lu  # noqa: E402

out = {}
for res in (4, 16):
    dirs = []
    for s in range(6):
        gy, gx = torch.meshgrid(torch.linspace(-1.0 + 1.0 / res, 1.0 - 1.0 / res, res),
                                torch.linspace(-1.0 + 1.0 / res, 1.0 - 1.0 / res, res), indexing="ij")
        dirs.append(lu.safe_normalize(lu.cube_to_dir(s, gx, gy)))
    out[f"dirs_{res}"] = torch.stack(dirs).numpy()      # [6, res, res, 3]: direction of texel (face, y, x)
np.savez_compressed(Path(__file__).resolve().parent / "cube_dirs.npz", **out)
print({k: v.shape for k, v in out.items()})

"""Generates tests/golden/geomloss_*.npz with the REFERENCE's own utils/loss_utils.py calculate_loss,
first_order_edge_aware_loss and get_img_grad_weight (imported from /root/reference; run in the build container).

kornia (requirements.txt:45, kornia==0.7.3) and matplotlib are not installed: matplotlib is an empty stub (unused on
this path); `kornia.filters.spatial_gradient` is the ONE symbol the path needs and is supplied by the restatement in
oracle/losses_oracle.py — so these vectors pin everything the reference itself does around it (term selection,
lambdas, abs/exp/sum/mean, the image-gradient weight), not kornia's Sobel filter (see the oracle's header)."""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, str(ROOT))
from oracle import losses_oracle as lo  # noqa: E402


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return lambda *a, **kw: None


for name in ("kornia", "matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = _Stub(name)
kf = types.ModuleType("kornia.filters")
kf.spatial_gradient = lambda x, order=1: lo.spatial_gradient(x)
sys.modules["kornia.filters"] = kf
from utils import loss_utils as ref  # noqa: E402  (the reference's)


class _Image:
    """camera.original_image: the reference calls .cuda() on it (loss_utils.py:153); there is no GPU here."""
    def __init__(self, t):
        self.t = t

    def cuda(self):
        return self.t


CASES = {
    # name: (H, W, seed, iteration, weighted, opt overrides)
    "all_terms_weighted": (61, 83, 11, 5000, True, dict(lambda_dist=100.0, lambda_normal_smooth=0.01, lambda_depth_smooth=0.02)),
    "cosine_normal_only": (48, 40, 12, 100, False, dict()),
    "dist_and_depth": (35, 97, 13, 20000, True, dict(lambda_dist=10.0, lambda_normal_render_depth=0.0, lambda_depth_smooth=0.05,
                                                   lambda_normal_smooth=0.01)),
}


def make_opt(**kw):
    o = types.SimpleNamespace(lambda_dssim=0.2, lambda_dist=0.0, lambda_normal_render_depth=0.05, lambda_normal_smooth=0.0,
                              lambda_depth_smooth=0.0, normal_loss_start=0, dist_loss_start=3000, normal_smooth_from_iter=0,
                              normal_smooth_until_iter=18000, use_perceptual_loss=False, perceptual_loss_start_iter=18000,
                              lambda_perceptual_loss=0.0)
    o.__dict__.update(kw)
    return o


def main():
    for name, (H, W, seed, iteration, weighted, over) in CASES.items():
        pkg, gt = lo.synthetic_render_pkg(H, W, seed)
        leaves = {k: pkg[k].clone().requires_grad_(True) for k in ("render", "rend_normal", "surf_normal", "surf_depth", "rend_dist")}
        pkg_l = dict(pkg, **leaves)
        opt = make_opt(**over)
        iw = None
        if weighted:   # train_refnerf.py:1178-1179
            iw = (1.0 - ref.get_img_grad_weight(gt)).clamp(0, 1).detach() ** 2
        cam = types.SimpleNamespace(original_image=_Image(gt))
        pc = types.SimpleNamespace(get_xyz=torch.zeros(4, 3))
        loss, tb = ref.calculate_loss(cam, pc, pkg_l, opt, iteration, iw, None)
        loss.backward()
        out = dict(H=H, W=W, iteration=iteration, weighted=weighted, loss=loss.item(), gt=gt.numpy(),
                   grad_weight=ref.get_img_grad_weight(gt).numpy(),
                   edge_normal=ref.first_order_edge_aware_loss(pkg["rend_normal"], gt).item(),
                   edge_depth=ref.first_order_edge_aware_loss(pkg["surf_depth"], gt).item(),
                   opt_keys=np.array(sorted(over)), opt_vals=np.array([over[k] for k in sorted(over)], dtype=np.float64))
        for k, v in leaves.items():
            out[k] = v.detach().numpy()
            out["grad_" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
        path = ROOT / "tests" / "golden" / f"geomloss_{name}.npz"
        np.savez_compressed(path, **out)
        print("wrote", path, loss.item(), {k: float(v) for k, v in tb.items() if k.startswith("loss_")})


if __name__ == "__main__":
    main()

"""Fused L1+SSIM (mrgs_photometric_*) vs the reference formulation in eager torch on the GPU. Dev tool."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from materialrefgs_b200 import losses  # noqa: E402
from oracle import losses_oracle as lo  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for H, W in ((800, 800), (1080, 1920)):
    img, gt = lo.synthetic_pair(3, H, W, seed=1)
    img, gt = img.to(dev).requires_grad_(True), gt.to(dev)

    def run(f):
        img.grad = None
        f(img, gt, 0.2).backward()
    t_f = timeit(lambda: run(losses.photometric_loss))
    t_e = timeit(lambda: run(lo.photometric_loss))
    print(f"{W}x{H}: fused fwd+bwd {t_f:.3f} ms, eager torch (cuDNN grouped convs) {t_e:.3f} ms, x{t_e / t_f:.1f}")

# the whole calculate_loss (photometric + the four geometric regularisers), fused vs the eager restatement
import types  # noqa: E402

opt = types.SimpleNamespace(lambda_dssim=0.2, lambda_dist=100.0, lambda_normal_render_depth=0.05, lambda_normal_smooth=0.01,
                            lambda_depth_smooth=0.02, normal_loss_start=0, dist_loss_start=3000, normal_smooth_from_iter=0,
                            normal_smooth_until_iter=18000, use_perceptual_loss=False, perceptual_loss_start_iter=18000)
LEAVES = ("render", "rend_normal", "surf_normal", "surf_depth", "rend_dist")
for H, W in ((800, 800), (1080, 1920)):
    pkg, gt = lo.synthetic_render_pkg(H, W, seed=2)
    gt = gt.to(dev)
    leaves = {k: pkg[k].to(dev).requires_grad_(True) for k in LEAVES}
    iw = (1.0 - losses.get_img_grad_weight(gt)).clamp(0, 1) ** 2

    def fused():
        for v in leaves.values():
            v.grad = None
        t = losses.geometry_losses(leaves["rend_normal"], leaves["surf_normal"], leaves["rend_dist"], leaves["surf_depth"], gt,
                                   iw, normal=True, dist=True, normal_smooth=True, depth_smooth=True)
        (losses.photometric_loss(leaves["render"], gt, 0.2) + 0.05 * t[0] + 100.0 * t[1] + 0.01 * t[2] + 0.02 * t[3]).backward()

    def eager():
        for v in leaves.values():
            v.grad = None
        lo.calculate_loss(gt, leaves, opt, 5000, iw).backward()
    t_f, t_e = timeit(fused), timeit(eager)
    print(f"{W}x{H}: calculate_loss fwd+bwd fused {t_f:.3f} ms, eager torch {t_e:.3f} ms, x{t_e / t_f:.1f}; "
          f"get_img_grad_weight fused {timeit(lambda: losses.get_img_grad_weight(gt)):.3f} ms, "
          f"eager {timeit(lambda: lo.get_img_grad_weight(gt)):.3f} ms")

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

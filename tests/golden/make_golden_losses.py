"""Generates tests/golden/losses_*.npz with the REFERENCE's own utils/loss_utils.py l1_loss / ssim (imported from
/root/reference; run in the build container). That module also imports kornia and matplotlib, which are not
installed and which l1_loss / ssim never touch: they are replaced by empty stub modules for the import."""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, str(ROOT))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return lambda *a, **kw: None


for name in ("kornia", "kornia.filters", "matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = _Stub(name)
from utils import loss_utils as ref  # noqa: E402  (the reference's)
from oracle import losses_oracle as lo  # noqa: E402  (only for the synthetic inputs)


def main():
    for (C, H, W), seed in (((3, 70, 45), 1), ((3, 128, 96), 2), ((1, 33, 64), 3)):
        img, gt = lo.synthetic_pair(C, H, W, seed)
        img = img.clone().requires_grad_(True)
        l1 = ref.l1_loss(img, gt)
        s = ref.ssim(img, gt)
        g_l1, = torch.autograd.grad(l1, img, retain_graph=True)
        g_s, = torch.autograd.grad(s, img)
        path = ROOT / "tests" / "golden" / f"losses_c{C}_{H}x{W}.npz"
        np.savez_compressed(path, img=img.detach().numpy(), gt=gt.numpy(), l1=l1.item(), ssim=s.item(),
                            grad_l1=g_l1.numpy(), grad_ssim=g_s.numpy())
        print("wrote", path, l1.item(), s.item())


if __name__ == "__main__":
    main()

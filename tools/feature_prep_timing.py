"""Fused feature preparation (mrgs_surfel_features_*) vs the same chain in eager torch on the GPU. Dev tool."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from materialrefgs_b200.features import surfel_features  # noqa: E402
from oracle import features_oracle as fo  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=20):
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


for P in (1_000_000, 5_000_000):
    raw, campos = fo.synthetic_params(P, seed=3)
    raw = {k: v.to(dev).requires_grad_(True) for k, v in raw.items()}
    campos = campos.to(dev)
    ups = [torch.randn(P, w, device=dev) for w in (2, 4, 1, 8)]
    args = [raw[k] for k, _ in fo.RAW_FIELDS]

    def run(f):
        for t in args:
            t.grad = None
        outs = f()
        torch.autograd.backward(outs, ups)
    t_fused = timeit(lambda: run(lambda: surfel_features(campos, *args)))
    t_eager = timeit(lambda: run(lambda: fo.prepare_features(*args, campos)))
    bytes_fb = P * 4 * ((63 + 15) + (63 + 15 + 63))
    print(f"P={P}: fused fwd+bwd {t_fused:.3f} ms ({bytes_fb / t_fused / 1e6:.0f} GB/s algorithmic), eager torch {t_eager:.3f} ms, x{t_eager / t_fused:.1f}")

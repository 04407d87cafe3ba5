"""Differentiable EnvLight queries at per-surfel scale (render_volume's two queries): fused fwd+bwd time. Dev tool."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from materialrefgs_b200.shading import EnvLight  # noqa: E402

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


env = EnvLight(device=dev, max_res=512, min_res=16, trainable=True)
with torch.no_grad():
    env.base.normal_()
env.build_mips()
levels = [l.detach().requires_grad_(True) for l in env.specular]
env.set_chain(levels)
env.diffuse = env.diffuse.detach().requires_grad_(True)
for P in (1_000_000, 5_000_000):
    d = torch.nn.functional.normalize(torch.randn(P, 3, device=dev), dim=-1).requires_grad_(True)
    r = torch.rand(P, 1, device=dev).requires_grad_(True)
    w = torch.randn(P, 3, device=dev)

    def run(mode):
        out = env(d, mode=mode, roughness=None if mode else r)
        (out * w).sum().backward()
    print(f"P={P}: diffuse query fwd+bwd {timeit(lambda: run('diffuse')):.3f} ms, "
          f"specular (6-level chain) fwd+bwd {timeit(lambda: run(None)):.3f} ms")

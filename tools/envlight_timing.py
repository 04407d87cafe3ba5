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

# get_full_color_volume: the fused per-surfel kernel pair vs the same chain in eager torch (the oracle's restatement on the GPU)
from materialrefgs_b200 import synthetic  # noqa: E402
from materialrefgs_b200.shading import get_full_color_volume  # noqa: E402
from oracle import shading_oracle as so  # noqa: E402

cam = synthetic.orbit_camera(2, 8, 64, 64)
oracle_env = so.EnvLightOracle(levels, diffuse=env.diffuse)
lut = so.load_lut(dev)
for P in (1_000_000,):
    g = torch.Generator(device=dev).manual_seed(3)
    t = dict(xyz=1.3 * (2 * torch.rand(P, 3, device=dev, generator=g) - 1),
             n=torch.nn.functional.normalize(torch.randn(P, 3, device=dev, generator=g), dim=-1),
             albedo=torch.rand(P, 3, device=dev, generator=g), rs=torch.rand(P, 1, device=dev, generator=g),
             ro=torch.rand(P, 1, device=dev, generator=g))
    t = {k: v.requires_grad_(True) for k, v in t.items()}
    w = torch.randn(P, 3, device=dev)

    def fused():
        d, s = get_full_color_volume(env, t["xyz"], t["albedo"], cam.HWK, cam.R, cam.T, t["n"], None, refl_strength=t["rs"],
                                     roughness=t["ro"])
        ((d + s) * w).sum().backward()

    def eager():
        d, s = so.get_full_color_volume(oracle_env, lut, t["xyz"], t["albedo"], cam, t["n"], t["rs"], t["ro"])
        ((d + s) * w).sum().backward()
    print(f"P={P}: get_full_color_volume fwd+bwd fused {timeit(fused):.3f} ms, eager torch {timeit(eager, n=5):.3f} ms")

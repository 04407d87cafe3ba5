"""Rasterizer throughput at BASELINE.json's configurations C2..C5, ours next to the unmodified reference extension
(oracle/_ref) on the same GPU and inputs: forward+backward (training) and forward only under no_grad (evaluation).
CUDA events, 3 warm-ups, views cycled so that consecutive frames differ. Dev tool; prints one JSON object per line."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from materialrefgs_b200 import synthetic  # noqa: E402
from tests import refimpl  # noqa: E402

dev = torch.device("cuda:0")
CONFIGS = {
    "C2": dict(P=300_000, W=800, H=800, opacity="init", unbounded=False, radius=4.0),
    "C3": dict(P=1_000_000, W=800, H=800, opacity="trained", unbounded=False, radius=4.0),
    "C4": dict(P=3_000_000, W=1920, H=1080, opacity="trained", unbounded=True, radius=3.0),
    "C5": dict(P=5_000_000, W=800, H=800, opacity="trained", unbounded=False, radius=4.0),
}
S = 8


def run_config(name, cfg, mod, iters):
    cloud = synthetic.make_cloud(cfg["P"], S=S, opacity=cfg["opacity"], unbounded=cfg["unbounded"]).to(dev)
    cams = [synthetic.orbit_camera(i, 8, cfg["W"], cfg["H"], radius=cfg["radius"]).to(dev) for i in range(8)]
    gc, gf, go = (t.to(dev) for t in synthetic.upstream_grads(S, cfg["H"], cfg["W"]))
    bg = torch.zeros(3, device=dev)
    leaves = {k: getattr(cloud, k).clone().requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs", "features")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)

    def frame(i, train):
        cam = cams[i % len(cams)]
        rs = mod.GaussianRasterizationSettings(cfg["H"], cfg["W"], cam.tanfovx, cam.tanfovy, bg, 1.0, cam.world_view_transform,
                                               cam.full_proj_transform, 3, cam.camera_center, False, False)
        _, color, feat, radii, allmap = mod.GaussianRasterizer(rs)(
            means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
            features=leaves["features"], scales=leaves["scales"], rotations=leaves["rotations"])
        if train:
            for t in (*leaves.values(), m2d):
                t.grad = None
            torch.autograd.backward((color, feat, allmap), (gc, gf, go))

    out = {}
    for mode, train in (("train_fps", True), ("eval_fps", False)):
        with torch.set_grad_enabled(train):
            for i in range(3):
                frame(i, train)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(iters):
                frame(3 + i, train)
            b.record()
            torch.cuda.synchronize()
        out[mode] = round(iters * 1000.0 / a.elapsed_time(b), 1)
    return out


def main():
    import materialrefgs_b200.diff_surfel_rasterization as ours
    ref = refimpl.load_reference()
    only = sys.argv[1:] or list(CONFIGS)
    for name in only:
        cfg = CONFIGS[name]
        iters = 16 if cfg["P"] <= 1_000_000 else 8
        row = {"config": name, **{k: cfg[k] for k in ("P", "W", "H")}, "S": S, "ours": run_config(name, cfg, ours, iters)}
        if ref is not None:
            row["reference"] = run_config(name, cfg, ref, iters)
            row["speedup_train"] = round(row["ours"]["train_fps"] / row["reference"]["train_fps"], 2)
            row["speedup_eval"] = round(row["ours"]["eval_fps"] / row["reference"]["eval_fps"], 2)
        print(json.dumps(row), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

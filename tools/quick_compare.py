"""Ad-hoc timing of reference (oracle/_ref) vs libmrgs on one synthetic view. Dev tool only."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from materialrefgs_b200 import synthetic  # noqa: E402
from tests import refimpl  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1_000_000)
    ap.add_argument("--S", type=int, default=8)
    ap.add_argument("--W", type=int, default=800)
    ap.add_argument("--H", type=int, default=800)
    ap.add_argument("--opacity", default="trained")
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cloud = synthetic.make_cloud(a.P, S=a.S, opacity=a.opacity).to(dev)
    cam = synthetic.orbit_camera(1, 8, a.W, a.H).to(dev)
    gc, gf, go = (t.to(dev) for t in synthetic.upstream_grads(a.S, a.H, a.W))
    bg = torch.zeros(3, device=dev)
    e = torch.empty(0, device=dev)
    fargs = (bg, cloud.means3D, e, cloud.features, cloud.opacities, cloud.scales, cloud.rotations, 1.0, e,
             cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, a.H, a.W,
             cloud.shs, 3, cam.camera_center, False, False)

    import materialrefgs_b200.rasterizer as ours
    mods = [("mrgs", ours.rasterize_forward_raw, ours.rasterize_backward_raw)]
    if not a.no_ref:
        ref = refimpl.load_reference()
        if ref is not None:
            mods.append(("ref", ref._C.rasterize_gaussians, ref._C.rasterize_gaussians_backward))
    for name, fwd, bwd in mods:
        out = fwd(*fargs)
        R, contrib, color, feat, others, radii, geom, binning, img = out
        def b():
            return bwd(bg, cloud.means3D, radii, e, cloud.features, cloud.scales, cloud.rotations, 1.0, e,
                       cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, gc, gf, go,
                       cloud.shs, 3, cam.camera_center, geom, R, binning, img, contrib, False)
        tf = timeit(lambda: fwd(*fargs))
        tb = timeit(b)
        pv = int((radii > 0).sum())
        print(f"{name}: P={a.P} Pv={pv} R={R} fwd median {tf[0]:.3f} ms (min {tf[1]:.3f})  "
              f"bwd median {tb[0]:.3f} ms (min {tb[1]:.3f})", flush=True)


if __name__ == "__main__":
    main()

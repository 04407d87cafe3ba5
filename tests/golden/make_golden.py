"""Generates tests/golden/raster_*.npz from the UNMODIFIED reference CUDA extension.

Run on a GPU box where oracle/_ref has been built (oracle/build_ref.sh):
    python tests/golden/make_golden.py --out gpurun_out/golden
and copy the resulting files into tests/golden/. Each file is self-contained: the inputs and
every decoded artefact of one reference forward+backward (geometry state, sorted keys, tile
ranges, per-pixel contributor state, images, the nine gradient tensors).
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from materialrefgs_b200 import synthetic  # noqa: E402
from tests import refimpl  # noqa: E402

CASES = {
    # name: (P, S, W, H, opacity, bg, scale_modifier, view, scale_mult)
    "a_s8_trained": (2500, 8, 96, 80, "trained", (0.0, 0.0, 0.0), 1.0, 1, 1.0),
    "b_s0_init_bg": (1500, 0, 64, 64, "init", (0.2, 0.5, 0.9), 1.0, 3, 1.5),
    "c_s11_ragged_mod": (2000, 11, 50, 37, "trained", (0.1, 0.1, 0.1), 0.7, 5, 1.2),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden")
    a = ap.parse_args()
    out = Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    ref = refimpl.load_reference()
    assert ref is not None, "oracle/_ref is not built"
    dev = torch.device("cuda:0")
    for name, (P, S, W, H, opacity, bg, mod, view, smult) in CASES.items():
        cloud = synthetic.make_cloud(P, S=S, opacity=opacity, scale_mult=smult, seed=synthetic.SEED + len(name))
        cam = synthetic.orbit_camera(view, 8, W, H)
        gc, gf, go = synthetic.upstream_grads(S, H, W)
        cl, cm = cloud.to(dev), cam.to(dev)
        bgt = torch.tensor(bg, device=dev)
        e = torch.empty(0, device=dev)
        fargs = (bgt, cl.means3D, e, cl.features, cl.opacities, cl.scales, cl.rotations, mod, e,
                 cm.world_view_transform, cm.full_proj_transform, cm.tanfovx, cm.tanfovy, H, W, cl.shs, 3,
                 cm.camera_center, False, False)
        R, contrib, color, feat, others, radii, geom, binning, img = ref._C.rasterize_gaussians(*fargs)
        grads = ref._C.rasterize_gaussians_backward(
            bgt, cl.means3D, radii, e, cl.features, cl.scales, cl.rotations, mod, e, cm.world_view_transform,
            cm.full_proj_transform, cm.tanfovx, cm.tanfovy, gc.to(dev), gf.to(dev), go.to(dev), cl.shs, 3,
            cm.camera_center, geom, R, binning, img, contrib, False)
        torch.cuda.synchronize()
        g = refimpl.decode_ref_geom(geom, P)
        b = refimpl.decode_ref_binning(binning, R)
        im = refimpl.decode_ref_image(img, H, W)
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        vis = (radii > 0)
        n = lambda t: t.detach().cpu().numpy()

        def vis_only(t):  # culled rows are uninitialised memory in the reference: zero them
            t = t.clone()
            t[~vis] = 0
            return n(t)
        names = ("dL_dmeans2D", "dL_dcolors", "dL_dfeatures", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat",
                 "dL_dsh", "dL_dscales", "dL_drotations")
        np.savez_compressed(
            out / f"raster_{name}.npz",
            # inputs
            W=W, H=H, S=S, bg=np.array(bg, np.float32), scale_modifier=np.float32(mod), sh_degree=3,
            tan_fovx=np.float64(cam.tanfovx), tan_fovy=np.float64(cam.tanfovy),
            means3D=n(cloud.means3D), scales=n(cloud.scales), rotations=n(cloud.rotations),
            opacities=n(cloud.opacities), shs=n(cloud.shs), features=n(cloud.features),
            viewmatrix=n(cam.world_view_transform), projmatrix=n(cam.full_proj_transform),
            campos=n(cam.camera_center), dL_dcolor=n(gc), dL_dfeature=n(gf), dL_dothers=n(go),
            # reference outputs
            R=R, radii=n(radii), tiles_touched=n(g["tiles_touched"]), depths=vis_only(g["depths"]),
            means2D=vis_only(g["means2D"]), transMat=vis_only(g["transMat"]),
            normal_opacity=vis_only(g["normal_opacity"]), rgb=vis_only(g["rgb"]), clamped=vis_only(g["clamped"]),
            keys=n(b["keys"]), point_list=n(b["point_list"]), ranges=n(im["ranges"][:tiles]),
            final_T=n(im["accum_alpha"]), n_contrib=n(im["n_contrib"]), color=n(color), feature=n(feat),
            others=n(others), **{k: n(v) for k, v in zip(names, grads)})
        print(f"{name}: P={P} Pv={int(vis.sum())} R={R} contrib max {int(im['n_contrib'][0].max())}")


def cubemap_golden(out):
    """Golden vectors of the cubemap prefilter from the unmodified renderutils_plugin."""
    import importlib
    plug_dir = ROOT / "oracle" / "_ref" / "renderutils_plugin"
    if not (plug_dir / "renderutils_plugin.so").exists():
        print("renderutils_plugin not built; skipping cubemap golden")
        return
    sys.path.insert(0, str(plug_dir))
    plug = importlib.import_module("renderutils_plugin")
    sys.path.insert(0, str(ROOT))
    from oracle import cubemap_oracle as co
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(21)
    for N, rough in ((16, 0.3), (16, 1.0), (32, 0.08)):
        cube = torch.randn(6, N, N, 3, generator=g).to(dev)
        dout = torch.randn(6, N, N, 4, generator=g).to(dev)
        ct = co.ndf_cutoff_costheta(rough, 0.99)
        bounds = plug.specular_bounds(N, ct)
        spec = plug.specular_cubemap_fwd(cube, bounds, rough, ct)
        dspec = plug.specular_cubemap_bwd(cube, bounds, dout, rough, ct)
        diff = plug.diffuse_cubemap_fwd(cube)
        ddiff = plug.diffuse_cubemap_bwd(cube, dout[..., :3].contiguous())
        n = lambda t: t.detach().cpu().numpy()
        np.savez_compressed(out / f"cubemap_n{N}_r{int(rough * 100):03d}.npz", N=N, roughness=rough, cutoff=0.99,
                            costheta=ct, cube=n(cube), dout=n(dout), bounds=n(bounds), spec4=n(spec),
                            dspec=n(dspec), diffuse=n(diff), ddiffuse=n(ddiff))
        print(f"cubemap golden N={N} r={rough}: wsum mean {float(spec[..., 3].mean()):.5f}")


if __name__ == "__main__":
    main()
    cubemap_golden(Path(sys.argv[sys.argv.index("--out") + 1]) if "--out" in sys.argv else Path("gpurun_out/golden"))

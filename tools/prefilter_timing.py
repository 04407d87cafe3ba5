"""EnvLight.build_mips forward + backward at the BASELINE chain (6x512^2, 6 levels): prefilter plans vs the
unmodified reference plugin (oracle/_ref/renderutils_plugin) on the same GPU. Dev tool; prints JSON.

    python tools/prefilter_timing.py [--res 512] [--shape 32x1|32x2|16x1|16x2|8x1|8x2] [--no-ref] [--out profiles/x.json]
"""
import argparse
import importlib
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--min-res", type=int, default=16)
ap.add_argument("--shape", default=None, help="WxG, default: automatic per level")
ap.add_argument("--no-ref", action="store_true")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--out", default=None)
a = ap.parse_args()
if a.shape:
    os.environ["MRGS_PREFILTER_SHAPE"] = a.shape

from materialrefgs_b200 import _lib, prefilter as pf  # noqa: E402
from materialrefgs_b200.shading import EnvLight  # noqa: E402

dev = torch.device("cuda:0")
peak = 6542.1
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(n):
        flush.zero_()           # evict L2 between iterations
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


t0 = time.perf_counter()
env = EnvLight(device=dev, max_res=a.res, min_res=a.min_res, trainable=True)
torch.cuda.synchronize()
build_s = time.perf_counter() - t0
with torch.no_grad():
    env.base.normal_()
chain = env._chain
assert chain is not None
sink = torch.randn(chain.texels, 4, device=dev)
gd = torch.randn(6, a.min_res, a.min_res, 3, device=dev)

fwd_ms = timeit(lambda: chain.forward(env.base), a.iters)
bwd_ms = timeit(lambda: chain.backward(sink, None), a.iters)
bwd_d_ms = timeit(lambda: chain.backward(sink, gd), a.iters)

lib = _lib.load()
lib.mrgs_profile_enable(1)
lib.mrgs_profile_reset()
for _ in range(5):
    chain.forward(env.base)
    chain.backward(sink, None)
torch.cuda.synchronize()
prof = _lib.profile_read()
lib.mrgs_profile_enable(0)

levels = []
wbytes_f = wbytes_b = 0
for r, ro, (pfw, pbw) in zip(chain.sizes, chain.roughnesses, chain.spec):
    levels.append({"res": r, "roughness": ro, "taps": pfw.taps, "shape": f"{pfw.patch_width}x{pfw.rows_per_lane}",
                   "weight_slots_fwd": pfw.rows * pfw.rows_per_lane * 32, "weight_slots_bwd": pbw.rows * pbw.rows_per_lane * 32,
                   "segments_fwd": int(pfw.seg_desc.shape[0]), "plan_MB_fwd": pfw.nbytes / 1e6, "plan_MB_bwd": pbw.nbytes / 1e6})
    wbytes_f += pfw.nbytes
    wbytes_b += pbw.nbytes
wbytes_f += chain.diff[0].nbytes
wbytes_b += chain.diff[1].nbytes
taps = sum(l["taps"] for l in levels) + chain.diff[0].taps
# algorithmic bytes of one gather launch: every tap's weight once (4 B) + source and destination texels once
alg_f = 4 * taps + 16 * chain.texels + 12 * chain.texels
out = {
    "chain": f"6x{a.res}^2, {chain.n} levels, shape={a.shape or "auto"}", "plan_build_s": build_s,
    "taps": taps, "plan_bytes_fwd": wbytes_f, "plan_bytes_bwd": wbytes_b,
    "forward_ms": fwd_ms, "backward_ms": bwd_ms, "backward_with_diffuse_ms": bwd_d_ms,
    "stage_ms": {k: (v[0] / v[1] if v[1] else 0.0) for k, v in prof.items() if k.startswith("prefilter")},
    "gather_fwd_GBps_algorithmic": alg_f / (fwd_ms * 1e-3) / 1e9, "gather_fwd_GBps_plan_bytes": wbytes_f / (fwd_ms * 1e-3) / 1e9,
    "gather_bwd_GBps_plan_bytes": wbytes_b / (bwd_ms * 1e-3) / 1e9, "hbm_peak_GBps": peak,
    "levels": levels,
}

if not a.no_ref:
    d = ROOT / "oracle" / "_ref" / "renderutils_plugin"
    if (d / "renderutils_plugin.so").exists():
        sys.path.insert(0, str(d))
        ref = importlib.import_module("renderutils_plugin")
        from oracle import shading_oracle as so
        raw = [env.base.detach()]
        while raw[-1].shape[1] > a.min_res:
            raw.append(so.cubemap_mip(raw[-1]))
        keys = []
        t0 = time.perf_counter()
        for lvl, ro in zip(raw, chain.roughnesses):
            ct = pf.cutoff_costheta(ro, 0.99)
            keys.append((ro, ct, ref.specular_bounds(lvl.shape[1], ct)))
        torch.cuda.synchronize()
        out["reference_bounds_s"] = time.perf_counter() - t0
        douts = [torch.randn(6, l.shape[1], l.shape[1], 4, device=dev) for l in raw]

        def ref_fwd():
            r_ = [env.base.detach()]
            while r_[-1].shape[1] > a.min_res:
                r_.append(torch.nn.functional.avg_pool2d(r_[-1].permute(0, 3, 1, 2), (2, 2)).permute(0, 2, 3, 1).contiguous())
            ref.diffuse_cubemap_fwd(r_[-1])
            for lvl, (ro, ct, b) in zip(r_, keys):
                o4 = ref.specular_cubemap_fwd(lvl, b, ro, ct)
                _ = o4[..., :3] / o4[..., 3:]

        def ref_bwd():
            ref.diffuse_cubemap_bwd(raw[-1], gd)
            for lvl, (ro, ct, b), d4 in zip(raw, keys, douts):
                ref.specular_cubemap_bwd(lvl, b, d4, ro, ct)

        out["reference_forward_ms"] = timeit(ref_fwd, 5)
        out["reference_backward_ms"] = timeit(ref_bwd, 5)
        out["reference_note"] = ("unmodified plugin ops per level; its cubemap_mip backward (nvdiffrast dr.texture) "
                                 "is not available and not timed")
print(json.dumps(out, indent=1))
if a.out:
    Path(a.out).write_text(json.dumps(out, indent=1))

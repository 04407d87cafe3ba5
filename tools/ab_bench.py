"""A/B harness for kernel variants selected by environment variables: runs bench.py once per value and prints the
frames/s and the per-stage CUDA-event times side by side. Dev tool (one gpurun call measures a whole sweep).

    python tools/ab_bench.py MRGS_BWD_SPLIT=0,82 --steps 12 --stages render_bwd render_fwd
    python tools/ab_bench.py MRGS_OPTIMISTIC_BINNING=1,0 --repeat 2
"""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run(env_name, value, steps, extra):
    env = dict(os.environ)
    env[env_name] = value
    cmd = [sys.executable, str(ROOT / "bench.py"), "--no-cpu-baseline", "--steps", str(steps), *extra]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    if not lines:
        return {"error": (res.stderr or res.stdout)[-300:]}
    return json.loads(lines[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sweep", help="NAME=v1,v2,...")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--stages", nargs="*", default=["render_fwd", "render_bwd", "preprocess_fwd", "preprocess_bwd", "sort"])
    ap.add_argument("bench_args", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    name, values = a.sweep.split("=", 1)
    print(f"{name:>24s} {'frames/s':>9s} {'e2e':>8s} " + " ".join(f"{s:>15s}" for s in a.stages))
    for _ in range(a.repeat):
        for v in values.split(","):
            d = run(name, v, a.steps, a.bench_args)
            if "value" not in d:
                print(f"{v:>24s} {d.get('unavailable') or d.get('error')}")
                continue
            st = d.get("stage_ms", {})
            print(f"{v:>24s} {d['value']:9.1f} {d['e2e']['value']:8.1f} " + " ".join(f"{st.get(s, float('nan')):15.4f}" for s in a.stages))


if __name__ == "__main__":
    main()

"""Per-SASS-instruction hot spots of one kernel from an .ncu-rep captured with --import-source on: executed warp
instructions, lane utilisation and stall samples, grouped into contiguous regions. Dev tool (runs without a GPU).

    python tools/ncu_hotspots.py gpurun_out/r01_v8_bwd.ncu-rep render_bwd [--top 25]
"""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-name", f"regex:{a.kernel}"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    rows = [r for r in rows if r.get("Instructions Executed", "").replace(",", "").isdigit()]
    tot_i = sum(int(r["Instructions Executed"]) for r in rows)
    tot_s = sum(int(r["# Samples"]) for r in rows)
    print(f"{len(rows)} SASS instructions, {tot_i / 1e6:.1f} M warp instructions executed, {tot_s} stall samples")
    print("\n-- top instructions by stall samples")
    for r in sorted(rows, key=lambda r: -int(r["# Samples"]))[:a.top]:
        n, s = int(r["Instructions Executed"]), int(r["# Samples"])
        stalls = {k[6:]: int(v) for k, v in r.items() if k.startswith("stall_") and "(" not in k and v.isdigit() and int(v) > 0}
        top = ", ".join(f"{k} {v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100 * s / tot_s:5.1f}% samples {100 * n / tot_i:5.2f}% instr  thr {r['Avg. Threads Executed']:>5s}  {r['Source'].strip()[:60]:60s} | {top}")
    # opcode classes
    agg = {}
    for r in rows:
        op = r["Source"].strip().split()[0]
        if op.startswith("@"):
            op = r["Source"].strip().split()[1]
        op = op.split(".")[0]
        e = agg.setdefault(op, [0, 0])
        e[0] += int(r["Instructions Executed"])
        e[1] += int(r["# Samples"])
    print("\n-- by opcode")
    for op, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:18]:
        print(f"{op:10s} {100 * n / tot_i:5.1f}% instr {100 * s / tot_s:5.1f}% samples")


if __name__ == "__main__":
    main()

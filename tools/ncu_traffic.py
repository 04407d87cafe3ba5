"""DRAM traffic per launch of every kernel in an `ncu --set full` capture -> profiles/ncu_traffic.json, the tracked file
bench.py reads `roofline.traffic` from. Dev tool (runs without a GPU).

    python tools/ncu_traffic.py gpurun_out/r02_full.ncu-rep [more.ncu-rep ...] [--out profiles/ncu_traffic.json]

Kernel keys are the stage names bench.py uses (render_fwd, render_bwd, preprocess_fwd, preprocess_bwd,
prefilter_gather, ...) mapped from the demangled kernel names; several launches of one kernel are averaged."""
import argparse
import csv
import io
import json
import subprocess
from pathlib import Path

STAGE_OF = (("render_bwd_kernel", "render_bwd"), ("render_fwd_kernel", "render_fwd"), ("preprocess_bwd", "preprocess_bwd"),
            ("preprocess_fwd", "preprocess_fwd"), ("prefilter_gather", "prefilter_gather"), ("radix_scatter", "radix_scatter"),
            ("shade_bwd", "shade_bwd"), ("shade_fwd", "shade_fwd"), ("emit_instances", "emit_instances"))
METRICS = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
           "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3,
         "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--out", default="profiles/ncu_traffic.json")
    a = ap.parse_args()
    acc = {}
    for rep in a.reports:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            stage = next((s for k, s in STAGE_OF if k in name), None)
            if stage is None:
                continue
            e = acc.setdefault(stage, {"launches": 0, "kernel": name.split("(")[0], "source": []})
            e["launches"] += 1
            if Path(rep).name not in e["source"]:
                e["source"].append(Path(rep).name)
            for m in METRICS:
                if m in col and r[col[m]] not in ("", "n/a"):
                    v = float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)
                    e[m] = e.get(m, 0.0) + v
    kernels = {}
    for stage, e in acc.items():
        n = e["launches"]
        k = {"kernel": e["kernel"], "launches_in_capture": n, "source": e["source"]}
        for m in METRICS:
            if m in e:
                k[m.replace(".", "_")] = e[m] / n
        k["dram_bytes_per_launch"] = (e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)) / n
        kernels[stage] = k
    doc = {"source": ", ".join(Path(r).name for r in a.reports),
           "what": "per-launch averages from ncu --set full --clock-control none captures of `python bench.py` (C3)",
           "kernels": kernels}
    Path(a.out).write_text(json.dumps(doc, indent=1))
    for s, k in kernels.items():
        print(f"{s:18s} {k['launches_in_capture']:3d} launches  dram {k['dram_bytes_per_launch'] / 1e6:9.1f} MB  "
              f"time {k.get('gpu__time_duration_sum', 0):.3f} ms")


if __name__ == "__main__":
    main()

#!/bin/bash
# forward/backward time of the 6x512^2 chain for every patch shape (dev tool, run on a GPU box)
for s in 32x1 32x2 16x1 16x2 8x1 8x2; do
  timeout 200 python tools/prefilter_timing.py --shape $s --no-ref --iters 10 --out gpurun_out/prefilter_$s.json > /dev/null 2>&1
  python - <<PY
import json
d=json.load(open("gpurun_out/prefilter_$s.json"))
print("$s", round(d["forward_ms"],3), round(d["backward_ms"],3), d["plan_bytes_fwd"], [ (l["res"], round(l["weight_slots_fwd"]/l["taps"],3)) for l in d["levels"]])
PY
done

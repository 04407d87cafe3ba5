#!/bin/bash
# bench.py at N GPUs of one box, launched like the driver launches it: C3 (weak), C5 (8-view batch, strong) and the
# 200-camera evaluation shard. Usage: tools/scale_run.sh N [outfile]
N=${1:-2}
OUT=${2:-gpurun_out/scale_n$N.jsonl}
: > $OUT
run() {
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" >> $OUT 2>> gpurun_out/scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" >> $OUT 2>> gpurun_out/scale_n$N.err
  fi
}
run --steps 20 --warmup 3 --no-cpu-baseline
run --config C5 --steps 10 --warmup 3 --no-cpu-baseline
run --config C5-eval --steps 3 --warmup 3 --no-cpu-baseline
python - <<PY
import json
for l in open("$OUT"):
    l=l.strip()
    if not l.startswith("{"): continue
    a=json.loads(l)
    print(a["config"]["name"], "n=%d"%a["n_gpus"], "value=%.1f"%a["value"], "ms/step=%.2f"%a["ms_per_step"], "e2e=%.1f"%a["e2e"]["value"])
PY
tail -3 gpurun_out/scale_n$N.err

"""Kineto trace of a few bench steps: GPU busy time vs wall time, top kernels, idle gaps. Dev tool."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
st = bench.OursStep(dev, 0, 1)
for i in range(5):
    st.step(i)
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        st.step(5 + i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.elapsed_us() for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"GPU busy {busy / N:.1f} us/step, span {span / N:.1f} us/step, idle {(span - busy) / N:.1f} us/step, kernels/step {len(ev) / N:.1f}")
gaps = []
for a, b in zip(ev[:-1], ev[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 15:
        gaps.append((g, a.name[:50], b.name[:50]))
gaps.sort(reverse=True)
for g in gaps[:12]:
    print(f"gap {g[0]:.0f} us after {g[1]} before {g[2]}")
agg = {}
for e in ev:
    k = e.name[:60]
    agg.setdefault(k, [0.0, 0])
    agg[k][0] += e.time_range.elapsed_us()
    agg[k][1] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{v[0] / N:9.1f} us/step {v[1] / N:5.1f}x  {k}")
# ordered kernel list of the last view of the last step (names shortened), to see the torch glue
import re
last_fwd = max(i for i, e in enumerate(ev) if "preprocess_fwd" in e.name)
print("---- ordered kernels of the last view ----")
for e in ev[last_fwd - 3:]:
    nm = re.sub(r"void |at::native::|\(anonymous namespace\)::|mrgs::", "", e.name)[:110]
    print(f"{e.time_range.elapsed_us():8.1f} us  {nm}")

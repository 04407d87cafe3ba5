"""Host cost of one bench view: the same step on a workload so small that the GPU is never the limit. Dev tool."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

bench.WORKLOAD.update(P=2000, H=64, W=64, cube_res=64, min_res=16)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
st = bench.OursStep(dev, 0, 1)
for i in range(5):
    st.step(i)
torch.cuda.synchronize()
N = 50
t0 = time.perf_counter()
for i in range(N):
    st.step(5 + i)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / N
print(f"tiny workload: {dt * 1e3:.3f} ms/step = {dt * 1e3 / bench.VIEWS_PER_RANK:.3f} ms of host time per view")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    st.step(100 + i)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)

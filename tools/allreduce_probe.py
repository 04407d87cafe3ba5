"""Times the pieces of the multi-GPU step tail (dev tool): allreduce sizes, flatten copy, stats ops."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from materialrefgs_b200.parallel import GradArena
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
P = 1_000_000
arena = GradArena.create(P, dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
t_flat = timeit(lambda: dist.all_reduce(arena.flat))
t_stats = timeit(lambda: dist.all_reduce(arena.stats))
t_max = timeit(lambda: dist.all_reduce(arena.max_radii, op=dist.ReduceOp.MAX))
g = {n: torch.randn(P, w, device=dev) for n, w in (("means3D", 3), ("scales", 2), ("rotations", 4), ("opacities", 1), ("shs", 48), ("features", 8))}
def flatten():
    for name, v in arena.views.items(): v.copy_(g[name])
t_copy = timeit(flatten)
m2 = torch.randn(P, 3, device=dev); radii = torch.randint(0, 30, (P,), device=dev, dtype=torch.int32)
t_acc = timeit(lambda: arena.accumulate_view({}, m2, radii))
if rank == 0:
    print(f"world {world}: allreduce flat {arena.flat.numel()*4/1e6:.0f} MB {t_flat:.3f} ms (busbw {2*(world-1)/world*arena.flat.numel()*4/t_flat/1e6:.0f} GB/s), stats {t_stats:.3f} ms, max {t_max:.3f} ms, flatten copy {t_copy:.3f} ms, accumulate_view {t_acc:.3f} ms")
dist.destroy_process_group()

"""torchrun --nproc-per-node N tools/shard_check.py: EnvLight.build_mips split over the ranks (shard_build_mips) must
reproduce the unsharded chain and base gradient bit for bit on every rank. Dev tool for a multi-GPU box."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from materialrefgs_b200.shading import EnvLight  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator().manual_seed(7)
base = torch.randn(6, res, res, 3, generator=g).to(dev)
out = {}
for sharded in (False, True):
    env = EnvLight(device=dev, max_res=res, min_res=16, trainable=True)
    with torch.no_grad():
        env.base.copy_(base)
    if sharded:
        env.shard_build_mips()
    env.build_mips()
    sink = env.enable_level_grad_sink()
    gg = torch.Generator().manual_seed(9)
    sink[:, :3] = torch.randn(sink.shape[0], 3, generator=gg).to(dev)
    sink._mrgs_pending = True
    levels = [l.detach().clone() for l in env.specular] + [env.diffuse.detach().clone()]
    env.flush_level_grads()
    out[sharded] = (levels, env.base.grad.clone())
ok = all(torch.equal(a, b) for a, b in zip(out[False][0], out[True][0])) and torch.equal(out[False][1], out[True][1])
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if dist.get_rank() == 0:
    print("sharded build_mips == unsharded on all ranks:", bool(t.item()), "(res %d, world %d)" % (res, dist.get_world_size()))
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)

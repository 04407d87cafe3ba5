"""CPU, world_size 2, gloo: the host logic of the view-sharded step (materialrefgs_b200/parallel.py).
The per-view renderer is a deterministic stand-in; what is checked is the sharding, the arena layout,
the single sum-allreduce, and the densification semantics of scene/gaussian_model.py:1059-1061 +
train_refnerf.py:1416-1418 (norm per view, then summed; visibility count; max radii)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from materialrefgs_b200 import parallel

P, NV = 37, 5


def fake_view(i):
    g = torch.Generator().manual_seed(100 + i)
    grads = {name: torch.randn(P, w, generator=g) for name, w in parallel.GRAD_FIELDS}
    grads["shs"] = grads["shs"].view(P, 16, 3)
    radii = torch.randint(-1, 30, (P,), generator=g, dtype=torch.int32).clamp_min(0)
    return {"grads": grads, "viewspace_grad": torch.randn(P, 3, generator=g), "radii": radii}


def expected():
    flat = torch.zeros(P, sum(w for _, w in parallel.GRAD_FIELDS))
    stats = torch.zeros(P, 2)
    mx = torch.zeros(P, dtype=torch.int32)
    for i in range(NV):
        v = fake_view(i)
        flat += torch.cat([v["grads"][n].reshape(P, -1) for n, _ in parallel.GRAD_FIELDS], 1)
        vis = v["radii"] > 0
        stats[:, 0] += torch.linalg.norm(v["viewspace_grad"], dim=-1) * vis
        stats[:, 1] += vis.float()
        mx = torch.maximum(mx, v["radii"])
    return flat, stats, mx


def _cat(arena):
    return torch.cat([arena.views[n] for n, _ in parallel.GRAD_FIELDS], 1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        arena = parallel.GradArena.create(P, "cpu")
        seen = []

        def render(i):
            seen.append(i)
            return fake_view(i)
        views = parallel.train_step_view_sharded(render, NV, arena)
        ev = parallel.eval_views_sharded(lambda i: torch.tensor(float(i)), 7)
        # the same exchange in its two overlappable pieces, with a tail (the cubemap texel-gradient sink) in the SAME buffer
        a2 = parallel.GradArena.create(P, "cpu", extra_floats=4 * 11)
        assert a2.extra.numel() == 44 and a2.extra.data_ptr() == a2.flat.data_ptr() + 4 * (a2.flat.numel() - 44)
        assert (a2.extra.data_ptr() - a2.flat.data_ptr()) % 16 == 0 and a2.main.numel() == P * 68
        for i in parallel.shard_views(NV, rank, world):
            v = fake_view(i)
            a2.accumulate_view(v["grads"], v["viewspace_grad"], v["radii"])
        a2.extra += float(rank + 1)
        w_tail = a2.allreduce_extra_async()
        w_main = a2.allreduce_main_async()
        a2.wait(w_tail)
        a2.wait(w_main)
        assert torch.allclose(_cat(a2), _cat(arena), atol=1e-5) and torch.allclose(a2.stats, arena.stats, atol=1e-5)
        assert torch.equal(a2.max_radii, arena.max_radii)
        assert torch.equal(a2.extra, torch.full((44,), float(sum(range(1, world + 1)))))
        a3 = parallel.GradArena.create(P, "cpu", extra_floats=8)      # ... and as the ONE blocking collective
        a3.flat.fill_(1.0)
        a3.allreduce()
        assert torch.equal(a3.flat, torch.full_like(a3.flat, float(world)))
        # numpy arrays travel through the queue BY VALUE; torch tensors would go through shared-memory file descriptors
        # that vanish when this process exits before the parent has read them (ConnectionResetError)
        q.put((rank, seen, _cat(arena).numpy(), arena.stats.numpy().copy(), arena.max_radii.numpy().copy(),
               sorted(ev.keys()), tuple(views["shs"].shape)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(420)
def test_view_sharded_step_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    flat, stats, mx = expected()
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]          # round-robin, disjoint, complete
    for rank, seen, f, s, m, ev_keys, shs_shape in res:
        assert torch.allclose(torch.from_numpy(f), flat, atol=1e-5)
        assert torch.allclose(torch.from_numpy(s), stats, atol=1e-5)
        assert torch.equal(torch.from_numpy(m), mx)
        assert ev_keys == list(range(rank, 7, 2))
        assert tuple(shs_shape) == (P, 48)


def test_single_process_is_the_plain_sum():
    arena = parallel.GradArena.create(P, "cpu")
    parallel.train_step_view_sharded(fake_view, NV, arena)
    flat, stats, mx = expected()
    assert torch.allclose(_cat(arena), flat, atol=1e-5) and torch.allclose(arena.stats, stats, atol=1e-5)
    assert torch.equal(arena.max_radii, mx)
    assert parallel.shard_views(10, 3, 4) == [3, 7]


def test_cost_balanced_view_assignment():
    costs = [9.0, 1.0, 8.0, 2.0, 7.0, 3.0, 6.0, 4.0]
    parts = parallel.assign_views_balanced(costs, 4)
    assert sorted(i for p in parts for i in p) == list(range(8)) and all(len(p) == 2 for p in parts)
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 1e-9              # 9+1, 8+2, 7+3, 6+4
    rr = [sum(costs[i] for i in parallel.shard_views(8, r, 4)) for r in range(4)]
    assert max(rr) - min(rr) > 5.0                      # round-robin would leave one rank with 9+7 against 1+3
    uneven = parallel.assign_views_balanced([5.0, 4.0, 3.0, 2.0, 1.0], 2)
    assert sorted(len(p) for p in uneven) == [2, 3] and sorted(i for p in uneven for i in p) == list(range(5))
    assert parallel.assign_views_balanced(costs, 1) == [[0, 2, 4, 6, 7, 5, 3, 1]]


def test_bound_arena_receives_autograd_accumulation_in_place():
    """bind(): the parameters' .grad ARE the arena segments, autograd adds every view into them and the
    storage never moves (so no flattening copy is needed before the allreduce)."""
    arena = parallel.GradArena.create(P, "cpu")
    leaves = {n: torch.randn(P, 16, 3, requires_grad=True) if n == "shs" else torch.randn(P, w, requires_grad=True)
              for n, w in parallel.GRAD_FIELDS}
    arena.bind(leaves)
    arena.zero_()
    coef = [1.5, -2.0, 0.25]
    for c in coef:
        sum((l * l).sum() * c for l in leaves.values()).backward()
    assert arena.bound(leaves)
    for n, l in leaves.items():
        assert torch.allclose(arena.views[n], (2 * sum(coef) * l.detach()).reshape(P, -1), atol=1e-5)
    assert arena.grads.numel() == P * 66 and arena.flat.numel() == P * 68
    arena.zero_()
    assert float(leaves["shs"].grad.abs().max()) == 0.0


# ---- densification after a view-sharded step: every rank must arrive at the same cloud without a broadcast ----
def _densify_worker(rank, world, port, q):
    from materialrefgs_b200 import surfel_model as sm
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        fields = {n: 0.3 * torch.randn((P, *f.shape), generator=g) for n, f in sm.FIELDS.items()}
        fields["scaling"] = torch.log(0.02 * torch.exp(0.8 * torch.randn((P, 2), generator=g)))
        store = sm.SurfelStore(fields)
        arena = parallel.GradArena.create(P, "cpu")

        def render(i):
            v = fake_view(i)
            v["viewspace_grad"] = 4e-4 * v["viewspace_grad"]
            v["viewspace_grad"][:, 2] = 0.0          # the rasterizer's screen-space gradient has no z part
            return v
        parallel.train_step_view_sharded(render, NV, arena)
        store.load_reduced_stats(arena.stats, arena.max_radii)
        gen = torch.Generator().manual_seed(1234)    # same seed on every rank
        store.densify_and_prune(0.0002, 0.05, 2.5, 20, generator=gen)
        q.put((rank, store.num_points, store.checksum().tolist(), store["xyz"].detach().numpy().copy()))
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.timeout(420)
def test_ranks_densify_to_the_same_cloud_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_densify_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    single = ctx.Queue()
    _densify_worker(0, 1, 0, single)
    ref = single.get(timeout=10)
    assert res[0][1] == res[1][1] == ref[1] != P
    assert res[0][2] == res[1][2]
    assert (res[0][3] == res[1][3]).all()
    assert abs(res[0][3] - ref[3]).max() <= 1e-6             # sharded sum order differs from the serial one by rounding only

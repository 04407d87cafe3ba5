"""CUDA-graph replay of a whole training view (rasterize -> shade -> loss -> backward -> statistics).

A view of the render path is ~60 kernel launches (26 of them ours) with host work in between; replaying it as ONE graph
removes the launch gaps and the Python/autograd time per view (~9 % of a frame at C3). What makes a view
capturable:
  * the rasterizer's forward normally waits for the instance count R (it sizes the binning buffer, like the
    reference's rasterizer_impl.cu:287). Under capture it runs in the library's no_wait mode (include/mrgs.h,
    MrgsForwardArgs.no_wait): everything is enqueued against the capacity learnt from earlier frames, the kernels read R
    from device memory, and R is copied to a pinned int that `check()` compares with the capacity after a replay;
  * everything a camera contributes on the host (tan fov, ray / normal matrices) is baked into the kernel parameters, so
    a graph belongs to ONE camera: `ViewGraphs` keeps one graph per camera key, all sharing one memory pool;
  * every tensor the view reads must keep its address: parameters and the gradient arena do; the environment chain does
    with `EnvLight.static_chain = True`; per-view inputs (upstream gradients) live in fixed slots the caller fills
    before a replay.
Gradients never pass through autograd's accumulation: the per-surfel backward adds into the arena (grad_sink) and the
shading backward into the texel-gradient sink, so a replay accumulates exactly like an eager view does.
"""
from __future__ import annotations

from typing import Callable, Dict, Hashable

import torch

from . import rasterizer as _rz


class ViewGraphs:
    def __init__(self, device):
        self.device = torch.device(device)
        self.pool = None
        self.graphs: Dict[Hashable, tuple] = {}
        # Eager first calls and captures run on ONE side stream: autograd ties a leaf's gradient accumulator to the
        # stream that was current when the accumulator was created; an accumulator born on the legacy default stream would
        # make the captured backward synchronise with that stream, which CUDA forbids during capture.
        self.side = torch.cuda.Stream(device=self.device)

    def run(self, key: Hashable, fn: Callable[[], object]):
        """Result of fn() for camera `key`. The first call runs fn eagerly - that is the call's result, and it teaches the
        rasterizer its capacity and fills every lazily built cache - and then captures fn for later (capturing executes
        nothing, so gradients are accumulated exactly once). Later calls replay the graph and return the graph's own
        output tensors, rewritten in place by every replay."""
        hit = self.graphs.get(key)
        if hit is not None:
            hit[0].replay()
            return hit[1]
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            out = fn()
        main.wait_stream(self.side)
        g = torch.cuda.CUDAGraph()
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        with torch.cuda.graph(g, pool=self.pool, stream=self.side):
            captured = fn()
        self.graphs[key] = (g, captured)
        return out

    def check(self) -> None:
        """After a synchronisation point: raise if a replayed view overflowed its instance capacity."""
        if not _rz.captured_counts_ok():
            raise RuntimeError("a captured view produced more (tile, surfel) instances than the capacity it was captured "
                               "with: its outputs are invalid; render the view eagerly once and re-capture")

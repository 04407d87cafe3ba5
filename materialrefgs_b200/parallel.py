"""View-sharded multi-GPU use of the render path. New relative to the reference,
which is single-GPU and renders one view per iteration (train_refnerf.py:1168-1173).

  * evaluation: cameras are dealt round-robin to ranks, no collective (eval.py:23-74 renders
    independent cameras);
  * training step: a batch of views is split across ranks; every rank accumulates the per-surfel
    parameter gradients of its views into ONE flat fp32 arena (a contiguous segment per field,
    bound as the parameters' .grad so autograd accumulates in place) with the densification
    statistics at its tail; a single sum-allreduce over that arena and one max-allreduce over
    max_radii2D follow. Densification semantics follow
    scene/gaussian_model.py:1059-1061 + train_refnerf.py:1416-1418: the norm of the screen-space
    gradient is taken PER VIEW, then summed; denom counts the views in which the surfel was visible.

Host logic only (torch.distributed); the kernels are the ones in libmrgs.so.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Sequence

import torch
import torch.distributed as dist

# parameter-space gradient fields of one surfel, in arena order (floats per field)
GRAD_FIELDS = (("means3D", 3), ("scales", 2), ("rotations", 4), ("opacities", 1), ("shs", 48), ("features", 8))


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of view indices to a rank (eval and train batches alike)."""
    return list(range(rank, n_views, world))


def assign_views_balanced(costs: Sequence[float], world: int) -> List[List[int]]:
    """Deal len(costs) views to `world` ranks by cost (e.g. the instance count R of each camera's last render, which the
    forward returns to the host anyway): longest-processing-time-first with equal counts per rank (the remainder goes to
    the lightest ranks), so that no rank waits for the slowest one at the step's collective. Deterministic: every
    rank computes the same assignment from the same cost table. Returns the item indices of every rank."""
    n = len(costs)
    cap = [n // world + (1 if r < n % world else 0) for r in range(world)]
    order = sorted(range(n), key=lambda i: (-costs[i], i))
    load, out = [0.0] * world, [[] for _ in range(world)]
    for i in order:
        r = min((r for r in range(world) if len(out[r]) < cap[r]), key=lambda r: (load[r], r))
        out[r].append(i)
        load[r] += costs[i]
    return out


@dataclass
class GradArena:
    """ONE flat fp32 buffer holding, back to back, a contiguous [P, w] segment per gradient field and
    the densification statistics (xyz_gradient_accum, denom packed as [P, 2]) at the tail, so that a
    single sum-allreduce covers all of it; max_radii2D [P] (int32) needs a max-allreduce of its own.

    Segments are contiguous so they can BE the parameters' .grad (bind()): autograd then accumulates
    every view's gradient in place and no flattening copy precedes the allreduce."""
    flat: torch.Tensor
    views: Dict[str, torch.Tensor]
    stats: torch.Tensor
    max_radii: torch.Tensor
    extra: torch.Tensor = None     # optional tail summed by the same collective (EnvLight's texel-gradient sink)

    @staticmethod
    def create(P: int, device, fields: Sequence = GRAD_FIELDS, extra_floats: int = 0) -> "GradArena":
        """extra_floats: further fp32 values at the END of the same flat buffer (after a 16-byte aligned offset), e.g.
        4 * texels for the environment map's [texels, 4] gradient sink (EnvLight.use_level_grad_sink), so that ONE
        sum-allreduce covers per-surfel gradients, densification statistics and the cubemap gradients."""
        F = sum(w for _, w in fields)
        main = P * (F + 2)
        pad = (-main) % 4
        flat = torch.zeros((main + pad + extra_floats,), dtype=torch.float32, device=device)
        views, o = {}, 0
        for name, w in fields:
            views[name] = flat[o:o + P * w].view(P, w)
            o += P * w
        return GradArena(flat, views, flat[o:o + 2 * P].view(P, 2),
                         torch.zeros((P,), dtype=torch.int32, device=device),
                         flat[main + pad:] if extra_floats else None)

    @property
    def main(self) -> torch.Tensor:
        """Gradients + statistics (everything but the extra tail), 1-D."""
        return self.flat if self.extra is None else self.flat[:self.flat.numel() - self.extra.numel()]

    @property
    def grads(self) -> torch.Tensor:
        """The gradient part of the buffer (without the statistics and the extra tail), 1-D."""
        return self.main[:self.main.numel() - self.stats.numel()]

    def zero_(self):
        self.flat.zero_()
        self.max_radii.zero_()
        if self.extra is not None:       # (the sink's owner tracks "holds unflushed gradients" on the tensor)
            self.extra._mrgs_pending = False

    def bind(self, leaves: Dict[str, torch.Tensor]):
        """Make each parameter's .grad a view of its arena segment (call once; zero_() per step).
        With .grad present and grad mode off, autograd's accumulation is an in-place add, so the
        storage stays the arena's."""
        for name, v in self.views.items():
            leaf = leaves.get(name)
            if leaf is not None:
                leaf.grad = v.view(leaf.shape)

    def bound(self, leaves: Dict[str, torch.Tensor]) -> bool:
        return all(leaves[n].grad is not None and leaves[n].grad.data_ptr() == v.data_ptr()
                   for n, v in self.views.items() if n in leaves)

    def accumulate_view(self, grads: Dict[str, torch.Tensor], viewspace_grad: torch.Tensor, radii: torch.Tensor):
        """Add one view's gradients (those not already accumulated through bind()); the densification
        statistics use that view's own norm. On the GPU the statistics are one libmrgs kernel
        (mrgs_densify_stats); the torch expression below serves host tensors (the gloo tests)."""
        for name, v in self.views.items():
            g = grads.get(name)
            if g is not None:
                v.add_(g.reshape(v.shape))
        if viewspace_grad.is_cuda:
            import ctypes as C
            from . import _lib
            lib = _lib.load()
            P = radii.shape[0]
            g = viewspace_grad.contiguous()
            assert g.dtype == torch.float32 and g.shape == (P, 3) and radii.dtype == torch.int32
            stream = torch.cuda.current_stream(g.device).cuda_stream
            _lib.check(lib.mrgs_densify_stats(P, g.data_ptr(), radii.contiguous().data_ptr(), self.stats.data_ptr(),
                                              self.max_radii.data_ptr(), C.c_void_p(stream)), "mrgs_densify_stats")
            return
        vis = radii > 0
        norm = torch.linalg.norm(viewspace_grad, dim=-1)      # all components, like gaussian_model.py:1060
        self.stats[:, 0].add_(torch.where(vis, norm, torch.zeros_like(norm)))
        self.stats[:, 1].add_(vis.to(torch.float32))
        torch.maximum(self.max_radii, torch.where(vis, radii, torch.zeros_like(radii)), out=self.max_radii)

    def allreduce(self, group=None, extra=()):
        """ONE sum-allreduce for gradients + statistics and one max-allreduce for the radii. `extra`: further
        buffers to sum across ranks in the same breath (the environment map's texel-gradient sink, 33 MB at
        6 x 512^2: EnvLight.level_grad_sink)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)     # incl. the extra tail, if any
        for t in extra:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=group)

    # ---- the same exchange in two overlappable pieces -------------------------------------------------------
    # The extra tail (cubemap texel gradients) is complete as soon as the LAST view's shading backward is enqueued,
    # i.e. before that view's rasterizer backward: its allreduce can run under those kernels. The main part is
    # complete after the last per-surfel backward; its allreduce can run under the build_mips backward, which
    # only consumes the (already reduced) tail. NCCL work is enqueued on NCCL's own stream, ordered after
    # everything enqueued so far on the current stream; wait() orders the current stream after the collective.
    @staticmethod
    def _active(group) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def allreduce_extra_async(self, group=None):
        if not self._active(group) or self.extra is None:
            return None
        return dist.all_reduce(self.extra, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def allreduce_main_async(self, group=None):
        if not self._active(group):
            return None
        return [dist.all_reduce(self.main, op=dist.ReduceOp.SUM, group=group, async_op=True),
                dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=group, async_op=True)]

    @staticmethod
    def wait(work):
        for w in (work if isinstance(work, (list, tuple)) else [work]):
            if w is not None:
                w.wait()


def train_step_view_sharded(render_view: Callable[[int], Dict], n_views: int, arena: GradArena,
                            group=None) -> Dict[str, torch.Tensor]:
    """render_view(i) must run forward+backward of view i and return
    {'grads': {field: tensor}, 'viewspace_grad': [P,3], 'radii': [P]}. After the call every rank
    holds the batch-summed gradients in arena.views (== the sum of n_views reference iterations
    without an optimizer step in between)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    arena.zero_()
    for i in shard_views(n_views, rank, world):
        out = render_view(i)
        arena.accumulate_view(out["grads"], out["viewspace_grad"], out["radii"])
    arena.allreduce(group)
    return arena.views


def eval_views_sharded(render_view: Callable[[int], torch.Tensor], n_views: int, group=None) -> Dict[int, torch.Tensor]:
    """Each rank renders its own cameras; results stay local (gather on the host if needed)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    return {i: render_view(i) for i in shard_views(n_views, rank, world)}

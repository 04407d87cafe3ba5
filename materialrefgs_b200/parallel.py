"""View-sharded multi-GPU use of the render path. New relative to the reference,
which is single-GPU and renders one view per iteration (train_refnerf.py:1168-1173).

  * evaluation: cameras are dealt round-robin to ranks, no collective (eval.py:23-74 renders
    independent cameras);
  * training step: a batch of views is split across ranks; every rank accumulates the per-surfel
    parameter gradients of its views into ONE flat fp32 arena [P, F] and the densification
    statistics into a second small arena; a single sum-allreduce over the gradient arena and one
    max-allreduce over max_radii2D follow. Densification semantics follow
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


@dataclass
class GradArena:
    """Flat [P, F] fp32 gradient buffer with named column views, plus densification statistics
    (xyz_gradient_accum, denom packed as [P, 2]) and max_radii2D [P]."""
    flat: torch.Tensor
    views: Dict[str, torch.Tensor]
    stats: torch.Tensor
    max_radii: torch.Tensor

    @staticmethod
    def create(P: int, device, fields: Sequence = GRAD_FIELDS) -> "GradArena":
        F = sum(w for _, w in fields)
        flat = torch.zeros((P, F), dtype=torch.float32, device=device)
        views, o = {}, 0
        for name, w in fields:
            views[name] = flat[:, o:o + w]
            o += w
        return GradArena(flat, views, torch.zeros((P, 2), dtype=torch.float32, device=device),
                         torch.zeros((P,), dtype=torch.int32, device=device))

    def zero_(self):
        self.flat.zero_()
        self.stats.zero_()
        self.max_radii.zero_()

    def accumulate_view(self, grads: Dict[str, torch.Tensor], viewspace_grad: torch.Tensor, radii: torch.Tensor):
        """Add one view's gradients; densification statistics use that view's own norm."""
        for name, v in self.views.items():
            g = grads.get(name)
            if g is not None:
                v.add_(g.reshape(v.shape))
        vis = radii > 0
        norm = torch.linalg.norm(viewspace_grad[:, :2], dim=-1)
        self.stats[:, 0].add_(torch.where(vis, norm, torch.zeros_like(norm)))
        self.stats[:, 1].add_(vis.to(torch.float32))
        torch.maximum(self.max_radii, torch.where(vis, radii, torch.zeros_like(radii)), out=self.max_radii)

    def allreduce(self, group=None):
        """ONE sum-allreduce for gradients (+stats appended) and one max-allreduce for the radii."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.stats, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=group)


def train_step_view_sharded(render_view: Callable[[int], Dict], n_views: int, arena: GradArena,
                            group=None) -> Dict[str, torch.Tensor]:
    """render_view(i) must run forward+backward of view i and return
    {'grads': {field: tensor}, 'viewspace_grad': [P,3], 'radii': [P]}. After the call every rank
    holds the batch-summed gradients in arena.flat (== the sum of n_views reference iterations
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
